#!/usr/bin/env python
"""Generates tests/golden/ref_stats_kw_positions.json from the UNMODIFIED reference (oracle/_ref/ref_harness):
sixteen more four-player Knockout Whist positions of the position stream (deals 40-55 advanced by 2 ... 56 random
plies), searched by the reference's SOSolver<Card, Sequential>{10000}, 30 repetitions each — together with the eight
positions of ref_stats.json the "at least 20 positions" SURVEY.md 8(c) asks for.  Same record format as
oracle/gen_golden.py; tests/test_ref_stats.py compares the oracle against both files.

    python oracle/gen_golden_kw_positions.py        # ~2 minutes
"""
from __future__ import annotations

import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle_lib as O  # noqa: E402
from gen_golden import kw_position, summarise  # noqa: E402


def main():
    cases = []
    for k in range(16):
        stream, plies = 40 + k, 2 + (k * 18) % 56
        s = kw_position(stream, plies)
        legal = O.kw_legal(s)
        ctr = [5000]
        while len(legal) == 1:                     # a forced move says nothing about the search
            O.kw_apply(s, legal[0], 99, stream, ctr)
            legal = O.kw_legal(s)
        if not legal:
            continue
        rows = json.loads(O.ref_harness("kw-stats", s.hex(), "so", "seq", 10000, 1, 30))
        cases.append({"name": f"deal{stream}-ply{plies}", "game": "kw", "state": s.hex(), "solver": "so", "policy": "seq",
                      "iterations": 10000, "threads": 1, "reps": 30, **summarise(rows, 64)})
        print(cases[-1]["name"], len(legal), "legal", cases[-1]["best_counts"], flush=True)
    out = {"generator": "oracle/gen_golden_kw_positions.py",
           "reference": "sfranzen/ismcsolver (unmodified, g++ -O3 -DNDEBUG), SOSolver<Card,Sequential>{10000}", "cases": cases}
    path = os.path.join(ROOT, "tests", "golden", "ref_stats_kw_positions.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
