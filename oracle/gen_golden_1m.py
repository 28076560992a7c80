#!/usr/bin/env python
"""Generates tests/golden/ref_stats_1m.json from the UNMODIFIED reference (oracle/_ref/ref_harness):
the reference's own headline search — SOSolver<Card, RootParallel>{1,000,000 iterations, 16 threads}
(execution.h:181-204: one tree per thread, i.e. 16 trees of about 62,500 iterations) — on 24
four-player Knockout Whist positions of the synthetic position stream (openings and mid-game roots).

Recorded per position: mean and standard deviation, over the repetitions, of the visit share and of
the win rate (score / visits) of every root move, and how often operator() returned each move.
tests/test_gpu_geometry.py compares the CUDA engine at the reference's geometry (16 trees x 62,500)
AND at the geometry bench.py runs against these numbers.

Run in the build container, where /root/reference exists (`make -C oracle` builds the harness):

    python oracle/gen_golden_1m.py          # ~15 minutes on 8 cores
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

ITERATIONS = 1_000_000
THREADS = 16            # the host cores of the box BENCH_r01.json was measured on
REPS = 16
# (deal stream, random plies played from the opening): SURVEY.md 8(d) position stream
POSITIONS = [(10, 0), (11, 0), (12, 0), (13, 0), (14, 0), (15, 0),
             (16, 3), (17, 5), (18, 8), (19, 10), (20, 13), (21, 15),
             (22, 18), (23, 21), (24, 24), (25, 26), (26, 30), (27, 33),
             (28, 36), (29, 40), (30, 43), (31, 47), (32, 52), (33, 58)]


def main():
    cases = []
    for stream, plies in POSITIONS:
        s = kw_position(stream, plies)
        legal = O.kw_legal(s)
        if len(legal) < 2:
            # a forced move says nothing about the search: move on until there is a choice
            import numpy as np
            rng = np.random.default_rng(stream)
            ctr = [1000]
            while len(legal) == 1:
                O.kw_apply(s, legal[0], 77, stream, ctr)
                legal = O.kw_legal(s)
            if not legal:
                continue
        rows = json.loads(O.ref_harness("kw-stats", s.hex(), "so", "root", ITERATIONS, THREADS, REPS))
        cases.append({"name": f"deal{stream}-ply{plies}", "game": "kw", "state": s.hex(), "solver": "so",
                      "policy": "root", "iterations": ITERATIONS, "threads": THREADS, "reps": REPS,
                      "seconds_mean": sum(r["seconds"] for r in rows) / len(rows), **summarise(rows, 64)})
        print(cases[-1]["name"], len(legal), "legal", cases[-1]["best_counts"], flush=True)
    out = {"generator": "oracle/gen_golden_1m.py",
           "reference": "sfranzen/ismcsolver (unmodified, g++ -O3 -DNDEBUG), SOSolver<Card,RootParallel>{1000000, 16}",
           "cases": cases}
    path = os.path.join(ROOT, "tests", "golden", "ref_stats_1m.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
