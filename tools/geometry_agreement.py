#!/usr/bin/env python
"""Decisions and root statistics of the engine at several search geometries against the golden positions of the
unmodified reference (tests/golden/ref_stats_1m.json), with the device time of each geometry.

    python tools/geometry_agreement.py --geometries 16,2368,16x128 --out gpurun_out/geometry.jsonl

A geometry is `trees` (private trees) or `trees x lanes` (every tree searched by that many lanes with virtual loss)."""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--geometries", default="16,2368,16x128")
    ap.add_argument("--seeds", type=int, default=8)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    from ismcsolver_b200 import agreement as A
    from ismcsolver_b200 import capi
    eng = capi.Engine(0)
    golden = A.load_golden()
    for item in args.geometries.split(","):
        parts = [int(x) for x in item.split("x")]
        trees, lanes = parts[0], parts[1] if len(parts) > 1 else 0
        t0 = time.perf_counter()
        res = A.search_golden_positions(eng, golden, trees, lanes, seeds=args.seeds)
        dt = time.perf_counter() - t0
        summary = A.agreement(golden, res)
        line = {"geometry": item, "trees": trees, "lanes_per_tree": lanes or 1, "seeds": args.seeds, "call_seconds": dt,
                **{k: v for k, v in summary.items() if k != "rows"},
                "disagreements": [{k: r[k] for k in ("name", "decisive", "reference", "engine", "agree_per_seed", "regret_max")}
                                  for r in summary["rows"] if not r["agree_modal"] or r["agree_per_seed"] < 1.0]}
        print(json.dumps(line), flush=True)
        if args.out:
            with open(args.out, "a") as f:
                f.write(json.dumps(line) + "\n")
    eng.close()


if __name__ == "__main__":
    main()
