#!/usr/bin/env python
"""Head-to-head playing strength of two search geometries at EQUAL iterations per decision
("new engine vs reference head-to-head win rate 50 % +- 3 sigma over >= 1k games at equal
iterations"; test/benchmark.cpp:73-121 is the reference's own harness of this kind).

Four-player Knockout Whist.  Seats 0 and 2 search, seats 1 and 3 move uniformly at random; in the first
half of the games seat 0 decides with geometry A and seat 2 with geometry B, in the second half they swap
(same deals), so the seat order cancels.  A geometry is "trees x iterations-per-decision[ x lanes]":
`2368x1000000` is what bench.py runs (2368 private trees of 422 iterations), `16x1000000` is the
reference's RootParallel on a 16-core box (execution.h:181-204: one tree per thread), `16x1000000x148` is 16
trees each searched by 148 lanes with virtual loss.

    python tools/head_to_head.py --games 1000 --a 2368x1000000 --b 16x1000000 --out gpurun_out/h2h.jsonl
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def parse_geometry(text: str, max_batch: int):
    from ismcsolver_b200.selfplay import SeatSearch
    parts = [int(p) for p in text.lower().split("x")]
    trees, iterations = parts[0], parts[1]
    lanes = parts[2] if len(parts) > 2 else 0
    return SeatSearch(trees=trees, iterations=iterations, lanes=lanes, max_batch=max_batch, label=text)


def batch_limit(cfg_text: str, pool_gb: float) -> int:
    """Positions per call so that the node pools (32 B slots, load <= 0.5, a power of two per tree) fit `pool_gb`."""
    parts = [int(p) for p in cfg_text.lower().split("x")]
    trees, iterations = parts[0], parts[1]
    lanes = parts[2] if len(parts) > 2 else 0
    per_tree = iterations // trees + 2 + lanes
    cap = 16
    while cap < 2 * per_tree:
        cap *= 2
    return max(1, int(pool_gb * 2**30 // (trees * cap * 32)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--games", type=int, default=1000, help="games per half (A/B, then B/A on the same deals)")
    ap.add_argument("--a", default="2368x1000000")
    ap.add_argument("--b", default="16x1000000")
    ap.add_argument("--seed", type=int, default=2026)
    ap.add_argument("--pool-gb", type=float, default=64.0)
    ap.add_argument("--out", default="")
    ap.add_argument("--search-forced", action="store_true", help="search decisions with one legal move too")
    args = ap.parse_args()

    from ismcsolver_b200 import capi
    from ismcsolver_b200.selfplay import play_match
    eng = capi.Engine(0)
    a = parse_geometry(args.a, batch_limit(args.a, args.pool_gb))
    b = parse_geometry(args.b, batch_limit(args.b, args.pool_gb))
    total = {"a": 0, "b": 0, "random": 0}
    halves = []
    t_all = time.perf_counter()
    for half, (s0, s2) in enumerate(((a, b), (b, a))):
        t0 = time.perf_counter()
        rep = play_match(eng, args.games, 4, {0: s0, 2: s2}, seed=args.seed, skip_forced=not args.search_forced)
        wall = time.perf_counter() - t0
        wins_a = rep.wins[0] if s0 is a else rep.wins[2]
        wins_b = rep.wins[2] if s0 is a else rep.wins[0]
        total["a"] += wins_a; total["b"] += wins_b; total["random"] += rep.wins[1] + rep.wins[3]
        line = {"half": half, "seat0": s0.label, "seat2": s2.label, "wall_seconds": wall, **rep.as_dict()}
        halves.append(line)
        print(json.dumps(line), flush=True)
        if args.out:
            with open(args.out, "a") as f:
                f.write(json.dumps(line) + "\n")
    n = total["a"] + total["b"]
    share = total["a"] / max(1, n)
    sigma = 0.5 / math.sqrt(max(1, n))
    line = {"summary": True, "a": args.a, "b": args.b, "games": 2 * args.games, "wins_a": total["a"],
            "wins_b": total["b"], "wins_random_seats": total["random"],
            "a_share_of_decided": share, "sigma": sigma, "z": (share - 0.5) / sigma,
            "wall_seconds": time.perf_counter() - t_all,
            "a_it_per_s": sum(h["per_seat"]["0" if h["seat0"] == args.a else "2"]["iterations"] for h in halves)
                          / max(1e-9, sum(h["per_seat"]["0" if h["seat0"] == args.a else "2"]["search_seconds"] for h in halves)),
            "b_it_per_s": sum(h["per_seat"]["0" if h["seat0"] == args.b else "2"]["iterations"] for h in halves)
                          / max(1e-9, sum(h["per_seat"]["0" if h["seat0"] == args.b else "2"]["search_seconds"] for h in halves))}
    print(json.dumps(line), flush=True)
    if args.out:
        with open(args.out, "a") as f:
            f.write(json.dumps(line) + "\n")
    eng.close()


if __name__ == "__main__":
    main()
