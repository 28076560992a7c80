#!/usr/bin/env python
"""BASELINE config 5: full-game self-play with time-limited decisions (test/benchmark.cpp:73-121,144-160 as a
batch on the GPU), and the same decisions played against the unmodified reference.

  sweep   `--games` four-player Knockout Whist games, EVERY decision of every seat searched for `--time-ms` on the
          device clock.  The decisions that are due are searched `--share` at a time: a call gives each of its
          decisions trees = (resident lanes) / share trees, i.e. 1/share of one B200 for the whole budget
          (share 128: 1184 trees per decision).  Under torchrun the GAMES are sharded over the ranks (game g is
          played by rank g % world); there is no data-path collective, only the final sum of the counters.
  versus  `--vs-reference N`: 2 x N games, seat 0 / seat 2 = the engine with the same per-decision budget and share
          against the reference's SOSolver<Card, RootParallel> with the same `--time-ms` on all host cores
          (oracle/_ref/ref_harness kw-moves), seats 1 and 3 random, then the two swap seats.

    python tools/selfplay_sweep.py --games 1000 --time-ms 100 --vs-reference 24 --out gpurun_out/selfplay.jsonl
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.path.join(ROOT, "oracle", "_ref", "ref_harness")
RESIDENT_LANES = 148 * 8 * 128


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--games", type=int, default=1000)
    ap.add_argument("--time-ms", type=float, default=100.0)
    ap.add_argument("--share", type=int, default=128, help="decisions searched together: each gets 1/share of the GPU")
    ap.add_argument("--vs-reference", type=int, default=0, help="games per half of the engine-vs-reference match")
    ap.add_argument("--seed", type=int, default=5)
    ap.add_argument("--out", default="")
    args = ap.parse_args()

    import torch
    from ismcsolver_b200 import capi
    from ismcsolver_b200.selfplay import ReferenceSeat, SeatSearch, play_match

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    eng = capi.Engine(local_rank)
    trees = RESIDENT_LANES // args.share
    cfg = SeatSearch(trees=trees, iterations=0, time_ns=int(args.time_ms * 1e6), max_batch=args.share, label="engine")

    def emit(line):
        if rank == 0:
            print(json.dumps(line), flush=True)
            if args.out:
                with open(args.out, "a") as f:
                    f.write(json.dumps(line) + "\n")

    if args.games > 0:
        mine = len(range(rank, args.games, world))
        t0 = time.perf_counter()
        # rank r plays the deals r, r + world, ... (first_game/stride are deal ids: any disjoint split will do)
        rep = play_match(eng, mine, 4, {s: cfg for s in range(4)}, seed=args.seed, first_game=rank * 100_000)
        wall = time.perf_counter() - t0
        stats = torch.tensor(rep.wins + [rep.decisions, rep.iterations, rep.search_calls, rep.plies], dtype=torch.int64,
                             device="cuda")
        walls = torch.tensor([wall], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(stats)
            dist.all_reduce(walls, op=dist.ReduceOp.MAX)
        v = stats.tolist()
        wall = float(walls.item())
        emit({"workload": f"self-play sweep: {args.games} games x {args.time_ms} ms per decision, all 4 seats search, "
                          f"{world} GPU(s), games sharded over ranks",
              "games": args.games, "n_gpus": world, "time_ms_per_decision": args.time_ms, "gpu_share_per_decision": f"1/{args.share}",
              "trees_per_decision": trees, "wins": v[:4], "decisions": v[4], "iterations": v[5], "search_calls": v[6],
              "plies": v[7], "wall_seconds": wall, "decisions_per_second": v[4] / wall,
              "iterations_per_decision_mean": v[5] / max(1, v[4]),
              "iterations_per_second": v[5] / wall,
              "reference_decisions_per_second": 1e3 / args.time_ms,
              "reference_note": "the reference decides one position at a time: 1 / time-ms decisions per second on a box, whatever its cores"})

    if args.vs_reference > 0 and rank == 0:
        cores = int(subprocess.run([REF, "cores"], capture_output=True, text=True, check=True).stdout)
        total = {"engine": 0, "reference": 0, "random": 0}
        ref_it = []
        for half in range(2):
            ref = ReferenceSeat(REF, "ms", args.time_ms, cores)
            seats = {0: cfg, 2: ref} if half == 0 else {0: ref, 2: cfg}
            t0 = time.perf_counter()
            rep = play_match(eng, args.vs_reference, 4, seats, seed=args.seed + 1, skip_forced=True)
            ref.close()
            e_seat, r_seat = (0, 2) if half == 0 else (2, 0)
            total["engine"] += rep.wins[e_seat]; total["reference"] += rep.wins[r_seat]
            total["random"] += rep.wins[1] + rep.wins[3]
            emit({"half": half, "engine_seat": e_seat, "wins": rep.wins, "wall_seconds": time.perf_counter() - t0,
                  "engine_decisions": rep.per_seat[e_seat]["decisions"],
                  "engine_iterations_per_decision": rep.per_seat[e_seat]["iterations"] / max(1, rep.per_seat[e_seat]["decisions"]),
                  "reference_decisions": rep.per_seat[r_seat]["decisions"], "reference_seconds": ref.seconds})
        n = total["engine"] + total["reference"]
        emit({"summary": "engine (time mode, 1/%d of a B200 per decision) vs reference RootParallel on %d host threads, "
                         "%g ms per decision each" % (args.share, cores, args.time_ms),
              "games": 2 * args.vs_reference, **total, "engine_share_of_decided": total["engine"] / max(1, n),
              "sigma": 0.5 / math.sqrt(max(1, n))})
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
