#!/usr/bin/env python
"""Measurements of the other BASELINE.json configurations (bench.py measures configs[1]):

  mnk-tree   configs[2]: m,n,k-game 3x3..15x15, SO-ISMCTS, ONE shared tree per position, tree-parallel with virtual
             loss.  Lanes in flight are capped at iterations / 128 (profiles/r02_tree_parallel_quality.md: with
             fewer than about 100 iterations per lane the decision drifts away from the sequential one); a single
             1M-iteration decision can therefore use 7808 lanes = 5 % of a B200, and the GPU is filled by
             searching several positions at once (second line per board).
  kw-mo-exp3 configs[3]: Knockout Whist, MO-ISMCTS (one tree per player), EXP3, root-parallel
(configs[4], self-play with time-limited decisions: tools/selfplay_sweep.py)

Each prints one JSON line per case; with --reference the unmodified reference is timed beside it on the host
cores (oracle/_ref/ref_harness).  Device time is measured with CUDA events around ismcts_search_device."""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.path.join(ROOT, "oracle", "_ref", "ref_harness")
LANES_CAP_K = 128   # tree parallelisation: lanes in flight <= iterations / 128 (include/ismcts/cuda_execution.h: lanesFor)


def device_search(eng, torch, desc, roots_np, stride, reps=3):
    import numpy as np
    d_roots = torch.from_numpy(roots_np).cuda()
    n = desc.n_positions
    d_out = torch.zeros(3 * n * stride, dtype=torch.int64, device="cuda")
    d_it = torch.zeros(n * desc.trees, dtype=torch.int64, device="cuda")
    times = []
    for r in range(reps + 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.search_device(desc, d_roots.data_ptr(), d_out.data_ptr(), d_out.data_ptr() + 8 * n * stride,
                          d_out.data_ptr() + 16 * n * stride, d_it.data_ptr())
        e1.record()
        torch.cuda.synchronize()
        if r:
            times.append(e0.elapsed_time(e1))
    visits = d_out[: n * stride].cpu().numpy().reshape(n, stride)
    return min(times), int(d_it.sum().item()), visits


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("workload", choices=["mnk-tree", "kw-mo-exp3", "kw-hybrid"])
    ap.add_argument("--hybrid", default="16x128,16x148,64x32,148x8,2368x1")
    ap.add_argument("--reference", action="store_true")
    ap.add_argument("--iterations", type=int, default=1_000_000)
    args = ap.parse_args()

    import numpy as np
    import torch
    from ismcsolver_b200 import capi
    eng = capi.Engine(0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    eng.set_stream(stream.cuda_stream)
    cores = int(subprocess.run([REF, "cores"], capture_output=True, text=True).stdout or 1) if os.path.exists(REF) else None

    if args.workload == "mnk-tree":
        lanes = max(32, min(151552, args.iterations // LANES_CAP_K) // 32 * 32)
        for m, k in [(3, 3), (5, 4), (7, 4), (9, 5), (11, 5), (13, 5), (15, 5)]:
            s = capi.mnk_init(m, m, k)
            for batch in (1, max(1, 151552 // lanes)):
                desc = eng.make_desc(capi.GAME_MNK, batch, 1, iterations=args.iterations, seed=3, flags=1 | (lanes << 8))
                ms, done, visits = device_search(eng, torch, desc, capi.pack_states([s] * batch), 256)
                line = {"workload": f"m,n,k {m}x{m}x{k}, SO, one shared tree per position, virtual loss",
                        "positions_in_flight": batch, "lanes_per_tree": lanes, "iterations_per_lane": args.iterations / lanes,
                        "iterations": done, "ms": ms, "iterations_per_s": done / (ms * 1e-3),
                        "ms_per_decision": ms / batch, "best_moves": sorted(set(int(np.argmax(v)) for v in visits))}
                if args.reference and batch == 1 and os.path.exists(REF):
                    rows = json.loads(subprocess.run([REF, "mnk-stats", bytes(s).hex(), "so", "tree", "10000", str(cores), "3"],
                                                     capture_output=True, text=True, check=True).stdout)
                    line["reference_TreeParallel_it_per_s"] = 10000 / min(r["seconds"] for r in rows)
                    line["reference_cores"] = cores
                print(json.dumps(line), flush=True)
    elif args.workload == "kw-mo-exp3":
        B = 64
        states = [capi.kw_deal(0x1535C75B200, i, 4) for i in range(B)]
        for trees in (592, 1184, 2368):
            desc = eng.make_desc(capi.GAME_KW, B, trees, iterations=args.iterations, seed=5, solver=capi.SOLVER_MO,
                                 policy=capi.POLICY_EXP3)
            ms, done, _ = device_search(eng, torch, desc, capi.pack_states(states), 64)
            line = {"workload": "Knockout Whist 4 players, MOSolver (4 trees per lane), EXP3, root-parallel",
                    "positions": B, "trees_per_position": trees, "iterations": done, "ms": ms,
                    "iterations_per_s": done / (ms * 1e-3)}
            print(json.dumps(line), flush=True)
        if args.reference and os.path.exists(REF):
            rows = json.loads(subprocess.run([REF, "kw-time", "4", "mo-exp3", "root", "200000", str(cores), "3"],
                                             capture_output=True, text=True, check=True).stdout)
            print(json.dumps({"reference": "MOSolver<Card,RootParallel,EXP3>{200000}", "cores": cores,
                              "iterations_per_s": 200000 / min(r["seconds"] for r in rows)}), flush=True)
    elif args.workload == "kw-hybrid":
        # root parallelisation with tree parallelisation inside: T trees per decision, each searched by W lanes with
        # virtual loss; positions in flight = resident lanes / (T x W)
        for item in args.hybrid.split(","):
            T, W = (int(x) for x in item.split("x"))
            B = max(1, 151552 // (T * W))
            states = [capi.kw_deal(0x1535C75B200, i, 4) for i in range(B)]
            flags = ((1 | (W << 8)) if W > 1 else 0) | capi.FLAG_KW_MAX4
            desc = eng.make_desc(capi.GAME_KW, B, T, iterations=args.iterations, seed=5, flags=flags)
            ms, done, _ = device_search(eng, torch, desc, capi.pack_states(states), 64)
            print(json.dumps({"workload": "Knockout Whist 4 players, SO-ISMCTS, UCB1: trees x lanes per tree",
                              "positions": B, "trees_per_position": T, "lanes_per_tree": W,
                              "iterations_per_tree": args.iterations / T, "iterations_per_lane": args.iterations / T / W,
                              "iterations": done, "ms": ms, "iterations_per_s": done / (ms * 1e-3)}), flush=True)
    eng.close()


if __name__ == "__main__":
    main()
