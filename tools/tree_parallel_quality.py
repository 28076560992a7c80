#!/usr/bin/env python
"""Decision quality of tree parallelisation (CudaTreeParallel: ONE shared tree, L iterations in flight with virtual
loss) against the number of lanes in flight, at a fixed budget.

For m,n,k boards (BASELINE config 3: empty board, SO-ISMCTS, UCB1) and the Knockout Whist golden opening, and for
every L in --lanes, `--seeds` independent searches (same position, different Philox streams) are run in one engine
call; L = 1 is the sequential search (bit-exact with the oracle, tests/test_gpu_parity.py).  Reported per
(position, L): the decisions, how many of them equal the most frequent decision of the sequential searches,
and the share of the root visits the decision got.  Every lane in flight adds a virtual loss, so with L close to
the budget the first iterations of every lane search an empty tree: the data behind the cap
`lanes <= iterations / k` in include/ismcts/cuda_execution.h (lanesFor) and tools/extra_bench.py.

    python tools/tree_parallel_quality.py --iterations 1000000 --out gpurun_out/tp_quality.jsonl
"""
from __future__ import annotations

import argparse
import collections
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def canonical_cell(move: int, m: int) -> int:
    """The smallest cell index among the images of `move` under the eight symmetries of an m x m board (an empty
    board does not distinguish them)."""
    r, c = divmod(move, m)
    images = []
    for rr, cc in ((r, c), (c, m - 1 - r), (m - 1 - r, m - 1 - c), (m - 1 - c, r)):
        images += [rr * m + cc, rr * m + (m - 1 - cc)]
    return min(images)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iterations", type=int, default=1_000_000)
    ap.add_argument("--seeds", type=int, default=8)
    ap.add_argument("--sizes", default="5x4,7x4,9x5,11x5,13x5")
    ap.add_argument("--lanes", default="1,32,256,2048,8192,32768,151552")
    ap.add_argument("--out", default="")
    args = ap.parse_args()

    import numpy as np
    from ismcsolver_b200 import capi
    eng = capi.Engine(0)
    golden = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_stats.json")))["cases"][0]["state"]
    kw = capi.KwState.from_buffer_copy(bytes.fromhex(golden))
    boards = []
    for item in args.sizes.split(","):
        m, k = (int(x) for x in item.split("x"))
        boards.append((f"mnk {m}x{m}x{k}", capi.mnk_init(m, m, k)))
    lanes_list = [int(x) for x in args.lanes.split(",")]
    sequential = {}

    def emit(line):
        print(json.dumps(line), flush=True)
        if args.out:
            with open(args.out, "a") as f:
                f.write(json.dumps(line) + "\n")

    for lanes in lanes_list:
        for game, named in ((capi.GAME_MNK, boards), (capi.GAME_KW, [("kw golden opening", kw)])):
            if not named:
                continue
            states = [s for _, s in named for _ in range(args.seeds)]
            desc = eng.make_desc(game, len(states), 1, iterations=args.iterations, seed=20262,
                                 flags=capi.FLAG_SHARED_TREE | (lanes << capi.FLAG_LANES_SHIFT))
            t0 = time.perf_counter()
            res = eng.search(desc, capi.pack_states(states))
            dt = time.perf_counter() - t0
            for i, (name, _) in enumerate(named):
                rows = range(i * args.seeds, (i + 1) * args.seeds)
                best = [int(res["best_move"][r]) for r in rows]
                if game == capi.GAME_MNK:
                    best = [canonical_cell(b, named[i][1].m) for b in best]
                share = [float(res["visits"][r].max()) / float(res["visits"][r].sum()) for r in rows]
                done = [int(res["iterations_done"][r].sum()) for r in rows]
                if lanes == 1:
                    sequential[name] = collections.Counter(best).most_common(1)[0][0]
                ref = sequential.get(name)
                emit({"position": name, "lanes": lanes, "iterations": args.iterations,
                      "iterations_per_lane": args.iterations / lanes, "decisions": best,
                      "sequential_decision": ref,
                      "agree_with_sequential": None if ref is None else sum(b == ref for b in best) / len(best),
                      "top_move_visit_share_mean": float(np.mean(share)), "iterations_done": done,
                      "call_seconds": dt})
    eng.close()


if __name__ == "__main__":
    main()
