#!/usr/bin/env python
"""bench.py — ISMCTS iterations/sec on Knockout Whist (BASELINE.json metric).

Workload (BASELINE.json configs[1]): Knockout Whist, 4 players, SO-ISMCTS root-parallel, 1,000,000
iterations per decision.  One step = one search call over a batch of `--positions` root positions
from the synthetic deal stream, each searched with 1M iterations spread over
`--trees` independent trees.  At N GPUs every rank searches the same positions with its own shard
of the trees (N x trees trees and N x 1M iterations per position: weak scaling) and the per-GPU root
arrays are merged by one NCCL all-reduce per step (RootParallel::compileVisitCounts,
execution.h:219-231).

`value`   : device-resident — positions and outputs stay in HBM, CUDA events on the launch stream.
`e2e`     : the same step through the host-buffer C ABI call (ismcts_search): host->device copy of
            the positions and device->host copy of the root arrays inside the timed region.
`--impl reference`: the UNMODIFIED reference (oracle/_ref/ref_harness, SOSolver<Card,RootParallel>
            on all host cores) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ISMCTS iterations/sec (Knockout Whist)"
UNIT = "iterations/s"
ITERATIONS_PER_DECISION = 1_000_000
DEAL_SEED = 0x1535C75B200           # printed with every line (config.deal_seed)
# algorithmic node traffic of one KW4 iteration (select 1.28 KB + expand 232 B + backprop 80 B)
ALGO_BYTES_PER_ITERATION = 1600.0
REF_HARNESS = os.path.join(ROOT, "oracle", "_ref", "ref_harness")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--positions", type=int, default=0, help="root positions per step (batch); 0 = eight waves of blocks")
    ap.add_argument("--trees", type=int, default=0,
                    help="trees per position per GPU; 0 = the host's hardware threads, i.e. the trees the reference's "
                         "RootParallel builds on this box (execution.h:160-167,181-189)")
    ap.add_argument("--lanes", type=int, default=128,
                    help="lanes per tree: every tree is searched by a block of lanes with virtual loss; 1 = one lane per tree")
    ap.add_argument("--wide-trees", type=int, default=2368, help="extra leg `wide`: private trees per position (one lane each)")
    ap.add_argument("--wide-positions", type=int, default=1024)
    ap.add_argument("--iterations", type=int, default=ITERATIONS_PER_DECISION, help="iterations per decision per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra workloads (position stream, MO+EXP3, ...)")
    ap.add_argument("--mo-trees", type=int, default=4736, help="extra leg MO+EXP3: lanes (sets of trees) per position")
    ap.add_argument("--single-trees", type=int, default=3906, help="extra leg single decision with --lanes 1: private trees")
    return ap.parse_args()


# ----------------------------------------------------------------------------- clocks

class ClockSampler:
    """Samples SM clocks and throttle reasons with nvidia-smi during the timed region."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) == 6:
                self.rows.append(parts)

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        clocks = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        reasons = set()
        for r in self.rows:
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": clocks[len(clocks) // 2] if clocks else None,
                "sm_max_mhz": int(self.rows[0][1]) if self.rows and self.rows[0][1].isdigit() else None,
                "samples": len(clocks), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------- reference arm

def host_cores() -> int:
    try:
        return int(subprocess.run([REF_HARNESS, "cores"], capture_output=True, text=True, check=True).stdout)
    except Exception:
        return os.cpu_count() or 1


def run_reference_sample(iterations: int, reps: int) -> dict:
    """Times the unmodified reference: SOSolver<Card,RootParallel>{iterations, all cores} on KW4 opening
    decisions (fresh deal per repetition), steady_clock around operator() only."""
    cores = host_cores()
    out = subprocess.run([REF_HARNESS, "kw-time", "4", "so", "root", str(iterations), str(cores), str(reps)],
                         capture_output=True, text=True, check=True, timeout=1800).stdout
    rows = json.loads(out)
    secs = [r["seconds"] for r in rows]
    return {"cores": cores, "seconds": secs, "iterations": iterations,
            "it_per_s_mean": iterations * len(secs) / sum(secs), "it_per_s_best": iterations / min(secs)}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if not os.path.exists(REF_HARNESS):
        # the reference harness is compiled by build() where /root/reference exists and travels in oracle/_ref
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/ref_harness was not built"}))
        return
    total = max(1, args.warmup) + args.steps
    rows = run_reference_sample(args.iterations, total)
    timed = rows["seconds"][-args.steps:]
    value = args.iterations * len(timed) / sum(timed)
    sample = (f"{args.steps} decisions x {args.iterations} iterations, SOSolver<Card,RootParallel> "
              f"on {rows['cores']} threads, KW4 opening deals")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(timed) / len(timed),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32+f64",
        "data": "synthetic (reference's own random deals)",
        "config": {"workload": "Knockout Whist 4 players, SOSolver RootParallel, 1M iterations per decision",
                   "iterations_per_decision": args.iterations, "positions_per_step": 1},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": rows["cores"], "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------- B200 arm

def position_stream(capi, n: int, seed: int = DEAL_SEED):
    """SURVEY.md 8(d) position stream: deal i advanced by d ~ U{0..60} uniformly random legal moves, drawn again
    if the game has ended by then."""
    import numpy as np
    rng = np.random.default_rng(seed & 0xFFFFFFFF)
    out = []
    i = 0
    while len(out) < n:
        s = capi.kw_deal(seed, 1_000_000 + i, 4)
        ctr = [0]
        for _ in range(int(rng.integers(0, 61))):
            mask = capi.kw_legal_mask(s)
            if not mask:
                break
            legal = [c for c in range(52) if mask >> c & 1]
            capi.kw_apply(s, int(rng.choice(legal)), seed, 1_000_000 + i, ctr)
        i += 1
        if capi.kw_legal_mask(s):
            out.append(s)
    return out


def ncu_figures():
    """Per-iteration counters of the newest ncu capture (profiles/ncu_latest.json, written by
    profiles/ncu_summary.py) — only if it profiled the code this library is built from."""
    from ismcsolver_b200.build import source_hash
    try:
        latest = json.load(open(os.path.join(ROOT, "profiles", "ncu_latest.json")))
    except Exception:
        return None, "no profiles/ncu_latest.json"
    if latest.get("source_hash") != source_hash():
        return None, f"profiles/{latest.get('report')} profiled other sources ({latest.get('source_hash')}) than this library"
    return latest, f"profiles/{latest.get('report')}"


def b200_arm(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from ismcsolver_b200 import capi
    from ismcsolver_b200.distributed import merge_root_arrays, shard_trees

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    eng = capi.Engine(local_rank)
    # a dedicated (non-default) stream: the engine launches on it and the CUDA events are recorded on it
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    eng.set_stream(stream.cuda_stream)
    stride = 64

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    class Leg:
        """One workload: `states` searched with `trees_total` trees per position sharing `iterations` iterations,
        the trees sharded over the ranks, root arrays merged by one all-reduce per step."""

        def __init__(self, states, trees_total, iterations, solver=capi.SOLVER_SO, policy=capi.POLICY_UCB1, lanes=0):
            self.B = len(states)
            self.flags = capi.FLAG_KW_MAX4 | ((capi.FLAG_SHARED_TREE | (lanes << capi.FLAG_LANES_SHIFT)) if lanes > 1 else 0)
            self.trees_total, self.iterations = trees_total, iterations
            self.tree_base, self.trees = shard_trees(trees_total, world, rank)
            self.solver, self.policy = solver, policy
            self.h_roots = torch.from_numpy(capi.pack_states(states)).pin_memory()
            self.d_roots = self.h_roots.cuda(non_blocking=True)
            self.per = self.B * stride
            self.d_out = torch.zeros(3 * self.per, dtype=torch.int64, device="cuda")     # visits | score2 | avail
            self.d_iters = torch.zeros(self.B * self.trees, dtype=torch.int64, device="cuda")
            # the iterations this rank's trees run (execution.h:42-52: an even split, the first trees one more)
            q, r = divmod(iterations, trees_total)
            self.my_iterations = self.B * (q * self.trees + max(0, min(r, self.tree_base + self.trees) - min(r, self.tree_base)))

        def desc(self, step):
            return eng.make_desc(capi.GAME_KW, self.B, self.trees, iterations=self.iterations, seed=0xB200 + step,
                                 solver=self.solver, policy=self.policy, trees_total=self.trees_total,
                                 tree_base=self.tree_base, flags=self.flags)

        def launch(self, step):
            p = self.d_out.data_ptr()
            eng.search_device(self.desc(step), self.d_roots.data_ptr(), p, p + 8 * self.per, p + 16 * self.per,
                              self.d_iters.data_ptr())

        def device_timed(self, steps, warmup, first_step=0):
            """(max-over-ranks ms for `steps` steps incl. the all-reduces, mean ms of this rank's search launches)"""
            eng.set_stream(stream.cuda_stream)
            for w in range(warmup):
                self.launch(first_step + w)
                merge_root_arrays(self.d_out)
            barrier()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            marks = []
            ev0.record(stream)
            for s in range(steps):
                k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                k0.record(stream)
                self.launch(first_step + warmup + s)
                k1.record(stream)
                merge_root_arrays(self.d_out)
                marks.append((k0, k1))
            ev1.record(stream)
            barrier()
            done = int(self.d_iters.sum().item())
            # every iteration asked for was run (`value` counts nominal iterations)
            assert done == self.my_iterations, f"rank {rank}: {done} iterations done, {self.my_iterations} expected"
            return max_over_ranks(ev0.elapsed_time(ev1)), sum(a.elapsed_time(b) for a, b in marks) / len(marks)

        def e2e_timed(self, steps, warmup, first_step=0):
            """Seconds (max over ranks) for `steps` host-buffer calls: H2D of the positions, search, D2H of the root
            arrays, and at N > 1 the all-reduce of visits, scores and availabilities."""
            eng.set_stream(None)
            roots_np = self.h_roots.numpy()
            for w in range(warmup):
                eng.search(self.desc(first_step + w), roots_np)
            barrier()
            t0 = time.perf_counter()
            for s in range(steps):
                res = eng.search(self.desc(first_step + warmup + s), roots_np)
                assert int(res["iterations_done"].sum()) == self.my_iterations
                if world > 1:
                    merged = torch.from_numpy(np.stack([res["visits"], res["score2"], res["avail"]]).astype(np.int64)).cuda()
                    dist.all_reduce(merged)
                    merged.cpu()
            barrier()
            return max_over_ranks(time.perf_counter() - t0)

    T = args.trees or host_cores()
    lanes = max(1, args.lanes)
    if lanes > 1:
        # never fewer than 128 iterations per lane (profiles/r02_tree_parallel_quality.md): on a host with very many
        # threads the trees are small and get fewer lanes
        cap = args.iterations // T // 128
        lanes = max(1, min(lanes, cap // 32 * 32 if cap >= 32 else cap))
    resident_blocks = 148 * 8
    B = args.positions or (max(1, 8 * resident_blocks * 128 // (T * lanes)) if lanes > 1 else 1024)
    iterations = args.iterations * world                      # weak scaling: every GPU adds its own 1M per decision
    main = Leg([capi.kw_deal(DEAL_SEED, i, 4) for i in range(B)], T * world, iterations, lanes=lanes)
    assert main.trees == T

    # -------- device-resident timing
    sampler = ClockSampler(local_rank)
    for w in range(args.warmup):
        main.launch(w)
        merge_root_arrays(main.d_out)
    barrier()
    if rank == 0:
        sampler.start()
    launches0 = eng.launch_count
    elapsed_ms, search_ms = main.device_timed(args.steps, 0, first_step=args.warmup)
    launches = eng.launch_count - launches0
    clocks = sampler.stop() if rank == 0 else None
    total_iterations = iterations * B * args.steps            # over all ranks
    value = total_iterations / (elapsed_ms * 1e-3)

    # -------- end to end through the host-buffer C ABI call
    e2e_s = main.e2e_timed(args.steps, min(2, args.warmup))
    e2e_value = total_iterations / e2e_s
    h2d = main.h_roots.numel()
    d2h = 3 * main.per * 8 + B * T * 8

    # -------- the other workloads north_star names, a few hundred ms each
    extra = {}
    if not args.no_extra:
        def leg_line(leg, steps=2, warmup=1, **more):
            ms, _ = leg.device_timed(steps, warmup, first_step=100)
            return {"value": leg.iterations * leg.B * steps / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / steps,
                    "positions_per_step": leg.B, "trees_per_position": leg.trees_total,
                    "iterations_per_decision": leg.iterations, **more}
        pos_states = position_stream(capi, B)
        extra["position_stream"] = leg_line(
            Leg(pos_states, T * world, iterations, lanes=lanes),
            workload="KW4 mid-game roots: deal i + d ~ U{0..60} random plies (SURVEY 8d), SO-ISMCTS root-parallel",
            mean_players_alive=float(np.mean([bin(s.alive_mask).count("1") for s in pos_states])))
        extra["mo_exp3"] = leg_line(
            Leg(main_states(capi, 32), args.mo_trees * world, iterations,
                solver=capi.SOLVER_MO, policy=capi.POLICY_EXP3),
            workload="KW4 MOSolver<Card, RootParallel, EXP3> (BASELINE config 4): one tree per player, trees sharded over the ranks")
        if T % world == 0:
            extra["strong_scaling"] = leg_line(
                Leg([capi.kw_deal(DEAL_SEED, i, 4) for i in range(B)], T, args.iterations, lanes=lanes),
                workload=f"the N=1 job (1M iterations per decision over {T} trees) split over {world} GPU(s)")
        # the throughput geometry: many private trees per decision, one lane each (profiles/r02_head_to_head.md: equal
        # playing strength, the reference's decision on 16 of 17 decisive golden positions)
        extra["wide"] = leg_line(
            Leg([capi.kw_deal(DEAL_SEED, i, 4) for i in range(args.wide_positions)], args.wide_trees * world, iterations),
            workload=f"1M iterations per decision over {args.wide_trees} private trees of "
                     f"{args.iterations // args.wide_trees} iterations, one lane per tree, persistent grid")
        # one decision at a time through the host-buffer call, as SOSolver<Card, CudaRootParallel>{1M}(game) issues it
        single_lanes = max(1, min(iterations // (T * world) // 128, 148 * 8 * 128 // (T * world))) if lanes > 1 else 1
        single = Leg([capi.kw_deal(DEAL_SEED, 0, 4)], (T if lanes > 1 else args.single_trees) * world, iterations, lanes=single_lanes)
        s_sec = single.e2e_timed(5, 2, first_step=200)
        extra["single_decision"] = {"value": iterations * 5 / s_sec, "unit": UNIT, "ms_per_decision": 1e3 * s_sec / 5,
                                    "trees": single.trees_total, "lanes_per_tree": single_lanes,
                                    "iterations_per_decision": iterations,
                                    "workload": "ONE position per call, host buffers (ismcts_search), wall clock: what "
                                                "SOSolver<Card, CudaRootParallel>{1M}(game) issues (lanes per tree capped at "
                                                "iterations / trees / 128)"}

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        per_launch_iters = main.my_iterations                            # this rank's launch
        achieved = per_launch_iters * ALGO_BYTES_PER_ITERATION / (search_ms * 1e-3) / 1e9
        sm_max = float(peaks.get("sm_max_mhz", 1965.0))
        ncu, ncu_source = ncu_figures()
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int32+f64", "data": "synthetic",
            "config": {
                "workload": "Knockout Whist 4 players, SOSolver RootParallel (CudaRootParallel), "
                            + (f"1M iterations per decision per GPU = {iterations} per decision on {world} GPUs (weak scaling)"
                               if world > 1 else "1M iterations per decision"),
                "iterations_per_decision": iterations, "positions_per_step": B,
                "trees_per_position": T * world, "iterations_per_tree": iterations / (T * world),
                "trees_source": ("the hardware threads of this box's host: the trees the reference's RootParallel builds here "
                                 "(execution.h:160-167,181-189), and what `--impl reference` runs") if not args.trees else "--trees",
                "lanes_per_tree": lanes, "iterations_per_lane": iterations / (T * world) / lanes,
                "wave_width": lanes, "deal_seed": hex(DEAL_SEED), "exploration": 0.7,
                "parallelism": ("root-parallel trees, each searched by a block of lanes with virtual loss"
                                if lanes > 1 else "root-parallel, one lane per tree")
                               + (f"; trees sharded over {world} GPUs, NCCL all-reduce of root arrays" if world > 1 else ""),
                "grid": ("one block per tree, %d waves of blocks" % max(1, B * T * lanes // (148 * 8 * 128))) if lanes > 1
                        else ("persistent: warps claim 32 trees at a time and reuse their node tables"
                              if B * T > 148 * 8 * 128 else "one tree per lane, one wave"),
                "l2": "the node tables in use (%.0f GB) are far larger than the 126 MB L2; every tree starts from a zeroed table"
                      % ((B * T * (32 << max(4, (2 * (iterations // (T * world) + 2 + lanes) - 1).bit_length())) / 2**30)
                         if lanes > 1 else 4.97),
                "geometry_evidence": "profiles/r02_head_to_head.md, tests/test_gpu_geometry.py: decisions equal the unmodified "
                                     "reference's on 17 of 17 decisive golden positions; `extra.wide` is round 1's geometry",
            },
            "search_kernel_ms": search_ms,
            "iterations_done_per_step": per_launch_iters,
            "roofline": {
                "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "traffic": ncu["dram_bytes_per_iteration"] * per_launch_iters if ncu else None,
                "traffic_source": "ncu dram__bytes_read.sum + dram__bytes_write.sum per iteration of the search kernel x "
                                  f"the iterations of this launch ({ncu_source})" if ncu else ncu_source,
                "algorithmic_bytes_per_launch": per_launch_iters * ALGO_BYTES_PER_ITERATION,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
                "note": "the path is bound by SM instruction issue, not HBM (see DESIGN.md); "
                        "issue-slot utilisation is in profiles/",
            },
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches,
            "clocks": clocks,
            "extra": extra,
        }
        if ncu:
            rate = per_launch_iters / (search_ms * 1e-3) * ncu["warp_inst_per_iteration"]
            line["issue_roofline"] = {
                "bound": "sm_issue", "unit": "warp-instructions/s", "peak": 148 * 4 * sm_max * 1e6, "achieved": rate,
                "frac": rate / (148 * 4 * sm_max * 1e6), "warp_inst_per_iteration": ncu["warp_inst_per_iteration"],
                "note": "the binding roofline: 148 SMs x 4 schedulers x SM clock; instructions per iteration from ncu "
                        f"smsp__inst_executed.sum ({ncu_source}), rate measured live",
            }
        if not args.no_cpu_baseline and world == 1 and os.path.exists(REF_HARNESS):
            try:
                ref = run_reference_sample(ITERATIONS_PER_DECISION, 6)
                line["cpu_baseline"] = {
                    "value": ref["it_per_s_mean"], "unit": UNIT, "cores": ref["cores"], "kind": "reference",
                    "sample": f"6 decisions x {ITERATIONS_PER_DECISION} iterations, unmodified reference "
                              f"SOSolver<Card,RootParallel> on {ref['cores']} threads (oracle/_ref/ref_harness)",
                    "best": ref["it_per_s_best"],
                }
            except Exception as ex:  # noqa: BLE001
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": None, "kind": "reference",
                                        "sample": f"failed: {ex}"}
        print(json.dumps(line))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def main_states(capi, n):
    return [capi.kw_deal(DEAL_SEED, i, 4) for i in range(n)]


def main():
    args = parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        b200_arm(args)


if __name__ == "__main__":
    main()
