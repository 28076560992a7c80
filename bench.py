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
# ncu (profiles/r01_final9_ncu_raw.txt): warp-instructions executed and DRAM bytes per iteration of the search kernel
NCU_PROFILE = "profiles/r01_final9_ncu_raw.txt"
NCU_WARP_INST_PER_ITERATION = 37755352886.0 / 64e6
NCU_DRAM_BYTES_PER_ITERATION = (80.701630e9 + 29.203759e9) / 64e6
REF_HARNESS = os.path.join(ROOT, "oracle", "_ref", "ref_harness")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--positions", type=int, default=64, help="root positions per step (batch)")
    ap.add_argument("--trees", type=int, default=2368, help="trees per position per GPU")
    ap.add_argument("--iterations", type=int, default=ITERATIONS_PER_DECISION, help="iterations per decision per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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

    B, T = args.positions, args.trees
    stride = 64
    states = [capi.kw_deal(DEAL_SEED, i, 4) for i in range(B)]
    h_roots = torch.from_numpy(capi.pack_states(states)).pin_memory()
    d_roots = h_roots.cuda(non_blocking=True)
    d_out = torch.zeros(3 * B * stride, dtype=torch.int64, device="cuda")     # visits | score2 | avail
    d_iters = torch.zeros(B * T, dtype=torch.int64, device="cuda")
    per = B * stride

    tree_base, my_trees = shard_trees(T * world, world, rank)
    assert my_trees == T

    def desc_for(step):
        # trees are sharded over ranks: rank r owns global trees [r*T, (r+1)*T) of world*T;
        # iterations per position scale with the world size (weak scaling)
        return eng.make_desc(capi.GAME_KW, B, T, iterations=args.iterations * world, seed=0xB200 + step,
                             trees_total=T * world, tree_base=tree_base, flags=capi.FLAG_KW_MAX4)

    def device_step(step):
        eng.search_device(desc_for(step), d_roots.data_ptr(), d_out.data_ptr(), d_out.data_ptr() + 8 * per,
                          d_out.data_ptr() + 16 * per, d_iters.data_ptr())
        merge_root_arrays(d_out)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # -------- device-resident timing
    for w in range(args.warmup):
        device_step(w)
    barrier()
    launches0 = eng.launch_count
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms = []
    ev0.record(stream)
    for s in range(args.steps):
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record(stream)
        eng.search_device(desc_for(args.warmup + s), d_roots.data_ptr(), d_out.data_ptr(), d_out.data_ptr() + 8 * per,
                          d_out.data_ptr() + 16 * per, d_iters.data_ptr())
        k1.record(stream)
        merge_root_arrays(d_out)
        kernel_ms.append((k0, k1))
    ev1.record(stream)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = eng.launch_count - launches0
    elapsed_ms = ev0.elapsed_time(ev1)
    search_ms = sum(a.elapsed_time(b) for a, b in kernel_ms) / len(kernel_ms)
    iters_done = int(d_iters.sum().item())
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    total_iterations = args.iterations * world * B * args.steps          # over all ranks
    value = total_iterations / (elapsed_ms * 1e-3)

    # -------- end to end through the host-buffer C ABI call
    eng.set_stream(None)
    roots_np = h_roots.numpy()
    for w in range(min(2, args.warmup)):
        eng.search(desc_for(w), roots_np)
    barrier()
    t0 = time.perf_counter()
    for s in range(args.steps):
        res = eng.search(desc_for(args.warmup + s), roots_np)
        if world > 1:
            merged = torch.from_numpy(res["visits"].astype(np.int64)).cuda()
            dist.all_reduce(merged)
            merged.cpu()
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = total_iterations / float(t.item())
    h2d = h_roots.numel()
    d2h = 3 * per * 8 + B * T * 8

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        per_launch_iters = args.iterations * world * B / world            # this rank's launch
        achieved = per_launch_iters * ALGO_BYTES_PER_ITERATION / (search_ms * 1e-3) / 1e9
        sm_max = float(peaks.get("sm_max_mhz", 1965.0))
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int32+f64", "data": "synthetic",
            "config": {
                "workload": "Knockout Whist 4 players, SOSolver RootParallel (CudaRootParallel), "
                            "1M iterations per decision",
                "iterations_per_decision": args.iterations * world, "positions_per_step": B,
                "trees_per_position": T * world, "iterations_per_tree": args.iterations / T,
                "wave_width": 1, "deal_seed": hex(DEAL_SEED), "exploration": 0.7,
                "parallelism": f"root-parallel trees sharded over {world} GPU(s), NCCL all-reduce of root arrays"
                               if world > 1 else "root-parallel, one tree per thread",
                "l2": "node pools (>4 GB per step) are far larger than the 126 MB L2; pools are re-zeroed every step",
            },
            "search_kernel_ms": search_ms,
            "iterations_done_last_step": iters_done,
            "roofline": {
                "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "traffic": NCU_DRAM_BYTES_PER_ITERATION * per_launch_iters,
                "traffic_source": "ncu dram__bytes_read.sum + dram__bytes_write.sum of this kernel and workload "
                                  f"({NCU_PROFILE}), bytes per launch",
                "algorithmic_bytes_per_launch": per_launch_iters * ALGO_BYTES_PER_ITERATION,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
                "note": "the path is bound by SM instruction issue, not HBM (see DESIGN.md); "
                        "issue-slot utilisation is in profiles/",
            },
            "issue_roofline": {
                "bound": "sm_issue", "unit": "warp-instructions/s",
                "peak": 148 * 4 * sm_max * 1e6,
                "achieved": per_launch_iters / (search_ms * 1e-3) * NCU_WARP_INST_PER_ITERATION,
                "frac": per_launch_iters / (search_ms * 1e-3) * NCU_WARP_INST_PER_ITERATION / (148 * 4 * sm_max * 1e6),
                "warp_inst_per_iteration": NCU_WARP_INST_PER_ITERATION,
                "note": "the binding roofline: 148 SMs x 4 schedulers x SM clock; instructions per iteration from ncu "
                        f"smsp__inst_executed.sum ({NCU_PROFILE})",
            },
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches,
            "clocks": clocks,
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


def main():
    args = parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        b200_arm(args)


if __name__ == "__main__":
    main()
