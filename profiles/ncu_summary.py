#!/usr/bin/env python
"""Summarises an ncu report: headline raw metrics of the first kernel in it.
usage: python profiles/ncu_summary.py gpurun_out/prof.ncu-rep [iterations_per_launch [latest.json]]
With a third argument the per-iteration figures bench.py quotes (warp-instructions, DRAM bytes) are written to
that JSON file together with the hash of the sources in the tree (ismcsolver_b200.build.source_hash): run it on
the tree the capture was taken from; bench.py quotes the figures only while the hash still matches."""
import csv, json, os, subprocess, sys, io

rep = sys.argv[1]
iters = float(sys.argv[2]) if len(sys.argv) > 2 else None
latest = sys.argv[3] if len(sys.argv) > 3 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__inst_executed.avg.per_cycle_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'lts__t_bytes.sum', 'l1tex__t_bytes.sum', 'smsp__cycles_active.avg',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active']
d = dict(zip(hdr, zip(units, vals)))
for k in want:
    if k in d:
        print(f"{k:85s} {d[k][1]:>18s} {d[k][0]}")
if iters and 'smsp__inst_executed.sum' in d:
    print("warp-instructions per iteration", float(d['smsp__inst_executed.sum'][1]) / iters)

if iters and latest:
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from ismcsolver_b200.build import source_hash
    f = lambda k: float(d[k][1].replace(",", ""))
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    dram = sum(f(k) * scale.get(d[k][0], 1.0) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    json.dump({"report": os.path.basename(rep), "source_hash": source_hash(), "iterations_per_launch": iters,
               "kernel": d.get("Kernel Name", ("", ""))[1],
               "warp_inst_per_iteration": f("smsp__inst_executed.sum") / iters,
               "dram_bytes_per_iteration": dram / iters,
               "issue_active_pct": f("smsp__issue_active.avg.pct_of_peak_sustained_active"),
               "warps_active_pct": f("sm__warps_active.avg.pct_of_peak_sustained_active"),
               "kernel_ms": f("gpu__time_duration.sum") * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(d["gpu__time_duration.sum"][0], 1.0)},
              open(latest, "w"), indent=1)
    print("wrote", latest)
