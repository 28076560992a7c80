#!/usr/bin/env python3
"""Executed warp-instructions of a kernel by source line: joins the SASS page of an ncu report
(`ncu --set full --import-source on`) with the line table of the library the report was taken from.

usage: tools/ncu_by_line.py <report.ncu-rep> <kernel-substring> [--site FILE] [--per N] [--lib PATH]
  --site FILE  group by the line of FILE in the inlining chain (default: innermost line)
  --per N      divide the counts by N (iterations per launch) -> warp-instructions per iteration
The library must be the build the report was taken from (addresses are joined)."""
import argparse
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ap = argparse.ArgumentParser()
ap.add_argument("report")
ap.add_argument("kernel")
ap.add_argument("--site", default="")
ap.add_argument("--per", type=float, default=1.0)
ap.add_argument("--lib", default=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                               "ismcsolver_b200", "libismcts_b200.so"))
ap.add_argument("--top", type=int, default=50)
ap.add_argument("--metric", default="Instructions Executed", help="column of the SASS page to sum (e.g. stall_long_sb, '# Samples')")
a = ap.parse_args()

# 1. address -> (innermost line, site line) from nvdisasm -gi of the kernel's cubin
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(a.lib)], cwd=tmp, check=True, capture_output=True)
line_of = {}
for cubin in os.listdir(tmp):
    dis = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    cur = None
    loc = site = "?"
    group_open = False
    for l in dis.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", l)
        if m:
            cur = m.group(1)
            continue
        if not cur or a.kernel not in cur:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            here = f"{os.path.basename(m.group(1))}:{m.group(2)}"
            if not group_open:
                loc, site, group_open = here, "?", True
            if a.site and os.path.basename(m.group(1)) == a.site and site == "?":
                site = here
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            group_open = False
            line_of[int(m.group(1), 16)] = (loc, site, cur)

# 2. executed counts per address from the report
src = subprocess.run(["ncu", "-i", a.report, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
ia, ie, it = hdr.index("Address"), hdr.index(a.metric), hdr.index("Thread Instructions Executed")
base = None
by = collections.Counter()
thr = collections.Counter()
tot = tot_thr = 0
for r in rows[hdr_i + 1:]:
    if len(r) <= it or not r[ia]:
        continue
    addr = int(r[ia], 16) if r[ia].startswith("0x") else int(r[ia])
    if base is None:
        base = addr
    e, t = float(r[ie] or 0), float(r[it] or 0)
    loc, site, _ = line_of.get(addr - base, ("?", "?", ""))
    key = site if a.site else loc
    by[key] += e
    thr[key] += t
    tot += e
    tot_thr += t
print(f"total warp-instructions {tot:.4g} ({tot / a.per:.1f} per unit), active lanes per instruction {tot_thr / max(tot, 1):.1f}")
for k, v in sorted(by.items(), key=lambda kv: -kv[1])[: a.top]:
    print(f"{v / a.per:12.1f} {100 * v / tot:6.2f}%  lanes {thr[k] / max(v, 1):5.1f}  {k}")
