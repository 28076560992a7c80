#!/usr/bin/env python3
"""Static SASS census of one kernel: instructions per source line (and per inlined function).

Usage: tools/sass_lines.py <file.cu> <kernel-name-substring> [--range 0xLO 0xHI] [--by func|line]
Compiles <file.cu> for sm_100a with -lineinfo into a cubin under /tmp and reads
`nvdisasm -gi` (line info with inlining).  Used to see what a loop of the search
kernel costs before spending GPU time (profiles/*.md quote its numbers).
"""
import argparse
import collections
import os
import re
import subprocess
import sys

ap = argparse.ArgumentParser()
ap.add_argument("cu")
ap.add_argument("kernel")
ap.add_argument("--range", nargs=2, default=None)
ap.add_argument("--by", default="line")
ap.add_argument("--flags", default="")
ap.add_argument("--dump", action="store_true")
ap.add_argument("--site", default="", help="group by the line of this file in the inlining chain")
a = ap.parse_args()

cubin = "/tmp/sass_lines.cubin"
cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-cubin",
       "-Xptxas=-v", "-o", cubin, a.cu] + a.flags.split()
r = subprocess.run(cmd, capture_output=True, text=True)
if r.returncode:
    sys.exit(r.stderr)
for l in r.stderr.splitlines():
    if a.kernel in l or "registers" in l or "spill" in l:
        print(l.strip()[:200])
dis = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout
# split into functions
cur = None
lines = []
for l in dis.splitlines():
    m = re.match(r"\s*\.text\.(\S+):", l)
    if m:
        cur = m.group(1)
        continue
    if cur and a.kernel in cur:
        lines.append(l)
lo, hi = (int(a.range[0], 16), int(a.range[1], 16)) if a.range else (0, 1 << 60)
loc = "?"
site = "?"
group_open = False
count = collections.Counter()
sites = collections.Counter()
total = 0
for l in lines:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        here = f"{os.path.basename(m.group(1))}:{m.group(2)}"
        if not group_open:
            loc = here
            site = "?"
            group_open = True
        if a.site and os.path.basename(m.group(1)) == a.site and site == "?":
            site = here
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        group_open = False
        addr = int(m.group(1), 16)
        if lo <= addr < hi:
            total += 1
            count[loc] += 1
            sites[site] += 1
            if a.dump:
                print(f"{addr:06x} {site:22s} {loc:22s} {m.group(2)[:90]}")
print("instructions in range:", total)
for k, v in sorted((sites if a.site else count).items(), key=lambda kv: -kv[1])[:80]:
    print(f"{v:6d}  {k}")
