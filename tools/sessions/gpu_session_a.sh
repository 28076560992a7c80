#!/bin/bash
# GPU session: tests, persistent-grid sweep, full bench, tree-parallel quality, ncu capture.
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out
python -m pytest tests -m gpu -q > $O/r02_pytest_b.log 2>&1; tail -15 $O/r02_pytest_b.log
for P in 64 128 256 512 1024; do
  python bench.py --no-extra --no-cpu-baseline --positions $P --steps 5 --warmup 3 > $O/r02_sweep_p$P.json 2> $O/r02_sweep_p$P.err
  python - <<PY
import json
try:
    d = json.load(open("$O/r02_sweep_p$P.json"))
    print("positions", $P, "value %.4g" % d["value"], "e2e %.4g" % d["e2e"]["value"], "ms/step %.2f" % d["ms_per_step"])
except Exception as e:
    print("positions", $P, "failed", e)
PY
done
python bench.py --positions 256 > $O/r02_bench_a.json 2> $O/r02_bench_a.err; tail -c 1500 $O/r02_bench_a.json
timeout 900 python tools/tree_parallel_quality.py --out $O/r02_tp_quality.jsonl > $O/r02_tp_quality.log 2>&1; tail -3 $O/r02_tp_quality.log | cut -c1-400
ncu --set full --clock-control none --import-source on -k regex:search_kernel -s 2 -c 1 -f -o $O/r02_a python bench.py --no-extra --no-cpu-baseline --positions 256 --steps 2 --warmup 1 > $O/r02_ncu_a.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_a_launches.csv python bench.py --no-extra --no-cpu-baseline --positions 256 --steps 2 --warmup 1 > $O/r02_ncu_a2.log 2>&1
ls -la $O/r02_a.ncu-rep
