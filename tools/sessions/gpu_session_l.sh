#!/bin/bash
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out
python -m pytest tests -m gpu -q > $O/r02_pytest_m.log 2>&1; tail -4 $O/r02_pytest_m.log
python bench.py > $O/r02_l_bench.json 2> $O/r02_l_bench.err; python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/r02_l_bench.json').read().splitlines() if l.startswith('{')][-1])
print('value %.4g e2e %.4g ms/step %.1f launches %d' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['gpu_launches']))
print({k: '%.4g' % v['value'] for k, v in d['extra'].items()})
PY
ncu --set full --clock-control none --import-source on -k regex:search_kernel -s 2 -c 1 -f -o $O/r02_l_hybrid python bench.py --no-extra --no-cpu-baseline --steps 2 --warmup 1 > $O/r02_ncu_l1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_l_launches.csv python bench.py --no-extra --no-cpu-baseline --steps 2 --warmup 1 > $O/r02_ncu_l2.log 2>&1
cp ismcsolver_b200/libismcts_b200.so $O/r02_l_lib.so
ls -la $O/r02_l_*.ncu-rep
