#!/bin/bash
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out
python -m pytest tests -m gpu -q > $O/r02_pytest_i.log 2>&1; tail -5 $O/r02_pytest_i.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | cut -c1-400
python bench.py > $O/r02_h_bench.json 2> $O/r02_h_bench.err; python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/r02_h_bench.json').read().splitlines() if l.startswith('{')][-1])
print('value %.4g e2e %.4g ms/step %.1f' % (d['value'], d['e2e']['value'], d['ms_per_step']), 'roofline', {k: d['roofline'][k] for k in ('achieved', 'frac', 'traffic')}, 'issue', d.get('issue_roofline', {}).get('frac'))
print({k: '%.4g' % v['value'] for k, v in d['extra'].items()}, 'cpu', d['cpu_baseline']['value'])
PY
ncu --set full --clock-control none --import-source on -k regex:search_kernel -s 2 -c 1 -f -o $O/r02_h_wide python bench.py --no-extra --no-cpu-baseline --positions 1024 --trees 2368 --lanes 1 --steps 2 --warmup 1 > $O/r02_ncu_h1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:search_kernel -s 9 -c 1 -f -o $O/r02_h_mo python tools/extra_bench.py kw-mo-exp3 > $O/r02_ncu_h2.log 2>&1
cp ismcsolver_b200/libismcts_b200.so $O/r02_h_lib.so
ls -la $O/r02_h_*.ncu-rep
