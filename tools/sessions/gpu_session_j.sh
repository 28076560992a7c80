#!/bin/bash
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out
python tools/extra_bench.py kw-mo-exp3 --reference > $O/r02_j_mo.jsonl 2> $O/r02_j_mo.err; cut -c100-300 $O/r02_j_mo.jsonl
python -m pytest tests -m gpu -q > $O/r02_pytest_k.log 2>&1; tail -5 $O/r02_pytest_k.log
python bench.py > $O/r02_j_bench.json 2> $O/r02_j_bench.err; python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/r02_j_bench.json').read().splitlines() if l.startswith('{')][-1])
print('value %.4g e2e %.4g ms/step %.1f' % (d['value'], d['e2e']['value'], d['ms_per_step']), 'traffic', d['roofline']['traffic'])
print({k: '%.4g' % v['value'] for k, v in d['extra'].items()})
PY
ncu --set full --clock-control none --import-source on -k regex:search_kernel -s 2 -c 1 -f -o $O/r02_j_hybrid python bench.py --no-extra --no-cpu-baseline --steps 2 --warmup 1 > $O/r02_ncu_j1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_j_launches.csv python bench.py --no-extra --no-cpu-baseline --steps 2 --warmup 1 > $O/r02_ncu_j2.log 2>&1
cp ismcsolver_b200/libismcts_b200.so $O/r02_j_lib.so
ls -la $O/r02_j_*.ncu-rep
