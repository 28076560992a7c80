#!/bin/bash
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out
for MB in 4 6 8; do
  ISMCTS_MIN_BLOCKS_MO=$MB python tools/extra_bench.py kw-mo-exp3 > $O/r02_i_mo_mb$MB.jsonl 2> $O/r02_i_mo_mb$MB.err
  echo "MO min blocks $MB:"; cut -c100-300 $O/r02_i_mo_mb$MB.jsonl
done
python -m pytest tests -m gpu -q > $O/r02_pytest_j.log 2>&1; tail -5 $O/r02_pytest_j.log
ISMCTS_MIN_BLOCKS_MO=8 python -m pytest tests -m gpu -q -k "mo or exp3 or players" > $O/r02_pytest_j8.log 2>&1; tail -3 $O/r02_pytest_j8.log
python bench.py > $O/r02_i_bench.json 2> $O/r02_i_bench.err; python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/r02_i_bench.json').read().splitlines() if l.startswith('{')][-1])
print('value %.4g e2e %.4g ms/step %.1f' % (d['value'], d['e2e']['value'], d['ms_per_step']))
print({k: '%.4g' % v['value'] for k, v in d['extra'].items()}, 'single ms', d['extra']['single_decision']['ms_per_decision'], d['extra']['single_decision'].get('lanes_per_tree'))
PY
ncu --set full --clock-control none --import-source on -k regex:search_kernel -s 2 -c 1 -f -o $O/r02_i_hybrid python bench.py --no-extra --no-cpu-baseline --steps 2 --warmup 1 > $O/r02_ncu_i1.log 2>&1
cp ismcsolver_b200/libismcts_b200.so $O/r02_i_lib.so
ls -la $O/r02_i_*.ncu-rep
