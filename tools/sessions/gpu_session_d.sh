#!/bin/bash
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out
python -m pytest tests -m gpu -q -x > $O/r02_pytest_e.log 2>&1; tail -4 $O/r02_pytest_e.log
for cfg in "64 2368 0" "1024 2368 0" "74 16 128" "296 16 128" "592 16 128" "592 64 32"; do
  set -- $cfg
  python bench.py --no-extra --no-cpu-baseline --positions $1 --trees $2 --lanes $3 --steps 4 --warmup 3 > $O/r02_d_p$1_t$2_l$3.json 2> $O/r02_d_p$1_t$2_l$3.err
  python -c "
import json
try:
    d = json.load(open('$O/r02_d_p$1_t$2_l$3.json')); print('positions $1 trees $2 lanes $3', 'value %.4g' % d['value'], 'e2e %.4g' % d['e2e']['value'], 'ms/step %.2f' % d['ms_per_step'])
except Exception as e: print('$cfg failed', e); print(open('$O/r02_d_p$1_t$2_l$3.err').read()[-600:])"
done
ncu --set full --clock-control none --import-source on -k regex:search_kernel -s 2 -c 1 -f -o $O/r02_d_hybrid python bench.py --no-extra --no-cpu-baseline --positions 296 --trees 16 --lanes 128 --steps 2 --warmup 1 > $O/r02_ncu_d1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:search_kernel -s 2 -c 1 -f -o $O/r02_d_wide python bench.py --no-extra --no-cpu-baseline --positions 1024 --steps 2 --warmup 1 > $O/r02_ncu_d2.log 2>&1
cp ismcsolver_b200/libismcts_b200.so $O/r02_d_lib.so
ls -la $O/r02_d_*.ncu-rep
