#!/bin/bash
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out
python -m pytest tests -m gpu -q -x > $O/r02_pytest_g.log 2>&1; tail -4 $O/r02_pytest_g.log
run() { name=$1; shift
  python bench.py --no-extra --no-cpu-baseline --steps 4 --warmup 3 "$@" > $O/r02_f_$name.json 2> $O/r02_f_$name.err
  python -c "
import json
try:
    d = json.load(open('$O/r02_f_$name.json')); print('$name', 'value %.4g' % d['value'], 'e2e %.4g' % d['e2e']['value'], 'ms/step %.2f' % d['ms_per_step'])
except Exception as e: print('$name failed', e); print(open('$O/r02_f_$name.err').read()[-600:])"
}
run hybrid_p74 --positions 74 --trees 16 --lanes 128
run hybrid_p592 --positions 592 --trees 16 --lanes 128
run hybrid32x128_p296 --positions 296 --trees 32 --lanes 128
run hybrid64x32_p592 --positions 592 --trees 64 --lanes 32
run wide_p1024 --positions 1024
python tools/geometry_agreement.py --geometries 16x128,32x128 --out $O/r02_f_geometry.jsonl > $O/r02_f_geometry.log 2>&1; cut -c1-330 $O/r02_f_geometry.log
ncu --set full --clock-control none --import-source on -k regex:search_kernel -s 2 -c 1 -f -o $O/r02_f_hybrid python bench.py --no-extra --no-cpu-baseline --positions 296 --trees 16 --lanes 128 --steps 2 --warmup 1 > $O/r02_ncu_f1.log 2>&1
cp ismcsolver_b200/libismcts_b200.so $O/r02_f_lib.so
ls -la $O/r02_f_*.ncu-rep
