#!/bin/bash
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out
python -m pytest tests -m gpu -q -x > $O/r02_pytest_f.log 2>&1; tail -4 $O/r02_pytest_f.log
run() { # name, lib, args...
  name=$1; lib=$2; shift 2
  python tools/bench_with_lib.py $lib --no-extra --no-cpu-baseline --steps 4 --warmup 3 "$@" > $O/r02_e_$name.json 2> $O/r02_e_$name.err
  python -c "
import json
try:
    d = json.load(open('$O/r02_e_$name.json')); print('$name', 'value %.4g' % d['value'], 'e2e %.4g' % d['e2e']['value'], 'ms/step %.2f' % d['ms_per_step'])
except Exception as e: print('$name failed', e); print(open('$O/r02_e_$name.err').read()[-600:])"
}
L=ismcsolver_b200/libismcts_b200.so; X=ismcsolver_b200/libismcts_b200_exp.so
run hybrid_p74 $L --positions 74 --trees 16 --lanes 128
run hybrid_p296 $L --positions 296 --trees 16 --lanes 128
run hybrid_p592 $L --positions 592 --trees 16 --lanes 128
run hybrid64x32_p592 $L --positions 592 --trees 64 --lanes 32
run wide_p64 $L --positions 64
run wide_p1024 $L --positions 1024
run wide_exp_p64 $X --positions 64
run wide_exp_p1024 $X --positions 1024
ncu --set full --clock-control none --import-source on -k regex:search_kernel -s 2 -c 1 -f -o $O/r02_e_hybrid python bench.py --no-extra --no-cpu-baseline --positions 296 --trees 16 --lanes 128 --steps 2 --warmup 1 > $O/r02_ncu_e1.log 2>&1
cp ismcsolver_b200/libismcts_b200.so $O/r02_e_lib.so
ls -la $O/r02_e_*.ncu-rep
