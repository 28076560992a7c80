#!/bin/bash
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out
python -m pytest tests -m gpu -q > $O/r02_pytest_h.log 2>&1; tail -6 $O/r02_pytest_h.log
run() { name=$1; shift
  python bench.py --no-extra --no-cpu-baseline --steps 4 --warmup 3 "$@" > $O/r02_g_$name.json 2> $O/r02_g_$name.err
  python -c "
import json
try:
    d = json.load(open('$O/r02_g_$name.json')); print('$name', 'value %.4g' % d['value'], 'e2e %.4g' % d['e2e']['value'], 'ms/step %.2f' % d['ms_per_step'])
except Exception as e: print('$name failed', e); print(open('$O/r02_g_$name.err').read()[-600:])"
}
run hybrid_default
run hybrid64x32_p592 --positions 592 --trees 64 --lanes 32
run wide_p1024 --positions 1024 --trees 2368 --lanes 1
run wide_p64 --positions 64 --trees 2368 --lanes 1
python bench.py > $O/r02_g_bench.json 2> $O/r02_g_bench.err; tail -c 2500 $O/r02_g_bench.json; tail -3 $O/r02_g_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/r02_g_bench_ref.json 2>&1; cut -c1-300 $O/r02_g_bench_ref.json
ncu --set full --clock-control none --import-source on -k regex:search_kernel -s 2 -c 1 -f -o $O/r02_g_hybrid python bench.py --no-extra --no-cpu-baseline --steps 2 --warmup 1 > $O/r02_ncu_g1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_g_launches.csv python bench.py --no-extra --no-cpu-baseline --steps 2 --warmup 1 > $O/r02_ncu_g2.log 2>&1
cp ismcsolver_b200/libismcts_b200.so $O/r02_g_lib.so
ls -la $O/r02_g_*.ncu-rep
