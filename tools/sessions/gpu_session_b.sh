#!/bin/bash
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out
python tools/geometry_agreement.py --geometries 16,2368,16x128,16x148,64x32 --out $O/r02_geometry.jsonl > $O/r02_geometry.log 2>&1; cut -c1-700 $O/r02_geometry.log
python tools/extra_bench.py kw-hybrid > $O/r02_hybrid.jsonl 2> $O/r02_hybrid.err; cat $O/r02_hybrid.jsonl | cut -c100-400
for B in 16x1000000x128 9472x1000000 37888x1000000 151552x1000000; do
  timeout 900 python tools/head_to_head.py --games 600 --a 2368x1000000 --b $B --pool-gb 80 --out $O/r02_h2h_$B.jsonl > $O/r02_h2h_$B.log 2>&1; tail -1 $O/r02_h2h_$B.log | cut -c1-500
done
for P in 128 1024 2048; do
  python bench.py --no-extra --no-cpu-baseline --positions $P --steps 4 --warmup 3 > $O/r02_sweep2_p$P.json 2> $O/r02_sweep2_p$P.err
  python -c "
import json
d = json.load(open('$O/r02_sweep2_p$P.json')); print('positions', $P, 'value %.4g' % d['value'], 'e2e %.4g' % d['e2e']['value'], 'ms/step %.2f' % d['ms_per_step'])"
done
timeout 900 python tools/selfplay_sweep.py --games 1000 --time-ms 100 --vs-reference 24 --out $O/r02_selfplay.jsonl > $O/r02_selfplay.log 2>&1; tail -4 $O/r02_selfplay.log | cut -c1-600
python tools/extra_bench.py mnk-tree --reference > $O/r02_mnk_tree.jsonl 2> $O/r02_mnk_tree.err; tail -3 $O/r02_mnk_tree.jsonl | cut -c1-400
python tools/extra_bench.py kw-mo-exp3 --reference > $O/r02_mo_exp3.jsonl 2> $O/r02_mo_exp3.err; tail -2 $O/r02_mo_exp3.jsonl | cut -c1-300
python -m pytest tests -m gpu -q > $O/r02_pytest_c.log 2>&1; tail -8 $O/r02_pytest_c.log
ncu --set full --clock-control none --import-source on -k regex:search_kernel -s 1 -c 1 -f -o $O/r02_mo python tools/extra_bench.py kw-mo-exp3 > $O/r02_ncu_mo.log 2>&1
ls -la $O/r02_mo.ncu-rep
