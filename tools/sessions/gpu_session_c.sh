#!/bin/bash
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out
python tools/extra_bench.py kw-hybrid --hybrid 16x128,16x148,64x32,148x8 > $O/r02_c_hybrid.jsonl 2> $O/r02_c_hybrid.err; cut -c100-420 $O/r02_c_hybrid.jsonl
ISMCTS_MIN_BLOCKS=4 python tools/extra_bench.py kw-hybrid --hybrid 16x128 > $O/r02_c_hybrid_mb4.jsonl 2>&1; cut -c100-420 $O/r02_c_hybrid_mb4.jsonl
python tools/geometry_agreement.py --geometries 16x128 --out $O/r02_c_geometry.jsonl > $O/r02_c_geometry.log 2>&1; cut -c1-330 $O/r02_c_geometry.log
python -m pytest tests -m gpu -q > $O/r02_pytest_d.log 2>&1; tail -6 $O/r02_pytest_d.log
ncu --set full --clock-control none --import-source on -k regex:search_kernel -s 2 -c 1 -f -o $O/r02_c_hybrid python tools/extra_bench.py kw-hybrid --hybrid 16x128 > $O/r02_ncu_c.log 2>&1
ls -la $O/r02_c_hybrid.ncu-rep
cp ismcsolver_b200/libismcts_b200.so $O/r02_c_lib.so
