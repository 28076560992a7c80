#!/bin/bash
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out
nvidia-smi -L | head -3
python -m pytest tests/test_cpp_conformance.py tests/test_gpu_parity.py -m gpu -q -k "devices or search_multi or conformance or lanes_cap" > $O/r02_pytest_2gpu.log 2>&1; tail -5 $O/r02_pytest_2gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 6 --warmup 3 > $O/r02_bench_2gpu.json 2> $O/r02_bench_2gpu.err; python -c "
import json
d=json.load(open('$O/r02_bench_2gpu.json')); print('2 GPUs value %.4g e2e %.4g ms/step %.1f' % (d['value'], d['e2e']['value'], d['ms_per_step'])); print({k: '%.4g' % v['value'] for k, v in d['extra'].items()})" || tail -5 $O/r02_bench_2gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/selfplay_sweep.py --games 256 --time-ms 100 --out $O/r02_selfplay_2gpu.jsonl > $O/r02_selfplay_2gpu.log 2>&1; tail -1 $O/r02_selfplay_2gpu.log | cut -c1-600
