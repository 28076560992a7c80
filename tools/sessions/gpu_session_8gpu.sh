#!/bin/bash
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out
nvidia-smi -L | wc -l; nproc
show() { python - "$1" <<'PY'
import json, sys
txt = open(sys.argv[1]).read()
lines = [l for l in txt.splitlines() if l.startswith('{')]
if not lines:
    print(sys.argv[1], 'no JSON'); sys.exit()
d = json.loads(lines[-1])
print(sys.argv[1], 'N', d['n_gpus'], 'value %.4g e2e %.4g ms/step %.1f' % (d['value'], d['e2e']['value'], d['ms_per_step']))
print('  ', {k: '%.4g' % v['value'] for k, v in d.get('extra', {}).items()})
PY
}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 6 --warmup 3 > $O/r02_bench_8gpu.json 2> $O/r02_bench_8gpu.err; show $O/r02_bench_8gpu.json; tail -2 $O/r02_bench_8gpu.err | cut -c1-300
timeout 120 python bench.py --impl reference --gpus 8 --steps 3 --warmup 1 > $O/r02_bench_8gpu_ref.json 2>&1; cut -c1-200 $O/r02_bench_8gpu_ref.json
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 tools/selfplay_sweep.py --games 1000 --time-ms 100 --out $O/r02_selfplay_8gpu.jsonl > $O/r02_selfplay_8gpu.log 2>&1; tail -1 $O/r02_selfplay_8gpu.log | cut -c1-700
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 4 --steps 4 --warmup 3 > $O/r02_bench_4gpu.json 2> $O/r02_bench_4gpu.err; show $O/r02_bench_4gpu.json
