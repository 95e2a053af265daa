#!/bin/bash
# 8-GPU evidence: the product's single-process devices=[...] path and the torchrun bench line (weak scaling + C5 24-step stack)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
{
nvidia-smi -L | wc -l
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "devices" 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8_err.txt
tail -2 gpurun_out/bench_n8_err.txt
python - <<'PY'
import json
r = json.loads(open('gpurun_out/bench_n8.json').read().strip().splitlines()[-1])
print('N', r['n_gpus'], 'value', r['value'], 'ms', r['ms_per_step'], 'e2e', r['e2e']['value'], r['e2e']['ms_per_step'])
print({k: v for k, v in r['e2e'].items() if isinstance(v, dict)})
print('strong', r.get('strong_scaling_C5'))
PY
} > gpurun_out/n8_lab.txt 2>&1
cat gpurun_out/n8_lab.txt
