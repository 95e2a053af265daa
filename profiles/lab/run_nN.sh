#!/bin/bash
# torchrun bench line at N GPUs (weak scaling + C5 24-step stack);  usage: run_nN.sh N
N=$1
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
{
nvidia-smi -L | wc -l
if [ "$N" = "8" ]; then timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "devices or multi_entry" 2>&1 | tail -2; fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n${N}_err.txt
tail -2 gpurun_out/bench_n${N}_err.txt
python - <<PY
import json
r = json.loads(open('gpurun_out/bench_n$N.json').read().strip().splitlines()[-1])
print('N', r['n_gpus'], 'value', r['value'], 'ms', r['ms_per_step'], 'e2e', r['e2e']['value'], r['e2e']['ms_per_step'])
print({k: (v.get('value'), v.get('ms_per_step')) for k, v in r['e2e'].items() if isinstance(v, dict)})
s = r.get('strong_scaling_C5'); print('strong', s['device_resident'], s['e2e']['ms_total'], s['e2e']['value'])
PY
} > gpurun_out/n${N}_lab.txt 2>&1
cat gpurun_out/n${N}_lab.txt
