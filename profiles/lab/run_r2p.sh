#!/bin/bash
# round-2 lab p: SRH tile kernel variants — parity tests and device-path time on the HRRR field
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "srh or stdheight or multi_entry" 2>&1 | tail -2
for rep in 1 2; do
python bench.py --workload C4 --no-extras --no-cpu-baseline --steps 10 --warmup 3 2>&1 | tail -1 | python -c "
import sys, json
r = json.loads(sys.stdin.read())
print('device path ms/step', round(r['ms_per_step'], 4), 'level-major kernel_ms', round(r['roofline']['kernel_ms'], 4), 'e2e ms', round(r['e2e']['ms_per_step'], 2))
"
done
} > gpurun_out/r2p_lab.txt 2>&1
cat gpurun_out/r2p_lab.txt
