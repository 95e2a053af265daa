#!/bin/bash
# round-2 lab p: SRH tile kernel geometry (levels per chunk x CTAs per SM), FP64 tables in shared memory
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
{

for lib in gpurun_lab/lib_kc*.so xcape_b200/libxcape_b200.so; do
  XCAPE_B200_LIB=$PWD/$lib python bench.py --workload C4 --no-extras --no-cpu-baseline --steps 10 --warmup 3 2>&1 | tail -1 | python -c "
import sys, json
r = json.loads(sys.stdin.read())
print('$lib', 'device path ms/step', round(r['ms_per_step'], 4), 'level-major kernel_ms', round(r['roofline']['kernel_ms'], 4))
"
done
} > gpurun_out/r2p_lab.txt 2>&1
cat gpurun_out/r2p_lab.txt
