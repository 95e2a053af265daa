#!/bin/bash
# round-2 lab k: ncu full capture + per-instruction stall sampling of the sorted two-column kernel
set -x
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
{
XCAPE_B200_SORT=1 XCAPE_B200_SORT_MODE=global ncu --set full --clock-control none --import-source on -k regex:cape_kernel2 -c 1 -o gpurun_out/r2k_cape2 python profiles/run_cape_once.py 1038240 1 2 C2 faithful
ncu -i gpurun_out/r2k_cape2.ncu-rep --page raw --csv > gpurun_out/r2k_cape2_raw.csv
ncu -i gpurun_out/r2k_cape2.ncu-rep --page source --csv > gpurun_out/r2k_cape2_source.csv
ls -la gpurun_out
} > gpurun_out/r2k_lab.txt 2>&1
tail -5 gpurun_out/r2k_lab.txt
