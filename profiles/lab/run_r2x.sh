#!/bin/bash
# ncu source-level capture of the sorted ascent kernel on the HRRR mixed-layer workload (C3, sigma grid, window order)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:cape_kernel2 -c 1 -o gpurun_out/r2x_cape2_c3 python profiles/run_cape_once.py 1905141 1 3 C3 faithful > gpurun_out/r2x.log 2>&1
ncu -i gpurun_out/r2x_cape2_c3.ncu-rep --page source --csv > gpurun_out/r2x_cape2_c3_source.csv
python profiles/summarise_ncu.py gpurun_out/r2x_cape2_c3.ncu-rep cape_kernel2 "x" > gpurun_out/r2x_summary.csv
tail -3 gpurun_out/r2x.log
