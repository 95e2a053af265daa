#!/bin/bash
# final ncu captures of the shipping kernels + the reference arm of the bench
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
{
ncu --set full --clock-control none --import-source on -k regex:cape_kernel2 -c 1 -o gpurun_out/r2w_cape2 python profiles/run_cape_once.py 1038240 1 2 C2 faithful > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:cape_source -c 1 -o gpurun_out/r2w_source python profiles/run_cape_once.py 1038240 1 2 C2 faithful > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:srh_tile -c 1 -o gpurun_out/r2w_srh_tile python bench.py --workload C4 --no-extras --no-cpu-baseline --steps 1 --warmup 1 > /dev/null 2>&1
python profiles/summarise_ncu.py gpurun_out/r2w_cape2.ncu-rep cape_kernel2 "x" | tail -32
python profiles/summarise_ncu.py gpurun_out/r2w_source.ncu-rep cape_source "x" | tail -32
python profiles/summarise_ncu.py gpurun_out/r2w_srh_tile.ncu-rep srh_tile "x" | tail -32
python bench.py --impl reference --steps 3 --warmup 1 | cut -c1-600
} > gpurun_out/r2w_lab.txt 2>&1
tail -5 gpurun_out/r2w_lab.txt | cut -c1-400
