#!/bin/bash
# quick A/B: kernel-alone and call times of the shipping library on C2 / C3 / C5 (+ bit-exactness spot tests)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "sorted_execution or full_field_bitexact or two_column or every_source or c1_ or garbage or edge or pinc or fast_mode or golden" 2>&1 | tail -2
python profiles/lab_time_kernel.py C2 2 10
LAB_SHUFFLE=1 python profiles/lab_time_kernel.py C2 2 10
python profiles/lab_time_kernel.py C3 3 5
python profiles/lab_time_kernel.py C5 2 5
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/quick_launches.csv python profiles/run_cape_once.py 1038240 2 2 C2 faithful > /dev/null 2>&1
grep -v "at::" gpurun_out/quick_launches.csv | awk -F'","' '{print $5, $NF}' | tail -5
} > gpurun_out/quick_lab.txt 2>&1
cat gpurun_out/quick_lab.txt
