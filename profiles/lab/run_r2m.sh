#!/bin/bash
# round-2 lab m: global (counting sort) vs windowed (bitonic) order for the sorted CAPE execution
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "sorted_execution or garbage" 2>&1 | tail -3
for mode in global window; do
  XCAPE_B200_SORT=1 XCAPE_B200_SORT_MODE=$mode python profiles/lab_time_kernel.py C2 2 10
  XCAPE_B200_SORT=1 XCAPE_B200_SORT_MODE=$mode LAB_SHUFFLE=1 python profiles/lab_time_kernel.py C2 2 10
  XCAPE_B200_SORT=1 XCAPE_B200_SORT_MODE=$mode LAB_ACTIVE=0 python profiles/lab_time_kernel.py C2 2 10
  XCAPE_B200_SORT=1 XCAPE_B200_SORT_MODE=$mode python profiles/lab_time_kernel.py C2 1 10
  XCAPE_B200_SORT=1 XCAPE_B200_SORT_MODE=$mode python profiles/lab_time_kernel.py C3 3 5
  XCAPE_B200_SORT=1 XCAPE_B200_SORT_MODE=$mode python profiles/lab_time_kernel.py C5 2 5
done
XCAPE_B200_SORT_MODE=global ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2m_launches.csv python profiles/run_cape_once.py 1038240 2 2 C2 faithful > gpurun_out/r2m_ncu.log 2>&1
grep -v "at::" gpurun_out/r2m_launches.csv | awk -F'","' '{print $5, $NF}' | tail -5
} > gpurun_out/r2m_lab.txt 2>&1
cat gpurun_out/r2m_lab.txt
