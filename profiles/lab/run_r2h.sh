#!/bin/bash
# round-2 lab h: sorted execution of the faithful kernel, A/B against storage order
set -x
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "sorted_execution or garbage or full_field_bitexact or two_column or c1_ or every_source" 2>&1 | tail -5
for sort in 0 1; do
  XCAPE_B200_SORT=$sort python profiles/lab_time_kernel.py C2 2 10
  XCAPE_B200_SORT=$sort LAB_SHUFFLE=1 python profiles/lab_time_kernel.py C2 2 10
  XCAPE_B200_SORT=$sort LAB_ACTIVE=0 python profiles/lab_time_kernel.py C2 2 10
  XCAPE_B200_SORT=$sort python profiles/lab_time_kernel.py C3 3 5
  XCAPE_B200_SORT=$sort python profiles/lab_time_kernel.py C5 2 5
done
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2h_launches.csv python profiles/run_cape_once.py 1038240 2 2 C2 faithful > gpurun_out/r2h_ncu.log 2>&1
tail -12 gpurun_out/r2h_launches.csv
} > gpurun_out/r2h_lab.txt 2>&1
tail -40 gpurun_out/r2h_lab.txt
