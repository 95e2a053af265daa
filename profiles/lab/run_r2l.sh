#!/bin/bash
# round-2 lab l: launch geometry of the sorted two-column kernel (threads per CTA x CTAs per SM -> register cap)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
{
for lib in gpurun_lab/lib_*.so xcape_b200/libxcape_b200.so; do
  XCAPE_B200_LIB=$PWD/$lib XCAPE_B200_SORT=1 python profiles/lab_time_kernel.py C2 2 10
  XCAPE_B200_LIB=$PWD/$lib XCAPE_B200_SORT=1 python profiles/lab_time_kernel.py C5 2 5
done
} > gpurun_out/r2l_lab.txt 2>&1
cat gpurun_out/r2l_lab.txt
