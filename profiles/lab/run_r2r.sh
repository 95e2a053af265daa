#!/bin/bash
# round-2 lab r: pageable-input end-to-end time vs host copy threads
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
{
nproc
for t in 2 4 8 12 16; do
  echo "COPY_THREADS=$t"
  XCAPE_B200_COPY_THREADS=$t XCAPE_B200_TRACE_ONE=1 python profiles/probe_e2e_layouts.py 10 2>&1 | grep -A3 "pageable level-last, all\|pinned level-last"
done
} > gpurun_out/r2r_lab.txt 2>&1
cat gpurun_out/r2r_lab.txt
