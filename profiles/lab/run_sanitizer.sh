#!/bin/bash
# compute-sanitizer over profiles/sanitizer_driver.py (memcheck / initcheck / racecheck)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
{
echo "# compute-sanitizer over profiles/sanitizer_driver.py, B200, end of round 2 (sorted execution in both orders, SRH tile kernel, multi-device entries added)"
for tool in memcheck initcheck racecheck; do
  echo "## $tool"
  timeout 1500 compute-sanitizer --tool $tool python profiles/sanitizer_driver.py 2>&1 | grep -v "^=========\s*$" | tail -6
done
} > gpurun_out/r2_compute_sanitizer_summary.txt 2>&1
cat gpurun_out/r2_compute_sanitizer_summary.txt
