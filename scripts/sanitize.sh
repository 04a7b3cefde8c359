#!/usr/bin/env bash
# Race / sync / memory checks of the CUDA kernels (the reference wires up none). One GPU, under gpurun:
#   gpurun --timeout 1200 -- 'bash scripts/sanitize.sh'
set -uo pipefail
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  compute-sanitizer --tool $tool --error-exitcode 1 python -m pytest tests/test_kernels_gpu.py -q -x -k "sgd or adam or minmax or nhwc or elementwise" \
      > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool exit=$?" | tee -a gpurun_out/sanitizer_summary.txt
done
# host side: the C++ scheduler under TSAN/ASAN runs in the CPU suite (tests/test_scheduler_tsan.py)
