#!/bin/bash
# compute-sanitizer over the small parity cases (memcheck + racecheck + synccheck); run under gpurun.
set -o pipefail
for tool in memcheck racecheck synccheck; do
  echo "== $tool"
  compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_g2_class_gpu.py tests/test_g4_gpu.py -m gpu -q -x --timeout 900 \
     -k "tiny or fish or class_vs_oracle or generic or constant" 2>&1 | tail -4
done
