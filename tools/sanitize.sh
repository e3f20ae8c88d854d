#!/bin/bash
# compute-sanitizer over the small parity cases (memcheck + racecheck + synccheck); run under gpurun.
set -o pipefail
SEL='tiny or fish or class_vs_oracle or generic or constant or fused_pyramid or band_equals or u8 or lines_u8 or to_u8 or dominant_orientation or fuzz'
for tool in memcheck racecheck synccheck; do
  echo "== $tool"
  compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_g2_class_gpu.py tests/test_g4_gpu.py tests/test_g2_batch_gpu.py tests/test_lines_u8_gpu.py \
     -m gpu -q -x --timeout 1800 -k "$SEL" --deselect tests/test_g2_batch_gpu.py::test_fused_pyramid_emission_bitwise[shape3] 2>&1 | tail -4
done
