#!/bin/bash
# compute-sanitizer over the small parity cases (memcheck + racecheck + synccheck); run under gpurun.
# Round 2 selection adds the pipelined G4 loop (with its L2 bulk prefetch and one-row-ahead angle loads), the static
# steer-at-angle kernels, the fused min/max statistics and the single-process band contexts.
set -o pipefail
mkdir -p gpurun_out
SEL='tiny or fish or class_vs_oracle or generic or constant or fused_pyramid or band_equals or u8 or lines_u8 or to_u8 or dominant_orientation or fuzz or steer or contexts_of_one_process or dev_multi or sweep or two_pixel_map or dense_rows'
for tool in memcheck racecheck synccheck; do
  echo "== $tool"
  compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_g2_class_gpu.py tests/test_g4_gpu.py tests/test_g2_batch_gpu.py tests/test_lines_u8_gpu.py tests/test_bands_gpu.py \
     -m gpu -q -x --timeout 1800 -k "$SEL" --deselect "tests/test_g2_batch_gpu.py::test_fused_pyramid_emission_bitwise[shape3]" \
     --deselect tests/test_lines_u8_gpu.py::test_cvsteer_run_cli 2>&1 | tail -6
done 2>&1 | tee gpurun_out/r02_sanitizer.txt
