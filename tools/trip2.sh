#!/bin/bash
# GPU trip 2 (round 2): full test suite, kernel table, bench line, ncu of the G4 steer kernel
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/t2_pytest.log
python tools/kernel_table.py > gpurun_out/t2_table.jsonl 2> gpurun_out/t2_table.err
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/t2_bench.json 2> gpurun_out/t2_bench.err
ncu --set full --clock-control none --import-source on -k regex:k_march -s 2 -c 1 -f -o gpurun_out/r02_prof_g4s_v2 python tools/prof_one.py g4s --n 4 --size 4k > gpurun_out/t2_ncu.log 2>&1
tail -12 gpurun_out/t2_pytest.log; cut -c1-200 gpurun_out/t2_table.jsonl; tail -3 gpurun_out/t2_bench.err; cut -c1-1500 gpurun_out/t2_bench.json
