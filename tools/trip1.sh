#!/bin/bash
# GPU trip 1 (round 2): tests, kernel table (default build vs the round-1 G4 loop), FFMA probe, ncu of the G4 kernels
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/t1_pytest.log
python tools/kernel_table.py > gpurun_out/t1_table.jsonl 2> gpurun_out/t1_table.err
CVS_LIB=$PWD/cvsteer_b200/variants/lib_g4pipe0.so python tools/kernel_table.py --only g4 > gpurun_out/t1_table_g4pipe0.jsonl 2>> gpurun_out/t1_table.err
python tools/quick_bench.py --ffma --n 8 > gpurun_out/t1_ffma.json 2>&1
for k in g4s g4b; do
  ncu --set full --clock-control none --import-source on -k regex:k_march -s 2 -c 1 -f -o gpurun_out/r02_prof_${k} python tools/prof_one.py $k --n 4 --size 4k > gpurun_out/t1_ncu_$k.log 2>&1
  CVS_LIB=$PWD/cvsteer_b200/variants/lib_g4pipe0.so ncu --set full --clock-control none --import-source on -k regex:k_march -s 2 -c 1 -f -o gpurun_out/r02_prof_${k}_pipe0 python tools/prof_one.py $k --n 4 --size 4k >> gpurun_out/t1_ncu_$k.log 2>&1
done
tail -3 gpurun_out/t1_pytest.log; cat gpurun_out/t1_table.jsonl | cut -c1-230; echo ---; cat gpurun_out/t1_table_g4pipe0.jsonl | cut -c1-230; cat gpurun_out/t1_ffma.json
