#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/t3_pytest.log
python tools/kernel_table.py --only g4 --sizes 4k > gpurun_out/t3_table.jsonl 2> gpurun_out/t3_table.err
python tools/kernel_table.py --only lines --sizes 1080p >> gpurun_out/t3_table.jsonl 2>> gpurun_out/t3_table.err
python bench.py --cfg1 > gpurun_out/t3_cfg1.json 2> gpurun_out/t3_cfg1.err
for k in g4s g4ss; do
ncu --set full --clock-control none --import-source on -k regex:k_march -s 2 -c 1 -f -o gpurun_out/r02_prof_${k}_v3 python tools/prof_one.py $k --n 4 --size 4k > gpurun_out/t3_ncu_$k.log 2>&1
done
tail -8 gpurun_out/t3_pytest.log; cut -c1-200 gpurun_out/t3_table.jsonl; cat gpurun_out/t3_cfg1.json
