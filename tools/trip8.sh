#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/t8_pytest.log
python tools/kernel_table.py --only g2 > gpurun_out/t8_table.jsonl 2> gpurun_out/t8_table.err
for k in m2 steer5s; do
  ncu --set full --clock-control none --import-source on -k regex:k_march -s 2 -c 1 -f -o gpurun_out/r02_v8_$k python tools/prof_one.py $k --n 4 --size 4k > gpurun_out/t8_ncu_$k.log 2>&1
done
tail -6 gpurun_out/t8_pytest.log; cut -c1-190 gpurun_out/t8_table.jsonl
