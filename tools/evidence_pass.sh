#!/bin/bash
# Final single-GPU evidence pass of round 2: tests (+ achieved parity errors), kernel table, bench line, launch list, ncu digests
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/t6_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/t6_smoke.log 2>&1
python tools/kernel_table.py > gpurun_out/r02_kernel_table.jsonl 2> gpurun_out/t6_table.err
SECONDS=0; timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/t6_bench.err; echo "bench.py default run: $SECONDS s" > gpurun_out/t6_bench_time.log
SECONDS=0; timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_reference_n1.json 2>> gpurun_out/t6_bench.err; echo "bench.py --impl reference: $SECONDS s" >> gpurun_out/t6_bench_time.log
python bench.py --cfg1 > gpurun_out/r02_cfg1.json 2>> gpurun_out/t6_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 20 --warmup 5 --configs "" --no-cpu --quick > gpurun_out/t6_launches.log 2>&1
for k in m0 m1 m2 lines steer5s steer5m g4b g4s pyr; do
  ncu --set full --clock-control none --import-source on -k regex:"k_march|k_pyr" -s 2 -c 1 -f -o gpurun_out/r02_final_$k python tools/prof_one.py $k --n 4 --size 4k > gpurun_out/t6_ncu_$k.log 2>&1
done
tail -4 gpurun_out/t6_pytest.log; tail -2 gpurun_out/t6_smoke.log | cut -c1-300; cut -c1-150 gpurun_out/r02_kernel_table.jsonl | head -30; cut -c1-600 gpurun_out/r02_bench_n1.json; tail -3 gpurun_out/t6_bench.err
cat gpurun_out/t6_bench_time.log
