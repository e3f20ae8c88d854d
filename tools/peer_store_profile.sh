#!/bin/bash
# 2-GPU trip: ncu of the fused kernel storing into the OTHER GPU's memory; cfg2-sized traffic captures of M0/M1/M2 on GPU 0
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on --devices 1 -k regex:k_march -c 1 -f -o gpurun_out/r02_final_peer_store python tools/prof_bands.py --reps 1 > gpurun_out/t7_peer.log 2>&1
python tools/prof_bands.py --reps 3 >> gpurun_out/t7_peer.log 2>&1
for k in m0 m1 m2; do
  ncu --set full --clock-control none -k regex:k_march -s 2 -c 1 -f -o gpurun_out/r02_traffic_$k python tools/prof_one.py $k --n 64 --size 1080p > gpurun_out/t7_ncu_$k.log 2>&1
done
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus 2 --steps 20 --warmup 5 --configs cfg5 --no-e2e --quick > gpurun_out/t7_bench_cfg5_n2.json 2> gpurun_out/t7_bench.err
cat gpurun_out/t7_peer.log | tail -8; tail -3 gpurun_out/t7_bench.err; cut -c1-100 gpurun_out/t7_bench_cfg5_n2.json
