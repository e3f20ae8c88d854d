#!/bin/bash
# 8-GPU trip: host<->device ceiling of the box with 8 GPUs streaming, then the contract bench at N=8 (all configs)
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621 tools/pcie_probe.py --seconds 0.7 > gpurun_out/t5_pcie_n8.json 2> gpurun_out/t5_pcie_n8.err
timeout 480 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29622 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/t5_bench_n8.json 2> gpurun_out/t5_bench_n8.err
nvidia-smi topo -m > gpurun_out/t5_topo.txt 2>&1; lscpu | head -30 >> gpurun_out/t5_topo.txt; cat /sys/bus/pci/devices/*/numa_node 2>/dev/null | sort | uniq -c >> gpurun_out/t5_topo.txt
grep -v "^$" gpurun_out/t5_bench_n8.err | tail -5 | cut -c1-300; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/t5_bench_n8.json').read().strip().splitlines()[-1])
    print('value',d['value'],'e2e',{k:v for k,v in d['e2e'].items() if k not in ('api','pcie_ceiling_how')})
    for k,v in d['configs'].items(): print(k, {kk:vv for kk,vv in v.items() if kk not in ('roofline','workload','clocks','collective') and not kk.endswith('_clocks')})
except Exception as e: print('bench parse failed', e)
d=json.load(open('gpurun_out/t5_pcie_n8.json'))
for k,v in d['cases'].items(): print(k, v['total_h2d_GB_s'], v['total_d2h_GB_s'], [ (x['h2d_GB_s'],x['d2h_GB_s']) for x in v['per_rank']][:3], v['per_rank'][0].get('numa_node'))
PY
