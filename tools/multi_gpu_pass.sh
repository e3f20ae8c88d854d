#!/bin/bash
# N-GPU trip (N = $1, default 8): the GPU tests that need more than one device, the host<->device ceiling of the box with N GPUs
# streaming, then the contract bench at N (all configs) exactly as the driver launches it, and its reference arm.
N=${1:-8}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/mg_pytest_n$N.log
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621 tools/pcie_probe.py --seconds 0.7 > gpurun_out/mg_pcie_n$N.json 2> gpurun_out/mg_pcie_n$N.err
SECONDS=0
timeout 480 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29622 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/mg_bench_n$N.json 2> gpurun_out/mg_bench_n$N.err
echo "bench.py --gpus $N: $SECONDS s" > gpurun_out/mg_time_n$N.log
SECONDS=0
timeout 480 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29623 bench.py --impl reference --gpus $N --steps 20 --warmup 5 > gpurun_out/mg_bench_reference_n$N.json 2>> gpurun_out/mg_bench_n$N.err
echo "bench.py --impl reference --gpus $N: $SECONDS s" >> gpurun_out/mg_time_n$N.log
nvidia-smi topo -m > gpurun_out/mg_topo_n$N.txt 2>&1; lscpu | head -30 >> gpurun_out/mg_topo_n$N.txt; cat /sys/bus/pci/devices/*/numa_node 2>/dev/null | sort | uniq -c >> gpurun_out/mg_topo_n$N.txt
cat gpurun_out/mg_pytest_n$N.log gpurun_out/mg_time_n$N.log
grep -v "^$" gpurun_out/mg_bench_n$N.err | tail -5 | cut -c1-300; N=$N python - <<'PY'
import json, os
N=os.environ["N"]
try:
    d=json.loads(open(f'gpurun_out/mg_bench_n{N}.json').read().strip().splitlines()[-1])
    print('value',d['value'],'e2e',{k:v for k,v in d['e2e'].items() if k not in ('api','pcie_ceiling_how')})
    for k,v in d['configs'].items(): print(k, {kk:vv for kk,vv in v.items() if kk not in ('roofline','workload','clocks','collective') and not kk.endswith('_clocks')})
except Exception as e: print('bench parse failed', e)
try:
    d=json.loads(open(f'gpurun_out/mg_pcie_n{N}.json').read().strip().splitlines()[-1])
    for k,v in d['cases'].items(): print(k, v['total_h2d_GB_s'], v['total_d2h_GB_s'], [ (x['h2d_GB_s'],x['d2h_GB_s']) for x in v['per_rank']][:3], v['per_rank'][0].get('numa_node'))
except Exception as e: print('pcie parse failed', e)
print(open(f'gpurun_out/mg_bench_reference_n{N}.json').read()[:300])
PY
