#!/bin/bash
# 2-GPU trip: multi-GPU tests (NCCL bands, peer stores), bench at N=2 (cfg5 gather modes), G4 BH=96 variant on GPU 0
mkdir -p gpurun_out
python -m pytest tests/test_bands_gpu.py tests/test_multi_gpu.py tests/test_lines_u8_gpu.py tests/test_g2_batch_gpu.py -m gpu -q -k "bands or multi or device or e2e or nccl or two" 2>&1 | tail -30 > gpurun_out/t4_pytest.log
CVS_LIB=$PWD/cvsteer_b200/variants/lib_g4bh96.so python tools/kernel_table.py --only g4 --sizes 4k > gpurun_out/t4_table_bh96.jsonl 2> gpurun_out/t4_table.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/t4_bench_n2.json 2> gpurun_out/t4_bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 tools/pcie_probe.py > gpurun_out/t4_pcie_n2.json 2> gpurun_out/t4_pcie_n2.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/t4_smoke.log 2>&1
tail -12 gpurun_out/t4_pytest.log; cut -c1-200 gpurun_out/t4_table_bh96.jsonl; tail -5 gpurun_out/t4_bench_n2.err; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/t4_bench_n2.json').read().strip().splitlines()[-1])
    print('value',d['value'],'e2e',d['e2e'])
    for k,v in d['configs'].items(): print(k, {kk:vv for kk,vv in v.items() if kk not in ('roofline','workload','clocks')})
except Exception as e: print('bench parse failed', e)
PY
cut -c1-1200 gpurun_out/t4_pcie_n2.json; tail -3 gpurun_out/t4_smoke.log
