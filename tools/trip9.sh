#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_bands_gpu.py tests/test_multi_gpu.py -m gpu -q 2>&1 | tail -8 > gpurun_out/t9_pytest.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus 2 --steps 20 --warmup 5 --configs cfg5 --no-e2e --quick > gpurun_out/t9_bench_cfg5_n2.json 2> gpurun_out/t9_bench.err
tail -3 gpurun_out/t9_pytest.log; tail -2 gpurun_out/t9_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/t9_bench_cfg5_n2.json').read().strip().splitlines()[-1])
print({k:v for k,v in d['configs']['cfg5'].items() if k not in ('roofline','workload','collective') and not k.endswith('clocks')})
PY
