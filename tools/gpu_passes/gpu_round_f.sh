#!/bin/bash
# 4-GPU pass: world-4 tensor-parallel tests + bench, then the world-2 cases that changed
set -u
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_tensor_parallel.py -m gpu -x -q -k "4" > gpurun_out/f_pytest_w4.log 2>&1; echo "pytest w4 rc=$?"; tail -6 gpurun_out/f_pytest_w4.log
echo "== tiny_tp4_persistent"
KLLM_TP_COMM=peer timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 \
   bench.py --gpus 4 --steps 512 --warmup 16 > gpurun_out/f_bench_tiny_tp4.json 2> gpurun_out/f_bench_tiny_tp4.err; echo "rc=$?"
tail -2 gpurun_out/f_bench_tiny_tp4.err | cut -c1-300
python -c "import json,sys; d=json.loads(open('gpurun_out/f_bench_tiny_tp4.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['config'].get('engine'), d['config'].get('tp_comm'), d['roofline']['frac'])"
timeout 500 python -m pytest tests/test_tensor_parallel.py -m gpu -x -q -k "int8 or qwen" > gpurun_out/f_pytest_w2.log 2>&1; echo "pytest w2 rc=$?"; tail -6 gpurun_out/f_pytest_w2.log
