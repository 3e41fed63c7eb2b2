#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_tensor_parallel.py -m gpu -x -q -k "decoder and int8" > gpurun_out/l_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/l_pytest.log
echo "== l7b_int8_tp2"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 \
   bench.py --gpus 2 --steps 512 --warmup 16 --workload llama2-7b-int8 > gpurun_out/l_bench_l7b_int8_tp2.json 2> gpurun_out/l_bench_l7b_int8_tp2.err; echo "rc=$?"
python -c "import json,sys; d=json.loads(open('gpurun_out/l_bench_l7b_int8_tp2.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['config'].get('engine'), d['config'].get('tp_comm'), d['roofline']['frac'])"
