#!/bin/bash
# 2-GPU pass after the barrier-free change: TP tests + bench lines
set -u
mkdir -p gpurun_out
timeout 700 python -m pytest tests/test_tensor_parallel.py -m gpu -x -q > gpurun_out/k_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/k_pytest.log
run_bench() {
  local name=$1; shift
  echo "== $name"
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 \
     bench.py --gpus 2 --steps 512 --warmup 16 "$@" > gpurun_out/k_bench_$name.json 2> gpurun_out/k_bench_$name.err; echo "rc=$?"
  python -c "import json,sys; d=json.loads(open('gpurun_out/k_bench_$name.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['config'].get('engine'), d['config'].get('tp_comm'), d['roofline']['frac'])"
}
run_bench tiny_tp2
run_bench l7b_int8_tp2 --workload llama2-7b-int8
run_bench l7b_tp2 --workload llama2-7b
