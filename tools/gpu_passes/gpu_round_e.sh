#!/bin/bash
# 2-GPU pass: tensor-parallel tests (persistent + graph engines, peer + NCCL) and bench lines
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tensor_parallel.py -m gpu -x -q > gpurun_out/e_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/e_pytest.log
run_bench() {  # name, env..., args
  local name=$1; shift
  echo "== $name"
  env "$@" timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 \
     bench.py --gpus 2 --steps 512 --warmup 16 $EXTRA > gpurun_out/e_bench_$name.json 2> gpurun_out/e_bench_$name.err; echo "rc=$?"
  tail -2 gpurun_out/e_bench_$name.err | cut -c1-300
  python -c "import json,sys; d=json.loads(open('gpurun_out/e_bench_$name.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['config'].get('engine'), d['config'].get('tp_comm'), d['roofline']['frac'])"
}
EXTRA="" run_bench tiny_tp2_persistent KLLM_TP_COMM=peer
EXTRA="" run_bench tiny_tp2_graph_peer KLLM_TP_COMM=peer KLLM_ENGINE=graph
EXTRA="--workload llama2-7b" run_bench l7b_tp2_persistent KLLM_TP_COMM=peer
