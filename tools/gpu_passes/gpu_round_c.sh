#!/bin/bash
# 2-GPU pass: tensor-parallel tests and bench (peer-memory and NCCL transports)
set -u
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/c_topo.txt 2>&1
timeout 900 python -m pytest tests/test_tensor_parallel.py -m gpu -x -q > gpurun_out/c_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/c_pytest.log
for comm in peer nccl; do
  echo "== tp2 $comm"
  KLLM_TP_COMM=$comm timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 \
     bench.py --gpus 2 --steps 512 --warmup 16 > gpurun_out/c_bench_tp2_$comm.json 2> gpurun_out/c_bench_tp2_$comm.err; echo "rc=$?"
  tail -3 gpurun_out/c_bench_tp2_$comm.err; cat gpurun_out/c_bench_tp2_$comm.json
done
