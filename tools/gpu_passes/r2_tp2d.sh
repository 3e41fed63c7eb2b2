#!/bin/bash
# round 2, 2 GPUs (budget-bound: ~6 minutes of box time): the C++ tensor-parallel runner (kuiper_tp_launch,
# load-time sharding, TCP rendezvous), the driver's own `bench.py --gpus 2` line, then as much of the Python TP
# suite as the time allows
set -u
mkdir -p gpurun_out
O=gpurun_out/r2tp2d
timeout 150 python -m pytest tests/test_cpp_tensor_parallel.py -m gpu -x -q > ${O}_pytest_cpp_tp.log 2>&1; echo "pytest cpp tp rc=$?"; tail -4 ${O}_pytest_cpp_tp.log | cut -c1-300
timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 64 --warmup 3 --reps 3 --no-secondary --no-cpu-baseline > ${O}_bench_tp2.json 2> ${O}_bench_tp2.err; echo "bench tp2 rc=$?"
python -c "
import json;d=json.load(open('${O}_bench_tp2.json'));x=d.get('exact') or {}
print('   tp2 int8', round(d['value'],1), round(d['e2e']['value'],1), {k:round(v) for k,v in d['by_position_tok_s'].items()}, round(d['roofline']['frac'],3), d.get('parity'), d['config'].get('classifier_rows_per_gpu'), '| exact', x.get('value'))"
timeout 100 python -m pytest tests/test_tensor_parallel.py -m gpu -x -q -k "matches_unsharded_oracle and (small-tp-2 or small-tp-int8-2)" > ${O}_pytest_tp.log 2>&1; echo "pytest tp rc=$?"; tail -4 ${O}_pytest_tp.log | cut -c1-300
