#!/bin/bash
# round 2, 2 GPUs: tensor-parallel tests (world 2) + bench.py --gpus 2 (Llama-2-7B int8 primary, fp32 secondary)
set -u
mkdir -p gpurun_out
O=gpurun_out/r2tp2
nvidia-smi topo -m > ${O}_topo.txt 2>&1
timeout 1200 python -m pytest tests/test_tensor_parallel.py -m gpu -x -q > ${O}_pytest.log 2>&1; echo "pytest tp rc=$?"; tail -8 ${O}_pytest.log | cut -c1-220
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 64 --warmup 3 --reps 3 > ${O}_bench.json 2> ${O}_bench.err; echo "bench tp2 rc=$?"; tail -3 ${O}_bench.err | cut -c1-300
python -c "
import json;d=json.load(open('${O}_bench.json'));print(round(d['value'],1),round(d['e2e']['value'],1),d['by_position_tok_s'],round(d['roofline']['frac'],3),d.get('parity'));x=d.get('exact');print('exact',x and (round(x['value'],1),round(x['roofline_frac'],3)));s=d.get('secondary');print('secondary',s and (round(s['value'],1),round(s['e2e']['value'],1),round(s['roofline']['frac'],3),s.get('parity')))"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --impl reference --steps 16 --warmup 3 > ${O}_bench_ref.json 2> ${O}_bench_ref.err; echo "bench ref rc=$?"; cut -c1-400 ${O}_bench_ref.json
