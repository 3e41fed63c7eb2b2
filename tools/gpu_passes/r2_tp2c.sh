#!/bin/bash
# round 2, 2 GPUs: A/B of the vocabulary-sharded classifier (same box, same run)
set -u
mkdir -p gpurun_out
O=gpurun_out/r2tp2c
run() { # name env
  name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 64 --warmup 3 --reps 3 --no-exact --no-secondary > ${O}_bench_${name}.json 2> ${O}_bench_${name}.err; echo "bench ${name} rc=$?"
  python -c "
import json;d=json.load(open('${O}_bench_${name}.json'));print('   ${name}',round(d['value'],1),round(d['e2e']['value'],1),{k:round(v) for k,v in d['by_position_tok_s'].items()},round(d['roofline']['frac'],3),d['config'].get('classifier_rows_per_gpu'))"
}
run sharded A=1
run replicated KLLM_TP_SHARD_CLS=0
run sharded_again A=1
