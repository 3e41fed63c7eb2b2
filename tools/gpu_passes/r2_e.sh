#!/bin/bash
# round 2, pass E: int8 consumer warps 6 / 8 / 16 (255 / 168 / 96 registers), whole -m gpu suite with the
# new prompt / prefill entry points
set -u
mkdir -p gpurun_out
O=gpurun_out/r2e
run() { # name, env..., -- bench args
  name=$1; shift
  env "$@" timeout 300 python bench.py --reps 3 --no-cpu-baseline ${BARGS} > ${O}_bench_${name}.json 2> ${O}_bench_${name}.err; echo "bench ${name} rc=$?"
  python -c "
import json;d=json.load(open('${O}_bench_${name}.json'));print('   ${name}',round(d['value'],1),round(d['e2e']['value'],1),{k:round(v) for k,v in d['by_position_tok_s'].items()},round(d['roofline']['frac'],3))"
}
BARGS="--workload llama2-7b-int8 --steps 256"
run int8_cw6 KLLM_CONSUMER_WARPS=6
run int8_cw8 KLLM_CONSUMER_WARPS=8
run int8_cw16 KLLM_CONSUMER_WARPS=16
timeout 1500 python -m pytest tests -m gpu -x -q > ${O}_pytest.log 2>&1; echo "pytest rc=$?"; tail -25 ${O}_pytest.log | cut -c1-220
KLLM_CONSUMER_WARPS=6 timeout 300 python tools/phase_timeline.py --workload llama2-7b-int8 --pos 64 > ${O}_timeline_int8_pos64_cw6.txt 2>${O}_timeline.err; cat ${O}_timeline_int8_pos64_cw6.txt
