#!/bin/bash
# round 2, pass D: consumer-warp study (fp32 6 vs 8 warps = 255 vs 168 registers), tcgen05 GEMM test
set -u
mkdir -p gpurun_out
O=gpurun_out/r2d
timeout 600 python -m pytest tests/test_prefill_gpu.py -m gpu -x -q > ${O}_pytest_prefill.log 2>&1; echo "pytest prefill rc=$?"; tail -15 ${O}_pytest_prefill.log
KLLM_CONSUMER_WARPS=6 timeout 900 python -m pytest tests/test_decoder_gpu.py -m gpu -x -q > ${O}_pytest_cw6.log 2>&1; echo "pytest cw6 rc=$?"; tail -3 ${O}_pytest_cw6.log
run() { # name, env..., -- bench args
  name=$1; shift
  env "$@" timeout 300 python bench.py --reps 3 --no-cpu-baseline ${BARGS} > ${O}_bench_${name}.json 2> ${O}_bench_${name}.err; echo "bench ${name} rc=$?"
  python -c "
import json;d=json.load(open('${O}_bench_${name}.json'));print('   ${name}',round(d['value'],1),round(d['e2e']['value'],1),{k:round(v) for k,v in d['by_position_tok_s'].items()},round(d['roofline']['frac'],3))"
}
BARGS="--steps 1024"
run tiny_cw8 KLLM_CONSUMER_WARPS=8
run tiny_cw6 KLLM_CONSUMER_WARPS=6
run tiny_cw6_s5 KLLM_CONSUMER_WARPS=6 KLLM_STAGES=5
BARGS="--workload llama2-7b-int8 --steps 256"
run int8_cw16 KLLM_CONSUMER_WARPS=16
run int8_cw14 KLLM_CONSUMER_WARPS=14
BARGS="--workload qwen2.5-0.5b --steps 1024"
run qwen_cw6 KLLM_CONSUMER_WARPS=6
BARGS="--workload llama2-7b --steps 256"
run l7b_cw6 KLLM_CONSUMER_WARPS=6
run l7b_cw8 KLLM_CONSUMER_WARPS=8
KLLM_CONSUMER_WARPS=6 timeout 200 python tools/phase_timeline.py --pos 256 > ${O}_timeline_tiny_pos256_cw6.txt 2>${O}_timeline.err; cat ${O}_timeline_tiny_pos256_cw6.txt
