#!/bin/bash
# round 2, pass C: spill study -- ring depth 6 (no L1 left) vs 5 (32 KB L1), int8 consumer warps 8 / 14 / 16
set -u
mkdir -p gpurun_out
O=gpurun_out/r2c
timeout 900 python -m pytest tests/test_decoder_gpu.py -m gpu -x -q > ${O}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 ${O}_pytest.log
run() { # name, env..., -- bench args
  name=$1; shift
  env "$@" timeout 300 python bench.py --reps 3 --no-cpu-baseline ${BARGS} > ${O}_bench_${name}.json 2> ${O}_bench_${name}.err; echo "bench ${name} rc=$?"
  python -c "
import json;d=json.load(open('${O}_bench_${name}.json'));print('   ${name}',round(d['value'],1),round(d['e2e']['value'],1),{k:round(v) for k,v in d['by_position_tok_s'].items()},round(d['roofline']['frac'],3))"
}
BARGS="--steps 1024"
run tiny_s6 KLLM_STAGES=6
run tiny_s5 KLLM_STAGES=5
run tiny_s5_pf12 KLLM_STAGES=5 KLLM_PREFETCH_STAGES=12
BARGS="--workload llama2-7b-int8 --steps 256"
run int8_cw16_s6 KLLM_CONSUMER_WARPS=16
run int8_cw16_s4 KLLM_CONSUMER_WARPS=16 KLLM_STAGES=4
run int8_cw14_s4 KLLM_CONSUMER_WARPS=14 KLLM_STAGES=4
run int8_cw8_s4 KLLM_CONSUMER_WARPS=8 KLLM_STAGES=4
BARGS="--workload qwen2.5-0.5b --steps 1024"
run qwen_s6 KLLM_STAGES=6
run qwen_s5 KLLM_STAGES=5
timeout 200 python tools/phase_timeline.py --pos 256 > ${O}_timeline_tiny_pos256.txt 2>${O}_timeline.err; cat ${O}_timeline_tiny_pos256.txt
KLLM_STAGES=5 timeout 200 python tools/phase_timeline.py --pos 256 > ${O}_timeline_tiny_pos256_s5.txt 2>>${O}_timeline.err; cat ${O}_timeline_tiny_pos256_s5.txt
