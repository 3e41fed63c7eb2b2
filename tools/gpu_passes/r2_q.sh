#!/bin/bash
# round 2, pass Q (1 GPU): register-hygiene kernel (no local-memory Params copy / poll buffers, real 6-way
# task round-robin, per-phase task rows, descriptor + norm-weight prefetch): parity suite, then A/B of the
# consumer-warp count and stage size on the same box
set -u
mkdir -p gpurun_out
O=gpurun_out/r2q
timeout 1500 python -m pytest tests -m gpu -x -q > ${O}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 ${O}_pytest_gpu.log | cut -c1-250
run() { # name, env..., uses BARGS
  name=$1; shift
  env "$@" timeout 300 python bench.py --reps 3 --no-cpu-baseline --no-exact ${BARGS} > ${O}_bench_${name}.json 2> ${O}_bench_${name}.err; rc=$?
  python -c "
import json
try:
    d=json.load(open('${O}_bench_${name}.json'));print('   ${name}',round(d['value'],1),round(d['e2e']['value'],1),{k:round(v) for k,v in d['by_position_tok_s'].items()},round(d['roofline']['frac'],3))
except Exception as e: print('   ${name} FAILED rc=$rc', e)"
}
BARGS="--steps 256"
run tiny_cw6 A=1
run tiny_cw8 KLLM_CONSUMER_WARPS=8
run tiny_cw8_rows4 KLLM_CONSUMER_WARPS=8 KLLM_TASK_ROWS_RT=4
run tiny_cw6_st44 KLLM_STAGE_BYTES=45056
run tiny_cw8_st44 KLLM_CONSUMER_WARPS=8 KLLM_STAGE_BYTES=45056
run tiny_cw8_pf12 KLLM_CONSUMER_WARPS=8 KLLM_PREFETCH_STAGES=12
BARGS="--workload llama2-7b-int8 --steps 128"
run int8_cw16 A=1
run int8_cw14 KLLM_CONSUMER_WARPS=14
run int8_cw8 KLLM_CONSUMER_WARPS=8
BARGS="--workload qwen2.5-0.5b --steps 256"
run qwen_cw6 A=1
run qwen_cw8 KLLM_CONSUMER_WARPS=8
BARGS="--workload llama2-7b --steps 64"
run l7b_cw6 A=1
run l7b_cw8 KLLM_CONSUMER_WARPS=8
KLLM_MODE=fast timeout 200 python tools/phase_timeline.py --pos 256 > ${O}_timeline_tiny_cw6_pos256.txt 2>>${O}_timeline.err; head -12 ${O}_timeline_tiny_cw6_pos256.txt
KLLM_MODE=fast KLLM_CONSUMER_WARPS=8 timeout 200 python tools/phase_timeline.py --pos 256 > ${O}_timeline_tiny_cw8_pos256.txt 2>>${O}_timeline.err; head -12 ${O}_timeline_tiny_cw8_pos256.txt
