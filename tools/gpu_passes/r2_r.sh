#!/bin/bash
# round 2, pass R (1 GPU): concurrent q/k/v polls + batched partial merges in the attention phases, new
# consumer-warp defaults (fp32 8, int8 14), ring-aware task rows; A/B of the ring stage size and the task rows
set -u
mkdir -p gpurun_out
O=gpurun_out/r2r
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
run tiny A=1
run tiny_rows4 KLLM_TASK_ROWS_RT=4
run tiny_rows2 KLLM_TASK_ROWS_RT=2
run tiny_st24 KLLM_STAGE_BYTES=24576
run tiny_st16 KLLM_STAGE_BYTES=16384
run tiny_st24_rows4 KLLM_STAGE_BYTES=24576 KLLM_TASK_ROWS_RT=4
run tiny_exact KLLM_MODE=exact
BARGS="--workload llama2-7b-int8 --steps 128"
run int8 A=1
run int8_rows4 KLLM_TASK_ROWS_RT=4
run int8_rows2 KLLM_TASK_ROWS_RT=2
run int8_st20 KLLM_STAGE_BYTES=20480
run int8_exact KLLM_MODE=exact
BARGS="--workload qwen2.5-0.5b --steps 256"
run qwen A=1
run qwen_st24 KLLM_STAGE_BYTES=24576
BARGS="--workload llama2-7b --steps 64"
run l7b A=1
run l7b_rows4 KLLM_TASK_ROWS_RT=4
KLLM_MODE=fast timeout 200 python tools/phase_timeline.py --pos 256 > ${O}_timeline_tiny_pos256.txt 2>>${O}_timeline.err; head -24 ${O}_timeline_tiny_pos256.txt
KLLM_MODE=fast timeout 200 python tools/phase_timeline.py --pos 1023 > ${O}_timeline_tiny_pos1023.txt 2>>${O}_timeline.err; tail -14 ${O}_timeline_tiny_pos1023.txt
KLLM_MODE=fast timeout 300 python tools/phase_timeline.py --workload llama2-7b-int8 --pos 256 > ${O}_timeline_int8_pos256.txt 2>>${O}_timeline.err; head -24 ${O}_timeline_int8_pos256.txt
