#!/bin/bash
# round 2, pass S (1 GPU): int8 fast rows on the tensor cores (mma.sync m16n8k32 s8, a warp pair per ring stage)
# against the dp4a form; task rows back to 4
set -u
mkdir -p gpurun_out
O=gpurun_out/r2s
timeout 1500 python -m pytest tests -m gpu -x -q > ${O}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 ${O}_pytest_gpu.log | cut -c1-250
grep -n "FAILED\|Error\|assert" ${O}_pytest_gpu.log | head -20
run() { # name, env..., uses BARGS
  name=$1; shift
  env "$@" timeout 300 python bench.py --reps 3 --no-cpu-baseline --no-exact ${BARGS} > ${O}_bench_${name}.json 2> ${O}_bench_${name}.err; rc=$?
  python -c "
import json
try:
    d=json.load(open('${O}_bench_${name}.json'));print('   ${name}',round(d['value'],1),round(d['e2e']['value'],1),{k:round(v) for k,v in d['by_position_tok_s'].items()},round(d['roofline']['frac'],3))
except Exception as e: print('   ${name} FAILED rc=$rc', e)"
}
BARGS="--workload llama2-7b-int8 --steps 128"
run int8_mma A=1
run int8_dp4a KLLM_INT8_MMA=0
run int8_mma_cw16 KLLM_CONSUMER_WARPS=16
run int8_mma_cw8 KLLM_CONSUMER_WARPS=8
run int8_mma_pf12 KLLM_PREFETCH_STAGES=14
BARGS="--steps 256"
run tiny A=1
BARGS="--workload qwen2.5-0.5b --steps 256"
run qwen A=1
KLLM_MODE=fast timeout 300 python tools/phase_timeline.py --workload llama2-7b-int8 --pos 256 > ${O}_timeline_int8_pos256.txt 2>>${O}_timeline.err; head -24 ${O}_timeline_int8_pos256.txt
