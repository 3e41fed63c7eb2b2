#!/bin/bash
# round 2, pass T (1 GPU): int8 fast rows in the TEAM form (four warps share a ring stage: stages are released one
# after the other, the refill overlaps the arithmetic) -- mma.sync for <= 8 short rows, dp4a for 1-2 long rows
set -u
mkdir -p gpurun_out
O=gpurun_out/r2t
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
run int8_team12 A=1
run int8_team16 KLLM_CONSUMER_WARPS=16
run int8_team8 KLLM_CONSUMER_WARPS=8
run int8_pairs14 KLLM_CONSUMER_WARPS=14
run int8_dp4a KLLM_INT8_MMA=0 KLLM_CONSUMER_WARPS=14
run int8_team12_st20 KLLM_STAGE_BYTES=20480
KLLM_MODE=fast timeout 300 python tools/phase_timeline.py --workload llama2-7b-int8 --pos 256 > ${O}_timeline_int8_pos256.txt 2>>${O}_timeline.err; head -12 ${O}_timeline_int8_pos256.txt | cut -c1-400
