#!/bin/bash
# round 2, pass H: int8 fast (dp4a fixed-point) mode -- tolerance test + bench
set -u
mkdir -p gpurun_out
O=gpurun_out/r2h
timeout 900 python -m pytest tests/test_decoder_gpu.py -m gpu -x -q -k "int8_fast or prompt_call" > ${O}_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 ${O}_pytest.log | cut -c1-220
run() { # name, env..., -- bench args
  name=$1; shift
  env "$@" timeout 300 python bench.py --reps 3 --no-cpu-baseline ${BARGS} > ${O}_bench_${name}.json 2> ${O}_bench_${name}.err; echo "bench ${name} rc=$?"
  python -c "
import json;d=json.load(open('${O}_bench_${name}.json'));print('   ${name}',round(d['value'],1),round(d['e2e']['value'],1),{k:round(v) for k,v in d['by_position_tok_s'].items()},round(d['roofline']['frac'],3))"
}
BARGS="--workload llama2-7b-int8 --steps 256"
run int8_fast_cw16 KLLM_INT8_MODE=fast KLLM_CONSUMER_WARPS=16
run int8_fast_cw8 KLLM_INT8_MODE=fast KLLM_CONSUMER_WARPS=8
run int8_exact_cw16 KLLM_INT8_MODE=exact KLLM_CONSUMER_WARPS=16
KLLM_INT8_MODE=fast timeout 300 python tools/phase_timeline.py --workload llama2-7b-int8 --pos 64 > ${O}_timeline_int8_fast.txt 2>${O}_timeline.err; cat ${O}_timeline_int8_fast.txt
timeout 600 python -m pytest tests/test_prefill_gpu.py -m gpu -q > ${O}_pytest_prefill.log 2>&1; echo "pytest prefill rc=$?"; tail -5 ${O}_pytest_prefill.log | cut -c1-200
