#!/bin/bash
# round 2, pass J: mbarrier try_wait with a suspend-time hint (20 us default build, 1 us variant), shallower polls
set -u
mkdir -p gpurun_out
O=gpurun_out/r2j
V=$PWD/kuiperllama_b200/lib/variants
run() { # name, env..., -- bench args
  name=$1; shift
  env "$@" timeout 300 python bench.py --reps 3 --no-cpu-baseline ${BARGS} > ${O}_bench_${name}.json 2> ${O}_bench_${name}.err; echo "bench ${name} rc=$?"
  python -c "
import json;d=json.load(open('${O}_bench_${name}.json'));print('   ${name}',round(d['value'],1),round(d['e2e']['value'],1),{k:round(v) for k,v in d['by_position_tok_s'].items()},round(d['roofline']['frac'],3))"
}
BARGS="--workload llama2-7b-int8 --steps 256"
run int8_fast KLLM_INT8_MODE=fast
run int8_fast_h1 KLLM_INT8_MODE=fast KLLM_LIB=$V/libkllm_hint1us.so
run int8_exact KLLM_INT8_MODE=exact
BARGS="--steps 1024"
run tiny A=1
run tiny_h1 KLLM_LIB=$V/libkllm_hint1us.so
timeout 600 python -m pytest tests/test_decoder_gpu.py -m gpu -x -q -k "not full_size" > ${O}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 ${O}_pytest.log
KLLM_INT8_MODE=fast timeout 300 python tools/phase_timeline.py --workload llama2-7b-int8 --pos 64 > ${O}_timeline_int8_fast.txt 2>${O}_timeline.err; cat ${O}_timeline_int8_fast.txt
timeout 200 python tools/phase_timeline.py --pos 256 > ${O}_timeline_tiny.txt 2>>${O}_timeline.err; cat ${O}_timeline_tiny.txt
