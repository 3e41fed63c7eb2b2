#!/bin/bash
# round 2, pass M: fast numerics (flash-decoding attention + dp4a rows), idle attention warps gated at the
# hardware barrier, probabilities read by ld.shared in the P.V chain; A/B against the ungated variant
set -u
mkdir -p gpurun_out
O=gpurun_out/r2m
V=$PWD/kuiperllama_b200/lib/variants
timeout 1800 python -m pytest tests/test_decoder_gpu.py tests/test_prefill_gpu.py tests/test_z_host_cpp.py -m gpu -x -q > ${O}_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 ${O}_pytest.log | cut -c1-250
run() { # name, env..., -- bench args
  name=$1; shift
  env "$@" timeout 400 python bench.py --reps 3 --no-cpu-baseline ${BARGS} > ${O}_bench_${name}.json 2> ${O}_bench_${name}.err; echo "bench ${name} rc=$?"
  python -c "
import json;d=json.load(open('${O}_bench_${name}.json'));x=d.get('exact');print('   ${name}',round(d['value'],1),round(d['e2e']['value'],1),{k:round(v) for k,v in d['by_position_tok_s'].items()},round(d['roofline']['frac'],3),'| exact:',x and (round(x['value'],1),{k:round(v) for k,v in x['by_position_tok_s'].items()},round(x['roofline_frac'],3)))"
}
BARGS="--steps 1024"
run tiny A=1
BARGS="--steps 1024 --numerics exact"
run tiny_exact_nogate KLLM_LIB=$V/libkllm_nogate.so
BARGS="--workload llama2-7b-int8 --steps 256"
run int8 A=1
BARGS="--workload llama2-7b-int8 --steps 256 --numerics exact"
run int8_exact_nogate KLLM_LIB=$V/libkllm_nogate.so
BARGS="--workload qwen2.5-0.5b --steps 1024"
run qwen A=1
BARGS="--workload llama2-7b --steps 256"
run l7b A=1
timeout 200 python tools/phase_timeline.py --pos 1023 > ${O}_timeline_tiny_pos1023.txt 2>${O}_timeline.err; cat ${O}_timeline_tiny_pos1023.txt
KLLM_MODE=fast timeout 200 python tools/phase_timeline.py --pos 1023 > ${O}_timeline_tiny_fast_pos1023.txt 2>>${O}_timeline.err; cat ${O}_timeline_tiny_fast_pos1023.txt
timeout 200 python tools/phase_timeline.py --workload llama2-7b-int8 --pos 1023 > ${O}_timeline_int8_pos1023.txt 2>>${O}_timeline.err; cat ${O}_timeline_int8_pos1023.txt
KLLM_MODE=fast timeout 200 python tools/phase_timeline.py --workload llama2-7b-int8 --pos 1023 > ${O}_timeline_int8_fast_pos1023.txt 2>>${O}_timeline.err; cat ${O}_timeline_int8_fast_pos1023.txt
