#!/bin/bash
# round 2, pass G: rows per consumer task 1 / 2 / 4 (variant libraries), fp32 (6 warps) and int8 (16 / 8 warps)
set -u
mkdir -p gpurun_out
O=gpurun_out/r2g
V=$PWD/kuiperllama_b200/lib/variants
run() { # name, env..., -- bench args
  name=$1; shift
  env "$@" timeout 300 python bench.py --reps 3 --no-cpu-baseline ${BARGS} > ${O}_bench_${name}.json 2> ${O}_bench_${name}.err; echo "bench ${name} rc=$?"
  python -c "
import json;d=json.load(open('${O}_bench_${name}.json'));print('   ${name}',round(d['value'],1),round(d['e2e']['value'],1),{k:round(v) for k,v in d['by_position_tok_s'].items()},round(d['roofline']['frac'],3))"
}
BARGS="--steps 1024"
run tiny_t4 A=1
run tiny_t2 KLLM_LIB=$V/libkllm_t2.so
run tiny_t1 KLLM_LIB=$V/libkllm_t1.so
run tiny_t1_cw8 KLLM_LIB=$V/libkllm_t1.so KLLM_CONSUMER_WARPS=8
BARGS="--workload llama2-7b-int8 --steps 256"
run int8_t1_cw16 KLLM_LIB=$V/libkllm_t1.so KLLM_CONSUMER_WARPS=16
run int8_t2_cw16 KLLM_LIB=$V/libkllm_t2.so KLLM_CONSUMER_WARPS=16
run int8_t1_cw8 KLLM_LIB=$V/libkllm_t1.so KLLM_CONSUMER_WARPS=8
run int8_t1_cw6 KLLM_LIB=$V/libkllm_t1.so KLLM_CONSUMER_WARPS=6
BARGS="--workload qwen2.5-0.5b --steps 1024"
run qwen_t1 KLLM_LIB=$V/libkllm_t1.so
run qwen_t2 KLLM_LIB=$V/libkllm_t2.so
KLLM_LIB=$V/libkllm_t1.so timeout 600 python -m pytest tests/test_decoder_gpu.py -m gpu -x -q -k "not full_size" > ${O}_pytest_t1.log 2>&1; echo "pytest t1 rc=$?"; tail -3 ${O}_pytest_t1.log
KLLM_LIB=$V/libkllm_t1.so KLLM_CONSUMER_WARPS=16 timeout 300 python tools/phase_timeline.py --workload llama2-7b-int8 --pos 64 > ${O}_timeline_int8_t1.txt 2>${O}_timeline.err; cat ${O}_timeline_int8_t1.txt
