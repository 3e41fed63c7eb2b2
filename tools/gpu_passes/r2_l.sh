#!/bin/bash
# round 2, pass L: fused attention back for head_size < 128 (split stays for 128), software-pipelined P.V chain
set -u
mkdir -p gpurun_out
O=gpurun_out/r2l
timeout 1500 python -m pytest tests/test_decoder_gpu.py tests/test_prefill_gpu.py tests/test_z_host_cpp.py -m gpu -x -q > ${O}_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 ${O}_pytest.log | cut -c1-220
run() { # name, env..., -- bench args
  name=$1; shift
  env "$@" timeout 300 python bench.py --reps 3 --no-cpu-baseline ${BARGS} > ${O}_bench_${name}.json 2> ${O}_bench_${name}.err; echo "bench ${name} rc=$?"
  python -c "
import json;d=json.load(open('${O}_bench_${name}.json'));print('   ${name}',round(d['value'],1),round(d['e2e']['value'],1),{k:round(v) for k,v in d['by_position_tok_s'].items()},round(d['roofline']['frac'],3))"
}
BARGS="--steps 1024"
run tiny A=1
run tiny_sp4 KLLM_ATTN_SPLIT=4
BARGS="--workload llama2-7b-int8 --steps 256"
run int8_fast KLLM_INT8_MODE=fast
run int8_exact KLLM_INT8_MODE=exact
run int8_fast_sp1 KLLM_INT8_MODE=fast KLLM_ATTN_SPLIT=1
BARGS="--workload qwen2.5-0.5b --steps 1024"
run qwen A=1
BARGS="--workload llama2-7b --steps 256"
run l7b A=1
timeout 200 python tools/phase_timeline.py --pos 1023 > ${O}_timeline_tiny_pos1023.txt 2>${O}_timeline.err; cat ${O}_timeline_tiny_pos1023.txt
KLLM_INT8_MODE=fast timeout 200 python tools/phase_timeline.py --workload llama2-7b-int8 --pos 1023 > ${O}_timeline_int8_pos1023.txt 2>>${O}_timeline.err; cat ${O}_timeline_int8_pos1023.txt
