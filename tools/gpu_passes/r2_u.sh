#!/bin/bash
# round 2, FINAL evidence pass (1 GPU): whole -m gpu suite, the driver's default bench line + both reference
# arms, the other BASELINE configs, ncu --set full of the final kernels (both numerics), launch list of a short
# bench run, phase timelines
set -u
mkdir -p gpurun_out
O=gpurun_out/r2u
timeout 1500 python -m pytest tests -m gpu -x -q > ${O}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 ${O}_pytest_gpu.log | cut -c1-250
timeout 600 python bench.py > ${O}_bench_default.json 2> ${O}_bench_default.err; echo "bench default rc=$?"; cut -c1-400 ${O}_bench_default.json
timeout 300 python bench.py --impl reference --steps 32 --warmup 3 > ${O}_bench_reference_arm.json 2> ${O}_bench_reference_arm.err; echo "bench ref rc=$?"; cut -c1-300 ${O}_bench_reference_arm.json
timeout 300 python bench.py --impl reference-cuda --steps 512 > ${O}_reference_cuda_tinyllama.json 2> ${O}_refcuda.err; echo "refcuda tiny rc=$?"; cut -c1-200 ${O}_reference_cuda_tinyllama.json
run() { # name, uses BARGS
  name=$1; shift
  timeout 400 python bench.py --reps 3 --no-cpu-baseline ${BARGS} > ${O}_bench_${name}.json 2> ${O}_bench_${name}.err; rc=$?
  python -c "
import json
try:
    d=json.load(open('${O}_bench_${name}.json'));x=d.get('exact');print('   ${name}',round(d['value'],1),round(d['e2e']['value'],1),{k:round(v) for k,v in d['by_position_tok_s'].items()},round(d['roofline']['frac'],3),'| exact:',x and (round(x['value'],1),{k:round(v) for k,v in x['by_position_tok_s'].items()},round(x['roofline_frac'],3)))
except Exception as e: print('   ${name} FAILED rc=$rc', e)"
}
BARGS="--workload llama2-7b-int8 --steps 256"
run int8
BARGS="--workload qwen2.5-0.5b --steps 1024"
run qwen
BARGS="--workload llama2-7b --steps 128"
run l7b
cap() { # name mode workload steps start
  KLLM_MODE=$2 timeout 300 ncu --set full --clock-control none --import-source on -k regex:decode_megakernel --launch-skip 1 -c 1 -f -o ${O}_mega_$1 \
     python tools/run_decode_once.py --workload $3 --steps $4 --start $5 > ${O}_ncu_$1.log 2>&1; echo "ncu $1 rc=$?"; tail -1 ${O}_ncu_$1.log | cut -c1-160
}
cap tiny_fast fast tinyllama-1.1b 4 512
cap int8_fast fast llama2-7b-int8 2 512
cap tiny_exact exact tinyllama-1.1b 4 512
cap int8_exact exact llama2-7b-int8 2 512
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ${O}_launches.csv python bench.py --steps 8 --warmup 3 --reps 1 --no-cpu-baseline --no-exact > ${O}_launches_bench.log 2>&1; echo "launch list rc=$?"
KLLM_MODE=fast timeout 200 python tools/phase_timeline.py --pos 256 > ${O}_timeline_tiny_fast_pos256.txt 2>>${O}_timeline.err
KLLM_MODE=fast timeout 200 python tools/phase_timeline.py --pos 1023 > ${O}_timeline_tiny_fast_pos1023.txt 2>>${O}_timeline.err
KLLM_MODE=fast timeout 300 python tools/phase_timeline.py --workload llama2-7b-int8 --pos 1023 > ${O}_timeline_int8_fast_pos1023.txt 2>>${O}_timeline.err
tail -12 ${O}_timeline_tiny_fast_pos256.txt | cut -c1-500
ls -la gpurun_out/ | grep r2u | awk '{print $5, $9}'
