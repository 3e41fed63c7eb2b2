#!/bin/bash
# round 2, FINAL evidence pass, second take (the first one, r2_u.sh, ran green -- 161 passed, see
# profiles/r02u_final_pass_stdout.log -- but its four .ncu-rep files exceeded the 64 MiB that come back from the
# box): same measurements minus the test suite; the ncu reports are summarised ON THE BOX and only the two of
# the default (fast) numerics travel
set -u
mkdir -p gpurun_out
O=gpurun_out/r2v
timeout 600 python bench.py > ${O}_bench_default.json 2> ${O}_bench_default.err; echo "bench default rc=$?"; cut -c1-300 ${O}_bench_default.json
timeout 300 python bench.py --impl reference --steps 32 --warmup 3 > ${O}_bench_reference_arm.json 2> ${O}_bench_reference_arm.err; echo "bench ref rc=$?"; cut -c1-200 ${O}_bench_reference_arm.json
timeout 300 python bench.py --impl reference-cuda --steps 512 > ${O}_reference_cuda_tinyllama.json 2> ${O}_refcuda.err; echo "refcuda tiny rc=$?"; cut -c1-160 ${O}_reference_cuda_tinyllama.json
timeout 300 python bench.py --impl reference-cuda --workload llama2-7b-int8 --steps 128 > ${O}_reference_cuda_llama2-7b-int8.json 2>> ${O}_refcuda.err; echo "refcuda int8 rc=$?"; cut -c1-160 ${O}_reference_cuda_llama2-7b-int8.json
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
     python tools/run_decode_once.py --workload $3 --steps $4 --start $5 > ${O}_ncu_$1.log 2>&1; echo "ncu $1 rc=$?"
  python tools/ncu_summarize.py full ${O}_mega_$1.ncu-rep ${O}_megakernel_$1_ncu.txt > /dev/null 2>&1
  python tools/ncu_hot.py ${O}_mega_$1.ncu-rep 40 > ${O}_megakernel_$1_hot.txt 2>/dev/null
  head -12 ${O}_megakernel_$1_ncu.txt | cut -c1-200
}
cap tiny_fast fast tinyllama-1.1b 4 512
cap int8_fast fast llama2-7b-int8 2 512
cap tiny_exact exact tinyllama-1.1b 4 512
cap int8_exact exact llama2-7b-int8 2 512
cp profiles/dominant_kernel_traffic.json ${O}_dominant_kernel_traffic.json
python tools/ncu_summarize.py traffic ${O}_mega_tiny_fast.ncu-rep ${O}_dominant_kernel_traffic.json tinyllama-1.1b megakernel 4 "round 2 final pass: decode_megakernel<8,false,false>, fast numerics, 4 positions at context 513..516 (ncu --set full, launch 2 of tools/run_decode_once.py --steps 4 --start 512)" > /dev/null 2>&1
python tools/ncu_summarize.py traffic ${O}_mega_int8_fast.ncu-rep ${O}_dominant_kernel_traffic.json llama2-7b-int8 megakernel 2 "round 2 final pass: decode_megakernel<14,true,false>, fast numerics (dp4a rows), 2 positions at context 513..514" > /dev/null 2>&1
cat ${O}_dominant_kernel_traffic.json | head -30
rm -f ${O}_mega_tiny_exact.ncu-rep ${O}_mega_int8_exact.ncu-rep
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ${O}_launches.csv python bench.py --steps 8 --warmup 3 --reps 1 --no-cpu-baseline --no-exact > ${O}_launches_bench.log 2>&1; echo "launch list rc=$?"
KLLM_MODE=fast timeout 200 python tools/phase_timeline.py --pos 256 > ${O}_timeline_tiny_fast_pos256.txt 2>>${O}_timeline.err
KLLM_MODE=fast timeout 200 python tools/phase_timeline.py --pos 1023 > ${O}_timeline_tiny_fast_pos1023.txt 2>>${O}_timeline.err
KLLM_MODE=fast timeout 300 python tools/phase_timeline.py --workload llama2-7b-int8 --pos 1023 > ${O}_timeline_int8_fast_pos1023.txt 2>>${O}_timeline.err
du -sm gpurun_out | cut -f1
if [ "$(du -sm gpurun_out | cut -f1)" -gt 58 ]; then rm -f ${O}_mega_int8_fast.ncu-rep; fi
if [ "$(du -sm gpurun_out | cut -f1)" -gt 58 ]; then rm -f ${O}_mega_tiny_fast.ncu-rep; fi
ls -la gpurun_out/ | grep r2v | awk '{print $5, $9}'
