#!/bin/bash
# round 2, final evidence pass (1 GPU): whole -m gpu suite, default bench + reference arms, ncu --set full of the
# final kernels in both numerics, launch list of a short bench run
set -u
mkdir -p gpurun_out
O=gpurun_out/r2p
timeout 2400 python -m pytest tests -m gpu -x -q > ${O}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 ${O}_pytest_gpu.log | cut -c1-250
timeout 600 python bench.py > ${O}_bench_default.json 2> ${O}_bench_default.err; echo "bench default rc=$?"; cut -c1-300 ${O}_bench_default.json
timeout 400 python bench.py --impl reference --steps 32 --warmup 3 > ${O}_bench_reference_arm.json 2> ${O}_bench_reference_arm.err; echo "bench ref rc=$?"; cut -c1-200 ${O}_bench_reference_arm.json
timeout 400 python bench.py --impl reference-cuda --steps 1024 > ${O}_reference_cuda_tinyllama.json 2> ${O}_refcuda.err; echo "refcuda tiny rc=$?"; cut -c1-160 ${O}_reference_cuda_tinyllama.json
for wl in stories15m qwen2.5-0.5b; do
  timeout 300 python bench.py --impl reference --workload $wl --steps 32 --warmup 3 > ${O}_bench_reference_arm_${wl}.json 2>> ${O}_bench_reference_arm.err; echo "ref $wl rc=$?"; cut -c1-160 ${O}_bench_reference_arm_${wl}.json
done
cap() { # name mode workload steps start
  KLLM_MODE=$2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_megakernel --launch-skip 1 -c 1 -f -o ${O}_mega_$1 \
     python tools/run_decode_once.py --workload $3 --steps $4 --start $5 > ${O}_ncu_$1.log 2>&1; echo "ncu $1 rc=$?"; tail -1 ${O}_ncu_$1.log | cut -c1-200
}
cap tiny_fast fast tinyllama-1.1b 4 512
cap tiny_exact exact tinyllama-1.1b 4 512
cap int8_fast fast llama2-7b-int8 2 512
cap int8_exact exact llama2-7b-int8 2 512
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ${O}_launches.csv python bench.py --steps 8 --warmup 3 --reps 1 --no-cpu-baseline --no-exact > ${O}_launches_bench.log 2>&1; echo "launch list rc=$?"
ls -la gpurun_out/ | grep r2p | awk '{print $5, $9}'
