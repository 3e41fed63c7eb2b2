#!/bin/bash
# round 2, pass A: full-size parity tests (C2/C3/C4), phase timelines, new bench lines, the
# reference's own CUDA path timed on the same GPU, ncu --set full of the round-1 kernels (baseline)
set -u
mkdir -p gpurun_out
O=gpurun_out/r2a
timeout 900 python -m pytest tests/test_decoder_gpu.py -m gpu -x -q -k "full_size" > ${O}_pytest_full.log 2>&1; echo "pytest full rc=$?"; tail -5 ${O}_pytest_full.log
timeout 200 python tools/phase_timeline.py --pos 256 > ${O}_timeline_tiny_pos256.txt 2>${O}_timeline.err; echo "timeline rc=$?"; cat ${O}_timeline_tiny_pos256.txt
timeout 200 python tools/phase_timeline.py --pos 1023 > ${O}_timeline_tiny_pos1023.txt 2>>${O}_timeline.err; cat ${O}_timeline_tiny_pos1023.txt
timeout 300 python tools/phase_timeline.py --workload llama2-7b-int8 --pos 64 > ${O}_timeline_int8_pos64.txt 2>>${O}_timeline.err; cat ${O}_timeline_int8_pos64.txt
timeout 300 python bench.py --steps 1024 > ${O}_bench_tiny.json 2> ${O}_bench_tiny.err; echo "bench tiny rc=$?"; cut -c1-600 ${O}_bench_tiny.json
timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > ${O}_bench_tiny_k20.json 2> ${O}_bench_tiny_k20.err; echo "bench tiny k20 rc=$?"; cut -c1-300 ${O}_bench_tiny_k20.json
timeout 400 python bench.py --workload llama2-7b-int8 --steps 256 --reps 3 --no-cpu-baseline > ${O}_bench_int8.json 2> ${O}_bench_int8.err; echo "bench int8 rc=$?"; cut -c1-400 ${O}_bench_int8.json
timeout 200 python bench.py --impl reference-cuda --workload tinyllama-1.1b --steps 1024 > ${O}_refcuda_tiny.json 2> ${O}_refcuda_tiny.err; echo "refcuda tiny rc=$?"; cat ${O}_refcuda_tiny.json
timeout 400 python bench.py --impl reference-cuda --workload llama2-7b-int8 --steps 128 > ${O}_refcuda_int8.json 2> ${O}_refcuda_int8.err; echo "refcuda int8 rc=$?"; cat ${O}_refcuda_int8.json
timeout 400 ncu --set full --clock-control none --import-source on -k regex:decode_megakernel --launch-skip 1 -c 1 -f -o ${O}_mega_tiny \
   python tools/run_decode_once.py --steps 4 --start 256 > ${O}_ncu_tiny.log 2>&1; echo "ncu tiny rc=$?"; tail -2 ${O}_ncu_tiny.log
timeout 700 ncu --set full --clock-control none --import-source on -k regex:decode_megakernel --launch-skip 1 -c 1 -f -o ${O}_mega_int8 \
   python tools/run_decode_once.py --workload llama2-7b-int8 --steps 2 --start 64 > ${O}_ncu_int8.log 2>&1; echo "ncu int8 rc=$?"; tail -2 ${O}_ncu_int8.log
