#!/bin/bash
# round 2, pass F: ncu --set full of the current kernels (fp32 6 warps, int8 16 warps) + prefill tests + bench
set -u
mkdir -p gpurun_out
O=gpurun_out/r2f
timeout 600 python -m pytest tests/test_prefill_gpu.py -m gpu -q > ${O}_pytest_prefill.log 2>&1; echo "pytest prefill rc=$?"; tail -8 ${O}_pytest_prefill.log | cut -c1-200
timeout 300 python bench.py --steps 1024 --reps 3 --no-cpu-baseline > ${O}_bench_tiny.json 2> ${O}_bench_tiny.err; echo "bench tiny rc=$?"; python -c "
import json;d=json.load(open('${O}_bench_tiny.json'));print(round(d['value'],1),round(d['e2e']['value'],1),d['by_position_tok_s'],round(d['roofline']['frac'],3))"
timeout 200 python tools/phase_timeline.py --pos 256 > ${O}_timeline_tiny_pos256.txt 2>${O}_timeline.err; cat ${O}_timeline_tiny_pos256.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:decode_megakernel --launch-skip 1 -c 1 -f -o ${O}_mega_tiny \
   python tools/run_decode_once.py --steps 4 --start 256 > ${O}_ncu_tiny.log 2>&1; echo "ncu tiny rc=$?"; tail -2 ${O}_ncu_tiny.log
timeout 700 ncu --set full --clock-control none --import-source on -k regex:decode_megakernel --launch-skip 1 -c 1 -f -o ${O}_mega_int8 \
   python tools/run_decode_once.py --workload llama2-7b-int8 --steps 2 --start 64 > ${O}_ncu_int8.log 2>&1; echo "ncu int8 rc=$?"; tail -2 ${O}_ncu_int8.log
