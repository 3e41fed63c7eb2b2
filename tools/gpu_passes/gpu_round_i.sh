#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_megakernel -c 1 -f -o gpurun_out/i_mega_int8 \
   python tools/run_decode_once.py --workload llama2-7b-int8 --steps 4 --start 0 > gpurun_out/i_ncu_int8.log 2>&1; echo "ncu int8 rc=$?"; tail -3 gpurun_out/i_ncu_int8.log
timeout 300 python bench.py --steps 1024 --warmup 16 --no-cpu-baseline 2> gpurun_out/i_bench.err | tee gpurun_out/i_bench.json | python -c "import json,sys; d=json.load(sys.stdin); print(d['value'], d['e2e']['value'], d['config'].get('engine'), d['roofline']['frac'])"
