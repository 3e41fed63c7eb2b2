#!/bin/bash
# coalesced bulk copies: parity + speed on the three model sizes
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_decoder_gpu.py -x -q -m gpu > gpurun_out/h_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/h_pytest.log
for wl in tinyllama-1.1b llama2-7b-int8 qwen2.5-0.5b; do
  echo "== $wl"
  timeout 400 python bench.py --workload $wl --steps 512 --warmup 8 --no-cpu-baseline 2> gpurun_out/h_bench_$wl.err | tee gpurun_out/h_bench_$wl.json | python -c "import json,sys; d=json.load(sys.stdin); print(d['value'], d['e2e']['value'], d['config'].get('engine'), d['roofline']['frac'])"
done
