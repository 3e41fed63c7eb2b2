#!/bin/bash
# L2 prefetch run-ahead experiment on the persistent engine
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_decoder_gpu.py -x -q > gpurun_out/b_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/b_pytest.log
for pf in 0 6 12 24 48; do
  echo "== KLLM_PREFETCH_STAGES=$pf"
  KLLM_PREFETCH_STAGES=$pf timeout 300 python bench.py --steps 1024 --warmup 16 --no-cpu-baseline 2> gpurun_out/b_bench_pf$pf.err | tee gpurun_out/b_bench_pf$pf.json | python -c "import json,sys; d=json.load(sys.stdin); print(d['value'], d['e2e']['value'], d['roofline']['frac'])"
done
timeout 300 python tools/phase_timeline.py --pos 256 > gpurun_out/b_timeline_pos256.txt 2>&1; tail -12 gpurun_out/b_timeline_pos256.txt
