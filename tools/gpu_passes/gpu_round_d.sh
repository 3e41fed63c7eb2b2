#!/bin/bash
# tagged-exchange megakernel on one GPU: parity + speed against the barrier version
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_decoder_gpu.py tests/test_z_host_cpp.py -x -q -m gpu > gpurun_out/d_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/d_pytest.log
for tg in 1 0; do
  echo "== KLLM_MEGA_TAGGED=$tg"
  KLLM_MEGA_TAGGED=$tg timeout 300 python bench.py --steps 1024 --warmup 16 --no-cpu-baseline 2> gpurun_out/d_bench_tg$tg.err | tee gpurun_out/d_bench_tg$tg.json | python -c "import json,sys; d=json.load(sys.stdin); print(d['value'], d['e2e']['value'], d['roofline']['frac'])"
done
timeout 300 python tools/phase_timeline.py --pos 256 > gpurun_out/d_timeline_pos256.txt 2>&1; tail -12 gpurun_out/d_timeline_pos256.txt
