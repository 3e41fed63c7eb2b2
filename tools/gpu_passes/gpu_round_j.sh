#!/bin/bash
# barrier-free token (tagged hand-offs): parity + speed
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_decoder_gpu.py tests/test_z_host_cpp.py -x -q -m gpu > gpurun_out/j_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/j_pytest.log
for tg in 2 1; do
  echo "== KLLM_MEGA_TAGGED=$tg"
  KLLM_MEGA_TAGGED=$tg timeout 300 python bench.py --steps 1024 --warmup 16 --no-cpu-baseline 2> gpurun_out/j_bench_tg$tg.err | tee gpurun_out/j_bench_tg$tg.json | python -c "import json,sys; d=json.load(sys.stdin); print(d['value'], d['e2e']['value'], d['roofline']['frac'])"
done
for wl in llama2-7b-int8 qwen2.5-0.5b; do
  echo "== $wl"
  timeout 400 python bench.py --workload $wl --steps 512 --warmup 8 --no-cpu-baseline 2> gpurun_out/j_bench_$wl.err | tee gpurun_out/j_bench_$wl.json | python -c "import json,sys; d=json.load(sys.stdin); print(d['value'], d['e2e']['value'], d['config'].get('engine'), d['roofline']['frac'])"
done
timeout 300 python tools/phase_timeline.py --pos 256 > gpurun_out/j_timeline_pos256.txt 2>&1; tail -11 gpurun_out/j_timeline_pos256.txt | head -9
