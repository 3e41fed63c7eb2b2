#!/bin/bash
# 1-GPU pass: the other BASELINE.json configs, full test suite, default bench, ncu refresh
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/g_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/g_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/g_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/g_smoke.log
for wl in llama2-7b-int8 qwen2.5-0.5b llama2-7b stories15m; do
  echo "== $wl"
  steps=512; [ $wl = stories15m ] && steps=240
  timeout 400 python bench.py --workload $wl --steps $steps --warmup 8 --no-cpu-baseline 2> gpurun_out/g_bench_$wl.err | tee gpurun_out/g_bench_$wl.json | python -c "import json,sys; d=json.load(sys.stdin); print(d['value'], d['e2e']['value'], d['config'].get('engine'), d['roofline']['frac'])"
  tail -2 gpurun_out/g_bench_$wl.err | cut -c1-200
done
timeout 600 python bench.py > gpurun_out/g_bench.json 2> gpurun_out/g_bench.err; echo "bench rc=$?"; cat gpurun_out/g_bench.json | cut -c1-400
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/g_launches.csv \
   python bench.py --steps 64 --warmup 3 --no-cpu-baseline > gpurun_out/g_ncu_bench.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_megakernel -c 1 -f -o gpurun_out/g_mega \
   python tools/run_decode_once.py --steps 16 --start 504 > gpurun_out/g_ncu_mega.log 2>&1; echo "ncu mega rc=$?"
