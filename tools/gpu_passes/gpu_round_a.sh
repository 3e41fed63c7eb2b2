#!/bin/bash
# 1-GPU measurement pass: tests, bench, launch list, one full ncu capture of the megakernel.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?" 
tail -5 gpurun_out/a_pytest.log
timeout 600 python bench.py > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; echo "bench rc=$?"
cat gpurun_out/a_bench.json
timeout 300 python bench.py --impl reference --steps 64 --warmup 3 > gpurun_out/a_bench_ref.json 2> gpurun_out/a_bench_ref.err; echo "ref rc=$?"
cat gpurun_out/a_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/a_launches.csv \
   python bench.py --steps 64 --warmup 3 --no-cpu-baseline > gpurun_out/a_ncu_bench.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_megakernel -c 1 -f -o gpurun_out/a_mega \
   python tools/run_decode_once.py --steps 16 --start 504 > gpurun_out/a_ncu_mega.log 2>&1; echo "ncu mega rc=$?"
tail -3 gpurun_out/a_ncu_mega.log
