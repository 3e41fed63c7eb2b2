#!/bin/bash
# final 1-GPU pass of the round: default bench line + ncu evidence for the current kernel
set -u
mkdir -p gpurun_out
timeout 150 python bench.py > gpurun_out/m_bench.json 2> gpurun_out/m_bench.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/m_bench.json
timeout 100 ncu --set full --clock-control none --import-source on -k regex:decode_megakernel -c 1 -f -o gpurun_out/m_mega \
   python tools/run_decode_once.py --steps 8 --start 256 > gpurun_out/m_ncu_mega.log 2>&1; echo "ncu mega rc=$?"
timeout 80 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/m_launches.csv \
   python bench.py --steps 32 --warmup 3 --no-cpu-baseline > gpurun_out/m_ncu_bench.log 2>&1; echo "ncu launches rc=$?"
