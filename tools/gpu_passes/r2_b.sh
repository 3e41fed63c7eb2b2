#!/bin/bash
# round 2, pass B: task-based consumers (4 rows per warp task, packed reductions, 16 int8 warps) --
# whole -m gpu suite, phase timelines, bench lines
set -u
mkdir -p gpurun_out
O=gpurun_out/r2b
timeout 1200 python -m pytest tests -m gpu -x -q > ${O}_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 ${O}_pytest.log
timeout 200 python tools/phase_timeline.py --pos 256 > ${O}_timeline_tiny_pos256.txt 2>${O}_timeline.err; echo "timeline rc=$?"; cat ${O}_timeline_tiny_pos256.txt
timeout 300 python tools/phase_timeline.py --workload llama2-7b-int8 --pos 64 > ${O}_timeline_int8_pos64.txt 2>>${O}_timeline.err; cat ${O}_timeline_int8_pos64.txt
timeout 300 python bench.py --steps 1024 --no-cpu-baseline > ${O}_bench_tiny.json 2> ${O}_bench_tiny.err; echo "bench tiny rc=$?"; python -c "
import json;d=json.load(open('${O}_bench_tiny.json'));print(d['value'],d['e2e']['value'],d['by_position_tok_s'],d['roofline']['frac'])"
timeout 400 python bench.py --workload llama2-7b-int8 --steps 256 --reps 3 --no-cpu-baseline > ${O}_bench_int8.json 2> ${O}_bench_int8.err; echo "bench int8 rc=$?"; python -c "
import json;d=json.load(open('${O}_bench_int8.json'));print(d['value'],d['e2e']['value'],d['by_position_tok_s'],d['roofline']['frac'])"
KLLM_CONSUMER_WARPS=8 timeout 400 python bench.py --workload llama2-7b-int8 --steps 256 --reps 3 --no-cpu-baseline > ${O}_bench_int8_cw8.json 2> ${O}_bench_int8_cw8.err; echo "bench int8 cw8 rc=$?"; python -c "
import json;d=json.load(open('${O}_bench_int8_cw8.json'));print(d['value'],d['roofline']['frac'])"
timeout 300 python bench.py --workload qwen2.5-0.5b --steps 1024 --reps 3 --no-cpu-baseline > ${O}_bench_qwen.json 2> ${O}_bench_qwen.err; echo "bench qwen rc=$?"; python -c "
import json;d=json.load(open('${O}_bench_qwen.json'));print(d['value'],d['e2e']['value'],d['by_position_tok_s'],d['roofline']['frac'])"
timeout 300 python bench.py --workload llama2-7b --steps 256 --reps 3 --no-cpu-baseline > ${O}_bench_l7b.json 2> ${O}_bench_l7b.err; echo "bench l7b rc=$?"; python -c "
import json;d=json.load(open('${O}_bench_l7b.json'));print(d['value'],d['e2e']['value'],d['by_position_tok_s'],d['roofline']['frac'])"
