#!/bin/bash
# round 2, 2 GPUs: the whole Python tensor-parallel suite at world 2 on the final kernels (world 4 / 8 cases skip)
set -u
mkdir -p gpurun_out
O=gpurun_out/r2tp2e
timeout 230 python -m pytest tests/test_tensor_parallel.py -m gpu -q > ${O}_pytest_tp.log 2>&1; echo "pytest tp rc=$?"; tail -6 ${O}_pytest_tp.log | cut -c1-300
