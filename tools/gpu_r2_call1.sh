#!/bin/bash
# round 2, GPU call 1: parity (multi-sequence + audited continuous batching), then the reference's CUDA path on this box
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2c1_gpu.txt
timeout 900 python -m pytest tests -m gpu -q -rA --tb=short -p no:cacheprovider > gpurun_out/r2c1_tests.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/r2c1_tests.log
grep -E "^(FAILED|ERROR)" gpurun_out/r2c1_tests.log | head -20
timeout 900 python tools/ref_gpu_bench.py --sdpa > gpurun_out/r2c1_ref_gpu.json 2> gpurun_out/r2c1_ref_gpu.err
echo "ref rc=$?"; tail -c 2500 gpurun_out/r2c1_ref_gpu.json; tail -5 gpurun_out/r2c1_ref_gpu.err
