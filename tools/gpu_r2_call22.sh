#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gpt.py -x -q -m gpu 2>&1 | tail -2
for i in 1 2 3; do
BENCH_DEBUG=1 timeout 300 python bench.py --no-extra --no-cpu-baseline > gpurun_out/r2c22_bench_$i.json 2> gpurun_out/r2c22_bench_$i.err
grep "per-step\|slow step\|normal step" gpurun_out/r2c22_bench_$i.err | cut -c1-900
done
