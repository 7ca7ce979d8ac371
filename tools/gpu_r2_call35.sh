#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
for i in 1 2 3 4; do
BENCH_DEBUG=1 timeout 300 python bench.py --no-extra --no-cpu-baseline > gpurun_out/r2c35_bench_$i.json 2> gpurun_out/r2c35_bench_$i.err
grep "per-step\|slow step" gpurun_out/r2c35_bench_$i.err | cut -c1-700
done
