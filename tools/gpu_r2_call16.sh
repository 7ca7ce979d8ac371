#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
for i in 1 2; do
BENCH_DEBUG=1 timeout 300 python bench.py --no-extra --no-cpu-baseline > gpurun_out/r2c16_bench_$i.json 2> gpurun_out/r2c16_bench_$i.err
done
BENCH_DEBUG=1 BENCH_NO_SAMPLER=1 timeout 300 python bench.py --no-extra --no-cpu-baseline > gpurun_out/r2c16_bench_ns.json 2> gpurun_out/r2c16_bench_ns.err
tail -n 5 gpurun_out/r2c16_bench_*.err
