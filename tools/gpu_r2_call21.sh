#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 > gpurun_out/r2c21_tests.log; cat gpurun_out/r2c21_tests.log
BENCH_DEBUG=1 timeout 900 python bench.py > gpurun_out/r2c21_bench.json 2> gpurun_out/r2c21_bench.err
tail -c 1500 gpurun_out/r2c21_bench.json; tail -4 gpurun_out/r2c21_bench.err | cut -c1-300
