#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c11_bench.json 2> gpurun_out/r2c11_bench.err; echo "bench rc=$?"; tail -c 6000 gpurun_out/r2c11_bench.json; tail -5 gpurun_out/r2c11_bench.err
timeout 600 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r2c11_tests.log 2>&1; echo "all pytest rc=$?"; tail -5 gpurun_out/r2c11_tests.log
