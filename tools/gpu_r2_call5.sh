#!/bin/bash
mkdir -p gpurun_out
for cs in 8 16; do
  GSV_HX_CS=$cs timeout 300 python -m pytest tests/test_gpu_gpt.py -m gpu -q -x --tb=short -p no:cacheprovider -k "every_decode_kernel and hx or head_cluster or deterministic" > gpurun_out/r2c10_hx_tests_cs$cs.log 2>&1
  echo "cs=$cs hx pytest rc=$?"; tail -3 gpurun_out/r2c10_hx_tests_cs$cs.log
  GSV_HX_CS=$cs GSV_DECODE_IMPL=hx timeout 120 python tools/decode_speed.py 1
done 2>&1 | tee gpurun_out/r2c10_speed.log
GSV_HX_CS=8 GSV_B200_LIB=libgsv_b200_tl.so timeout 120 python tools/hx_timeline.py > gpurun_out/r2c10_hx_timeline_cs8.txt 2>&1; head -28 gpurun_out/r2c10_hx_timeline_cs8.txt; tail -5 gpurun_out/r2c10_hx_timeline_cs8.txt
