#!/bin/bash
# round 2, GPU call 2: the head-cluster single-sequence kernel -- parity first, then speed and timeline
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_gpt.py -m gpu -q -x --tb=short -p no:cacheprovider -k "every_decode_kernel and hx or head_cluster or deterministic" > gpurun_out/r2c3_hx_tests.log 2>&1
rc=$?; echo "hx pytest rc=$rc"; tail -15 gpurun_out/r2c3_hx_tests.log
if [ $rc -ne 0 ]; then exit 0; fi
for impl in hx ll1; do GSV_DECODE_IMPL=$impl timeout 120 python tools/decode_speed.py 1; done 2>&1 | tee gpurun_out/r2c3_speed.log
GSV_B200_LIB=libgsv_b200_tl.so timeout 120 python tools/hx_timeline.py > gpurun_out/r2c3_hx_timeline.txt 2>&1; tail -60 gpurun_out/r2c3_hx_timeline.txt
timeout 600 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r2c3_tests.log 2>&1; echo "all pytest rc=$?"; tail -5 gpurun_out/r2c3_tests.log
