#!/bin/bash
# round 2, GPU call 4: head-cluster kernel v3 -- cluster sizes 4 / 8 / 16: parity, speed, timeline
mkdir -p gpurun_out
for cs in 16 8 4; do
  GSV_HX_CS=$cs timeout 300 python -m pytest tests/test_gpu_gpt.py -m gpu -q -x --tb=short -p no:cacheprovider -k "every_decode_kernel and hx or head_cluster or deterministic" > gpurun_out/r2c4_hx_tests_cs$cs.log 2>&1
  echo "cs=$cs hx pytest rc=$?"; tail -3 gpurun_out/r2c4_hx_tests_cs$cs.log
  GSV_HX_CS=$cs GSV_DECODE_IMPL=hx timeout 120 python tools/decode_speed.py 1
done 2>&1 | tee gpurun_out/r2c4_speed.log
GSV_DECODE_IMPL=ll1 timeout 120 python tools/decode_speed.py 1 | tee -a gpurun_out/r2c4_speed.log
for cs in 16 8; do
GSV_HX_CS=$cs GSV_B200_LIB=libgsv_b200_tl.so timeout 120 python tools/hx_timeline.py > gpurun_out/r2c4_hx_timeline_cs$cs.txt 2>&1; tail -36 gpurun_out/r2c4_hx_timeline_cs$cs.txt | head -20
done
timeout 600 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r2c4_tests.log 2>&1; echo "all pytest rc=$?"; tail -5 gpurun_out/r2c4_tests.log
