#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_vocoder.py -x -q -m gpu -s 2>&1 | tail -40 > gpurun_out/r2c18_voc_tests.log
tail -25 gpurun_out/r2c18_voc_tests.log
for key in v2Pro v2ProPlus; do
timeout 300 python tools/voc_speed.py $key 2>&1 | tail -8 | tee gpurun_out/r2c18_voc_speed_$key.log
done
GSV_VOC_GRAPH=0 timeout 300 python tools/voc_speed.py v2Pro 2>&1 | tail -8 | head -3 | tee gpurun_out/r2c18_voc_speed_v2Pro_nograph.log
