#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_vocoder.py -x -q -m gpu -k "fused" 2>&1 | tail -25
timeout 900 python -m pytest tests/test_gpu_vocoder.py -x -q -m gpu 2>&1 | tail -3
for key in v2Pro v2ProPlus; do timeout 300 python tools/voc_speed.py $key 2>&1 | tail -8 | tail -3; done
GSV_VOC_FUSE=0 timeout 300 python tools/voc_speed.py v2Pro 2>&1 | tail -2
