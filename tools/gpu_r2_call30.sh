#!/bin/bash
# vocoder tests + speed + one full ncu capture of two persistent-kernel launches (stage 3: conv1-type and conv2-type)
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_vocoder.py -x -q -m gpu 2>&1 | tail -2
for key in v2Pro v2ProPlus; do timeout 300 python tools/voc_speed.py $key 2>&1 | tail -8 | tail -2; done
timeout 900 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:conv_umma_ws_kernel --launch-skip 62 --launch-count 2 -o gpurun_out/r02_conv_ws -f python tools/voc_ncu.py 16 500 v2Pro > gpurun_out/r02_conv_ws_ncu.log 2>&1
tail -2 gpurun_out/r02_conv_ws_ncu.log
