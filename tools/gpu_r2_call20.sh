#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:conv_umma_ws_kernel --launch-skip ${SKIP:-62} --launch-count 2 -o gpurun_out/r2c20_ws -f python tools/voc_ncu.py 16 500 v2Pro > gpurun_out/r2c20_ncu.log 2>&1
tail -3 gpurun_out/r2c20_ncu.log
