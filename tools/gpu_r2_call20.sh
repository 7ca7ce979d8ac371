#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 300 python tools/voc_speed.py v2Pro 2>&1 | tail -3 | tee gpurun_out/r2c20_voc_speed_v2Pro.log
timeout 900 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:conv_umma_ws_kernel --launch-skip 40 --launch-count 2 -o gpurun_out/r2c20_ws -f python tools/voc_ncu.py 16 500 v2Pro > gpurun_out/r2c20_ncu.log 2>&1
tail -5 gpurun_out/r2c20_ncu.log
ls -la gpurun_out/r2c20_ws.ncu-rep
