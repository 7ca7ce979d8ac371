#!/bin/bash
# per-kernel time and DRAM throughput of one flow + HiFi-GAN call (B = 16, T = 500), both decoders; vocoder speed table
cd /root/repo
mkdir -p gpurun_out
for key in v2Pro v2ProPlus; do
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,launch__grid_size,dram__bytes_read.sum,dram__bytes_write.sum,launch__shared_mem_per_block_dynamic --clock-control none --csv --log-file gpurun_out/r02_voc_$key.csv python tools/voc_ncu.py 16 500 $key > gpurun_out/r02_voc_$key.log 2>&1
python tools/voc_shares.py gpurun_out/r02_voc_$key.csv > gpurun_out/r02_voc_shares_$key.txt
head -16 gpurun_out/r02_voc_shares_$key.txt
timeout 300 python tools/voc_speed.py $key 2>&1 | tail -6 | tee gpurun_out/r02_voc_speed_$key.txt
done
timeout 900 python -m pytest tests/test_gpu_vocoder.py tests/test_gpu_encp.py tests/test_gpu_glue.py -x -q -m gpu 2>&1 | tail -3
