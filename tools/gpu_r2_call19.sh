#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
for key in ${KEYS:-v2Pro}; do
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,launch__grid_size,dram__bytes_read.sum,dram__bytes_write.sum,launch__block_size,launch__shared_mem_per_block_dynamic --clock-control none --csv --log-file gpurun_out/r2c19_voc_$key.csv python tools/voc_ncu.py 16 500 $key > gpurun_out/r2c19_voc_$key.log 2>&1
python tools/voc_shares.py gpurun_out/r2c19_voc_$key.csv > gpurun_out/r2c19_voc_shares_$key.txt
head -24 gpurun_out/r2c19_voc_shares_$key.txt
done
