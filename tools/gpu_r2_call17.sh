#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
BENCH_DEBUG=1 timeout 300 python bench.py --no-extra --no-cpu-baseline > gpurun_out/r2c17_bench.json 2> gpurun_out/r2c17_bench.err
tail -n 6 gpurun_out/r2c17_bench.err | cut -c1-600
for key in v2Pro v2ProPlus; do
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,launch__grid_size --clock-control none --csv --log-file gpurun_out/r2c17_voc_$key.csv python tools/voc_ncu.py 16 500 $key > gpurun_out/r2c17_voc_$key.log 2>&1
python tools/voc_shares.py gpurun_out/r2c17_voc_$key.csv > gpurun_out/r2c17_voc_shares_$key.txt
head -30 gpurun_out/r2c17_voc_shares_$key.txt
done
