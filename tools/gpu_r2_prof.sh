#!/bin/bash
# round 2 profile pass: launch list of the bench command, ncu --set full of the single-sequence and batched decode kernels
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r02_ncu_bench.log 2>&1
tail -2 gpurun_out/r02_ncu_bench.log | cut -c1-200
python tools/launch_shares.py gpurun_out/r02_launches.csv > gpurun_out/r02_launch_shares.txt 2>&1; head -12 gpurun_out/r02_launch_shares.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gpt_decode_hx -s 2 -c 1 -o gpurun_out/r02_prof_hx -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r02_ncu_hx.log 2>&1
tail -2 gpurun_out/r02_ncu_hx.log | cut -c1-200
ls -la gpurun_out | tail -8
