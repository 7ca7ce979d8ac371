#!/bin/bash
# One GPU round: bench line, ncu launch list of the same command, ncu --set full of the top kernel(s).
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log | cut -c1-300
ncu --set full --clock-control none --import-source on -k regex:gpt_decode_ll -s 2 -c 1 -o gpurun_out/prof_decode -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-300
# batched decode kernel (32 sequences, 25-token launch)
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gpt_decode_cl8 -s 1 -c 1 -o gpurun_out/prof_cl8 -f \
    python tools/decode_speed.py 32 > gpurun_out/ncu_cl8.log 2>&1
tail -2 gpurun_out/ncu_cl8.log | cut -c1-300
# continuous batching (GPT stage), both refill modes
timeout 120 python tools/batched_throughput.py 128 > gpurun_out/batched.log 2>&1; tail -2 gpurun_out/batched.log
ls -la gpurun_out
