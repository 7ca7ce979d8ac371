#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2c24_bench_n2.json 2> gpurun_out/r2c24_bench_n2.err
tail -c 600 gpurun_out/r2c24_bench_n2.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2c24_bench_n2.json').read().strip().splitlines()[-1])
print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['ttft_ms'], d['clocks'])
print(d['config3']); print(d['config4'])
P
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2c24_ref_n2.json 2> gpurun_out/r2c24_ref_n2.err
tail -c 400 gpurun_out/r2c24_ref_n2.json
