#!/bin/bash
# the bench line under torchrun on all GPUs of the box
cd /root/repo
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_final_bench_n$N.json 2> gpurun_out/r02_final_bench_n$N.err
tail -c 300 gpurun_out/r02_final_bench_n$N.err
python - <<P
import json
d=json.loads(open('gpurun_out/r02_final_bench_n$N.json').read().strip().splitlines()[-1])
print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['ttft_ms'], d['clocks'])
print(d.get('config3')); print(d.get('config4')); print(d.get('error'))
P
