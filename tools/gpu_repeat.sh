#!/bin/bash
# the whole GPU suite several times over on one box (intermittent failures: stream races, near-tie sampling) + the N=2 bench line
cd /root/repo
mkdir -p gpurun_out
for i in 1 2 3; do timeout 900 python -m pytest tests -x -q -m gpu -p no:cacheprovider 2>&1 | tail -2; done
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 > gpurun_out/r02_final_bench_n2.json 2> gpurun_out/r02_final_bench_n2.err
  python - <<'P'
import json
d=json.loads(open('gpurun_out/r02_final_bench_n2.json').read().strip().splitlines()[-1])
print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['ttft_ms'], d.get('config4'), d.get('error'))
P
fi
