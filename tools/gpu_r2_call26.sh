#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tts.py -x -q -m gpu -s 2>&1 | tail -15
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r2c26_bench.json 2> gpurun_out/r2c26_bench.err
tail -3 gpurun_out/r2c26_bench.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2c26_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['ttft_ms'])
print(d.get('config3')); print(d.get('config4')); print(d.get('error'))
P
