#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_gpt.py tests/test_gpu_tts.py -x -q -m gpu 2>&1 | tail -4
BENCH_DEBUG=1 timeout 900 python bench.py --no-cpu-baseline --no-extra > gpurun_out/r2c33_bench.json 2> gpurun_out/r2c33_bench.err
grep "per-step\|between clips" gpurun_out/r2c33_bench.err | cut -c1-420
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2c33_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['ttft_ms'], d['roofline']['frac'], d['roofline']['us_per_token'])
P
