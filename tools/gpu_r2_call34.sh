#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_gpt.py tests/test_gpu_tts.py -x -q -m gpu 2>&1 | tail -4
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r2c34_bench.json 2> gpurun_out/r2c34_bench.err; tail -2 gpurun_out/r2c34_bench.err | cut -c1-300
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2c34_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['ttft_ms'], d['roofline']['frac'])
print(d.get('config3')); print(d.get('config4')); print(d.get('error'))
for b,v in d['batches'].items(): print(b, round(v['decode_tok_s']), round(v['gpt_stage_tok_s']), round(v['gpt_stage_ms'],1), round(v['e2e_tok_s']))
P
