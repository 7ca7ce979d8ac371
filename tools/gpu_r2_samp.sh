#!/bin/bash
# sampler fast path: the GPT parity suite (every decode kernel replays the reference sampler), decode-only speed, headline
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gpt.py tests/test_gpu_tts.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -8
for b in 1 8 32; do timeout 120 python tools/decode_speed.py $b 2>&1 | tail -1; done
timeout 600 python bench.py --no-extra --no-cpu-baseline > gpurun_out/r2d3_bench.json 2> gpurun_out/r2d3_bench.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2d3_bench.json').read().strip().splitlines()[-1])
print('ms/step', round(d['ms_per_step'],2), 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ttft', round(d['ttft_ms'],2), 'us/token', round(d['roofline']['us_per_token'],1), 'frac', round(d['roofline']['frac'],4))
P
