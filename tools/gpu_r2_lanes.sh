#!/bin/bash
# prior-encoder lanes of the batched SoVITS stage: tests, then the full bench line
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tts.py tests/test_gpu_glue.py tests/test_gpu_encp.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -5
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r2d6_bench.json 2> gpurun_out/r2d6_bench.err; tail -2 gpurun_out/r2d6_bench.err | cut -c1-300
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2d6_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['ttft_ms'], d['roofline']['frac'], d['clocks'])
print(d.get('config3')); print(d.get('config4')); print({k:(round(v['ms'],1),round(v['tflops'],1)) for k,v in d['config5'].items()}); print(d.get('error'))
for b,v in d['batches'].items(): print(b, round(v['decode_tok_s']), round(v['roofline']['frac'],3), round(v['gpt_stage_tok_s']), round(v['sovits_stage_ms'],1), round(v['e2e_tok_s']), round(v['audio_s_per_s']))
P
