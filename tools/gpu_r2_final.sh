#!/bin/bash
# round-2 final pass: all GPU tests, smoke, the bench line (the expensive ncu launch list of the whole bench command is
# tools/gpu_r2_prof.sh)
cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > gpurun_out/r02_final_tests.log; cat gpurun_out/r02_final_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
timeout 900 python bench.py > gpurun_out/r02_final_bench.json 2> gpurun_out/r02_final_bench.err; tail -2 gpurun_out/r02_final_bench.err | cut -c1-300
python - <<'P'
import json
d=json.loads(open('gpurun_out/r02_final_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['ttft_ms'], d['roofline']['frac'], d['clocks'], d['cpu_baseline']['value'])
print(d.get('config3')); print(d.get('config4')); print({k:(round(v['ms'],1),round(v['tflops'],1)) for k,v in d['config5'].items()})
for b,v in d['batches'].items(): print(b, round(v['decode_tok_s']), round(v['roofline']['frac'],3), round(v['gpt_stage_tok_s']), round(v['e2e_tok_s']), round(v['audio_s_per_s']))
P
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_final_ref.json 2>/dev/null; tail -c 500 gpurun_out/r02_final_ref.json
