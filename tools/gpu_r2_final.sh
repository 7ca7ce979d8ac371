#!/bin/bash
# round-2 final pass: all GPU tests, smoke, the bench line, launch list of the bench command, ncu --set full of the head-cluster decode kernel
cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > gpurun_out/r02_final_tests.log; cat gpurun_out/r02_final_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/r02_final_bench.json 2> gpurun_out/r02_final_bench.err; tail -2 gpurun_out/r02_final_bench.err | cut -c1-300
python - <<'P'
import json
d=json.loads(open('gpurun_out/r02_final_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['ttft_ms'], d['roofline']['frac'], d['clocks'])
print(d.get('config3')); print(d.get('config4')); print({k:(round(v['ms'],1),round(v['tflops'],1)) for k,v in d['config5'].items()})
for b,v in d['batches'].items(): print(b, round(v['decode_tok_s']), round(v['roofline']['frac'],3), round(v['e2e_tok_s']))
P
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r02_ncu_bench.log 2>&1
python tools/launch_shares.py gpurun_out/r02_launches.csv > gpurun_out/r02_launch_shares.txt 2>&1; head -14 gpurun_out/r02_launch_shares.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gpt_decode_hx -s 2 -c 1 -o gpurun_out/r02_prof_hx -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r02_ncu_hx.log 2>&1
tail -2 gpurun_out/r02_ncu_hx.log | cut -c1-200
