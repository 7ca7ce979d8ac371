#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 > gpurun_out/r2c23_tests.log; cat gpurun_out/r2c23_tests.log
BENCH_DEBUG=1 timeout 900 python bench.py > gpurun_out/r2c23_bench.json 2> gpurun_out/r2c23_bench.err
grep "per-step\|slow step" gpurun_out/r2c23_bench.err | cut -c1-400
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2c23_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['ttft_ms'], d['clocks'])
print(d['config3']); print(d['config4'])
for b,v in d['batches'].items(): print(b, {k:(round(x,1) if isinstance(x,float) else x) for k,x in v.items() if k!='roofline'})
P
