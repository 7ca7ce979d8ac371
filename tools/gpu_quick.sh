#!/bin/bash
# quick GPU iteration: GPT tests, bench line, icc hit rate + duration of the decode kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gpt.py -m gpu -q -x 2>&1 | tail -5
python bench.py --steps 5 --warmup 3 --no-cpu-baseline ${BENCH_ARGS} 2>&1 | tail -1 > gpurun_out/bench_quick.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_quick.json'))
print('value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'ms/step',round(d['ms_per_step'],2),'decode tok/s',round(d['roofline']['decode_only_tok_s'],1),'launch_ms',round(d['roofline']['launch_ms'],3),'frac',round(d['roofline']['frac'],4))
print(d['extra'])
PY
ncu --metrics sm__icc_request_hit_rate.pct,gpu__time_duration.sum,smsp__inst_executed.sum,sm__cycles_elapsed.max -k regex:gpt_decode -s 2 -c 1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra 2>&1 | grep -E "gpt_decode|icc|duration|inst_executed|cycles_elapsed" | head -8
