#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 300 python tools/voc_speed.py v2Pro 2>&1 | tail -8 | head -2
GSV_VOC_GRAPH=0 timeout 300 python tools/voc_speed.py v2Pro 2>&1 | tail -8 | head -2
timeout 900 python bench.py --no-cpu-baseline --no-extra > gpurun_out/r2c27_bench.json 2> gpurun_out/r2c27_bench.err
GSV_VOC_GRAPH=0 timeout 900 python bench.py --no-cpu-baseline --no-extra > gpurun_out/r2c27_bench_nograph.json 2> gpurun_out/r2c27_bench_nograph.err
python - <<'P'
import json
for f in ('gpurun_out/r2c27_bench.json','gpurun_out/r2c27_bench_nograph.json'):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d['ttft_ms'], d['gpu_launches'])
P
