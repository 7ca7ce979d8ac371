#!/bin/bash
# decode-ahead streaming: its tests, the per-chunk SoVITS latency breakdown, the headline with the debug trace
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_glue.py tests/test_gpu_tts.py tests/test_gpu_encp.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -5
timeout 300 python tools/sovits_chunk_latency.py > gpurun_out/sovits_chunk_latency.txt 2> gpurun_out/sovits_chunk_latency.err; tail -3 gpurun_out/sovits_chunk_latency.err; cat gpurun_out/sovits_chunk_latency.txt
BENCH_DEBUG=1 timeout 600 python bench.py --no-extra --no-cpu-baseline > gpurun_out/r2d1_bench.json 2> gpurun_out/r2d1_bench.err
tail -c 1500 gpurun_out/r2d1_bench.json; grep -a "host ms between clips" gpurun_out/r2d1_bench.err | head -1 | cut -c1-300
