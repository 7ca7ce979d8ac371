#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_glue.py -m gpu -q -rA --tb=short -p no:cacheprovider > gpurun_out/r2c14_encp.log 2>&1; echo "encp pytest rc=$?"
grep -E "passed|failed|vs oracle|vs reference|decode|Error|error|assert" gpurun_out/r2c14_encp.log | cut -c1-250 | head -40
