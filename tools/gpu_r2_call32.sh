#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
GSV_VOC_WS=2 GSV_VOC_FUSE=2 timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/voc_ncu.py 2 40 v2Pro > gpurun_out/r2c32_memcheck_voc.log 2>&1; echo "memcheck voc rc=$?"; grep -c "Invalid\|out of bounds\|misaligned" gpurun_out/r2c32_memcheck_voc.log; tail -3 gpurun_out/r2c32_memcheck_voc.log
GSV_VOC_WS=2 GSV_VOC_FUSE=2 timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/voc_ncu.py 2 40 v2ProPlus > gpurun_out/r2c32_memcheck_vocplus.log 2>&1; echo "memcheck vocplus rc=$?"; tail -3 gpurun_out/r2c32_memcheck_vocplus.log
