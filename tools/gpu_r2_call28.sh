#!/bin/bash
cd /root/repo
for B in 8 16 32; do
  for NCL in 1 2 4 6 7; do
    if [ $((NCL*8)) -ge $B ]; then
      echo -n "B=$B clusters=$NCL: "; GSV_CL8_CLUSTERS=$NCL timeout 120 python tools/decode_speed.py $B 2>&1 | tail -1
    fi
  done
done
timeout 600 python -m pytest tests/test_gpu_gpt.py -x -q -m gpu 2>&1 | tail -3
