#!/bin/bash
cd /root/repo
for B in 2 4 6; do
  echo -n "B=$B cl: "; timeout 120 python tools/decode_speed.py $B 2>&1 | tail -1
  echo -n "B=$B cl8 default clusters: "; GSV_DECODE_IMPL=cl8 timeout 120 python tools/decode_speed.py $B 2>&1 | tail -1
  echo -n "B=$B cl8 one per cluster: "; GSV_DECODE_IMPL=cl8 GSV_CL8_CLUSTERS=$B timeout 120 python tools/decode_speed.py $B 2>&1 | tail -1
done
