#!/bin/bash
for d in 2 3 0; do echo "== DBG=$d"; GSV_HX_KVDIRECT=$d GSV_HX_CS=8 GSV_DECODE_IMPL=hx timeout 120 python tools/hx_debug.py 2>&1 | tail -2; done
