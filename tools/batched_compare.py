#!/usr/bin/env python
"""Continuous batching: per-request tokens of the overlapped schedule against the reference order (same seeds)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gsv-tts-lite_b200"))
import torch
from gsv_tts import _synthetic as syn
from tests import gpu_harness as H
R = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dev = torch.device("cuda:0")
cfg = syn.GPT_CONFIG
m = H.build_gpt(cfg, syn.gpt_state_dict(cfg, 0), torch.bfloat16, dev, [(32, 1024)])
if os.environ.get('GSV_SPARE') is not None:
    m.SPARE_SLOTS = int(os.environ['GSV_SPARE'])
g = torch.Generator().manual_seed(1234)
xs, ys, bs, mx = [], [], [], []
for r in range(R):
    nx = int(torch.randint(40, 121, (1,), generator=g)); ny = int(torch.randint(75, 251, (1,), generator=g))
    xs.append(torch.randint(0, 732, (nx,), generator=g).to(dev)); ys.append(torch.randint(0, 1024, (ny,), generator=g).to(dev))
    bs.append(torch.zeros(nx, 1024, device=dev, dtype=torch.bfloat16)); mx.append(int(torch.randint(50, 251, (1,), generator=g)))
if len(sys.argv) > 2:
    m.debug_seed = 5
    m.infer_batched(xs[:40], ys[:40], bs[:40], max_new=[20] * 40)          # the warm-up of batched_throughput.py
    torch.cuda.synchronize()
runs = []
seq = [True, False, True, False] if len(sys.argv) > 2 else [False, True]
for overlap in seq:
    m.overlap_refill = overlap
    m.debug_seed = 5
    outs, order = m.infer_batched(xs, ys, bs, max_new=mx)
    torch.cuda.synchronize()
    runs.append({r: t.cpu().tolist() for t, r in zip(outs, order.tolist())})
for i in range(len(runs)):
    for j in range(i + 1, len(runs)):
        bad = [r for r in range(R) if runs[i][r] != runs[j][r]]
        print(f"run {i} ({'overlap' if seq[i] else 'reference order'}) vs run {j} ({'overlap' if seq[j] else 'reference order'}): {len(bad)} requests differ {bad[:12]}")
