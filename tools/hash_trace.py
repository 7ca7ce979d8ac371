#!/usr/bin/env python
"""Find the first GEMV phase whose staged operand differs between two identical teacher-forced runs
(library built with EXTRA=-DGSV_HASHTRACE)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gsv-tts-lite_b200"))
import numpy as np, torch
from gsv_tts import _native as N, _synthetic as syn
from tests import gpu_harness as H
cfg = syn.GPT_CONFIG
dev = torch.device("cuda:0")
m = H.build_gpt(cfg, syn.gpt_state_dict(cfg, 0), torch.bfloat16, dev, [(1, 256)])
g = torch.Generator().manual_seed(1)
x = torch.randint(0, 732, (1, 40), generator=g); y = torch.randint(0, 1024, (1, 30), generator=g)
bert = torch.randn(1, 40, 1024, generator=g)
n = 16
forced = torch.randint(0, 1024, (n,), generator=g).to(torch.int32).to(dev)
lib = N.lib()
runs = []
for r in range(6):
    rec = torch.zeros(2 * 8192, dtype=torch.int64, device=dev)
    N.check(lib.gsv_gpt_set_forced(m._ctx, forced.data_ptr(), n))
    N.check(lib.gsv_gpt_set_timeline(m._ctx, rec.data_ptr(), 8192, 0))
    m._single_setup(x, y, bert, 15, 1.0, 1.0, 1.35, 10, None)
    m._decode(n); torch.cuda.synchronize()
    a = rec.cpu().numpy().reshape(-1, 2); k = int(a[0, 0]); runs.append(a[1:k + 1].copy())
names = ["QKV", "O", "MLP1", "MLP2", "HEAD"]
per_step = 24 * 4 + 1
for r in range(1, len(runs)):
    a, b = runs[0], runs[r]
    assert len(a) == len(b) and (a[:, 0] == b[:, 0]).all()
    diff = np.where(a[:, 1] != b[:, 1])[0]
    if len(diff) == 0: print(f"run {r}: identical"); continue
    i = int(diff[0]); step, rem = divmod(i, per_step)
    print(f"run {r}: first differing staged operand: record {i} = step {step}, layer {rem // 4}, phase {names[int(a[i, 0])]}; {len(diff)} of {len(a)} differ")
