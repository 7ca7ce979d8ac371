#!/usr/bin/env python
"""Decode-only speed of the small-batch kernel (GPU box): us/token at batch 1 (and B via argv)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gsv-tts-lite_b200"))
import torch
from gsv_tts import _native as N, _synthetic as syn
if os.environ.get('GSV_B200_LIB'):      # A/B builds of the library (tools only)
    N.LIB_PATH = os.path.join(os.path.dirname(N.LIB_PATH), os.environ['GSV_B200_LIB'])
from tests import gpu_harness as H
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
dev = torch.device("cuda:0")
cfg = syn.GPT_CONFIG
m = H.build_gpt(cfg, syn.gpt_state_dict(cfg, 0), torch.bfloat16, dev, [(max(B, 1), 512)])
g = torch.Generator().manual_seed(1)
m._release_all()
for s in range(B):
    samp = N.GptSampling(top_k=15, top_p=1.0, temperature=1.0, repetition_penalty=1.35, suppress_steps=10, suppress_first=1, max_new_tokens=0, mask_eos=1, max_kv=512, seed=s + 1)
    m._prefill(s, torch.randint(0, 732, (64,), generator=g), torch.randint(0, 1024, (100,), generator=g), torch.zeros(64, 1024), samp)
m._decode(25); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); 
for _ in range(4): m._decode(25)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"B={B}: {ms*1e3/100:.1f} us/step, {B*100/(ms/1e3):.0f} tok/s")
