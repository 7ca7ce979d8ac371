#!/usr/bin/env python
"""In-kernel timeline of the 8-sequence cluster kernel (needs a -DGSV_TIMELINE build: either `make EXTRA=-DGSV_TIMELINE`,
or objects built into csrc/build_tl and linked as gsv_tts/libgsv_b200_tl.so; GSV_DECODE_IMPL=cl8)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gsv-tts-lite_b200"))
import numpy as np, torch
from gsv_tts import _native as N, _synthetic as syn
if os.path.exists(os.path.join(os.path.dirname(N.LIB_PATH), 'libgsv_b200_tl.so')):
    N.LIB_PATH = os.path.join(os.path.dirname(N.LIB_PATH), 'libgsv_b200_tl.so')     # local -DGSV_TIMELINE build (see docstring)
from tests import gpu_harness as H
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda:0")
cfg = syn.GPT_CONFIG
m = H.build_gpt(cfg, syn.gpt_state_dict(cfg, 0), torch.bfloat16, dev, [(B, 512)])
g = torch.Generator().manual_seed(1)
m._release_all()
for s in range(B):
    samp = N.GptSampling(top_k=15, top_p=1.0, temperature=1.0, repetition_penalty=1.35, suppress_steps=10, suppress_first=1, max_new_tokens=0, mask_eos=1, max_kv=512, seed=s + 1)
    m._prefill(s, torch.randint(0, 732, (64,), generator=g), torch.randint(0, 1024, (100,), generator=g), torch.zeros(64, 1024), samp)
m._decode(25); torch.cuda.synchronize()
G, MAXR = 148, 1024
rec = torch.zeros(G * 2 * MAXR, dtype=torch.int64, device=dev)
N.check(N.lib().gsv_gpt_set_timeline(m._ctx, rec.data_ptr(), MAXR, 0))
m._decode(3); torch.cuda.synchronize()
r = rec.cpu().numpy().reshape(G, MAXR, 2)[:16]
names = {1: "A.start", 49: " ln staged", 41: " ln1 staged", 31: " O mma+epi done", 42: " M1 mma+epi done", 51: " M2 mma+epi done", 50: " qkv done", 52: "A.end (att pushed)", 3: "O.start (att arrived)", 53: "O.end", 4: "M1.start (y1 arrived)", 54: "M1.end",
         5: "M2.start (h arrived)", 55: "M2.end", 6: "HEAD.start", 20: "S.start", 21: "S.done"}
n = int(r[0, 0, 0])
ids = r[0, 1:n + 1, 0]
T = r[:, 1:n + 1, 1].astype(np.int64)
sd = np.where(ids == 21)[0]
lo, hi = sd[0] + 1, sd[1] + 1
print(f"one token ({B} sequences): {T[:, hi - 1].max() - T[:, lo - 1].max()} ns")
starts = np.where(ids[lo:hi] == 1)[0] + lo
base, nxt = starts[5], starts[6]
t0 = T[:, base].min()
for k in range(base, nxt + 1):
    v = T[:, k] - t0
    print(f"  {names.get(int(ids[k]), ids[k]):24s} {v.min():7d} {int(np.median(v)):7d} {v.max():7d}")
k0 = np.where(ids[lo:hi] == 6)[0][0] + lo
t0 = T[:, k0].min()
for k in range(k0, hi):
    v = T[:, k] - t0
    print(f"  {names.get(int(ids[k]), ids[k]):24s} {v.min():7d} {int(np.median(v)):7d} {v.max():7d}")
