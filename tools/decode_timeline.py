#!/usr/bin/env python
"""In-kernel timeline of the small-batch decode kernel (tuning tool, GPU box only; build the library
with `make EXTRA=-DGSV_TIMELINE`).  Every CTA records {marker, globaltimer ns}; this prints, per
marker of one steady-state token, the spread over CTAs and the step from the previous marker."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gsv-tts-lite_b200"))
import numpy as np, torch
from gsv_tts import _native as N, _synthetic as syn
from tests import gpu_harness as H

dev = torch.device("cuda:0")
cfg = syn.GPT_CONFIG
m = H.build_gpt(cfg, syn.gpt_state_dict(cfg, 0), torch.bfloat16, dev, [(1, 512)])
g = torch.Generator().manual_seed(1)
x = torch.randint(0, 732, (1, 64), generator=g); y = torch.randint(0, 1024, (1, 100), generator=g)
m.debug_seed = 1
m._single_setup(x, y, torch.zeros(1, 64, 1024), 15, 1.0, 1.0, 1.35, 10, 400)
m._decode(25); torch.cuda.synchronize()
G, MAXR = 148, 1024
rec = torch.zeros(G * 2 * MAXR, dtype=torch.int64, device=dev)
N.check(N.lib().gsv_gpt_set_timeline(m._ctx, rec.data_ptr(), MAXR, 0))
m._decode(3); torch.cuda.synchronize()
r = rec.cpu().numpy().reshape(G, MAXR, 2)
names = {1: "A.start", 40: "A.staged", 2: "ATT.start", 3: "O.start", 41: "O.staged", 4: "MLP1.start", 42: "MLP1.staged",
         5: "MLP2.start", 43: "MLP2.staged", 6: "HEAD.start", 44: "HEAD.staged", 20: "S.start", 21: "S.done", 50: " att.qkv", 51: " att.loop", 52: " att.pub", 60: " QKV.dot", 61: " O.dot", 62: " MLP1.dot", 63: " MLP2.dot", 64: " HEAD.dot", 70: " QKV.pub", 71: " O.pub", 72: " MLP1.pub", 73: " MLP2.pub", 74: " HEAD.pub"}
cnt = r[:, 0, 0]
n = int(cnt.max())
full = np.where(cnt == n)[0]                     # CTAs that take part in every phase (others skip some markers)
print(f"{len(full)} of {G} CTAs recorded all {n} markers")
ids = r[full[0], 1:n + 1, 0]
T = r[full][:, 1:n + 1, 1].astype(np.int64)      # [len(full)][n]
s_done = np.where(ids == 21)[0]
lo, hi = s_done[0] + 1, s_done[1] + 1            # second token
print(f"one token: {T[:, hi - 1].max() - T[:, lo - 1].max()} ns")
per_layer = (np.where(ids[lo:hi] == 1)[0])
nl = len(per_layer)
mk = (per_layer[1] - per_layer[0]) if nl > 1 else 9
base = lo + 5 * mk
t0 = T[:, base].min()
print("layer 5 (ns relative to the earliest CTA entering QKV): marker  min / median / max over CTAs   [argmax CTA]")
for k in range(base, base + mk + 1):
    v = T[:, k] - t0
    print(f"  {names.get(int(ids[k]), ids[k]):12s} {v.min():7d} {int(np.median(v)):7d} {v.max():7d}   [{int(full[v.argmax()])}]")
# tail of the token: head + sampling
k0 = np.where(ids[lo:hi] == 6)[0][0] + lo
t0 = T[:, k0].min()
print("head + sampling:")
for k in range(k0, hi):
    v = T[:, k] - t0
    print(f"  {names.get(int(ids[k]), ids[k]):12s} {v.min():7d} {int(np.median(v)):7d} {v.max():7d}")
