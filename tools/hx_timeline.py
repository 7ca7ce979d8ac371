#!/usr/bin/env python
"""In-kernel timeline of the head-cluster decode kernel (gpt_decode_hx.cu).  Build the tuning library first:
    make -C gsv-tts-lite_b200/csrc BUILD=build_tl OUT=../gsv_tts/libgsv_b200_tl.so EXTRA=-DGSV_TIMELINE
and run with GSV_B200_LIB=libgsv_b200_tl.so."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gsv-tts-lite_b200"))
import numpy as np, torch
from gsv_tts import _native as N, _synthetic as syn
if os.environ.get('GSV_B200_LIB'):
    N.LIB_PATH = os.path.join(os.path.dirname(N.LIB_PATH), os.environ['GSV_B200_LIB'])
from tests import gpu_harness as H
dev = torch.device("cuda:0")
cfg = syn.GPT_CONFIG
m = H.build_gpt(cfg, syn.gpt_state_dict(cfg, 0), torch.bfloat16, dev, [(1, 512)])
g = torch.Generator().manual_seed(1)
x = torch.randint(0, 732, (1, 64), generator=g); y = torch.randint(0, 1024, (1, 100), generator=g)
m.debug_seed = 1
m._single_setup(x, y, torch.zeros(1, 64, 1024), 15, 1.0, 1.0, 1.35, 10, 400)
m._decode(25); torch.cuda.synchronize()
G, MAXR = 64, 2048
rec = torch.zeros((G + 1) * 2 * MAXR, dtype=torch.int64, device=dev)      # + the sampler's own region
N.check(N.lib().gsv_gpt_set_timeline(m._ctx, rec.data_ptr(), MAXR, 0))
m._decode(3); torch.cuda.synchronize()
rs = rec.cpu().numpy().reshape(G + 1, MAXR, 2)[G]
r = rec.cpu().numpy().reshape(G + 1, MAXR, 2)[:G]
names = {40: "  (attn: warp partial in smem)", 41: "  (attn: barrier passed)", 42: "  (attn: CTA partial merged)", 43: "  (all-read: poll done)",
         44: "  (all-read: barrier passed)", 45: "  (y1 inbox full)", 30: " S0 landed", 31: " S2 landed", 1: "layer start", 2: " qkv rows done, pushed", 3: " q/k/v gathered", 4: " attention done, pushed", 5: " att merged",
         6: " O partial published", 7: " all-read 1 done, pushed", 8: " y1 gathered, LN1 done", 9: " MLP-up done", 10: " MLP-down done, pushed",
         11: " reduce-scatter done, published", 12: " all-read 2 done, pushed", 13: "HEAD start (LN2 done)", 20: "head rows published", 21: "token done"}
n = int(r[0, 0, 0])
ids = r[0, 1:n + 1, 0]
T = r[:, 1:n + 1, 1].astype(np.int64)
sd = np.where(ids == 21)[0]
lo, hi = sd[0] + 1, sd[1] + 1
print(f"one token: {T[:, hi - 1].max() - T[:, lo - 1].max()} ns   (records per CTA: {n})")
starts = np.where(ids[lo:hi] == 1)[0] + lo
for li in (5, 12):
    base, nxt = starts[li], starts[li + 1]
    t0 = T[:, base].min()
    print(f"layer {li}: (ns after the first CTA entered the layer: min / median / max over the 64 CTAs)")
    for k in range(base, nxt + 1):
        v = T[:, k] - t0
        print(f"  {names.get(int(ids[k]), ids[k]):34s} {v.min():7d} {int(np.median(v)):7d} {v.max():7d}")
k0 = np.where(ids[lo:hi] == 13)[0][0] + lo
t0 = T[:, k0].min()
print("head + sampling:")
for k in range(k0, hi):
    v = T[:, k] - t0
    print(f"  {names.get(int(ids[k]), ids[k]):34s} {v.min():7d} {int(np.median(v)):7d} {v.max():7d}")

# the sampler CTA's own markers (gpt_sample.cuh mark_sampler): one token's worth, ns after its head rows were published
snames = {58: "sampler CTA: head rows published", 59: "logits polled out of L2", 60: "sample_slot entered", 61: "columns in registers, group maxima published",
          62: "candidates compacted", 63: "pivot known", 64: "token known (warp 0: exp, sum, noise, arg-max)", 65: "bookkeeping stored", 66: "next input published"}
ns = int(rs[0, 0])
sid, st = rs[1:ns + 1, 0], rs[1:ns + 1, 1]
b = np.where(sid == 58)[0]
if len(b) >= 2:
    print("sampler (second token of the launch):")
    for k in range(b[1], b[2] if len(b) > 2 else ns):
        print(f"  {snames.get(int(sid[k]), sid[k]):62s} {int(st[k] - st[b[1]]):7d}")
