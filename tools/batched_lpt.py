#!/usr/bin/env python
"""Config 3 (GPT stage, 128 mixed requests through 32 slots): the queue in the caller's order against longest predicted
first (what TTS.infer_features_batched(queue_order="longest_first") hands to the scheduler)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gsv-tts-lite_b200"))
import torch
import bench
from gsv_tts import _synthetic as syn
from tests import gpu_harness as H
dev = torch.device("cuda:0")
cfg = syn.GPT_CONFIG
sd = syn.gpt_state_dict(cfg, 0)
sd["ar_predict_layer.weight"][cfg["model"]["EOS"]] = 0.0
m = H.build_gpt(cfg, sd, torch.bfloat16, dev, [(32, 1024)])
xs, ys, bs, mx = bench.mixed_requests(128, 1234, dev, torch.bfloat16)
m.debug_seed = 5
m.infer_batched(xs[:40], ys[:40], bs[:40], max_new=[20] * 40)
torch.cuda.synchronize()
for name, perm in (("caller's order", list(range(128))), ("longest first", sorted(range(128), key=lambda i: -mx[i])),
                   ("caller's order", list(range(128))), ("longest first", sorted(range(128), key=lambda i: -mx[i]))):
    m.debug_seed = 5
    t0 = time.perf_counter()
    outs, order = m.infer_batched([xs[i] for i in perm], [ys[i] for i in perm], [bs[i] for i in perm], max_new=[mx[i] for i in perm])
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    tok = sum(int(o.numel()) for o in outs)
    ok = sorted(order.tolist()) == list(range(128)) and all(int(o.numel()) == mx[perm[r]] for o, r in zip(outs, order.tolist()))
    print(f"{name:15s}: {tok} tokens in {dt*1e3:.1f} ms -> {tok/dt:.0f} tok/s; every request once with its length: {ok}")
