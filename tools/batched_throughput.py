#!/usr/bin/env python
"""SURVEY.md 8d config 3 (GPT stage): continuous batching over 32 slots, 128 requests of mixed length
(Nx ~ U{40..120}, Ny ~ U{75..250}, target length ~ U{50..250} enforced through max_new), generated tokens / wall."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gsv-tts-lite_b200"))
import torch
from gsv_tts import _native as N, _synthetic as syn
from tests import gpu_harness as H
R = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dev = torch.device("cuda:0")
cfg = syn.GPT_CONFIG
m = H.build_gpt(cfg, syn.gpt_state_dict(cfg, 0), torch.bfloat16, dev, [(32, 1024)])
g = torch.Generator().manual_seed(1234)
xs, ys, bs, mx = [], [], [], []
for r in range(R):
    nx = int(torch.randint(40, 121, (1,), generator=g)); ny = int(torch.randint(75, 251, (1,), generator=g))
    xs.append(torch.randint(0, 732, (nx,), generator=g)); ys.append(torch.randint(0, 1024, (ny,), generator=g))
    bs.append(torch.zeros(nx, 1024)); mx.append(int(torch.randint(50, 251, (1,), generator=g)))
xs = [t.to(dev) for t in xs]; ys = [t.to(dev) for t in ys]; bs = [t.to(dev, torch.bfloat16) for t in bs]
if os.environ.get('GSV_BATCH_INTERVAL'):
    m.BATCH_INTERVAL = int(os.environ['GSV_BATCH_INTERVAL'])
m.debug_seed = 5
m.infer_batched(xs[:40], ys[:40], bs[:40], max_new=[20] * 40)          # warm-up (kernel selection, weight re-tiling)
torch.cuda.synchronize()
for overlap in (True, False):
    m.overlap_refill = overlap
    m.debug_seed = 5
    l0 = int(N.lib().gsv_gpt_launch_count(m._ctx))
    t0 = time.perf_counter()
    outs, order = m.infer_batched(xs, ys, bs, max_new=mx)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    tok = sum(int(o.numel()) for o in outs)
    print(f"{R} requests, 32 slots, refill prompts {'on a second stream' if overlap else 'between decode launches'}: {tok} tokens in "
          f"{dt*1e3:.1f} ms -> {tok/dt:.0f} tok/s, {tok*0.04/dt:.0f} audio-s/s (GPT stage), "
          f"{int(N.lib().gsv_gpt_launch_count(m._ctx)) - l0} launches; mean length {tok/R:.1f}")

if os.environ.get("GSV_BATCH_TRACE"):
    # where the time of the overlapped run goes: launches, live slots per launch, device time of the decode launches
    m.overlap_refill = True
    m.debug_seed = 5
    stats = {"launch": 0, "live": [], "ev": []}
    dec, rd = m._decode, m._read_enqueue

    def decode(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(torch.cuda.current_stream(dev)); dec(n); b.record(torch.cuda.current_stream(dev))
        stats["launch"] += 1; stats["ev"].append((a, b))

    def read_enqueue(k):
        rd(k)
    m._decode, m._read_enqueue = decode, read_enqueue
    wait = m._read_wait

    def read_wait():
        wait(); stats["live"].append(int(m._h_active.sum()))
    m._read_wait = read_wait
    t0 = time.perf_counter()
    outs, order = m.infer_batched(xs, ys, bs, max_new=mx)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    ms = [a.elapsed_time(b) for a, b in stats["ev"]]
    print(f"trace: {dt*1e3:.1f} ms wall, {stats['launch']} decode launches of {m.BATCH_INTERVAL} steps, device time in decode launches {sum(ms):.1f} ms "
          f"(mean {sum(ms)/len(ms):.2f} ms, max {max(ms):.2f}), mean live slots at harvest {sum(stats['live'])/len(stats['live']):.1f}")
    import collections
    print("live histogram:", sorted(collections.Counter(stats["live"]).items()))
    print("launch ms deciles:", [round(sorted(ms)[int(len(ms) * q / 10)], 2) for q in range(10)])
