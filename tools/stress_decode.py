#!/usr/bin/env python
"""Race hunt for the small-batch decode kernel (GPU box): the arithmetic is order-deterministic, so
repeated teacher-forced runs must be bit-identical; any difference is a synchronisation bug."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gsv-tts-lite_b200"))
import torch
from gsv_tts import _native as N, _synthetic as syn
from tests import gpu_harness as H
which = sys.argv[1] if len(sys.argv) > 1 else "tiny"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cfg = syn.GPT_CONFIG_TINY if which == "tiny" else syn.GPT_CONFIG
dev = torch.device("cuda:0")
dt = torch.bfloat16
m = H.build_gpt(cfg, syn.gpt_state_dict(cfg, 0), dt, dev, [(1, 256)])
g = torch.Generator().manual_seed(1)
x = torch.randint(0, 732, (1, 40), generator=g); y = torch.randint(0, 1024, (1, 30), generator=g)
bert = torch.randn(1, 40, 1024, generator=g)
n = 24
forced = torch.randint(0, 1024, (n,), generator=g).to(torch.int32).to(dev)
V = cfg["model"]["vocab_size"]
lib = N.lib()
ref, bad = None, 0
import hashlib, collections
hashes = collections.Counter()
for r in range(reps):
    trace = torch.zeros(n + 1, V, dtype=torch.float32, device=dev)
    N.check(lib.gsv_gpt_set_forced(m._ctx, forced.data_ptr(), n))
    N.check(lib.gsv_gpt_set_logits_trace(m._ctx, trace.data_ptr(), n + 1))
    m._single_setup(x, y, bert, 15, 1.0, 1.0, 1.35, 10, None)
    m._decode(n); torch.cuda.synchronize()
    t = trace.cpu()
    hashes[hashlib.md5(t.numpy().tobytes()).hexdigest()[:8]] += 1
    if ref is None: ref = t
    elif not torch.equal(ref, t):
        bad += 1
        rows = (ref != t).any(1).nonzero().flatten().tolist()
        print(f"rep {r}: MISMATCH first row {rows[0]} ({len(rows)} rows), max diff {(ref - t).abs().max():.3f}")
print(f"{which}: {bad} mismatching runs out of {reps}; distinct outputs: {dict(hashes)}")
