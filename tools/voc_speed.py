"""Vocoder-only timing: streaming chunk (B=1, T=50) and batch shapes; prints ms per call and TFLOP/s."""
import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gsv-tts-lite_b200")]
from tests import gpu_harness as H
dev = torch.device("cuda:0")
key = sys.argv[1] if len(sys.argv) > 1 else "v2Pro"
fd, sd, model = H.build_vocoder(key, torch.bfloat16, dev)
flop_frame = (813.1e6 if key != "v2ProPlus" else 1828.4e6) + 14.2e6
side = torch.cuda.Stream(dev)      # a capturable stream: B=1, T<=64 shapes replay a CUDA graph from their second call on
torch.cuda.set_stream(side)
for B, T in [(1, 50), (1, 55), (1, 500), (8, 100), (16, 500), (64, 500)]:
    z = torch.randn(B, 192, T, device=dev, dtype=torch.bfloat16)
    mk = torch.ones(B, 1, T, device=dev, dtype=torch.bfloat16)
    ge = torch.randn(B, model["gin_channels"], 1, device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        a = fd.flow_dec(z, mk, ge)
    torch.cuda.synchronize()
    n = 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = fd.launch_count()
    e0.record()
    for _ in range(n):
        a = fd.flow_dec(z, mk, ge)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"{key} B={B} T={T}: {ms:.3f} ms/call, {B*T*flop_frame/ms/1e9:.1f} TFLOP/s, {B*T*0.02/(ms/1e3):.0f} audio-s/s, "
          f"{(fd.launch_count()-l0)//n} launches, finite={bool(torch.isfinite(a.float()).all())}")
