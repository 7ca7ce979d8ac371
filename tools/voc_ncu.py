"""One flow_dec call at a given shape (for ncu launch lists): python tools/voc_ncu.py B T"""
import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gsv-tts-lite_b200")]
from tests import gpu_harness as H
dev = torch.device("cuda:0")
B, T = int(sys.argv[1]), int(sys.argv[2])
key = sys.argv[3] if len(sys.argv) > 3 else "v2Pro"
fd, sd, model = H.build_vocoder(key, torch.bfloat16, dev)
z = torch.randn(B, 192, T, device=dev, dtype=torch.bfloat16)
mk = torch.ones(B, 1, T, device=dev, dtype=torch.bfloat16)
ge = torch.randn(B, model["gin_channels"], 1, device=dev, dtype=torch.bfloat16)
fd.flow_dec(z, mk, ge)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
fd.flow_dec(z, mk, ge)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
