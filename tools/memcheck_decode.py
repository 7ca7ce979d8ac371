import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gsv-tts-lite_b200"))
import torch
from gsv_tts import _native as N, _synthetic as syn
from tests import gpu_harness as H
cfg = syn.GPT_CONFIG_TINY if len(sys.argv) < 2 else syn.GPT_CONFIG
dev = torch.device("cuda:0")
m = H.build_gpt(cfg, syn.gpt_state_dict(cfg, 0), torch.float16, dev, [(1, 128)])
g = torch.Generator().manual_seed(1)
m.debug_seed = 1
m._single_setup(torch.randint(0, 732, (1, 20), generator=g), torch.randint(0, 1024, (1, 20), generator=g), torch.zeros(1, 20, 1024), 15, 1.0, 1.0, 1.35, 10, 10)
m._decode(2); torch.cuda.synchronize(); m._read(1); print("ok", m._h_tokens[0, :4].tolist())
