#!/usr/bin/env python
"""Per-step logit errors of the single-sequence kernel against the oracle (debugging aid)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gsv-tts-lite_b200"))
import torch
from gsv_tts import _synthetic as syn
from tests import gpu_harness as H
dev = torch.device("cuda:0")
for name, cfg in (("tiny", syn.GPT_CONFIG_TINY), ("full", syn.GPT_CONFIG)):
    e = H.gpt_teacher_forced_error(cfg, name, torch.float16, dev)
    print(name, "per-row max|kernel - oracle|:", ["%.3g" % v for v in e["per_row_vs_oracle"]])
