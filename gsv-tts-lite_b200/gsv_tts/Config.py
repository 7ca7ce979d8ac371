"""Runtime configuration of the B200 backend (counterpart of reference gsv_tts/Config.py:85-108).

The reference picks device and dtype at import time (bf16 on sm >= 8.0, Config.py:3-37, 55-82).  Here
the device must be an sm_100 GPU and the storage dtype is fp16 or bf16 (fp32 accumulation in every
kernel); both are explicit constructor arguments of ``TTS`` with the reference's defaults.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Tuple

import torch


@dataclass
class Config:
    device: torch.device = field(default_factory=lambda: torch.device("cuda", 0))
    dtype: torch.dtype = torch.bfloat16
    gpt_cache: List[Tuple[int, int]] = field(default_factory=lambda: [(1, 512), (1, 768), (1, 1024), (4, 512), (4, 1024)])
    sovits_cache: List[int] = field(default_factory=lambda: [50, 55])
    use_flash_attn: bool = False       # accepted for signature compatibility; the native kernels replace both reference backends
    use_bert: bool = False
    samplerate: int = 32000            # reference TTS.py:140-142
    gpt_hz: int = 25
    sovits_hz: int = 50
