"""ORACLE support (test infrastructure): import the reference's hot-path modules in the
BUILD container without executing ``gsv_tts/__init__.py`` (which needs ``av``, ``pysbd``
... that are absent here).

A stub parent package named ``gsv_ref`` is registered whose ``__path__`` points at the
reference tree; the reference only uses relative imports below that, so its GPT and
SoVITS modules load unmodified.  ``/root/reference`` does not exist on the GPU box:
nothing that runs there may call this; ``available()`` is the guard.
"""
from __future__ import annotations

import os
import sys
import types

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _find_root() -> str:
    """/root/reference in the build container; on the GPU box the unmodified copy of the hot-path files that
    ``baseline/install_ref.py`` placed under the git-ignored ``baseline/_ref`` (same-box GPU baseline only)."""
    cands = [os.environ.get("GSV_REFERENCE_ROOT"), "/root/reference", os.path.join(_REPO, "baseline", "_ref")]
    for c in cands:
        if c and os.path.isdir(os.path.join(c, "gsv_tts", "GPT_SoVITS")):
            return c
    return "/root/reference"


REF_ROOT = _find_root()


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "gsv_tts", "GPT_SoVITS"))


def _ensure():
    if "gsv_ref" not in sys.modules:
        if not available():
            raise RuntimeError(f"reference tree not found at {REF_ROOT}")
        pkg = types.ModuleType("gsv_ref")
        pkg.__path__ = [os.path.join(REF_ROOT, "gsv_tts")]
        sys.modules["gsv_ref"] = pkg


def gpt_module(flash: bool = False):
    _ensure()
    if flash:
        from gsv_ref.GPT_SoVITS.GPT import t2s_model_flash_attn as m
    else:
        from gsv_ref.GPT_SoVITS.GPT import t2s_model as m
    return m


def gpt_utils():
    _ensure()
    from gsv_ref.GPT_SoVITS.GPT import utils as u
    return u


def sovits_models():
    _ensure()
    from gsv_ref.GPT_SoVITS.SoVITS import models as m
    return m


def build_reference_gpt(state_dict, config, dtype, device, gpt_cache, flash=False):
    """Loader.get_gpt_weights tail (reference Loader.py:156-166) on an in-memory state dict."""
    import torch
    m = gpt_module(flash).Text2SemanticDecoder(config)
    m.load_state_dict(state_dict)
    m = m.to(device, dtype).eval()
    m.initialize_runtime(dtype, torch.device(device), gpt_cache)
    return m


def build_reference_flow_dec(state_dict, model: dict, dtype, device):
    """The ``flow`` and ``dec`` sub-modules exactly as ``SynthesizerTrn.__init__`` builds them
    (reference models.py:293-303), with ``dec.remove_weight_norm()`` as Loader.py:95 does."""
    import torch
    M = sovits_models()
    dec = M.Generator(
        model["inter_channels"], model["resblock"], model["resblock_kernel_sizes"],
        model["resblock_dilation_sizes"], model["upsample_rates"], model["upsample_initial_channel"],
        model["upsample_kernel_sizes"], gin_channels=model["gin_channels"])
    flow = M.ResidualCouplingBlock(model["inter_channels"], model["hidden_channels"], 5, 1, 4,
                                   gin_channels=model["gin_channels"])
    dec.remove_weight_norm()
    flow.load_state_dict({k[len("flow."):]: v for k, v in state_dict.items() if k.startswith("flow.")})
    dec.load_state_dict({k[len("dec."):]: v for k, v in state_dict.items() if k.startswith("dec.")})
    return flow.to(device, dtype).eval(), dec.to(device, dtype).eval()
