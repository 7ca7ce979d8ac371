"""Checkpoint ingestion for the B200 backend.

Accepts what the reference's loader accepts (reference gsv_tts/Loader.py:42-170) and hands the tensors to
the native mirrors instead of the PyTorch modules:

* GPT: an upstream GPT-SoVITS ``.ckpt`` (``{"config", "weight"}`` with ``model.h.layers.{i}.self_attn.*`` /
  ``linear1`` / ``linear2`` names), or a directory with ``config.json`` + ``model.safetensors`` whose keys are
  already in the Lite layout (written by ``TTS.to_safetensors`` there).
* SoVITS: an upstream ``.pth`` (``{"config", "weight"}``; v2Pro / v2ProPlus files carry a 2-byte version tag in
  place of the zip magic), or a directory with ``hps.json`` + ``model.safetensors``.  Only ``flow.*`` and
  ``dec.*`` are consumed by the native vocoder; ``enc_p`` / quantizer stay with the reference modules
  (SURVEY.md 8 f-1).

The parsing functions are pure CPU code (tested without a GPU); ``get_gpt_weights`` / ``get_sovits_weights``
additionally build the native contexts and therefore need an sm_100 device.
"""
from __future__ import annotations

import hashlib
import io
import json
import os
import re
import sys
import types
import contextlib
from typing import Dict, Optional, Tuple

import torch

from .Config import Config

# 2-byte tags that replace b"PK" at the start of versioned .pth files, and md5 of the first 8 KiB of the
# official pretrained files (the reference keeps the same two tables, Loader.py:17-27)
VERSION_TAGS = {b"01": "v2", b"05": "v2Pro", b"06": "v2ProPlus"}
PRETRAINED_MD5 = {
    "dc3c97e17592963677a4a1681f30c653": "v2",
    "6642b37f3dbb1f76882b69937c95a5f3": "v2",
    "c7e9fce2223f3db685cdfa1e6368728a": "v2Pro",
    "66b313e39455b57ab1b0bc0b239c9d0a": "v2ProPlus",
}
SUPPORTED_VERSIONS = ("v2", "v2Pro", "v2ProPlus")

# upstream transformer-layer parameter names -> Lite names (reference Loader.py:130-154)
_LAYER_RENAMES = (
    (r"self_attn\.in_proj_(weight|bias)$", r"qkv.\1"),
    (r"self_attn\.out_proj\.(weight|bias)$", r"out_proj.\1"),
    (r"linear1\.(weight|bias)$", r"mlp.0.\1"),
    (r"linear2\.(weight|bias)$", r"mlp.2.\1"),
)
_LAYER_KEY = re.compile(r"^model\.h\.layers\.(\d+)\.(.+)$")


def remap_gpt_keys(weights: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Upstream GPT-SoVITS key names -> the key set of ``Text2SemanticDecoder``."""
    out = {}
    for key, value in weights.items():
        m = _LAYER_KEY.match(key)
        if m:
            tail = m.group(2)
            for pat, rep in _LAYER_RENAMES:
                tail, n = re.subn(pat, rep, tail)
                if n:
                    break
            out[f"t2s_transformer.blocks.{m.group(1)}.{tail}"] = value
        elif key.startswith("model."):
            out[key[len("model."):]] = value
        else:
            out[key] = value
    return out


def read_gpt_checkpoint(path: str) -> Tuple[dict, Dict[str, torch.Tensor]]:
    """-> (config, state dict in the Lite key layout)."""
    if os.path.isdir(path):
        from safetensors.torch import load_file
        with open(os.path.join(path, "config.json")) as f:
            config = json.load(f)
        return config, load_file(os.path.join(path, "model.safetensors"))
    with _upstream_pickle_modules():
        blob = torch.load(path, map_location="cpu", weights_only=False)
    return _to_plain(blob["config"]), remap_gpt_keys(blob["weight"])


class _AttrBag:
    """Stand-in for the classes upstream checkpoints pickle their ``config`` with (``utils.HParams``,
    ``utils.DictToAttrRecursive``): unpickling only needs ``__new__`` + ``__dict__`` / ``__setstate__``."""

    def __setstate__(self, state):
        self.__dict__.update(state if isinstance(state, dict) else {})

    def __setitem__(self, key, value):          # DictToAttrRecursive is a dict subclass: items arrive through SETITEMS
        self.__dict__[key] = value


@contextlib.contextmanager
def _upstream_pickle_modules():
    """Upstream ``.pth`` files pickle ``config`` as ``utils.HParams`` (top-level module ``utils``); the reference
    registers its own ``GPT_SoVITS/utils.py`` under that name before ``torch.load`` (Loader.py:13-14).  Do the same
    with a minimal shim for the duration of the load and put ``sys.modules`` back afterwards."""
    shim = types.ModuleType("utils")
    shim.HParams = type("HParams", (_AttrBag,), {})
    shim.DictToAttrRecursive = type("DictToAttrRecursive", (_AttrBag,), {})
    prev = sys.modules.get("utils")
    sys.modules["utils"] = shim
    try:
        yield
    finally:
        if prev is None:
            sys.modules.pop("utils", None)
        else:
            sys.modules["utils"] = prev


def _to_plain(obj):
    """hps objects of upstream checkpoints are attribute bags; make them nested dicts."""
    if isinstance(obj, dict):
        return {k: _to_plain(v) for k, v in obj.items()}
    if hasattr(obj, "__dict__") and not isinstance(obj, torch.Tensor):
        return {k: _to_plain(v) for k, v in vars(obj).items()}
    return obj


def sniff_sovits_version(path: str) -> Tuple[Optional[str], bytes]:
    """Version from the 2-byte tag, else from the md5 of the first 8 KiB; also returns the file bytes with the
    zip magic restored."""
    with open(path, "rb") as f:
        data = f.read()
    version = VERSION_TAGS.get(data[:2])
    if version is None:
        version = PRETRAINED_MD5.get(hashlib.md5(data[:8192]).hexdigest())
    if data[:2] != b"PK":
        data = b"PK" + data[2:]
    return version, data


def read_sovits_checkpoint(path: str) -> Tuple[dict, Dict[str, torch.Tensor], str]:
    """-> (hps as nested dict, state dict, version)."""
    if os.path.isdir(path):
        from safetensors.torch import load_file
        with open(os.path.join(path, "hps.json")) as f:
            hps = json.load(f)
        sd = load_file(os.path.join(path, "model.safetensors"))
        version = hps.get("model", {}).get("version")
    else:
        version, data = sniff_sovits_version(path)
        with _upstream_pickle_modules():
            blob = torch.load(io.BytesIO(data), map_location="cpu", weights_only=False)
        hps = _to_plain(blob["config"])
        sd = blob["weight"]
        if version is None:
            version = hps.get("model", {}).get("version")
    if version not in SUPPORTED_VERSIONS:
        raise ValueError("The SoVITS model is not a v2 / v2Pro / v2ProPlus checkpoint")
    hps.setdefault("model", {})["version"] = version
    hps["model"]["semantic_frame_rate"] = "25hz"
    return hps, sd, version


def to_safetensors(checkpoint_path: str, output_dir: Optional[str] = None) -> str:
    """Reference ``TTS.to_safetensors`` (TTS.py:1482-1523): a ``.ckpt`` (GPT) or ``.pth`` (SoVITS) checkpoint becomes a directory
    with ``model.safetensors`` + ``config.json`` / ``hps.json`` -- the third format both loaders read.  The reference saves the
    state dict of the MODULE it built, so: GPT keys are already in the Lite layout, and the SoVITS ``dec.*`` weight-norm pairs
    are folded (``dec.remove_weight_norm()``, Loader.py:95) while ``flow.*`` keeps ``weight_g`` / ``weight_v``.  No device, no
    model construction: this is file conversion only.  Returns the directory."""
    from safetensors.torch import save_file
    if output_dir is None:
        output_dir, _ = os.path.splitext(checkpoint_path)
    os.makedirs(output_dir, exist_ok=True)
    suffix = os.path.splitext(checkpoint_path)[1]
    if suffix == ".ckpt":
        config, sd = read_gpt_checkpoint(checkpoint_path)
        meta_name, meta = "config.json", config
    elif suffix == ".pth":
        from .GPT_SoVITS.SoVITS.models_b200 import fold_weight_norm
        hps, raw, _ = read_sovits_checkpoint(checkpoint_path)
        sd = {}
        for k, v in raw.items():
            if k.startswith("dec.") and k.endswith(".weight_g"):
                continue
            if k.startswith("dec.") and k.endswith(".weight_v"):
                base = k[: -len("weight_v")]
                sd[base + "weight"] = fold_weight_norm(raw[base + "weight_g"].float(), v.float()).to(v.dtype)
            else:
                sd[k] = v
        meta_name, meta = "hps.json", hps
    else:
        raise ValueError(f"to_safetensors: expected a .ckpt (GPT) or .pth (SoVITS) file, got {checkpoint_path!r}")
    save_file({k: v.detach().cpu().contiguous() for k, v in sd.items() if isinstance(v, torch.Tensor)},
              os.path.join(output_dir, "model.safetensors"))
    with open(os.path.join(output_dir, meta_name), "w") as f:
        json.dump(meta, f, indent=4, ensure_ascii=False, default=str)
    return output_dir


class Gpt:
    def __init__(self, t2s_model, config):
        self.t2s_model = t2s_model
        self.config = config


class Sovits:
    def __init__(self, vq_model, hps):
        self.vq_model = vq_model
        self.hps = hps


def get_gpt_weights(gpt_path: str, tts_config: Config) -> Gpt:
    """Reference Loader.py:124-167.  Under ``torch.distributed`` only rank 0 reads the file; the tensors reach the other
    ranks by NCCL broadcast, GPU to GPU (the one collective of the path)."""
    from . import _shard
    from .GPT_SoVITS.GPT.t2s_model_b200 import Text2SemanticDecoder
    config, sd = _shard.broadcast_checkpoint(lambda: read_gpt_checkpoint(gpt_path), 0, tts_config.device)
    with torch.device("meta"):                 # parameter holders only: the tensors read / received above are assigned, not copied
        model = Text2SemanticDecoder(config)
    model.load_state_dict(sd, assign=True)
    model.eval()
    model.initialize_runtime(tts_config.dtype, tts_config.device, tts_config.gpt_cache)
    return Gpt(model, config)


def get_sovits_weights(sovits_path: str, tts_config: Config) -> Sovits:
    """Builds the native SoVITS half (flow + HiFi-GAN, and the prior encoder where the checkpoint carries it); same
    rank-0-reads / NCCL-broadcast rule as ``get_gpt_weights``."""
    from . import _shard
    from .GPT_SoVITS.SoVITS.models_b200 import SynthesizerTrn

    def read():
        hps, sd, _ = read_sovits_checkpoint(sovits_path)
        return hps, sd

    hps, sd = _shard.broadcast_checkpoint(read, 0, tts_config.device)
    data = hps.get("data", {})
    model = SynthesizerTrn(data.get("filter_length", 2048) // 2 + 1, hps.get("train", {}).get("segment_size", 20480) // data.get("hop_length", 640),
                           n_speakers=data.get("n_speakers", 0), **hps["model"])       # reference Loader.py:66-71, 87-92
    model.load_state_dict(sd)
    model.initialize_runtime(tts_config.dtype, tts_config.device, tts_config.sovits_cache)
    return Sovits(model, hps)
