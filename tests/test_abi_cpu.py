"""CPU: the C-ABI library loads and exports every symbol include/gsv_b200.h declares (no
compute calls), and the host-side structs match the header."""
import ctypes as C
import os
import re

import pytest

from gsv_tts import _native as N

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "gsv_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gsv_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    declared = _declared()
    bound = sorted(n for n, _, _ in N.SYMBOLS)
    assert declared == bound, (set(declared) ^ set(bound))


def test_library_exports_every_symbol():
    if not os.path.exists(N.LIB_PATH):
        pytest.skip("libgsv_b200.so not built (run __graft_entry__.build())")
    l = N.lib()
    for name, _, _ in N.SYMBOLS:
        assert hasattr(l, name)
    assert l.gsv_version() == 100


def test_struct_layouts():
    # gsv_gpt_sampling: 10 x 4-byte fields then a uint64 seed at offset 40
    assert C.sizeof(N.GptSampling) == 48 and N.GptSampling.seed.offset == 40
    assert C.sizeof(N.GptDims) == 48
    assert C.sizeof(N.GptWeights) == 19 * 8
    assert C.sizeof(N.VocDims) == 4 * (8 + 8 + 8 + 1 + 4 + 12 + 1)


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(N, "_lib", None)
    monkeypatch.setattr(N, "LIB_PATH", "/nonexistent/libgsv_b200.so")
    with pytest.raises(N.NativeError):
        N.lib()


def test_no_cpu_fallback():
    import torch
    from gsv_tts import _synthetic as syn
    from gsv_tts.GPT_SoVITS.GPT.t2s_model_b200 import Text2SemanticDecoder
    from gsv_tts.GPT_SoVITS.SoVITS.models_b200 import FlowDecoder
    m = Text2SemanticDecoder(syn.GPT_CONFIG_TINY)
    with pytest.raises(N.NativeError):
        m.initialize_runtime(torch.float32, "cpu", [(1, 64)])
    with pytest.raises(N.NativeError):
        FlowDecoder(**syn.SOVITS_MODEL["tiny"]).initialize_runtime(torch.float16, "cpu", [])


def test_state_dict_keys_match_reference_layout():
    """The B200 decoder class must accept exactly the key set Loader hands to load_state_dict."""
    from gsv_tts import _synthetic as syn
    from gsv_tts.GPT_SoVITS.GPT.t2s_model_b200 import Text2SemanticDecoder
    m = Text2SemanticDecoder(syn.GPT_CONFIG_TINY)
    sd = syn.gpt_state_dict(syn.GPT_CONFIG_TINY)
    assert sorted(m.state_dict().keys()) == sorted(sd.keys())
    m.load_state_dict(sd)      # strict
