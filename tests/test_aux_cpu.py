"""CPU: ``SynthesizerTrn.get_ge`` / ``extract_latent`` (once per cached speaker / prompt; plain tensor expressions over the
checkpoint's tensors) against the reference's own modules with the same weights, where the reference sources are present
(/root/reference here, baseline/_ref on the GPU box), and against goldens made from them otherwise."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gsv-tts-lite_b200"))
from gsv_tts import _synthetic as syn                                                      # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sovits_aux.npz")


def _ours(version):
    from gsv_tts.GPT_SoVITS.SoVITS.models_b200 import SynthesizerTrn
    model = dict(syn.SOVITS_MODEL["tiny"], version=version)
    sd = dict(syn.sovits_flow_dec_state_dict(model, 0))
    sd.update(syn.sovits_encp_state_dict(model, 0))
    sd.update(syn.sovits_aux_state_dict(model, 0))
    net = SynthesizerTrn(1025, 32, n_speakers=300, **model)
    net.load_state_dict(sd)
    return net, sd, model


def _inputs():
    g = torch.Generator().manual_seed(17)
    return torch.randn(1, 1025, 37, generator=g), torch.randn(1, 20480, generator=g) * 0.1, torch.randn(1, 768, 46, generator=g)


@pytest.mark.parametrize("version", ["v2", "v2Pro"])
def test_get_ge_and_extract_latent_match_golden(version):
    net, _sd, model = _ours(version)
    refer, sv, ssl = _inputs()
    g = np.load(GOLDEN)
    ge = net.get_ge(refer, sv if version == "v2Pro" else None)
    codes = net.extract_latent(ssl)
    assert tuple(ge.shape) == (1, model["gin_channels"], 1) and tuple(codes.shape) == (1, 1, 23)
    assert float(np.abs(ge.numpy() - g[f"ge_{version}"]).max()) < 1e-5
    assert np.array_equal(codes.numpy(), g[f"codes_{version}"])


def test_get_ge_and_extract_latent_match_the_reference_module():
    from oracle import ref_shim
    if not ref_shim.available():
        pytest.skip("reference sources not available")
    M = ref_shim.sovits_models()
    refer, sv, ssl = _inputs()
    for version in ("v2", "v2Pro"):
        net, sd, model = _ours(version)
        with torch.inference_mode():
            ref = M.SynthesizerTrn(1025, 32, n_speakers=300, **model).eval()
            full = dict(sd)
            full["quantizer.vq.layers.0._codebook.inited"] = torch.Tensor([True])   # a trained checkpoint: no k-means re-init
            ref.load_state_dict(full, strict=False)
            want_ge = ref.get_ge(refer, sv if version == "v2Pro" else None)
            want_codes = ref.extract_latent(ssl)
        ge = net.get_ge(refer, sv if version == "v2Pro" else None)
        codes = net.extract_latent(ssl)
        assert ge.shape == want_ge.shape and codes.shape == want_codes.shape
        assert float((ge - want_ge).abs().max()) < 1e-5, version
        assert torch.equal(codes, want_codes), version
