"""CPU: the oracle of the stage between the two hot paths (quantizer lookup, ge_to512,
TextEncoder + MRTE, streaming cross-fade, prior sample) against outputs of the REFERENCE's own modules
(tests/golden/encp_*.npz, written by oracle/make_golden.py encp) and, when /root/reference exists, against the reference
run live."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gsv-tts-lite_b200"))

from gsv_tts import _synthetic as syn          # noqa: E402
from oracle.encp_oracle import EncPOracle      # noqa: E402
from oracle import ref_shim                    # noqa: E402

TOL = 2e-5          # fp32 against fp32, different summation order


def golden(name):
    return np.load(os.path.join(ROOT, "tests", "golden", name))


@pytest.mark.parametrize("name,key", [("v2pro", "v2Pro"), ("v2", "v2")])
def test_encp_oracle_matches_reference_goldens(name, key):
    g = golden(f"encp_{name}.npz")
    model = dict(syn.SOVITS_MODEL[key])
    orc = EncPOracle(syn.sovits_encp_state_dict(model, 0), model)
    codes, text, ge = torch.from_numpy(g["codes"]), torch.from_numpy(g["text"]), torch.from_numpy(g["ge"])
    noise = torch.from_numpy(g["noise"])
    z_p, y_mask, m_p, logs_p, ge_out = orc.decode_front(codes, text, ge, noise)
    assert z_p.shape == (1, 192, 2 * codes.shape[-1]) and bool((y_mask == 1).all())
    assert np.abs(m_p.numpy() - g["m_p"]).max() < TOL
    assert np.abs(logs_p.numpy() - g["logs_p"]).max() < TOL
    assert np.abs(z_p.numpy() - g["z_p"]).max() < 5 * TOL
    # speed != 1: linear interpolation of the encoder output to int(T / speed) + 1 frames (models.py:217-219)
    _, _, m_s, logs_s, _ = orc.decode_front(codes, text, ge, None, speed=float(g["speed"]))
    assert m_s.shape == g["m_p_speed"].shape == (1, 192, int(2 * codes.shape[-1] / float(g["speed"])) + 1)
    assert np.abs(m_s.numpy() - g["m_p_speed"]).max() < TOL and np.abs(logs_s.numpy() - g["logs_p_speed"]).max() < TOL
    # MRTE restricted to a text window (mrte_model.py:26-32)
    _, _, m_w, _, _ = orc.decode_front(codes, text, ge, None, slice_indices=torch.from_numpy(g["slice_indices"]))
    assert np.abs(m_w.numpy() - g["m_p_slice"]).max() < TOL
    assert np.abs(m_w.numpy() - g["m_p"]).max() > 1e-3          # the window does change the result
    # streaming: the prefix is re-encoded every chunk, the first overlap_len frames are cross-faded with the kept tail
    orc.y_overlap = None
    for i, (n_codes, vs) in enumerate(g["stream_chunks"].tolist()):
        _, _, m_c, _, _ = orc.decode_front(codes[:, :, :n_codes], text, ge, None, stream_mode=True, valid_start_idx=vs,
                                           overlap_len=5)
        assert np.abs(m_c.numpy() - g[f"m_p_stream{i}"]).max() < TOL
    assert np.abs(g["m_p_stream1"][:, :, :5] - g["m_p"][:, :, 5:10]).max() > 1e-3      # cross-faded frames differ from a plain call


@pytest.mark.skipif(not ref_shim.available(), reason="/root/reference not present (GPU box)")
def test_encp_oracle_matches_reference_live_other_sizes():
    """Other lengths than the goldens' (one code, a text longer than the content, ge per frame) against the reference."""
    import torch.nn.functional as F
    M = ref_shim.sovits_models()
    model = dict(syn.SOVITS_MODEL["v2Pro"])
    sd = syn.sovits_encp_state_dict(model, 1)
    with torch.inference_mode():
        net = M.SynthesizerTrn(1025, 32, n_speakers=300, **model).eval()
        net.load_state_dict(sd, strict=False)
        orc = EncPOracle(sd, model)
        g = torch.Generator().manual_seed(9)
        for n, nt, ge_t in ((1, 3, False), (4, 30, False), (21, 6, True)):
            codes = torch.randint(0, 1024, (1, 1, n), generator=g)
            text = torch.randint(0, 732, (1, nt), generator=g)
            ge = torch.randn(1, model["gin_channels"], n if ge_t else 1, generator=g)
            q = F.interpolate(net.quantizer.decode(codes), size=2 * n, mode="nearest")
            ge2 = F.interpolate(ge, size=2 * n, mode="nearest") if ge_t else ge
            m, logs, _ = net.enc_p.infer(q, text, net.ge_to512(ge2.transpose(2, 1)).transpose(2, 1), 1)
            _, _, mo, lo, _ = orc.decode_front(codes, text, ge, None)
            assert float((m - mo).abs().max()) < TOL and float((logs - lo).abs().max()) < TOL
