"""GPU parity tests for the stage between the two hot paths: quantizer lookup, ge_to512,
TextEncoder + MRTE, streaming cross-fade, speed, prior sample and ``SynthesizerTrn.decode`` through the C ABI
(csrc/encp.cu), against the oracle on the same rounded weights and against outputs of the reference's own modules
(tests/golden/encp_*.npz); and the drop-in check of INTEGRATION.md run for real: the reference's ``SynthesizerTrn``
(unmodified, from baseline/_ref) on the same GPU with the same weights, its ``decode`` against ours."""
import numpy as np
import pytest
import torch

from gsv_tts import _synthetic as syn

pytestmark = pytest.mark.gpu

# m_p / logs_p are O(1) after 12 post-LN layers; 16-bit activations between every op (as the reference keeps them)
TOL = {torch.float16: 4e-2, torch.bfloat16: 2.5e-1}


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def _build(key, dtype, dev, seed=0):
    from gsv_tts.GPT_SoVITS.SoVITS.models_b200 import SynthesizerTrn
    model = dict(syn.SOVITS_MODEL[key])
    sd = dict(syn.sovits_flow_dec_state_dict(model, seed))
    sd.update(syn.sovits_encp_state_dict(model, seed))
    net = SynthesizerTrn(1025, 32, n_speakers=300, **model)
    net.load_state_dict(sd)
    net.initialize_runtime(dtype, dev, [50, 55])
    return net, sd, model


def _oracle(sd, model, dtype):
    from oracle.encp_oracle import EncPOracle
    return EncPOracle({k: v.to(dtype).float() for k, v in sd.items() if k.startswith(("enc_p.", "ge_to512.", "quantizer."))}, model)


@pytest.mark.parametrize("name,key", [("v2pro", "v2Pro"), ("v2", "v2")])
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_prior_encoder_matches_oracle_and_reference_golden(dev, name, key, dtype):
    from tests import gpu_harness as H
    g = H.golden(f"encp_{name}.npz")
    net, sd, model = _build(key, dtype, dev)
    orc = _oracle(sd, model, dtype)
    codes, text, ge = torch.from_numpy(g["codes"]), torch.from_numpy(g["text"]), torch.from_numpy(g["ge"])
    noise = torch.from_numpy(g["noise"])
    ge16 = ge.to(dtype)
    tol = TOL[dtype]
    # plain call with the reference's noise
    net._noise = noise[0]
    z_p, y_mask, ge_out, attn, m_p, logs_p = net.prior(codes, text, ge16, return_stats=True)
    net._noise = None
    zo, _, mo, lo, _ = orc.decode_front(codes, text, ge16.float(), noise)
    e_m = float((m_p.cpu() - mo[0]).abs().max())
    e_l = float((logs_p.cpu() - lo[0]).abs().max())
    e_z = float((z_p.float().cpu() - zo).abs().max())
    e_g = float(np.abs(m_p.cpu().numpy() - g["m_p"][0]).max())
    print(name, dtype, "m_p vs oracle", e_m, "logs_p", e_l, "z_p", e_z, "m_p vs reference golden", e_g)
    assert z_p.shape == (1, 192, 2 * codes.shape[-1]) and bool((y_mask == 1).all())
    assert e_m < tol and e_l < tol and e_z < 2 * tol and e_g < 1.5 * tol
    # the attention map decode() returns: probabilities over the text for every frame and head
    assert attn.shape == (4, 2 * codes.shape[-1], text.shape[-1])
    assert float((attn.sum(-1) - 1).abs().max()) < 1e-4
    # speed != 1
    sp = float(g["speed"])
    _, _, _, _, m_s, l_s = net.prior(codes, text, ge16, speed=sp, return_stats=True)
    assert tuple(m_s.shape) == tuple(g["m_p_speed"].shape[1:])
    assert float(np.abs(m_s.cpu().numpy() - g["m_p_speed"][0]).max()) < 1.5 * tol
    # MRTE restricted to a text window; masked text positions get no attention
    sl = torch.from_numpy(g["slice_indices"])
    _, _, _, attn_w, m_w, _ = net.prior(codes, text, ge16, slice_indices=sl, return_stats=True)
    assert float(np.abs(m_w.cpu().numpy() - g["m_p_slice"][0]).max()) < 1.5 * tol
    lo_, hi_ = int(sl[0, 0]), int(sl[0, 1])
    outside = [j for j in range(text.shape[-1] - 1) if not (lo_ <= j < hi_)]
    assert float(attn_w[:, :, outside].max()) < 1e-6
    # streaming: two chunks, the second cross-faded with the tail the first one left in the context
    net.enc_p.y_overlap = None
    for i, (n_codes, vs) in enumerate(g["stream_chunks"].tolist()):
        _, _, _, _, m_c, _ = net.prior(codes[:, :, :n_codes], text, ge16, stream_mode=True, valid_start_idx=vs, overlap_len=5,
                                       return_stats=True)
        assert float(np.abs(m_c.cpu().numpy() - g[f"m_p_stream{i}"][0]).max()) < 1.5 * tol, i
    net.enc_p.y_overlap = None
    # the text branch taken over from the previous call (the chunks of one utterance): bit-identical statistics; the flag is
    # ignored when the text length differs, and a dropped chunk's state can be rolled back
    got = {}
    for reuse in (False, True):
        net.enc_p.y_overlap = None
        got[reuse] = []
        for i, (n_codes, vs) in enumerate(g["stream_chunks"].tolist()):
            out = net.prior(codes[:, :, :n_codes], text, ge16, stream_mode=True, valid_start_idx=vs, overlap_len=5, return_stats=True,
                            text_unchanged=reuse and i > 0)
            got[reuse].append((out[0].clone(), out[4].clone()))
    for (z0, m0), (z1, m1) in zip(got[False], got[True]):
        assert torch.equal(m0, m1)
    shorter = text[:, :-2]
    _, _, _, _, m_a, _ = net.prior(codes, shorter, ge16, return_stats=True, text_unchanged=True)      # stale flag: another length
    _, _, _, _, m_b, _ = net.prior(codes, shorter, ge16, return_stats=True)
    assert torch.equal(m_a, m_b)
    net.enc_p.y_overlap = None
    (n0, v0), (n1, v1) = g["stream_chunks"].tolist()[:2]
    net.prior(codes[:, :, :n0], text, ge16, stream_mode=True, valid_start_idx=v0, overlap_len=5)
    net.prior(codes[:, :, :n1 - 1], text, ge16, stream_mode=True, valid_start_idx=v1, overlap_len=5)   # decoded ahead, then dropped
    net.enc_p.rollback()
    _, _, _, _, m_r, _ = net.prior(codes[:, :, :n1], text, ge16, stream_mode=True, valid_start_idx=v1, overlap_len=5, return_stats=True,
                                   text_unchanged=True)
    assert torch.equal(m_r, got[False][1][1])
    with pytest.raises(Exception):
        net.enc_p.rollback(); net.enc_p.rollback()
    net.enc_p.y_overlap = None


def test_decode_is_prior_then_flow_dec_and_other_sizes(dev):
    """decode() = prior + flow_dec on lengths the goldens do not cover (one code; text longer than the content; ge per
    frame), against the oracle chain."""
    from tests import gpu_harness as H
    from oracle.vocoder_oracle import VocoderOracle
    dtype = torch.float16
    net, sd, model = _build("v2Pro", dtype, dev, seed=1)
    orc = _oracle(sd, model, dtype)
    vo = VocoderOracle(H.folded_rounded_vocoder_sd({k: v for k, v in sd.items() if k.startswith(("flow.", "dec."))}, dtype), model)
    g = torch.Generator().manual_seed(9)
    for n, nt, ge_t in ((1, 3, False), (4, 30, False), (21, 6, True), (60, 40, False)):
        codes = torch.randint(0, 1024, (1, 1, n), generator=g)
        text = torch.randint(0, 732, (1, nt), generator=g)
        ge = torch.randn(1, model["gin_channels"], n if ge_t else 1, generator=g).to(dtype)
        net.debug_seed = 3
        audio, attn = net.decode(codes, text, ge, noise_scale=0.0)
        assert audio.shape == (1, 1, 2 * n * 640) and attn.shape == (4, 2 * n, nt)
        z_o, mask_o, m_o, _, ge_o = orc.decode_front(codes, text, ge.float(), None)
        a_o = vo.generator(vo.flow_reverse(z_o.to(dtype).float(), mask_o, ge_o) * mask_o, ge_o)
        err = float((audio.float().cpu() - a_o).abs().max())
        print("decode", n, nt, ge_t, "audio vs oracle chain", err)
        assert err < 2e-2          # 16-bit latents from a 12-layer 16-bit encoder feed the vocoder


def test_decode_against_the_reference_module_on_this_gpu(dev):
    """INTEGRATION.md, executed: the reference's own SynthesizerTrn (unmodified files under baseline/_ref or
    /root/reference) with the same weights on the same GPU and dtype; its decode() against ours (noise_scale = 0: the two
    draw their prior noise from different generators), and ours dropped in as its flow_dec."""
    from oracle import ref_shim
    if not ref_shim.available():
        pytest.skip("reference modules not available on this box")
    dtype = torch.float16
    net, sd, model = _build("v2Pro", dtype, dev, seed=2)
    M = ref_shim.sovits_models()
    with torch.inference_mode():
        ref = M.SynthesizerTrn(1025, 32, n_speakers=300, **model).eval()
        ref.dec.remove_weight_norm()                      # Loader.py:73, 95
        missing = ref.load_state_dict(sd, strict=False)
        assert not [k for k in missing.unexpected_keys]
        ref = ref.to(dev, dtype)
        g = torch.Generator().manual_seed(5)
        codes = torch.randint(0, 1024, (1, 1, 40), generator=g).to(dev)
        text = torch.randint(0, 732, (1, 25), generator=g).to(dev)
        ge = torch.randn(1, model["gin_channels"], 1, generator=g).to(dev, dtype)
        o_ref, attn_ref = ref.decode(codes, text, ge, noise_scale=0.0, cuda_graph=False)
        o_b200, attn_b200 = net.decode(codes, text, ge, noise_scale=0.0)
        e_audio = float((o_ref.float() - o_b200.float()).abs().max())
        e_attn = float((attn_ref.float() - attn_b200).abs().max())
        print("decode vs reference module on this GPU: audio", e_audio, "attn", e_attn, "audio peak", float(o_ref.abs().max()))
        assert o_ref.shape == o_b200.shape and attn_ref.shape == attn_b200.shape
        assert e_audio < 2e-2 and e_attn < 2e-2
        # the patch of INTEGRATION.md: the reference object keeps its enc_p, our kernels replace flow + dec
        ref.flow_dec = lambda z_p, y_mask, g_: net.flow_dec(z_p, y_mask, g_)
        o_mixed, _ = ref.decode(codes, text, ge, noise_scale=0.0, cuda_graph=False)
        e_mixed = float((o_ref.float() - o_mixed.float()).abs().max())
        print("reference enc_p + B200 flow_dec vs reference:", e_mixed)
        assert e_mixed < 5e-3
