"""CPU, build container only: the oracle restatement against the REFERENCE's own modules run live
(imported from /root/reference through oracle/ref_shim.py).  Skipped wherever the reference tree is
absent (the GPU box): there the committed goldens (tests/test_oracle_golden.py) are the pin."""
import numpy as np
import pytest
import torch

from gsv_tts import _synthetic as syn
from oracle import ref_shim
from oracle.gpt_oracle import GptOracle
from oracle.vocoder_oracle import VocoderOracle

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present")
torch.set_grad_enabled(False)


def test_gpt_prefill_and_decode_logits_match_reference_live():
    """Fresh seed (not one of the golden seeds): reference process_prompt + 6 teacher-forced
    decode_next_token steps (t2s_model.py:114-143) vs the oracle, fp32, tiny config."""
    cfg = syn.GPT_CONFIG_TINY
    sd = syn.gpt_state_dict(cfg, 3)
    g = torch.Generator().manual_seed(77)
    nx, ny, max_seq = 11, 17, 64
    x = torch.randint(0, 732, (nx,), generator=g)
    y = torch.randint(0, 1024, (ny,), generator=g)
    bert = torch.randn(nx, 1024, generator=g)
    forced = torch.randint(0, 1024, (6,), generator=g)
    from oracle.make_golden import gpt_teacher_forced
    with torch.inference_mode():
        ref = ref_shim.build_reference_gpt(sd, cfg, torch.float32, "cpu", [(1, max_seq)])
        want, _ = gpt_teacher_forced(ref, x, y, bert, forced, max_seq)
    want = want.numpy()

    orc = GptOracle(sd, cfg)
    K, V, kv_len = orc.new_cache(1, max_seq)
    hh = orc.prefill(x, y, bert, K, V, kv_len)
    got = [orc.logits(hh.unsqueeze(0))[0]]
    for t in forced.tolist():
        xin = orc.embed_next(torch.tensor([t]), kv_len - nx)
        got.append(orc.logits(orc.decode_step(xin, K, V, kv_len))[0])
    got = torch.stack(got).numpy()
    assert np.abs(got - want).max() < 2e-4


def test_flow_dec_matches_reference_live():
    """Fresh seed: reference flow(reverse) + dec (models.py:380-383) vs the oracle, fp32, tiny sizes."""
    model = syn.SOVITS_MODEL["tiny"]
    sd = syn.sovits_flow_dec_state_dict(model, 5)
    g = torch.Generator().manual_seed(78)
    T = 9
    z_p = torch.randn(1, 192, T, generator=g)
    mask = torch.ones(1, 1, T)
    ge = torch.randn(1, model["gin_channels"], 1, generator=g)
    with torch.inference_mode():
        flow, dec = ref_shim.build_reference_flow_dec(sd, model, torch.float32, "cpu")
        z = flow(z_p, mask, g=ge, reverse=True)
        want = dec(z * mask, g=ge).numpy()
    from tests.gpu_harness import folded_rounded_vocoder_sd
    vo = VocoderOracle(folded_rounded_vocoder_sd(sd, torch.float32), model)
    got = vo.flow_dec(z_p, mask, ge).numpy()
    assert np.abs(got - want).max() < 2e-4
