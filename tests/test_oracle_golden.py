"""CPU: the oracle restatement against the golden vectors produced by the REFERENCE's own
modules (oracle/make_golden.py).  This is what pins the oracle."""
import os

import numpy as np
import pytest
import torch

from gsv_tts import _synthetic as syn
from oracle.gpt_oracle import GptOracle, prompt_mask
from oracle.vocoder_oracle import VocoderOracle

torch.set_grad_enabled(False)


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def _tf_logits(orc, g):
    x, y = torch.from_numpy(g["x"]), torch.from_numpy(g["y"])
    bert = torch.from_numpy(g["bert"])
    K, V, kv_len = orc.new_cache(1, int(g["max_seq"]))
    h = orc.prefill(x, y, bert, K, V, kv_len)
    rows = [orc.logits(h.unsqueeze(0))[0]]
    for t in g["forced"].tolist():
        xin = orc.embed_next(torch.tensor([t]), kv_len - x.shape[0])
        rows.append(orc.logits(orc.decode_step(xin, K, V, kv_len))[0])
    return torch.stack(rows).numpy()


@pytest.mark.parametrize("name,cfg", [("tiny", syn.GPT_CONFIG_TINY), ("full", syn.GPT_CONFIG)])
def test_gpt_teacher_forced_logits(golden_dir, name, cfg):
    g = _load(golden_dir, f"gpt_{name}.npz")
    orc = GptOracle(syn.gpt_state_dict(cfg, 0, float(g["eos_boost"])), cfg)
    got = _tf_logits(orc, g)
    assert np.abs(got - g["tf_logits"]).max() < 2e-4      # fp32 vs fp32, different summation order


def test_gpt_infer_tokens_exact(golden_dir):
    g = _load(golden_dir, "gpt_tiny.npz")
    cfg = syn.GPT_CONFIG_TINY
    orc = GptOracle(syn.gpt_state_dict(cfg, 0, float(g["eos_boost"])), cfg)
    x, y, bert = (torch.from_numpy(g[k]) for k in ("x", "y", "bert"))
    torch.manual_seed(int(g["infer_seed"]))
    toks = orc.infer(x, y, bert, max_seq=int(g["max_seq"]))
    assert toks.shape == (1, 1, len(g["infer_tokens"]))
    assert (toks[0, 0].numpy() == g["infer_tokens"]).all()
    # streaming: same chunk boundaries, first and final chunk identical (incl. the quirk that the
    # final chunk after an EOS break carries the first sampled token, t2s_model.py:534-553)
    torch.manual_seed(int(g["infer_seed"]))
    chunks = list(orc.infer_stream(x, y, bert, stream_chunk=10, max_seq=int(g["max_seq"])))
    assert [c.shape[-1] for c, _ in chunks] == g["stream_lens"].tolist()
    assert (chunks[0][0][0, 0].numpy() == g["stream_first"]).all()
    assert (chunks[-1][0][0, 0].numpy() == g["stream_final"]).all()
    assert [f for _, f in chunks] == [False] * (len(chunks) - 1) + [True]


def test_prompt_mask_rule():
    m = prompt_mask(3, 2)
    assert m.tolist() == [
        [True, True, True, False, False],
        [True, True, True, False, False],
        [True, True, True, False, False],
        [True, True, True, True, False],
        [True, True, True, True, True],
    ]


def test_gpt_batched_contract(golden_dir):
    """Reference infer_batched on 7 requests through 4 slots: every request comes back once,
    without its first sampled token and cut before EOS.  The per-request oracle must agree on
    the invariants (lengths differ: the reference's RNG stream depends on its 5-step schedule)."""
    g = _load(golden_dir, "gpt_tiny_batched.npz")
    n = int(g["n_req"])
    assert sorted(g["order"].tolist()) == list(range(n))
    for i in range(n):
        t = g[f"tok{i}"]
        assert (t != 1024).all() and len(t) == g["lens"][i]
    cfg = syn.GPT_CONFIG_TINY
    orc = GptOracle(syn.gpt_state_dict(cfg, 0, float(g["eos_boost"])), cfg)
    xs = [torch.from_numpy(g[f"x{r}"]) for r in range(3)]
    ys = [torch.from_numpy(g[f"y{r}"]) for r in range(3)]
    bs = [torch.from_numpy(g[f"b{r}"].astype(np.float32)) for r in range(3)]
    torch.manual_seed(1)
    toks, order = orc.infer_batched(xs, ys, bs, slots=2, max_seq=int(g["max_seq"]))
    assert order == [0, 1, 2]
    for t, x, y in zip(toks, xs, ys):
        assert (t != 1024).all() and len(t) <= int(g["max_seq"]) - len(x) - len(y)


@pytest.mark.parametrize("name,key", [("tiny", "tiny"), ("tiny_ge_t", "tiny"), ("v2pro", "v2Pro"),
                                      ("v2proplus", "v2ProPlus"), ("v2", "v2")])
def test_vocoder_flow_dec(golden_dir, name, key):
    g = _load(golden_dir, f"vocoder_{name}.npz")
    model = syn.SOVITS_MODEL[key]
    vo = VocoderOracle(syn.sovits_flow_dec_state_dict(model, 0), model)
    z_p, mask, ge = (torch.from_numpy(g[k]) for k in ("z_p", "mask", "ge"))
    z = vo.flow_reverse(z_p, mask, ge)
    assert np.abs(z.numpy() - g["z"]).max() < 1e-4
    audio = vo.flow_dec(z_p, mask, ge)
    assert audio.shape == (z_p.shape[0], 1, z_p.shape[2] * 640)
    assert np.abs(audio.numpy() - g["audio"]).max() < 1e-4
