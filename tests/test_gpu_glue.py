"""GPU: the host glue of TTS.infer / infer_stream as device kernels (csrc/glue.cu) against outputs of the
reference's own method bodies (tests/golden/glue.npz) and the oracle, and the reference-flow entry points
``TTS.infer_phones`` / ``infer_phones_stream`` / ``infer`` (with a stand-in front end) against a restatement of
TTS.py:232-286 / 402-498 built from the same native model calls and the CPU oracle of the glue."""
import numpy as np
import pytest
import torch

from gsv_tts import _synthetic as syn

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tts(tmp_path_factory):
    from gsv_tts import TTS
    tmp = tmp_path_factory.mktemp("ckpt")
    cfg = syn.GPT_CONFIG_TINY
    gsd = syn.gpt_state_dict(cfg, 0, 6.0)
    gpath = tmp / "s1.ckpt"
    torch.save({"config": cfg, "weight": gsd}, gpath)
    model = dict(syn.SOVITS_MODEL["v2"])
    sd = dict(syn.sovits_flow_dec_state_dict(model, 0))
    sd.update(syn.sovits_encp_state_dict(model, 0))
    spath = tmp / "s2.pth"
    torch.save({"config": {"data": {"filter_length": 2048, "hop_length": 640, "n_speakers": 300}, "train": {"segment_size": 20480},
                           "model": model}, "weight": sd}, spath)
    t = TTS(gpt_cache=[(1, 256), (4, 256)], sovits_cache=[50, 55], device="cuda:0", dtype="float16")
    t.load_gpt_model(str(gpath))
    t.load_sovits_model(str(spath))
    return t


def test_viterbi_offsets_sola_match_reference_goldens(tts):
    from tests import gpu_harness as H
    from oracle import glue_oracle as G
    g = H.golden("glue.npz")
    dev = torch.device("cuda:0")
    for i in range(4):
        got = tts._viterbi_monotonic(torch.from_numpy(g[f"attn{i}"]).to(dev)).cpu().numpy()
        assert (got == g[f"assign{i}"]).all(), (i, np.nonzero(got != g[f"assign{i}"])[0][:8])
    audio = torch.from_numpy(g["audio"]).to(dev, torch.float16)
    assert tts._find_head_threshold_offsets(audio) == int(g["head_offset"])
    assert tts._find_tail_threshold_offsets(audio) == int(g["tail_offset"])
    silent = torch.zeros(5000, device=dev, dtype=torch.float16)
    assert tts._find_head_threshold_offsets(silent) == int(g["head_offset_silent"])
    assert tts._find_tail_threshold_offsets(silent) == int(g["tail_offset_silent"])
    short = torch.zeros(100, device=dev, dtype=torch.float16)            # shorter than one frame: nothing to find
    assert tts._find_head_threshold_offsets(short) == G.head_offset(np.zeros(100, np.float32)) == 100
    f1 = torch.from_numpy(g["sola_f1"]).to(dev, torch.float16)
    f2 = torch.from_numpy(g["sola_f2"]).to(dev, torch.float16)
    out, off = tts._sola_algorithm(f1.view(1, 1, -1), f2.view(1, 1, -1), 3200)
    assert int(off) == int(g["sola_offset"])
    assert out.shape == (1, 1, g["sola_out"].shape[0])
    assert float(np.abs(out[0, 0].float().cpu().numpy() - g["sola_out"]).max()) < 5e-3      # 16-bit inputs and cross-fade


class StandInFrontEnd:
    """What the text / audio front end supplies (out of scope: G2P, BERT, HuBERT, speaker embedding)."""

    def __init__(self, gin):
        g = torch.Generator().manual_seed(3)
        self.ge = torch.randn(1, gin, 1, generator=g)
        self.prompt_tokens = torch.randint(0, 1024, (1, 30), generator=g)
        self.phones1 = torch.randint(0, 732, (12,), generator=g).tolist()
        self.bert1 = torch.zeros(12, 1024)
        self.g = g

    def speaker(self, tts, sovits_model, spk_audio_path):
        return self.ge

    def prompt(self, tts, gpt_model, prompt_audio_path, prompt_audio_text):
        return self.prompt_tokens, self.phones1, self.bert1

    def phones_and_bert(self, text):
        n = max(4, len(text))
        gg = torch.Generator().manual_seed(len(text))
        phones2 = torch.randint(0, 732, (n,), generator=gg).tolist()
        word2ph = {"word": list(text[:n - 1]) + ["."], "ph": [1] * n}
        return phones2, word2ph, torch.zeros(n, 1024), text


def test_infer_phones_is_the_reference_flow(tts):
    from oracle import glue_oracle as G
    fe = StandInFrontEnd(512)
    tts.frontend = fe
    gpt = next(iter(tts.gpt_models.values())).t2s_model
    vq = next(iter(tts.sovits_models.values())).vq_model
    text = "hello b200 kernels"
    gpt.debug_seed, vq.debug_seed = 31, 7
    clip = tts.infer("spk.wav", "prompt.wav", "prompt text", text, return_subtitles=True)
    # restatement of TTS.py:232-286 from the same native model calls and the CPU oracle of the glue
    phones2, word2ph, bert2, _ = fe.phones_and_bert(text)
    dev = torch.device("cuda:0")
    ids = torch.tensor(fe.phones1 + phones2, device=dev).unsqueeze(0)
    gpt.debug_seed = 31
    pred = gpt.infer(ids, fe.prompt_tokens, torch.cat([fe.bert1, bert2]).unsqueeze(0))
    audio, attn = vq.decode(pred, torch.tensor(phones2, device=dev).unsqueeze(0), fe.ge)
    a = audio[0, 0].float().cpu().numpy()
    head = G.head_offset(a)
    want = a[head:]
    peak = np.abs(want).max()
    want = want / peak if peak > 1 else want
    want = np.concatenate([want, np.zeros(int(0.2 * 32000), np.float32)])
    assert clip.audio_data.shape == want.shape and float(np.abs(clip.audio_data - want).max()) == 0.0
    assert clip.orig_text == text and abs(clip.audio_len_s - len(want) / 32000) < 1e-9
    assign = G.viterbi_monotonic(attn.cpu().numpy())
    subs = G.get_subtitles(word2ph, assign, 1.0)
    assert [s["text"] for s in clip.subtitles] == [s["text"] for s in subs]
    assert abs(clip.subtitles[-1]["end_s"] - (subs[-1]["end_s"] + 0.2 - head / 32000)) < 1e-6
    tts.frontend = None
    with pytest.raises(Exception):
        tts.infer("spk.wav", "prompt.wav", "prompt text", text)


def test_infer_phones_stream_splices_chunks_like_the_reference(tts):
    from oracle import glue_oracle as G
    fe = StandInFrontEnd(512)
    gpt = next(iter(tts.gpt_models.values())).t2s_model
    vq = next(iter(tts.sovits_models.values())).vq_model
    dev = torch.device("cuda:0")
    text = "a streaming sentence."
    phones2, _, bert2, _ = fe.phones_and_bert(text)
    gpt.debug_seed, vq.debug_seed = 41, 9
    clips = list(tts.infer_phones_stream(fe.phones1, fe.bert1, fe.prompt_tokens, phones2, bert2, fe.ge, stream_chunk=10, overlap_len=5,
                                         force_steps=34))
    assert len(clips) == 3                                   # 10 tokens (boosted first chunk), 20 (one chunk late), the final 34
    # restatement of TTS.py:402-470 with the CPU oracle of SOLA / head trim around the same native decode calls
    ids = torch.tensor(fe.phones1 + phones2, device=dev).unsqueeze(0)
    ph2 = torch.tensor(phones2, device=dev).unsqueeze(0)
    gpt.debug_seed = 41
    vq.enc_p.y_overlap = None
    ov = 5 * 640
    last, vs, want = None, 0, []
    for c, (pred, final) in enumerate(gpt.infer_stream(ids, fe.prompt_tokens, torch.cat([fe.bert1, bert2]).unsqueeze(0), stream_chunk=10,
                                                      force_steps=34)):
        audio, attn = vq.decode(pred, ph2, fe.ge, stream_mode=True, valid_start_idx=vs, overlap_len=5)
        a = audio[0, 0].float().cpu().numpy()
        if last is not None:
            a, _ = G.sola(last, a, ov)
        last = a[-ov:].copy()
        if not final:
            a = a[:-ov]
            vs = attn.shape[1] - 5
        if c == 0:
            a = a[G.head_offset(a):]
        if final:
            a = np.concatenate([a, np.zeros(int(0.4 * 32000), np.float32)])
        want.append(a)
    vq.enc_p.y_overlap = None
    assert [len(c.audio_data) for c in clips] == [len(a) for a in want]
    for c, a in zip(clips, want):
        assert float(np.abs(c.audio_data - a).max()) < 5e-3      # the kernel cross-fades in the 16-bit storage type
    assert abs(clips[-1].audio_len_s - sum(len(a) for a in want) / 32000) < 1e-6


@pytest.mark.parametrize("force", [34, 40, 30, 9, None])
def test_chunks_decoded_ahead_give_the_same_clips(tts, force):
    """``decode_ahead`` starts a held-back chunk's SoVITS stage early and rolls the cross-chunk state back when the stream
    ends before the next boundary (34: the chunk of 30 is dropped; 40 / 30: ends on a boundary; 9: shorter than a chunk; None:
    until EOS or the cache cap): the clips must be the ones of the one-chunk-late order, bit for bit."""
    fe = StandInFrontEnd(512)
    gpt = next(iter(tts.gpt_models.values())).t2s_model
    vq = next(iter(tts.sovits_models.values())).vq_model
    phones2, _, bert2, _ = fe.phones_and_bert("a streaming sentence, decoded ahead.")
    runs = []
    for ahead in (False, True):
        gpt.debug_seed, vq.debug_seed = 43, 11
        runs.append(list(tts.infer_phones_stream(fe.phones1, fe.bert1, fe.prompt_tokens, phones2, bert2, fe.ge, stream_chunk=10,
                                                 overlap_len=5, force_steps=force, decode_ahead=ahead)))
    a, b = runs
    assert len(a) == len(b) and len(a) >= 1
    for ca, cb in zip(a, b):
        assert ca.audio_data.shape == cb.audio_data.shape
        assert float(np.abs(ca.audio_data - cb.audio_data).max()) == 0.0 if ca.audio_data.size else True
        assert abs(ca.audio_len_s - cb.audio_len_s) < 1e-9
