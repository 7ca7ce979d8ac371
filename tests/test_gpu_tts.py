"""GPU: the kept public surface end to end on synthetic checkpoints written in the reference's file formats:
TTS(...) -> load_gpt_model / load_sovits_model (Loader) -> infer_features (GPT decode + flow + HiFi-GAN)."""
import pytest
import torch

from gsv_tts import _synthetic as syn

pytestmark = pytest.mark.gpu


def test_tts_surface_on_synthetic_checkpoints(tmp_path):
    from gsv_tts import TTS, AudioClip
    from tests.test_loader_cpu import _upstream_gpt_names
    cfg = syn.GPT_CONFIG_TINY
    gsd = syn.gpt_state_dict(cfg, 0, 6.0)
    gpt_path = tmp_path / "s1.ckpt"
    torch.save({"config": cfg, "weight": _upstream_gpt_names(gsd, cfg["model"]["n_layer"])}, gpt_path)
    model = dict(syn.SOVITS_MODEL["tiny"])
    vsd = syn.sovits_flow_dec_state_dict(model, 0)
    pth = tmp_path / "s2.pth"
    torch.save({"config": {"model": model}, "weight": vsd}, pth)

    tts = TTS(gpt_cache=[(1, 256), (4, 256)], sovits_cache=[50, 55], device="cuda:0", dtype="float16")
    tts.load_gpt_model(str(gpt_path))
    tts.load_sovits_model(str(pth))
    assert tts.get_gpt_list() == [str(gpt_path)] and tts.get_sovits_list() == [str(pth)]

    g = torch.Generator().manual_seed(3)
    x = torch.randint(0, 732, (1, 24), generator=g)
    y = torch.randint(0, 1024, (1, 30), generator=g)
    bert = torch.zeros(1, 24, 1024)
    T = 20
    z_p = torch.randn(1, 192, T, generator=g)
    ge = torch.randn(1, model["gin_channels"], 1, generator=g)
    torch.manual_seed(11)
    tokens, clip = tts.infer_features(x, bert, y, z_p, torch.ones(1, 1, T), ge)
    assert tokens.dtype == torch.int64 and tokens.shape[:2] == (1, 1) and 0 < tokens.shape[-1] < 256
    assert isinstance(clip, AudioClip) and clip.samplerate == 32000
    assert clip.audio_data.shape == (T * 640,) and abs(clip.audio_len_s - T * 0.02) < 1e-9
    assert float(abs(clip.audio_data).max()) <= 1.0
    # batched GPT stage: results come back in request order
    outs = tts.infer_features_batched([x[0]] * 3, [bert[0]] * 3, [y[0]] * 3)
    assert len(outs) == 3 and all(o.dtype == torch.int64 and o.numel() > 0 for o in outs)
    with pytest.raises(Exception):
        tts.infer("text needs the front end")
    tts.unload_gpt_model(str(gpt_path))
    assert tts.get_gpt_list() == []


def test_streaming_overlap_matches_back_to_back(tmp_path):
    """infer_features_stream (vocoder of chunk c on a second stream while chunk c+1 decodes on 128 SMs) yields the
    token chunks of Text2SemanticDecoder.infer_stream and the audio of a plain flow_dec call on the same features."""
    from gsv_tts import TTS
    from tests.test_loader_cpu import _upstream_gpt_names
    cfg = syn.GPT_CONFIG_TINY
    gsd = syn.gpt_state_dict(cfg, 0, 6.0)
    gpt_path = tmp_path / "s1.ckpt"
    torch.save({"config": cfg, "weight": _upstream_gpt_names(gsd, cfg["model"]["n_layer"])}, gpt_path)
    model = dict(syn.SOVITS_MODEL["tiny"])
    pth = tmp_path / "s2.pth"
    torch.save({"config": {"model": model}, "weight": syn.sovits_flow_dec_state_dict(model, 0)}, pth)
    tts = TTS(gpt_cache=[(1, 256)], sovits_cache=[50, 55], device="cuda:0", dtype="float16")
    tts.load_gpt_model(str(gpt_path))
    tts.load_sovits_model(str(pth))
    gpt = tts.gpt_models[str(gpt_path)].t2s_model
    voc = tts.sovits_models[str(pth)].vq_model

    g = torch.Generator().manual_seed(5)
    x = torch.randint(0, 732, (1, 24), generator=g)
    y = torch.randint(0, 1024, (1, 30), generator=g)
    bert = torch.zeros(1, 24, 1024)
    ge = torch.randn(1, model["gin_channels"], 1, generator=g)
    T = 10
    seen = []

    def features_of_chunk(tokens, final):
        # stands for enc_p: frames derived from the tokens handed over (so a wrong or late chunk changes the audio)
        seen.append((tokens.cpu().clone(), final))
        zg = torch.Generator().manual_seed(int(tokens.sum().item()) % 1000 + len(seen))
        return torch.randn(1, 192, T, generator=zg), torch.ones(1, 1, T), ge

    gpt.debug_seed = 21
    clips = list(tts.infer_features_stream(x, bert, y, features_of_chunk, stream_chunk=10))
    assert len(clips) == len(seen) >= 2 and seen[-1][1]
    # the same request, decode and vocoder back to back on one stream
    gpt.debug_seed = 21
    ref_chunks = [(t.cpu().clone(), f) for t, f in gpt.infer_stream(x, y, bert, stream_chunk=10)]
    assert len(ref_chunks) == len(seen)
    for (t, f), (tr, fr) in zip(seen, ref_chunks):
        assert f == fr and torch.equal(t, tr)
    for i, ((t, f), clip) in enumerate(zip(seen, clips)):
        zg = torch.Generator().manual_seed(int(t.sum().item()) % 1000 + i + 1)
        ref = voc.flow_dec(torch.randn(1, 192, T, generator=zg), torch.ones(1, 1, T), ge)[0, 0].float().cpu().numpy()
        assert clip.audio_data.shape == ref.shape
        assert float(abs(clip.audio_data - ref).max()) == 0.0


def test_batched_pipeline_overlap_matches_back_to_back(tmp_path):
    """infer_phones_batched: the SoVITS stage of the requests harvested at one read runs on the second stream behind the next
    decode launch.  Same tokens and, with one utterance per vocoder group, bit-equal audio as the back-to-back order; default
    grouping (padded groups) gives clips of the right lengths."""
    from gsv_tts import TTS
    from tests.test_loader_cpu import _upstream_gpt_names
    cfg = syn.GPT_CONFIG_TINY
    gsd = syn.gpt_state_dict(cfg, 0, 6.0)
    gpt_path = tmp_path / "s1.ckpt"
    torch.save({"config": cfg, "weight": _upstream_gpt_names(gsd, cfg["model"]["n_layer"])}, gpt_path)
    model = dict(syn.SOVITS_MODEL["tiny"], version="v2Pro")
    sd = dict(syn.sovits_flow_dec_state_dict(model, 0))
    sd.update(syn.sovits_encp_state_dict(model, 0))
    pth = tmp_path / "s2.pth"
    torch.save({"config": {"model": model}, "weight": sd}, pth)
    tts = TTS(gpt_cache=[(1, 256), (4, 256)], sovits_cache=[50, 55], device="cuda:0", dtype="float16")
    tts.load_gpt_model(str(gpt_path))
    tts.load_sovits_model(str(pth))
    gpt = tts.gpt_models[str(gpt_path)].t2s_model
    vq = tts.sovits_models[str(pth)].vq_model
    g = torch.Generator().manual_seed(8)
    n_req = 9
    xs = [torch.randint(0, 732, (int(torch.randint(12, 30, (1,), generator=g)),), generator=g) for _ in range(n_req)]
    ys = [torch.randint(0, 1024, (int(torch.randint(20, 40, (1,), generator=g)),), generator=g) for _ in range(n_req)]
    bs = [torch.zeros(x.numel(), 1024) for x in xs]
    ph2 = [x[x.numel() // 2:].cuda().view(1, -1) for x in xs]
    ges = [torch.randn(1, model["gin_channels"], 1, generator=g).cuda().half() for _ in range(n_req)]
    mx = [int(torch.randint(6, 40, (1,), generator=g)) for _ in range(n_req)]
    runs = {}
    for overlap in (False, True):
        gpt.debug_seed, vq.debug_seed = 31, 5
        runs[overlap] = tts.infer_phones_batched(xs, bs, ys, ph2, ges, max_new=mx, overlap=overlap, max_frames=1)
    (t0, c0), (t1, c1) = runs[False], runs[True]
    assert all(torch.equal(a, b) for a, b in zip(t0, t1)) and sum(t.numel() > 0 for t in t0) >= 5
    for i, (a, b) in enumerate(zip(c0, c1)):
        assert a.audio_data.shape == b.audio_data.shape == (2 * t0[i].numel() * 640,), i
        assert a.audio_data.size == 0 or float(abs(a.audio_data - b.audio_data).max()) == 0.0, i     # EOS first: an empty clip
    gpt.debug_seed, vq.debug_seed = 31, 5
    t2, c2 = tts.infer_phones_batched(xs, bs, ys, ph2, ges, max_new=mx)
    assert all(torch.equal(a, b) for a, b in zip(t0, t2))
    for i, c in enumerate(c2):
        assert c.audio_data.shape == (2 * t0[i].numel() * 640,)
        n = c.audio_data.shape[0] - 20 * 640                       # the tail sees the group's padded frames instead of the signal's end
        if n > 0:
            assert float(abs(c.audio_data[:n] - c0[i].audio_data[:n]).max()) < 2e-3, i
    # the prior encoder of the utterances of one SoVITS stage dealt over several native contexts / streams: bit-equal clips
    toks = [t for t in t0]
    per_lane = {}
    for lanes in (1, 3, 4):
        vq.debug_seed = 5
        per_lane[lanes] = tts.decode_batched(toks, ph2, ges, max_frames=1, lanes=lanes)
    for lanes in (3, 4):
        for i, (a, b) in enumerate(zip(per_lane[1], per_lane[lanes])):
            assert a.audio_data.shape == b.audio_data.shape, (lanes, i)
            assert a.audio_data.size == 0 or float(abs(a.audio_data - b.audio_data).max()) == 0.0, (lanes, i)

