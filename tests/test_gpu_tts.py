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
