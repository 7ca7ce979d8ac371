"""CPU: checkpoint ingestion of the B200 backend accepts the reference's three formats
(reference gsv_tts/Loader.py:42-170): upstream .ckpt / .pth (incl. the 2-byte version tag) and
safetensors directories; the public package surface imports without the front-end dependencies."""
import json
import os

import pytest
import torch

from gsv_tts import Loader, _synthetic as syn


def _upstream_gpt_names(sd, n_layer):
    """Inverse of the remap: Lite names -> upstream GPT-SoVITS names (what a user's .ckpt holds)."""
    back = {"qkv.weight": "self_attn.in_proj_weight", "qkv.bias": "self_attn.in_proj_bias",
            "out_proj.weight": "self_attn.out_proj.weight", "out_proj.bias": "self_attn.out_proj.bias",
            "mlp.0.weight": "linear1.weight", "mlp.0.bias": "linear1.bias",
            "mlp.2.weight": "linear2.weight", "mlp.2.bias": "linear2.bias"}
    out = {}
    for k, v in sd.items():
        if k.startswith("t2s_transformer.blocks."):
            _, _, i, tail = k.split(".", 3)
            out[f"model.h.layers.{i}.{back.get(tail, tail)}"] = v
        else:
            out["model." + k] = v
    return out


def test_gpt_ckpt_and_safetensors_dir(tmp_path):
    cfg = syn.GPT_CONFIG_TINY
    sd = syn.gpt_state_dict(cfg, 0)
    ckpt = tmp_path / "s1.ckpt"
    torch.save({"config": cfg, "weight": _upstream_gpt_names(sd, cfg["model"]["n_layer"])}, ckpt)
    config, got = Loader.read_gpt_checkpoint(str(ckpt))
    assert config == cfg and sorted(got) == sorted(sd)
    assert all(torch.equal(got[k], sd[k]) for k in sd)
    # the mirror class accepts exactly this key set (strict load)
    from gsv_tts.GPT_SoVITS.GPT.t2s_model_b200 import Text2SemanticDecoder
    Text2SemanticDecoder(config).load_state_dict(got)
    # safetensors directory (Lite layout, as TTS.to_safetensors writes it)
    from safetensors.torch import save_file
    d = tmp_path / "gpt_dir"
    d.mkdir()
    save_file({k: v.contiguous() for k, v in sd.items()}, str(d / "model.safetensors"))
    (d / "config.json").write_text(json.dumps(cfg))
    config2, got2 = Loader.read_gpt_checkpoint(str(d))
    assert config2 == cfg and all(torch.equal(got2[k], sd[k]) for k in sd)


@pytest.mark.parametrize("tag,version", [(b"05", "v2Pro"), (b"06", "v2ProPlus"), (None, "v2")])
def test_sovits_pth_version_tag_and_dir(tmp_path, tag, version):
    model = dict(syn.SOVITS_MODEL["tiny"], version=version)
    sd = syn.sovits_flow_dec_state_dict(model, 0)
    hps = {"data": {"filter_length": 2048, "hop_length": 640, "n_speakers": 300}, "train": {"segment_size": 20480},
           "model": {k: v for k, v in model.items() if k != "version" or tag is None}}
    pth = tmp_path / "s2.pth"
    torch.save({"config": hps, "weight": sd}, pth)
    if tag is not None:                                    # versioned files carry the tag instead of the zip magic
        raw = pth.read_bytes()
        assert raw[:2] == b"PK"
        pth.write_bytes(tag + raw[2:])
    got_hps, got_sd, got_version = Loader.read_sovits_checkpoint(str(pth))
    assert got_version == version and got_hps["model"]["version"] == version
    assert got_hps["model"]["semantic_frame_rate"] == "25hz"
    assert all(torch.equal(got_sd[k], sd[k]) for k in sd)
    from gsv_tts.GPT_SoVITS.SoVITS.models_b200 import FlowDecoder
    fd = FlowDecoder(**got_hps["model"])
    fd.load_state_dict(got_sd)
    assert sorted(fd.state_dict()) == sorted(sd)


def test_sovits_rejects_unknown_version(tmp_path):
    model = dict(syn.SOVITS_MODEL["tiny"])
    model.pop("version")
    pth = tmp_path / "old.pth"
    torch.save({"config": {"model": model}, "weight": {}}, pth)
    with pytest.raises(ValueError):
        Loader.read_sovits_checkpoint(str(pth))


def test_public_surface_imports_and_refuses_cpu():
    import gsv_tts
    from gsv_tts import AudioClip, TTS, cut_text
    assert cut_text("你好。今天天气不错！Hello there. Bye.", max_len=8) == ["你好。", "今天天气不错！", "Hello there.", " Bye."]
    clip = AudioClip([0.0, 0.5, -0.5, 0.25], samplerate=4)
    assert clip.audio_len_s == 1.0 and clip.audio_data.dtype.name == "float32"
    with pytest.raises(RuntimeError):
        TTS(device="cpu")


def test_sovits_pth_with_config_pickled_from_module_utils(tmp_path):
    """Real upstream .pth files pickle ``config`` as ``utils.HParams`` (a top-level module named ``utils``; the reference
    registers its own under that name before torch.load, Loader.py:13-14).  Write such a file from a throw-away module
    of that name, make sure the name is gone again, and load it."""
    import subprocess
    import sys
    import textwrap
    model = dict(syn.SOVITS_MODEL["tiny"], version="v2")
    pth = tmp_path / "hp.pth"
    writer = tmp_path / "write_it.py"
    (tmp_path / "utils.py").write_text(textwrap.dedent("""
        class HParams:
            def __init__(self, **kwargs):
                for k, v in kwargs.items():
                    if type(v) == dict:
                        v = HParams(**v)
                    setattr(self, k, v)
    """))
    writer.write_text(textwrap.dedent(f"""
        import sys, torch
        sys.path.insert(0, {str(tmp_path)!r})
        import utils
        hps = utils.HParams(data=dict(hop_length=640), train=dict(segment_size=20480), model={model!r})
        torch.save({{"config": hps, "weight": {{"dec.conv_post.weight": torch.ones(1, 4, 7)}}}}, {str(pth)!r})
    """))
    subprocess.run([sys.executable, str(writer)], check=True)
    assert "utils" not in sys.modules or not hasattr(sys.modules["utils"], "HParams")
    before = sys.modules.get("utils")
    hps, sd, version = Loader.read_sovits_checkpoint(str(pth))
    assert sys.modules.get("utils") is before                 # the shim does not leak
    assert version == "v2" and hps["model"]["upsample_rates"] == model["upsample_rates"]
    assert hps["data"]["hop_length"] == 640 and isinstance(hps["model"], dict)
    assert torch.equal(sd["dec.conv_post.weight"], torch.ones(1, 4, 7))


def test_to_safetensors_round_trip(tmp_path):
    """TTS.to_safetensors (reference TTS.py:1482-1523): .ckpt / .pth -> directory the loaders read back; GPT tensors equal,
    SoVITS dec.* weight-norm folded (the reference saves the module after dec.remove_weight_norm()), flow.* untouched."""
    from gsv_tts.GPT_SoVITS.SoVITS.models_b200 import fold_weight_norm
    cfg = syn.GPT_CONFIG_TINY
    sd = syn.gpt_state_dict(cfg, 0)
    ckpt = tmp_path / "s1.ckpt"
    torch.save({"config": cfg, "weight": _upstream_gpt_names(sd, cfg["model"]["n_layer"])}, ckpt)
    out = Loader.to_safetensors(str(ckpt))
    assert out == str(tmp_path / "s1")
    config2, got2 = Loader.read_gpt_checkpoint(out)
    assert config2 == cfg and sorted(got2) == sorted(sd) and all(torch.equal(got2[k], sd[k]) for k in sd)

    model = dict(syn.SOVITS_MODEL["tiny"], version="v2Pro")
    s2 = dict(syn.sovits_flow_dec_state_dict(model, 0))
    s2.update(syn.sovits_encp_state_dict(model, 0))
    hps = {"data": {"filter_length": 2048, "hop_length": 640, "n_speakers": 300}, "train": {"segment_size": 20480}, "model": model}
    pth = tmp_path / "s2.pth"
    torch.save({"config": hps, "weight": s2}, pth)
    out2 = Loader.to_safetensors(str(pth), str(tmp_path / "conv"))
    hps2, got, version = Loader.read_sovits_checkpoint(out2)
    assert version == "v2Pro" and hps2["model"]["upsample_rates"] == list(model["upsample_rates"])
    assert not any(k.startswith("dec.") and k.endswith(("weight_g", "weight_v")) for k in got)
    for k, v in s2.items():
        if k.startswith("dec.") and k.endswith("weight_v"):
            base = k[: -len("weight_v")]
            assert torch.allclose(got[base + "weight"], fold_weight_norm(s2[base + "weight_g"], v), atol=1e-6)
        elif not (k.startswith("dec.") and k.endswith("weight_g")):
            assert torch.equal(got[k], v), k
    # and the native class loads the converted directory like the original file
    from gsv_tts.GPT_SoVITS.SoVITS.models_b200 import SynthesizerTrn
    a, b = SynthesizerTrn(1025, 32, n_speakers=300, **hps2["model"]), SynthesizerTrn(1025, 32, n_speakers=300, **model)
    a.load_state_dict(got)
    b.load_state_dict(s2)
    sa, sb = a.state_dict(), b.state_dict()
    assert sorted(sa) == sorted(sb) and all(torch.allclose(sa[k].float(), sb[k].float(), atol=1e-6) for k in sa)
    with pytest.raises(ValueError):
        Loader.to_safetensors(str(tmp_path / "x.bin"))


def test_native_sovits_class_survives_the_reference_loader_call_sequence():
    """Reference Loader.py:87-99 with only the class swapped: ctor keywords, load_state_dict(strict=False),
    dec.remove_weight_norm(), .to(device, dtype), .eval() (initialize_runtime needs the GPU and is covered there)."""
    from gsv_tts.GPT_SoVITS.SoVITS.models_b200 import SynthesizerTrn
    model = dict(syn.SOVITS_MODEL["tiny"], version="v2Pro")
    sd = dict(syn.sovits_flow_dec_state_dict(model, 0))
    sd.update(syn.sovits_encp_state_dict(model, 0))
    sd.update(syn.sovits_aux_state_dict(model, 0))
    vq = SynthesizerTrn(2048 // 2 + 1, 20480 // 640, n_speakers=300, **model)
    vq.load_state_dict(sd, strict=False)
    vq.dec.remove_weight_norm()
    assert vq.to("cpu", torch.float16) is vq and vq.eval() is vq
    assert vq.samples_per_frame == 640 and vq.enc_p.y_overlap is None
    assert set(k.split(".")[0] for k in vq.state_dict()) >= {"flow", "dec", "enc_p", "quantizer", "ref_enc", "ssl_proj", "sv_emb"}
