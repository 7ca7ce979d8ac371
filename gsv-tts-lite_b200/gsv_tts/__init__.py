"""gsv_tts -- B200-native drop-in for the two hot paths of GSV-TTS-Lite.

Same import name and public surface as the reference package (reference
gsv_tts/__init__.py:1-11): ``TTS``, ``AudioClip``, ``cut_text``.  Submodules are imported
lazily so that ``import gsv_tts`` works without the optional frontend dependencies.
"""
__all__ = ["TTS", "AudioClip", "cut_text"]


def __getattr__(name):
    # `from . import TTS` would re-enter this hook (the submodule and the class share the name)
    import importlib
    if name in ("TTS", "cut_text"):
        mod = importlib.import_module(__name__ + ".TTS")
        value = getattr(mod, name)
        globals()[name] = value
        return value
    if name == "AudioClip":
        value = importlib.import_module(__name__ + ".Player").AudioClip
        globals()[name] = value
        return value
    raise AttributeError(name)
