"""gsv_tts -- B200-native drop-in for the two hot paths of GSV-TTS-Lite.

Same import name and public surface as the reference package (reference
gsv_tts/__init__.py:1-11): ``TTS``, ``AudioClip``, ``cut_text``.  Submodules are imported
lazily so that ``import gsv_tts`` works without the optional frontend dependencies.
"""
__all__ = ["TTS", "AudioClip", "cut_text"]


def __getattr__(name):
    if name in ("TTS", "cut_text"):
        from . import TTS as _t
        return getattr(_t, name)
    if name == "AudioClip":
        from .Player import AudioClip
        return AudioClip
    raise AttributeError(name)
