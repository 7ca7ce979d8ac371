"""``AudioClip`` -- the return type of every ``TTS.infer*`` call (reference gsv_tts/Player.py:70-99).

Only the value type is kept: live playback (``AudioQueue``, sounddevice thread, Player.py:13-67) is a
caller-side concern and out of scope."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np


@dataclass
class AudioClip:
    audio_data: np.ndarray                 # float32 mono, peak-normalised by the caller
    samplerate: int = 32000
    audio_len_s: float = 0.0
    subtitles: Optional[List[dict]] = None
    orig_text: str = ""
    _queue: object = field(default=None, repr=False)

    def __post_init__(self):
        self.audio_data = np.asarray(self.audio_data, dtype=np.float32)
        if not self.audio_len_s:
            self.audio_len_s = float(self.audio_data.shape[-1]) / float(self.samplerate)

    def save(self, path: str) -> None:
        """16-bit PCM WAV with the standard library (the reference uses soundfile, Player.py:92-99)."""
        import wave
        pcm = np.clip(self.audio_data, -1.0, 1.0)
        pcm = (pcm * 32767.0).astype("<i2")
        with wave.open(path, "wb") as w:
            w.setnchannels(1)
            w.setsampwidth(2)
            w.setframerate(int(self.samplerate))
            w.writeframes(pcm.tobytes())

    def play(self):
        raise NotImplementedError("live playback is out of scope of the B200 hot-path package (SURVEY.md 2, row 10)")
