"""``PcmSegment``: the handful of ``pydub.AudioSegment`` members the mastering chain
touches (ENG:43,54,63,80,89,96-99,117-126), for hosts without pydub/FFmpeg.

Decode / encode are host I/O outside the hot path.  When pydub is importable it is used
for every container format; otherwise only RIFF/WAV (stdlib ``wave``) is available.
"""
from __future__ import annotations

import array
import io
import wave


class PcmSegment:
    def __init__(self, data=b"", sample_width=2, frame_rate=44100, channels=2):
        self._data = bytes(data)
        self.sample_width = int(sample_width)
        self.frame_rate = int(frame_rate)
        self.channels = int(channels)
        self.frame_width = self.sample_width * self.channels

    # pydub surface ----------------------------------------------------------------
    def _spawn(self, data, overrides=None):
        if isinstance(data, list):
            data = b"".join(data)
        if isinstance(data, array.array):
            data = data.tobytes()
        if hasattr(data, "read"):
            data = data.read()
        kw = dict(sample_width=self.sample_width, frame_rate=self.frame_rate, channels=self.channels)
        kw.update(overrides or {})
        return PcmSegment(data, **kw)

    def get_array_of_samples(self):
        return array.array({1: "b", 2: "h", 4: "i"}[self.sample_width], self._data)

    def frame_count(self, ms=None):
        if ms is not None:
            return ms * (self.frame_rate / 1000.0)
        return float(len(self._data) // self.frame_width)

    def __len__(self):
        return round(1000 * (self.frame_count() / self.frame_rate))

    @property
    def raw_data(self):
        return self._data

    @classmethod
    def from_file(cls, f, format=None):
        with wave.open(f, "rb") as w:
            return cls(w.readframes(w.getnframes()), w.getsampwidth(), w.getframerate(), w.getnchannels())

    from_wav = from_file

    def export(self, out_f=None, format="wav", **_kw):
        if format != "wav":
            raise ValueError("PcmSegment writes WAV only; install pydub + FFmpeg for other formats")
        close = False
        if out_f is None:
            out_f = io.BytesIO()
        elif isinstance(out_f, (str, bytes)) or hasattr(out_f, "__fspath__"):
            out_f, close = open(out_f, "wb"), True
        with wave.open(out_f, "wb") as w:
            w.setnchannels(self.channels)
            w.setsampwidth(self.sample_width)
            w.setframerate(self.frame_rate)
            w.writeframesraw(self._data)
        if close:
            out_f.close()
        elif hasattr(out_f, "seek"):
            out_f.seek(0)
        return out_f


def segment_class():
    """pydub.AudioSegment when installed (any container via FFmpeg), else PcmSegment."""
    try:
        from pydub import AudioSegment  # type: ignore
        return AudioSegment
    except Exception:
        return PcmSegment
