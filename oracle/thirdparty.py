"""Restatement of the un-vendored third-party code the reference reaches on the hot path.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Neither pydub nor pyloudnorm is under ``/root/reference`` or installable in this
image (no network; ``requirements.txt:2,5`` leaves both unpinned).  What follows
restates the PUBLISHED behaviour of

* pydub 0.25.1  ``pydub/audio_segment.py`` (``AudioSegment``: ``frame_count``,
  ``__len__``, ms slicing with <=2 ms silence padding, ``get_sample_slice``,
  ``get_frame``, ``rms``, ``max_possible_amplitude``, ``_spawn``, ``overlay``,
  ``append``/``__add__``/``__radd__``) and ``pydub/effects.py``
  (``compress_dynamic_range``), ``pydub/utils.py`` (``ratio_to_db``/``db_to_float``);
* pyloudnorm 0.1.1  ``pyloudnorm/meter.py`` (``Meter.integrated_loudness``),
  ``pyloudnorm/iirfilter.py`` (``IIRfilter``), ``pyloudnorm/util.py`` (``valid_audio``)

only as far as ``worker/audio_mastering_engine.py`` uses them (call sites
ENG:43,54,63,80,82,89,126,207-210,213,218), on top of the REAL CPython ``audioop``
(``rms`` / ``mul`` / ``add``) and the REAL ``scipy.signal.lfilter``.  Parity at
these two boundaries is therefore "unpinned by the reference" (SURVEY.md 8c).
"""
from __future__ import annotations

import array
import math
import warnings

import numpy as np
import scipy.signal

with warnings.catch_warnings():
    warnings.simplefilter("ignore", DeprecationWarning)
    import audioop  # CPython <= 3.12 C module; pydub calls exactly this


# --------------------------------------------------------------------------- pydub.utils
def db_to_float(db, using_amplitude=True):
    """pydub/utils.py db_to_float: ``10 ** (db / 20)`` (C ``pow``)."""
    db = float(db)
    return 10 ** (db / 20) if using_amplitude else 10 ** (db / 10)


def ratio_to_db(ratio, val2=None, using_amplitude=True):
    """pydub/utils.py ratio_to_db: ``20 * math.log(ratio, 10)`` (= log(x)/log(10))."""
    ratio = float(ratio)
    if val2 is not None:
        ratio = ratio / val2
    if ratio == 0:
        return -float("inf")
    return (20 if using_amplitude else 10) * math.log(ratio, 10)


# --------------------------------------------------------------------------- pydub.AudioSegment
class TooManyMissingFrames(Exception):
    pass


class AudioSegment:
    """Raw-PCM subset of pydub.AudioSegment (no codecs: decode/encode are host I/O,
    outside the hot path)."""

    def __init__(self, data=b"", sample_width=2, frame_rate=44100, channels=2):
        self._data = bytes(data)
        self.sample_width = int(sample_width)
        self.frame_rate = int(frame_rate)
        self.channels = int(channels)
        self.frame_width = self.channels * self.sample_width

    # -- construction helpers ------------------------------------------------
    @classmethod
    def from_numpy(cls, pcm, frame_rate):
        """pcm: int16 array, shape (N,) mono or (N, C) interleaved."""
        pcm = np.ascontiguousarray(pcm, dtype=np.int16)
        ch = 1 if pcm.ndim == 1 else pcm.shape[1]
        return cls(pcm.tobytes(), 2, frame_rate, ch)

    def to_numpy(self):
        a = np.frombuffer(self._data, dtype=np.int16)
        return a.reshape(-1, self.channels) if self.channels > 1 else a.copy()

    def _spawn(self, data, overrides=None):
        if isinstance(data, list):
            data = b"".join(data)
        if isinstance(data, array.array):
            data = data.tobytes()
        if hasattr(data, "read"):
            if hasattr(data, "seek"):
                data.seek(0)
            data = data.read()
        md = dict(sample_width=self.sample_width, frame_rate=self.frame_rate, channels=self.channels)
        md.update(overrides or {})
        return self.__class__(data, **md)

    # -- geometry ------------------------------------------------------------
    def frame_count(self, ms=None):
        if ms is not None:
            return ms * (self.frame_rate / 1000.0)
        return float(len(self._data) // self.frame_width)

    def __len__(self):
        return round(1000 * (self.frame_count() / self.frame_rate))

    @property
    def array_type(self):
        return {1: "b", 2: "h", 4: "i"}[self.sample_width]

    def get_array_of_samples(self):
        return array.array(self.array_type, self._data)

    @property
    def max_possible_amplitude(self):
        return (2 ** (self.sample_width * 8)) / 2

    @property
    def rms(self):
        return audioop.rms(self._data, self.sample_width)

    def get_frame(self, index):
        s = index * self.frame_width
        return self._data[s:s + self.frame_width]

    def get_sample_slice(self, start_sample=None, end_sample=None):
        max_val = int(self.frame_count())

        def bounded(val, default):
            if val is None:
                return default
            if val < 0:
                return 0
            if val > max_val:
                return max_val
            return val

        s = bounded(start_sample, 0) * self.frame_width
        e = bounded(end_sample, max_val) * self.frame_width
        return self._spawn(self._data[s:e])

    def _parse_position(self, val):
        if val < 0:
            val = len(self) - abs(val)
        val = self.frame_count(ms=len(self)) if val == float("inf") else self.frame_count(ms=val)
        return int(val)

    def __getitem__(self, ms):
        if not isinstance(ms, slice):
            start, end = ms, ms + 1
        else:
            if ms.step:
                raise NotImplementedError("stepped slicing is not on the reference's path")
            start = ms.start if ms.start is not None else 0
            end = ms.stop if ms.stop is not None else len(self)
            start = min(start, len(self))
            end = min(end, len(self))
        start = self._parse_position(start) * self.frame_width
        end = self._parse_position(end) * self.frame_width
        data = self._data[start:end]
        expected = end - start
        missing = (expected - len(data)) // self.frame_width
        if missing:
            if missing > self.frame_count(ms=2):
                raise TooManyMissingFrames("more than 2 ms of frames missing")
            silence = audioop.mul(data[:self.frame_width], self.sample_width, 0)
            data += silence * missing
        return self._spawn(data)

    # -- combination ---------------------------------------------------------
    def append(self, seg, crossfade=100):
        if crossfade:
            raise NotImplementedError("crossfades are not on the reference's path")
        return self._spawn(self._data + seg._data)

    def __add__(self, arg):
        if isinstance(arg, AudioSegment):
            return self.append(arg, crossfade=0)
        raise NotImplementedError("gain via + is not on the reference's path")

    def __radd__(self, rarg):
        if rarg == 0:  # lets the builtin sum() work (ENG:80)
            return self
        raise TypeError("Gains must be the second addend after the AudioSegment")

    def overlay(self, seg, position=0):
        """position-0, play-once overlay = saturating ``audioop.add`` over the overlap."""
        out = [self[:position]._data]
        seg1 = self[position:]._data
        seg2 = seg._data
        remaining = max(0, len(seg1))
        if len(seg2) >= remaining:
            seg2 = seg2[:remaining]
        out.append(audioop.add(seg1[:len(seg2)], seg2, self.sample_width))
        out.append(seg1[len(seg2):])
        return self._spawn(b"".join(out))


# --------------------------------------------------------------------------- pydub.effects
def compress_dynamic_range(seg, threshold=-20.0, ratio=4.0, attack=5.0, release=50.0):
    """pydub/effects.py compress_dynamic_range, frame-by-frame, with the real audioop.

    This is the faithful (slow) loop: one ``audioop.rms`` over the look-back window
    and one ``audioop.mul`` per frame, exactly as pydub does.
    """
    thresh_rms = seg.max_possible_amplitude * db_to_float(threshold)
    look_frames = int(seg.frame_count(ms=attack))

    def rms_at(frame_i):
        return seg.get_sample_slice(frame_i - look_frames, frame_i).rms

    def db_over_threshold(rms):
        if rms == 0:
            return 0.0
        db = ratio_to_db(rms / thresh_rms)
        return max(db, 0)

    output = []
    attenuation = 0.0
    attack_frames = seg.frame_count(ms=attack)
    release_frames = seg.frame_count(ms=release)
    for i in range(int(seg.frame_count())):
        rms_now = rms_at(i)
        max_attenuation = (1 - (1.0 / ratio)) * db_over_threshold(rms_now)
        attenuation_inc = max_attenuation / attack_frames
        attenuation_dec = max_attenuation / release_frames
        if rms_now > thresh_rms and attenuation <= max_attenuation:
            attenuation += attenuation_inc
            attenuation = min(attenuation, max_attenuation)
        else:
            attenuation -= attenuation_dec
            attenuation = max(attenuation, 0)
        frame = seg.get_frame(i)
        if attenuation != 0.0:
            frame = audioop.mul(frame, seg.sample_width, db_to_float(-attenuation))
        output.append(frame)
    return seg._spawn(data=b"".join(output))


# --------------------------------------------------------------------------- pyloudnorm
class IIRfilter:
    """pyloudnorm/iirfilter.py: RBJ-cookbook biquad, applied with scipy lfilter."""

    def __init__(self, G, Q, fc, rate, filter_type, passband_gain=1.0):
        self.G, self.Q, self.fc, self.rate = G, Q, fc, rate
        self.filter_type = filter_type
        self.passband_gain = passband_gain
        self.b, self.a = self.generate_coefficients()

    def generate_coefficients(self):
        A = 10 ** (self.G / 40.0)
        w0 = 2.0 * np.pi * (self.fc / self.rate)
        alpha = np.sin(w0) / (2.0 * self.Q)
        if self.filter_type == "high_shelf":
            b0 = A * ((A + 1) + (A - 1) * np.cos(w0) + 2 * np.sqrt(A) * alpha)
            b1 = -2 * A * ((A - 1) + (A + 1) * np.cos(w0))
            b2 = A * ((A + 1) + (A - 1) * np.cos(w0) - 2 * np.sqrt(A) * alpha)
            a0 = (A + 1) - (A - 1) * np.cos(w0) + 2 * np.sqrt(A) * alpha
            a1 = 2 * ((A - 1) - (A + 1) * np.cos(w0))
            a2 = (A + 1) - (A - 1) * np.cos(w0) - 2 * np.sqrt(A) * alpha
        elif self.filter_type == "high_pass":
            b0 = (1 + np.cos(w0)) / 2
            b1 = -(1 + np.cos(w0))
            b2 = (1 + np.cos(w0)) / 2
            a0 = 1 + alpha
            a1 = -2 * np.cos(w0)
            a2 = 1 - alpha
        else:
            raise ValueError("only the K-weighting filter types are on the reference's path")
        return np.array([b0, b1, b2]) / a0, np.array([a0, a1, a2]) / a0

    def apply_filter(self, data):
        return self.passband_gain * scipy.signal.lfilter(self.b, self.a, data)


class Meter:
    """pyloudnorm/meter.py Meter with the default "K-weighting" filter class."""

    last_loudness = None  # test hook: the most recent integrated_loudness() result

    def __init__(self, rate, filter_class="K-weighting", block_size=0.400):
        if filter_class != "K-weighting":
            raise ValueError("only K-weighting is on the reference's path (ENG:213)")
        self.rate = rate
        self.block_size = block_size
        self._filters = {
            "high_shelf": IIRfilter(4.0, 1 / np.sqrt(2), 1500.0, rate, "high_shelf"),
            "high_pass": IIRfilter(0.0, 0.5, 38.0, rate, "high_pass"),
        }

    def integrated_loudness(self, data):
        input_data = data.copy()
        # pyloudnorm/util.py valid_audio
        if not isinstance(input_data, np.ndarray):
            raise ValueError("Data must be of type numpy.ndarray.")
        if not np.issubdtype(input_data.dtype, np.floating):
            raise ValueError("Data must be floating point.")
        if input_data.ndim == 2 and input_data.shape[1] > 5:
            raise ValueError("Audio must have five channels or less.")
        if input_data.shape[0] < self.block_size * self.rate:
            raise ValueError("Audio must have length greater than the block size.")

        if input_data.ndim == 1:
            input_data = np.reshape(input_data, (input_data.shape[0], 1))
        numChannels = input_data.shape[1]
        numSamples = input_data.shape[0]

        for _name, stage in self._filters.items():
            for ch in range(numChannels):
                input_data[:, ch] = stage.apply_filter(input_data[:, ch])

        G = [1.0, 1.0, 1.0, 1.41, 1.41]
        T_g = self.block_size
        Gamma_a = -70.0
        overlap = 0.75
        step = 1.0 - overlap

        T = numSamples / self.rate
        numBlocks = int(np.round(((T - T_g) / (T_g * step))) + 1)
        j_range = np.arange(0, numBlocks)
        z = np.zeros(shape=(numChannels, numBlocks))
        for i in range(numChannels):
            for j in j_range:
                lo = int(T_g * (j * step) * self.rate)
                hi = int(T_g * (j * step + 1) * self.rate)
                z[i, j] = (1.0 / (T_g * self.rate)) * np.sum(np.square(input_data[lo:hi, i]))

        with warnings.catch_warnings():
            warnings.simplefilter("ignore", category=RuntimeWarning)
            with np.errstate(divide="ignore", invalid="ignore"):
                l = [-0.691 + 10.0 * np.log10(np.sum([G[i] * z[i, j] for i in range(numChannels)])) for j in j_range]
                J_g = [j for j, l_j in enumerate(l) if l_j >= Gamma_a]
                z_avg_gated = [np.mean([z[i, j] for j in J_g]) for i in range(numChannels)]
                Gamma_r = -0.691 + 10.0 * np.log10(np.sum([G[i] * z_avg_gated[i] for i in range(numChannels)])) - 10.0
                J_g = [j for j, l_j in enumerate(l) if (l_j > Gamma_r and l_j > Gamma_a)]
                z_avg_gated = np.nan_to_num(np.array([np.mean([z[i, j] for j in J_g]) for i in range(numChannels)]))
                LUFS = -0.691 + 10.0 * np.log10(np.sum([G[i] * z_avg_gated[i] for i in range(numChannels)]))
        Meter.last_loudness = float(LUFS)
        return LUFS
