"""``settings`` dict -> ``b200m_plan`` (host-side filter design).

Design is O(1) per job and stays on the host.  The coefficient
expressions are evaluated with numpy / scipy in the reference's own operation order
(ENG = /root/reference/worker/audio_mastering_engine.py) so they round identically:
shelves ENG:170-182, peaks ENG:185-193, crossover ENG:197-198 (``scipy.signal.butter``),
compressor parameters as pydub derives them (ENG:207-209), K-weighting as pyloudnorm
derives it (ENG:213).  ``libb200master`` carries an independent C implementation of the
same design (``b200m_plan_from_settings``) for non-Python hosts.
"""
from __future__ import annotations

import hashlib
import math
import threading

import numpy as np
from scipy.signal import butter

from . import lib as L

# ENG:128-134 sees only x = int16 / 32768, so the exciter is a pure function of the 65536 possible samples.
# numpy's float32 tanh is a SIMD polynomial (neither correctly rounded nor reproducible by a device routine),
# so the table is made HERE, by the very numpy expression the reference evaluates, and the kernels gather from it
# (b200m_plan.sat_lut).  Tables are kept alive for the life of the process: the plan holds a raw pointer.
_LUT_LOCK = threading.Lock()
_LUT_CACHE: dict = {}          # saturation value -> (float32[65536] indexed by the uint16 bit pattern, content key)
_LUT_MAX = 64


def exciter_table(saturation) -> np.ndarray:
    """(1 - mix) * x + mix * tanh(x * (1 + 4 mix)) in float32 for every int16 sample, indexed by ``uint16(sample)``;
    the statements of ENG:131-134 on the array ENG:121 produces."""
    s16 = np.arange(65536, dtype=np.uint16).view(np.int16)
    x = s16.astype(np.float32) / (2 ** 15)                    # ENG:121
    mix = (saturation / 100.0) ** 2                            # ENG:131
    saturated = np.tanh(x * (1 + mix * 4))                     # ENG:133
    return np.ascontiguousarray((1 - mix) * x + mix * saturated, dtype=np.float32)   # ENG:134


def _lut_key(table: np.ndarray) -> int:
    return int.from_bytes(hashlib.blake2b(table.tobytes(), digest_size=8).digest(), "little") or 1


def install_exciter_table(saturation, table) -> None:
    """Use ``table`` for plans with this ``saturation`` from now on (``None`` forgets it).  Lets a test replay
    the table of the host that wrote a golden fixture; production code never calls this."""
    with _LUT_LOCK:
        if table is None:
            _LUT_CACHE.pop(saturation, None)
            return
        t = np.ascontiguousarray(table, dtype=np.float32)
        if t.shape != (65536,):
            raise ValueError("an exciter table has 65536 float32 entries")
        _LUT_CACHE[saturation] = (t, _lut_key(t))


def _exciter_lut(saturation):
    with _LUT_LOCK:
        hit = _LUT_CACHE.get(saturation)
        if hit is None:
            if len(_LUT_CACHE) >= _LUT_MAX:
                _LUT_CACHE.pop(next(iter(_LUT_CACHE)))        # plans made earlier keep their own reference (Plan._lut)
            t = exciter_table(saturation)
            hit = _LUT_CACHE[saturation] = (t, _lut_key(t))
        return hit

# ENG:207-209: (attack_ms, release_ms) of the low / mid / high compressor
BAND_TIMES = ((10.0, 200.0), (5.0, 150.0), (1.0, 50.0))
# ENG:67-72 defaults; the GUI sends `*_band_threshold` / `*_band_ratio` instead
# (mastering_gui.py:187-189) -- both spellings are accepted.
BAND_KEYS = (("low_thresh", "low_band_threshold", -25.0, "low_ratio", "low_band_ratio", 6.0),
             ("mid_thresh", "mid_band_threshold", -20.0, "mid_ratio", "mid_band_ratio", 3.0),
             ("high_thresh", "high_band_threshold", -15.0, "high_ratio", "high_band_ratio", 4.0))


def normalize_settings(settings: dict) -> dict:
    """Resolve defaults and aliases exactly where the reference reads them."""
    s = dict(settings or {})
    out = {
        "saturation": s.get("saturation", 0),                # ENG:58
        "bass_boost": s.get("bass_boost", 0.0),              # ENG:147
        "mid_cut": s.get("mid_cut", 0.0),                    # ENG:148
        "presence_boost": s.get("presence_boost", 0.0),      # ENG:149
        "treble_boost": s.get("treble_boost", 0.0),          # ENG:150
        "width": s.get("width", 1.0),                        # ENG:60
        "multiband": bool(s.get("multiband")),               # ENG:65
        "lufs": s.get("lufs"),                               # ENG:84
    }
    for tk, tk_gui, td, rk, rk_gui, rd in BAND_KEYS:
        out[tk] = s.get(tk, s.get(tk_gui, td))
        out[rk] = s.get(rk, s.get(rk_gui, rd))
    return out


def _biquad(b, a):
    return L.Biquad(b[0] / a[0], b[1] / a[0], b[2] / a[0], a[1] / a[0], a[2] / a[0])


def shelf_biquad(rate, cutoff_hz, gain_db, kind, q=0.707):
    """ENG:170-182.  None when bypassed (gain_db == 0, ENG:171)."""
    if gain_db == 0:
        return None
    wn = cutoff_hz / (0.5 * rate)
    g = 10.0 ** (gain_db / 20.0)
    w = wn * 2 * np.pi
    alpha = np.sin(w) / (2.0 * q)
    c, rt = np.cos(w), 2 * np.sqrt(g) * alpha
    if kind == "low":
        b = (g * ((g + 1) - (g - 1) * c + rt), 2 * g * ((g - 1) - (g + 1) * c), g * ((g + 1) - (g - 1) * c - rt))
        a = ((g + 1) + (g - 1) * c + rt, -2 * ((g - 1) + (g + 1) * c), (g + 1) + (g - 1) * c - rt)
    else:
        b = (g * ((g + 1) + (g - 1) * c + rt), -2 * g * ((g - 1) + (g + 1) * c), g * ((g + 1) + (g - 1) * c - rt))
        a = ((g + 1) - (g - 1) * c + rt, 2 * ((g - 1) - (g + 1) * c), (g + 1) - (g - 1) * c - rt)
    return _biquad(b, a)


def peak_biquad(rate, center_hz, gain_db, q=1.0):
    """ENG:185-193.  None when bypassed (ENG:186)."""
    if gain_db == 0:
        return None
    wn = center_hz / (0.5 * rate)
    g = 10.0 ** (gain_db / 20.0)
    w = wn * 2 * np.pi
    alpha = np.sin(w) / (2.0 * q)
    b = (1 + alpha * g, -2 * np.cos(w), 1 - alpha * g)
    a = (1 + alpha / g, -2 * np.cos(w), 1 - alpha / g)
    return _biquad(b, a)


def butter4(rate, freq, btype):
    """ENG:197-198 ``butter(4, f, btype, fs=rate, output='sos')`` -> two Biquads."""
    sos = butter(4, freq, btype=btype, fs=rate, output="sos")
    return [L.Biquad(r[0], r[1], r[2], r[4], r[5]) for r in sos]


def kweight_biquads(rate):
    """pyloudnorm Meter(rate): high shelf (+4 dB, Q 1/sqrt2, 1500 Hz) then high pass
    (Q 0.5, 38 Hz), RBJ forms with A = 10^(G/40) (pyloudnorm/iirfilter.py)."""
    G, Q, fc = 4.0, 1 / np.sqrt(2), 1500.0
    A = 10 ** (G / 40.0)
    w0 = 2.0 * np.pi * (fc / rate)
    alpha = np.sin(w0) / (2.0 * Q)
    b0 = A * ((A + 1) + (A - 1) * np.cos(w0) + 2 * np.sqrt(A) * alpha)
    b1 = -2 * A * ((A - 1) + (A + 1) * np.cos(w0))
    b2 = A * ((A + 1) + (A - 1) * np.cos(w0) - 2 * np.sqrt(A) * alpha)
    a0 = (A + 1) - (A - 1) * np.cos(w0) + 2 * np.sqrt(A) * alpha
    a1 = 2 * ((A - 1) - (A + 1) * np.cos(w0))
    a2 = (A + 1) - (A - 1) * np.cos(w0) - 2 * np.sqrt(A) * alpha
    shelf = _biquad((b0, b1, b2), (a0, a1, a2))
    Q, fc = 0.5, 38.0
    w0 = 2.0 * np.pi * (fc / rate)
    alpha = np.sin(w0) / (2.0 * Q)
    hp = _biquad(((1 + np.cos(w0)) / 2, -(1 + np.cos(w0)), (1 + np.cos(w0)) / 2),
                 (1 + alpha, -2 * np.cos(w0), 1 - alpha))
    return [shelf, hp]


def make_band(rate, threshold_db, ratio, attack_ms, release_ms) -> L.Band:
    """pydub compress_dynamic_range set-up (effects.py): thresh_rms, look/attack/release
    frame counts (``ms * (rate / 1000.0)``), slope ``1 - 1/ratio``."""
    if ratio == 0:
        raise ZeroDivisionError("float division by zero")      # pydub: 1.0 / ratio
    thresh_rms = 32768.0 * (10 ** (float(threshold_db) / 20))
    attack_frames = attack_ms * (rate / 1000.0)
    release_frames = release_ms * (rate / 1000.0)
    return L.Band(thresh_rms, attack_frames, release_frames, 1 - (1.0 / ratio), int(attack_frames), 0)


def make_plan(settings: dict, rate: int, channels: int, low_crossover=250, high_crossover=4000) -> L.Plan:
    s = normalize_settings(settings)
    if channels not in (1, 2):
        raise ValueError("only mono and stereo PCM are supported (the reference reshapes to (-1, 2), ENG:119-120)")
    p = L.Plan()
    p.sample_rate, p.channels = int(rate), int(channels)
    sat = s["saturation"]
    p.sat_on = int(sat != 0)
    mix = (sat / 100.0) ** 2                                    # ENG:131
    p.sat_clean, p.sat_mix, p.sat_drive = np.float32(1 - mix), np.float32(mix), np.float32(1 + mix * 4)
    if p.sat_on:
        table, key = _exciter_lut(sat)
        p._lut = table                                          # keeps the table alive as long as the plan
        p.sat_lut, p.sat_lut_key = table.ctypes.data, key
    secs = [shelf_biquad(rate, 250, s["bass_boost"], "low"),    # ENG:154-161, in order
            peak_biquad(rate, 1000, -s["mid_cut"]),
            peak_biquad(rate, 4000, s["presence_boost"]),
            shelf_biquad(rate, 8000, s["treble_boost"], "high")]
    secs = [q for q in secs if q is not None]
    p.n_eq = len(secs)
    for i, q in enumerate(secs):
        p.eq[i] = q
    p.width = float(s["width"])
    p.width_on = int(channels == 2 and s["width"] != 1.0)
    p.multiband = int(s["multiband"])
    if p.multiband:
        for i, q in enumerate(butter4(rate, low_crossover, "lowpass")):
            p.lp[i] = q
        for i, q in enumerate(butter4(rate, high_crossover, "highpass")):
            p.hp[i] = q
        for i, ((tk, _, _, rk, _, _), (att, rel)) in enumerate(zip(BAND_KEYS, BAND_TIMES)):
            p.band[i] = make_band(rate, s[tk], s[rk], att, rel)
    for i, q in enumerate(kweight_biquads(rate)):
        p.kw[i] = q
    p.has_lufs = int(s["lufs"] is not None)
    p.lufs = float(s["lufs"]) if s["lufs"] is not None else math.nan
    return p


def to_c_settings(settings: dict) -> L.Settings:
    s = normalize_settings(settings)
    c = L.Settings()
    for k in ("saturation", "bass_boost", "mid_cut", "presence_boost", "treble_boost", "width",
              "low_thresh", "low_ratio", "mid_thresh", "mid_ratio", "high_thresh", "high_ratio"):
        setattr(c, k, float(s[k]))
    c.multiband = int(s["multiband"])
    c.has_lufs = int(s["lufs"] is not None)
    c.lufs = float(s["lufs"]) if s["lufs"] is not None else 0.0
    return c
