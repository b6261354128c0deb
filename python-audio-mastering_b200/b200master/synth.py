"""Deterministic synthetic programme material.

The reference ships no audio (``.gitignore:23-26`` excludes every audio format), so
every test, fixture and benchmark uses these generated tracks: pink noise (70 %
common / 30 % independent between L and R), five sines, a 60 Hz kick + click every
0.5 s and a 6-12 kHz "hi-hat" burst on the off-beats, peak-normalised to -1 dBFS and
rounded to int16.  Seed = ``0xB200 + track_index``.

``make_track`` is the host (numpy) generator used for fixtures and parity tests;
``make_tracks_torch`` builds the same recipe on a torch device for the big
benchmark batches (different RNG stream, same statistics).
"""
from __future__ import annotations

import numpy as np

SEED_BASE = 0xB200
SINES_HZ = (55.0, 220.0, 997.0, 3500.0, 9000.0)


def _db(x):
    return 10.0 ** (x / 20.0)


# hi-hat pattern: (period s, decay tau s, level dBFS, hold s): full level for `hold`, then an exponential decay.
# HAT_SPARSE is the recipe the golden fixtures were written with (one short burst per off-beat: the high band is
# above its -15 dB RMS threshold ~1.5 % of the time); HAT_DENSE is the benchmark workload's (loud 50 ms bursts on
# every eighth note: every compressor band is above its threshold for >= 10 % of the frames, SURVEY 8d --
# bench.py measures and asserts it).
HAT_SPARSE = (0.5, 0.015, -10.0, 0.0)
HAT_DENSE = (0.25, 0.020, 9.0, 0.050)


def make_track(index: int, seconds: float, rate: int, channels: int = 2, hat_cfg=HAT_SPARSE) -> np.ndarray:
    """int16 array of shape (N, 2) (or (N,) for mono), N = round(seconds * rate)."""
    n = int(round(seconds * rate))
    rng = np.random.Generator(np.random.PCG64(SEED_BASE + index))
    t = np.arange(n, dtype=np.float64) / rate
    nfreq = n // 2 + 1
    f = np.fft.rfftfreq(n, 1.0 / rate)

    def pink():
        spec = np.fft.rfft(rng.standard_normal(n))
        shape = np.ones(nfreq)
        shape[1:] = 1.0 / np.sqrt(f[1:])
        shape[0] = 0.0
        y = np.fft.irfft(spec * shape, n)
        return y / np.sqrt(np.mean(y * y))

    common, ind_l, ind_r = pink(), pink(), pink()
    noise_l = 0.7 * common + 0.3 * ind_l
    noise_r = 0.7 * common + 0.3 * ind_r
    noise_l *= _db(-20.0) / np.sqrt(np.mean(noise_l ** 2))
    noise_r *= _db(-20.0) / np.sqrt(np.mean(noise_r ** 2))

    sines_l = np.zeros(n)
    sines_r = np.zeros(n)
    for hz in SINES_HZ:
        sines_l += _db(-18.0) * np.sin(2 * np.pi * hz * t)
        sines_r += _db(-18.0) * np.sin(2 * np.pi * hz * t + np.pi / 5)

    beat = np.mod(t, 0.5)
    kick = _db(-3.0) * np.exp(-beat / 0.040) * np.sin(2 * np.pi * 60.0 * beat)
    click = np.zeros(n)
    click[(np.arange(0, n, int(round(0.5 * rate))))] = _db(-3.0)

    hat_spec = np.fft.rfft(rng.standard_normal(n))
    hat_spec[(f < 6000.0) | (f > 12000.0)] = 0.0
    hat = np.fft.irfft(hat_spec, n)
    hat /= max(np.max(np.abs(hat)), 1e-12)
    off = t - 0.5 * hat_cfg[0]
    hat_env = np.where(off >= 0, np.exp(-np.maximum(np.mod(off, hat_cfg[0]) - hat_cfg[3], 0.0) / hat_cfg[1]), 0.0)
    hat = _db(hat_cfg[2]) * hat * hat_env

    left = noise_l + sines_l + kick + click + hat
    right = noise_r + sines_r + kick + click + 0.8 * hat
    peak = max(np.max(np.abs(left)), np.max(np.abs(right)))
    scale = _db(-1.0) / peak * 32767.0
    pcm = np.stack([np.rint(left * scale), np.rint(right * scale)], axis=1).astype(np.int16)
    if channels == 1:
        return np.ascontiguousarray(pcm[:, 0])
    return pcm


def make_tracks_torch(first_index: int, n_tracks: int, seconds: float, rate: int, device, hat_cfg=HAT_SPARSE):
    """(n_tracks, N, 2) int16 tensor on ``device`` built with torch ops (benchmark input
    only -- generating 10^2..10^3 three-minute tracks with numpy would take minutes)."""
    import torch

    n = int(round(seconds * rate))
    out = torch.empty((n_tracks, n, 2), dtype=torch.int16, device=device)
    t = torch.arange(n, dtype=torch.float32, device=device) / rate
    f = torch.fft.rfftfreq(n, 1.0 / rate).to(device)
    shape = torch.zeros_like(f)
    shape[1:] = torch.rsqrt(f[1:])
    band = ((f >= 6000.0) & (f <= 12000.0)).to(torch.float32)
    beat = torch.remainder(t, 0.5)
    kick = _db(-3.0) * torch.exp(-beat / 0.040) * torch.sin(2 * np.pi * 60.0 * beat)
    click = torch.zeros(n, device=device)
    click[torch.arange(0, n, int(round(0.5 * rate)), device=device)] = _db(-3.0)
    off = t - 0.5 * hat_cfg[0]
    hat_env = torch.where(off >= 0, torch.exp(-torch.clamp_min(torch.remainder(off, hat_cfg[0]) - hat_cfg[3], 0.0) / hat_cfg[1]), torch.zeros_like(off))
    t64 = torch.arange(n, dtype=torch.float64, device=device) / rate
    sines_l = torch.zeros(n, device=device)
    sines_r = torch.zeros(n, device=device)
    for hz in SINES_HZ:
        sines_l += (_db(-18.0) * torch.sin(2 * np.pi * hz * t64)).float()
        sines_r += (_db(-18.0) * torch.sin(2 * np.pi * hz * t64 + np.pi / 5)).float()
    for k in range(n_tracks):
        g = torch.Generator(device=device)
        g.manual_seed(SEED_BASE + first_index + k)

        def pink():
            y = torch.fft.irfft(torch.fft.rfft(torch.randn(n, generator=g, device=device)) * shape, n)
            return y / y.square().mean().sqrt()

        common, il, ir = pink(), pink(), pink()
        nl = 0.7 * common + 0.3 * il
        nr = 0.7 * common + 0.3 * ir
        nl = nl * (_db(-20.0) / nl.square().mean().sqrt())
        nr = nr * (_db(-20.0) / nr.square().mean().sqrt())
        hat = torch.fft.irfft(torch.fft.rfft(torch.randn(n, generator=g, device=device)) * band, n)
        hat = _db(hat_cfg[2]) * hat / hat.abs().max().clamp_min(1e-12) * hat_env
        left = nl + sines_l + kick + click + hat
        right = nr + sines_r + kick + click + 0.8 * hat
        peak = torch.maximum(left.abs().max(), right.abs().max())
        scale = _db(-1.0) / peak * 32767.0
        out[k, :, 0] = torch.round(left * scale).to(torch.int16)
        out[k, :, 1] = torch.round(right * scale).to(torch.int16)
    return out
