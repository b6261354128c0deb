"""CPU tests of the host side: the C-ABI library loads and exports every declared symbol,
the host-side filter design matches the oracle's, framing and settings logic."""
import ctypes as C
import os
import re

import numpy as np
import pytest
from scipy.signal import butter

from conftest import ROOT
from b200master import lib as L
from b200master import ms_framing, normalize_settings
from b200master.plan import make_plan, to_c_settings
from oracle import port


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "b200_master.h")).read()
    return sorted(set(re.findall(r"\b(b200m_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    lib = L.load()
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/b200_master.h but not exported"
    assert set(names) == set(L.EXPORTS)
    assert lib.b200m_abi_version() == L.ABI_VERSION == 2


def test_no_cpu_fallback_without_device(has_cuda):
    if has_cuda:
        pytest.skip("a CUDA device is present")
    lib = L.load()
    h = C.c_void_p()
    assert lib.b200m_create(0, C.byref(h)) == L.ERR_CUDA
    assert b"no CPU fallback" in lib.b200m_last_error(None)
    from b200master import Engine
    with pytest.raises(RuntimeError):
        Engine(0)


def _bq(b):
    return np.array([b.b0, b.b1, b.b2, b.a1, b.a2])


@pytest.mark.parametrize("rate", [44100, 48000, 96000])
def test_python_plan_matches_oracle_design(rate):
    st = dict(bass_boost=4.0, mid_cut=3.0, presence_boost=1.0, treble_boost=3.0, saturation=25, width=1.2,
              multiband=True, lufs=-14.0)
    p = make_plan(st, rate, 2)
    secs = [s for s in port.eq_sections(rate, st) if s is not None]
    assert p.n_eq == 4
    for i, s in enumerate(secs):
        assert np.array_equal(_bq(p.eq[i]), s[0, [0, 1, 2, 4, 5]])
    lp = butter(4, 250, btype="lowpass", fs=rate, output="sos")
    hp = butter(4, 4000, btype="highpass", fs=rate, output="sos")
    for i in range(2):
        assert np.array_equal(_bq(p.lp[i]), lp[i, [0, 1, 2, 4, 5]])
        assert np.array_equal(_bq(p.hp[i]), hp[i, [0, 1, 2, 4, 5]])
    from oracle import thirdparty
    m = thirdparty.Meter(rate)
    for i, f in enumerate(m._filters.values()):
        assert np.array_equal(_bq(p.kw[i]), np.r_[f.b, f.a[1:]])
    mix = (25 / 100.0) ** 2
    assert p.sat_clean == np.float32(1 - mix) and p.sat_mix == np.float32(mix) and p.sat_drive == np.float32(1 + mix * 4)
    # pydub compressor set-up
    assert p.band[0].thresh_rms == 32768.0 * (10 ** (-25.0 / 20))
    assert [p.band[i].look_frames for i in range(3)] == [int(10 * rate / 1000.0), int(5 * (rate / 1000.0)), int(1 * (rate / 1000.0))]


@pytest.mark.parametrize("rate", [44100, 48000, 96000])
def test_c_design_matches_python_design(rate):
    """b200m_plan_from_settings (C, libm) vs the numpy/scipy design the reference uses."""
    st = dict(bass_boost=5.0, mid_cut=-2.0, presence_boost=2.0, treble_boost=3.5, saturation=40, width=0.7,
              multiband=True, lufs=-9.0, mid_thresh=-22.0, mid_ratio=2.5)
    p = make_plan(st, rate, 2)
    q = L.Plan()
    s = to_c_settings(st)
    assert L.load().b200m_plan_from_settings(C.byref(s), rate, 2, C.byref(q)) == 0
    assert (p.n_eq, p.width_on, p.multiband, p.has_lufs, p.sat_on) == (q.n_eq, q.width_on, q.multiband, q.has_lufs, q.sat_on)
    for a, b in [(p.eq[i], q.eq[i]) for i in range(4)] + [(p.lp[i], q.lp[i]) for i in range(2)] + \
                [(p.hp[i], q.hp[i]) for i in range(2)] + [(p.kw[i], q.kw[i]) for i in range(2)]:
        assert np.allclose(_bq(a), _bq(b), rtol=1e-13, atol=0)
    for i in range(3):
        assert p.band[i].thresh_rms == pytest.approx(q.band[i].thresh_rms, rel=1e-15)
        assert (p.band[i].look_frames, p.band[i].attack_frames, p.band[i].release_frames, p.band[i].slope) == \
               (q.band[i].look_frames, q.band[i].attack_frames, q.band[i].release_frames, q.band[i].slope)
    assert (p.sat_clean, p.sat_mix, p.sat_drive, p.lufs, p.width) == (q.sat_clean, q.sat_mix, q.sat_drive, q.lufs, q.width)


def test_bypassed_sections_are_dropped():
    p = make_plan(dict(bass_boost=0, mid_cut=0.0, presence_boost=3.5, treble_boost=0), 44100, 2)
    assert p.n_eq == 1 and p.width_on == 0 and p.multiband == 0 and p.has_lufs == 0 and p.sat_on == 0
    assert make_plan(dict(width=1.3), 44100, 1).width_on == 0          # mono: widener is a no-op (ENG:137)


def test_settings_defaults_and_gui_aliases():
    s = normalize_settings({"low_band_threshold": -30.0, "mid_ratio": 2.0, "mid_band_ratio": 9.0, "compress": False,
                            "original_filename": "x.wav"})
    assert s["low_thresh"] == -30.0 and s["low_ratio"] == 6.0 and s["mid_ratio"] == 2.0
    assert s["high_thresh"] == -15.0 and s["high_ratio"] == 4.0 and s["lufs"] is None and s["multiband"] is False
    assert normalize_settings({"multiband": 1})["multiband"] is True


@pytest.mark.parametrize("n,rate", [(44077, 44100), (44079, 44100), (7938021, 44100), (7938023, 44100),
                                    (8640000, 48000), (341775, 11025), (1, 44100)])
def test_ms_framing_matches_oracle_chunking(n, rate):
    bounds = port.chunk_bounds(n, rate)
    total = bounds[-1][1] if bounds else 0
    assert ms_framing(n, rate) == total
    for (s, e) in bounds[:-1]:
        assert e - s == 30 * rate                                      # chunk = 30 * rate frames exactly


def test_drop_in_module_imports_without_gpu():
    import audio_mastering_engine as ame
    assert set(ame.EQ_PRESETS) == {"techno", "dubstep", "pop", "rock"}
    assert ame.EQ_PRESETS["rock"]["mid_cut"] == -2.0
    for name in ["process_audio_from_gcs", "process_audio", "batch_process_audio", "audio_segment_to_float_array",
                 "float_array_to_audio_segment", "apply_saturation", "apply_stereo_width", "apply_eq_to_samples",
                 "apply_shelf_filter", "apply_peak_filter", "apply_multiband_compressor", "normalize_to_lufs", "soft_limiter"]:
        assert callable(getattr(ame, name))
    x = np.zeros((4, 2), dtype=np.float32)
    assert ame.apply_saturation(x, 0) is x                             # ENG:129 bypass returns its argument
    assert ame.apply_shelf_filter(x[:, 0], 44100, 250, 0, "low") is not None


def test_product_does_not_import_oracle():
    """The product path must never route through the oracle."""
    pkg = os.path.join(ROOT, "python-audio-mastering_b200")
    for dirpath, _d, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f"{f} imports oracle"


def test_struct_layouts_match_the_header(tmp_path):
    """ctypes mirrors of include/b200_master.h have the C compiler's sizes (INTEGRATION.md quotes them)."""
    import ctypes as C
    import subprocess
    from b200master import lib as L
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "%s"\nint main(void){printf("%%zu %%zu %%zu %%zu\\n", sizeof(b200m_plan), '
                   'sizeof(b200m_band), sizeof(b200m_biquad), sizeof(b200m_settings));return 0;}\n'
                   % os.path.join(ROOT, "include", "b200_master.h"))
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", str(src), "-o", str(exe)])
    sizes = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    assert sizes == [C.sizeof(L.Plan), C.sizeof(L.Band), C.sizeof(L.Biquad), C.sizeof(L.Settings)]
    assert sizes[0] == 592


def test_wav_header_is_the_wave_modules():
    """b200m_wav_header (host side of b200m_master_batch_wav): the 44 bytes the stdlib wave module writes
    for 16-bit PCM, which is what pydub's export(format="wav") emits (ENG:96-99).  No device work."""
    import ctypes as C
    import io
    import wave
    from b200master import lib as L
    lib = L.load()
    for rate, ch, frames in [(44100, 2, 0), (48000, 2, 1440000), (96000, 1, 12345), (8000, 1, 1)]:
        out = (C.c_ubyte * 44)()
        assert lib.b200m_wav_header(rate, ch, frames, out) == 0
        f = io.BytesIO()
        with wave.open(f, "wb") as w:
            w.setnchannels(ch); w.setsampwidth(2); w.setframerate(rate)
            w.writeframesraw(b"\0" * (frames * ch * 2))
        assert bytes(out) == f.getvalue()[:44]
    assert lib.b200m_wav_header(44100, 3, 10, (C.c_ubyte * 44)()) != 0
    assert lib.b200m_wav_header(44100, 2, 1 << 31, (C.c_ubyte * 44)()) != 0      # does not fit a RIFF file


def test_gain_routine_constants_and_accuracy():
    """k_comp's 10^x (exp10_gain, csrc/b200m_kernels.cuh): the constants in the header are the correctly rounded
    ones (Taylor coefficients of 10^r, 32 log2 10, the two pieces of -log10(2)/32, the table 2^(j/32)), and the
    operation sequence -- simulated with exact FMA semantics by scripts/exp10_check.py -- stays within 1.05 ulp of
    60-digit arithmetic over the exponents a compressor produces (the class of libm's exp10, which pydub's
    db_to_float goes through)."""
    mpmath = pytest.importorskip("mpmath")
    import importlib.util, random
    spec = importlib.util.spec_from_file_location("exp10_check", os.path.join(ROOT, "scripts", "exp10_check.py"))
    chk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(chk)
    src = open(os.path.join(ROOT, "python-audio-mastering_b200", "csrc", "b200m_kernels.cuh")).read()
    body = re.search(r"__constant__ double c_exp10\[11\] = \{(.*?)\};", src, re.S).group(1)
    consts = [float.fromhex(t) if "0x" in t else float(t) for t in re.findall(r"-?0x[0-9a-fA-F.]+p[+-]?\d+|\d+\.\d+", re.sub(r"//.*", "", body))]
    assert len(consts) == 11
    assert consts[:7] == chk.TAYLOR[:7]
    assert consts[7] == 32.0 * chk.LOG2_10 and consts[8] == chk.NEG_LOG10_2_HI / 32.0 and consts[9] == chk.NEG_LOG10_2_LO / 32.0
    assert consts[10] == chk.MAGIC
    tab = re.search(r"g_exp10_tab\[32\] = \{(.*?)\};", src, re.S).group(1)
    tab = [float.fromhex(t) for t in re.findall(r"0x[0-9a-fA-F.]+p[+-]?\d+", tab)]
    assert tab == [float(mpmath.mpf(2) ** (mpmath.mpf(j) / 32)) for j in range(32)]
    rnd = random.Random(11)
    worst = 0.0
    for i in range(3000):
        att = rnd.random() * (60.0 if i % 4 else 5000.0)
        q = att * 0.05
        x = -chk.fma(chk.fma(-q, 20.0, att), 0.05, q)          # att / 20 rounded like CPython's true division
        assert x == -(att / 20.0)
        if x > -300.0:
            worst = max(worst, chk.ulp_err(chk.exp10_tf(x, 32, 6), x))
    assert chk.exp10_tf(-0.0, 32, 6) == 1.0                      # pydub multiplies only when att != 0: 10^-0 must be exactly 1
    assert worst <= 1.05, worst
