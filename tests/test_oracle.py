"""CPU tests pinning the oracle (``oracle/port.py`` + ``compressor.c``) against the golden
vectors produced by the UNCHANGED reference engine (``oracle/make_golden.py``), and
against standards-derived known answers."""
import math
import os

import numpy as np
import pytest
from scipy.signal import sosfilt

from conftest import golden_names, load_golden
from oracle import port, refload, thirdparty


@pytest.mark.parametrize("name", golden_names())
def test_port_matches_reference_golden(name):
    g = load_golden(name)
    out, info = port.master(g["pcm"], g["rate"], g["settings"], impl="c")
    assert out.shape == g["out"].shape
    assert np.array_equal(out, g["out"]), f"{name}: port differs from the reference output"
    if g["settings"].get("lufs") is not None:
        assert info["loudness"] == g["loudness"] or (math.isinf(info["loudness"]) and math.isinf(g["loudness"]))


def test_port_stage_goldens():
    g = load_golden("stages")
    rate, st = g["rate"], g["settings"]
    x = port.pcm_to_float(g["pcm"])
    assert np.array_equal(x, g["to_float"])
    sat = port.exciter(x, 35)
    assert np.array_equal(sat, g["saturation35"])
    eqd = port.eq(sat, rate, st)
    assert np.array_equal(eqd, g["eq"])
    wid = port.widen(eqd, 1.4)
    assert np.array_equal(wid, g["width14"])
    q1 = port.float_to_pcm16(wid)
    assert np.array_equal(q1, g["q1"])
    mb = port.multiband(q1, rate, port.band_params({}))
    assert np.array_equal(mb, g["multiband"])
    assert np.array_equal(sosfilt(port.shelf_sos(rate, 250, 4.0, "low"), x[:, 0]), g["lowshelf_L"])
    assert np.array_equal(sosfilt(port.peak_sos(rate, 4000, -3.0), x[:, 1]), g["peak_R"])
    proc = port.pcm_to_float(mb)
    norm, loud, _gain = port.normalize(proc.copy(), rate, -14.0)
    assert loud == g["loudness"]
    assert np.array_equal(norm, g["normalized"])
    assert np.array_equal(port.limiter(norm.copy()), g["limited"])
    assert np.array_equal(port.limiter((proc * np.float32(1.7)).copy()), g["limited32"])
    assert np.array_equal(port.float_to_pcm16(port.limiter(norm.copy())), g["final"])


def test_c_compressor_matches_faithful_audioop_loop():
    """compressor.c vs the frame-by-frame pydub/audioop loop: bit-identical output."""
    g = load_golden("stages")
    rate = g["rate"]
    bands = port.split_bands(g["q1"][:12000], rate)
    for band, (thr, ratio), (att, rel) in zip(bands, port.band_params({"high_thresh": -40.0}), port.BAND_TIMES):
        a = port.compress_band(band, rate, thr, ratio, att, rel, impl="py")
        b = port.compress_band(band, rate, thr, ratio, att, rel, impl="c")
        assert np.array_equal(a, b)
    mono = np.ascontiguousarray(bands[0][:, 0])
    assert np.array_equal(port.compress_band(mono, rate, -30.0, 2.0, 3.3, 77.0, impl="py"),
                          port.compress_band(mono, rate, -30.0, 2.0, 3.3, 77.0, impl="c"))


def test_audioop_known_answers():
    """the three audioop primitives pydub uses."""
    audioop = thirdparty.audioop
    h = lambda v: np.array(v, dtype=np.int16).tobytes()
    assert audioop.rms(b"", 2) == 0
    assert audioop.rms(h([3, 4]), 2) == 3
    got = np.frombuffer(audioop.mul(h([3, -3, 5, -5, 1, -1, 32767, -32768]), 2, 0.5), dtype=np.int16)
    assert got.tolist() == [1, -2, 2, -3, 0, -1, 16383, -16384]
    got = np.frombuffer(audioop.mul(h([20000, -20000]), 2, 2.0), dtype=np.int16)
    assert got.tolist() == [32767, -32768]
    got = np.frombuffer(audioop.add(h([30000, -30000, 5]), h([10000, -10000, 6]), 2), dtype=np.int16)
    assert got.tolist() == [32767, -32768, 11]


def test_quantiser_truncates_and_wraps():
    """truncation toward zero and the +1.0 -> -32768 wrap (ENG:124-125)."""
    for dt in (np.float32, np.float64):
        x = np.array([1.0, -1.0, 0.99999, 1.5, 32767.9 / 32768, -32767.9 / 32768, 0.3 / 32768, -0.3 / 32768], dtype=dt)
        assert port.float_to_pcm16(x).tolist() == [-32768, -32768, 32767, -32768, 32767, -32767, 0, 0]


def test_k_weighting_matches_bs1770_table():
    """restated pyloudnorm coefficients vs the ITU-R BS.1770 48 kHz table."""
    m = thirdparty.Meter(48000)
    sh, hp = m._filters["high_shelf"], m._filters["high_pass"]
    assert np.allclose(sh.b, [1.53512485958697, -2.69169618940638, 1.19839281085285], atol=2e-4)
    assert np.allclose(sh.a, [1.0, -1.69065929318241, 0.73248077421585], atol=2e-4)
    assert np.allclose(hp.a, [1.0, -1.99004745483398, 0.99007225036621], atol=2e-4)


def test_full_scale_sine_loudness():
    """997 Hz full-scale mono sine reads -3.05 LUFS under the restated meter."""
    for rate in (44100, 48000, 96000):
        t = np.arange(5 * rate) / rate
        x = np.sin(2 * np.pi * 997 * t).astype(np.float32)
        assert abs(port.integrated_loudness(x, rate) - (-3.05)) < 0.01


def test_short_audio_raises():
    with pytest.raises(ValueError):
        port.integrated_loudness(np.zeros(1000, dtype=np.float32), 48000)


def test_chunks_are_independent():
    """ENG:48-54: every 30-s chunk restarts from zero state."""
    from b200master import synth
    rate = 12000
    pcm = synth.make_track(21, 31.0, rate)
    st = dict(bass_boost=4.0, mid_cut=3.0, width=1.2, multiband=True)
    _, info = port.master(pcm, rate, st)
    a = port.process_chunk(pcm[:30 * rate], rate, st)
    b = port.process_chunk(pcm[30 * rate:], rate, st)
    assert np.array_equal(info["processed"], np.concatenate([a, b]))


@pytest.mark.skipif(not refload.reference_available(), reason="needs /root/reference (authoring container)")
def test_reference_engine_runs_unchanged():
    """The unchanged reference file, driven through its own entry point, equals the port."""
    from b200master import synth
    pcm = synth.make_track(33, 0.6, 44100)
    st = dict(bass_boost=2.0, presence_boost=3.5, treble_boost=2.5, saturation=20, width=1.2, multiband=True, lufs=-14.0)
    ref, _ = refload.run_reference(pcm, 44100, st)
    out, _ = port.master(pcm, 44100, st, impl="c")
    assert np.array_equal(ref, out)


# ---------------------------------------------------------------------------------------------------
# Cross-checks of the restated third-party arithmetic against the REAL packages.  pydub and pyloudnorm are
# unpinned dependencies of the reference (requirements.txt:2,5) that are absent from this image (no index
# access): these tests skip today and pin rows a10 / a13 of SURVEY.md section 8 the day the wheels appear.
# ---------------------------------------------------------------------------------------------------
def _real_package(name):
    """The REAL package ``name`` (never the shim oracle/refload.py may have put into sys.modules), or skip."""
    import importlib
    import sys
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if (k == name or k.startswith(name + ".")) and
             getattr(sys.modules[k], "__file__", None) is None}          # the shims are bare ModuleType objects
    try:
        mod = importlib.import_module(name)
        if getattr(mod, "__file__", None) is None:
            raise ImportError(name)
        return mod
    except ImportError:
        pytest.skip(f"{name} is not installed here (unpinned dependency of the reference, requirements.txt); the restatement "
                    f"in oracle/thirdparty.py stays unpinned at this boundary")
    finally:
        for k, v in saved.items():
            sys.modules.setdefault(k, v)


def test_restated_compressor_matches_real_pydub():
    pydub = _real_package("pydub")
    real_compress = _real_package("pydub.effects").compress_dynamic_range
    from b200master import synth
    rate = 44100
    q1 = port.process_chunk(synth.make_track(40, 1.0, rate), rate, dict(bass_boost=4.0, treble_boost=3.0))
    for band, (thr, ratio), (att, rel) in zip(port.split_bands(q1, rate), port.band_params({"high_thresh": -32.0}), port.BAND_TIMES):
        seg = pydub.AudioSegment(band.tobytes(), sample_width=2, frame_rate=rate, channels=2)
        real = np.frombuffer(real_compress(seg, threshold=thr, ratio=ratio, attack=att, release=rel)._data, dtype=np.int16).reshape(-1, 2)
        assert np.array_equal(real, port.compress_band(band, rate, thr, ratio, att, rel, impl="py")), "thirdparty.compress_dynamic_range"
        assert np.array_equal(real, port.compress_band(band, rate, thr, ratio, att, rel, impl="c")), "compressor.c"
    # the AudioSegment members the chain touches (ENG:43,54,80,126,210)
    seg = pydub.AudioSegment(q1.tobytes(), sample_width=2, frame_rate=rate, channels=2)
    mine = thirdparty.AudioSegment.from_numpy(q1, rate)
    assert len(seg) == len(mine) and seg.frame_count() == mine.frame_count()
    assert seg[100:350]._data == mine[100:350]._data
    assert seg.overlay(seg)._data == mine.overlay(mine)._data
    assert seg.rms == mine.rms and seg.max_possible_amplitude == mine.max_possible_amplitude


def test_restated_meter_matches_real_pyloudnorm():
    pyln = _real_package("pyloudnorm")
    from b200master import synth
    for rate, seconds in ((44100, 3.0), (48000, 2.5), (96000, 1.0)):
        x = port.pcm_to_float(synth.make_track(41, seconds, rate))
        mono = x.mean(axis=1)                                           # ENG:215
        assert pyln.Meter(rate).integrated_loudness(mono) == thirdparty.Meter(rate).integrated_loudness(mono)
        real, mine = pyln.Meter(rate), thirdparty.Meter(rate)
        for (_n, fr), (_m, fm) in zip(real._filters.items(), mine._filters.items()):
            assert np.array_equal(fr.b, fm.b) and np.array_equal(fr.a, fm.a)
    with pytest.raises(ValueError):
        pyln.Meter(44100).integrated_loudness(np.zeros(1000, dtype=np.float32))
