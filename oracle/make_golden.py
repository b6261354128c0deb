"""Generate ``tests/golden/*.npz`` by running the UNCHANGED reference engine.

TEST INFRASTRUCTURE ONLY.  Run in the authoring container (needs ``/root/reference``):

    python -m oracle.make_golden

Each fixture holds the seeded input PCM, the settings (JSON), and what the
reference's own ``process_audio_from_gcs`` (ENG:24-113) produced for it, with the
faithful pydub/audioop compressor loop of ``thirdparty.py``.  ``stages.npz`` also
holds per-helper outputs obtained by calling the reference's helper functions
(ENG:117-227) directly.
"""
from __future__ import annotations

import json
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "python-audio-mastering_b200"))

from oracle import refload, thirdparty  # noqa: E402
from b200master import synth  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")

POP = dict(bass_boost=2.0, mid_cut=0.0, presence_boost=3.5, treble_boost=2.5)
TECHNO = dict(bass_boost=4.0, mid_cut=3.0, presence_boost=1.0, treble_boost=3.0)
DUBSTEP = dict(bass_boost=5.0, mid_cut=4.0, presence_boost=2.0, treble_boost=3.5)
ROCK = dict(bass_boost=1.5, mid_cut=-2.0, presence_boost=2.5, treble_boost=1.0)


def cases():
    sq = np.zeros((30000, 2), dtype=np.int16)
    sq[:, 0] = np.where((np.arange(30000) // 50) % 2 == 0, 32767, -32768)
    sq[:, 1] = np.where((np.arange(30000) // 37) % 2 == 0, -32768, 32767)
    quiet = (synth.make_track(11, 1.0, 44100).astype(np.int32) // 4096).astype(np.int16)
    return {
        # BASELINE config 1 in miniature: Pop, multiband off, -14 LUFS
        "cfg1_pop_44k": (synth.make_track(0, 2.0, 44100), 44100,
                         dict(POP, saturation=0, width=1.0, multiband=False, lufs=-14.0)),
        # BASELINE config 2 in miniature: same track, full chain
        "cfg2_full_44k": (synth.make_track(0, 2.0, 44100), 44100,
                          dict(POP, saturation=25, width=1.2, multiband=True, lufs=-14.0)),
        # BASELINE config 3 shape: 48 kHz Techno, multiband on
        "cfg3_techno_48k": (synth.make_track(1, 1.5, 48000), 48000,
                            dict(TECHNO, saturation=0, width=1.0, multiband=True, lufs=-14.0)),
        # two chunks at a low rate: filter/compressor state reset at 30 s, whole-track loudness.
        # (12 kHz keeps the file small; the 4 kHz / 8 kHz EQ sections are bypassed because the
        # reference's doubled centre frequencies would put them beyond Nyquist = unstable.)
        "two_chunks_12k": (synth.make_track(2, 31.0, 12000), 12000,
                           dict(bass_boost=5.0, mid_cut=4.0, presence_boost=0.0, treble_boost=0.0,
                                saturation=10, width=1.3, multiband=True, lufs=-9.0)),
        # mono input
        "mono_44k": (synth.make_track(3, 1.0, 44100, channels=1), 44100,
                     dict(ROCK, saturation=40, width=1.5, multiband=True, lufs=-23.0)),
        # lufs None -> float32 limiter path; EQ fully bypassed -> float32 width
        "no_lufs_no_eq": (synth.make_track(4, 1.0, 44100), 44100,
                          dict(saturation=60, width=0.5, multiband=False)),
        # rock preset (negative mid_cut => boost), custom band settings, no width
        "rock_custom_bands": (synth.make_track(5, 1.5, 44100), 44100,
                              dict(ROCK, saturation=80, width=1.0, multiband=True, lufs=-14.0,
                                   low_thresh=-30.0, low_ratio=2.5, mid_thresh=-28.0, mid_ratio=8.0,
                                   high_thresh=-35.0, high_ratio=10.0)),
        # non-ms-aligned length (tail dropped / padded by pydub's ms slicing)
        "ragged_tail_a": (synth.make_track(6, 1.0, 44100)[:44100 - 23], 44100,
                          dict(TECHNO, saturation=0, width=1.1, multiband=True, lufs=-14.0)),
        "ragged_tail_b": (synth.make_track(6, 1.0, 44100)[:44100 - 21], 44100,
                          dict(TECHNO, saturation=0, width=1.1, multiband=False, lufs=-16.0)),
        # full-scale square waves: clip wrap (+1.0 -> -32768) and limiter everywhere
        "fullscale_square": (sq, 48000, dict(DUBSTEP, saturation=100, width=2.0, multiband=True, lufs=-6.0)),
        # silence -> loudness -inf -> gain inf -> NaN -> int16 0 (reference emits this, no error)
        "silence": (np.zeros((22050, 2), dtype=np.int16), 44100,
                    dict(POP, saturation=10, width=1.2, multiband=True, lufs=-14.0)),
        # a few LSBs of signal: bands never reach threshold, loudness gate edge
        "quiet": (quiet, 44100, dict(POP, saturation=5, width=1.0, multiband=True, lufs=-14.0)),
        # 96 kHz (config 4's rate; 16-bit because the reference's 24-bit path is broken)
        "hi_rate_96k": (synth.make_track(7, 1.0, 96000), 96000,
                        dict(TECHNO, saturation=15, width=1.2, multiband=True, lufs=-14.0)),
    }


def write_exciter_tables(saturations):
    """``exciter_tables.npz``: the exciter (ENG:128-134) of THIS host's numpy over the 65536 int16 samples, for every
    saturation value the fixtures use.  numpy's float32 tanh is a SIMD polynomial that may differ by an ulp between
    CPU dispatch targets, and the fixtures' outputs were computed with this host's; a GPU test replays the stored
    table (``plan.install_exciter_table``) so that it compares like with like on any box.  Stored compactly: the ulp
    difference of np.tanh(float32) from float32(np.tanh(float64)) (int8, mostly 0 / +-1), plus a SHA-256 of the
    table the loader must reproduce (``tests/conftest.py: authoring_exciter_table``)."""
    import hashlib
    from b200master.plan import exciter_table
    out = {}
    for sat in sorted(set(saturations)):
        if sat == 0:
            continue
        s16 = np.arange(65536, dtype=np.uint16).view(np.int16)
        x = s16.astype(np.float32) / (2 ** 15)
        mix = (sat / 100.0) ** 2
        arg = x * (1 + mix * 4)
        t32 = np.tanh(arg)
        t0 = np.tanh(arg.astype(np.float64)).astype(np.float32)
        d = t32.view(np.int32).astype(np.int64) - t0.view(np.int32).astype(np.int64)
        assert np.abs(d).max() <= 100
        table = exciter_table(sat)
        assert np.array_equal(table.view(np.int32), ((1 - mix) * x + mix * t32).astype(np.float32).view(np.int32))
        tag = repr(float(sat))
        out["d_" + tag] = d.astype(np.int8)
        out["sha_" + tag] = np.array(hashlib.sha256(table.tobytes()).hexdigest())
        print(f"exciter table sat={sat}: {int((d != 0).sum())} entries differ from the correctly rounded tanh")
    np.savez_compressed(os.path.join(GOLDEN, "exciter_tables.npz"), **out)


def main():
    warnings.simplefilter("ignore", DeprecationWarning)
    os.makedirs(GOLDEN, exist_ok=True)
    manifest = {}
    for name, (pcm, rate, settings) in cases().items():
        thirdparty.Meter.last_loudness = None
        out, _log = refload.run_reference(pcm, rate, settings)
        loud = thirdparty.Meter.last_loudness
        np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), pcm=pcm, out=out,
                            rate=np.int64(rate), settings=np.array(json.dumps(settings)),
                            loudness=np.float64(np.nan if loud is None else loud))
        manifest[name] = dict(frames=int(pcm.shape[0]), rate=rate, loudness=loud,
                              settings=settings)
        print(f"{name}: {pcm.shape} @ {rate} -> {out.shape}, loudness {loud}")

    # per-helper outputs of the reference functions themselves
    eng = refload.load_reference_engine()
    rate = 48000
    pcm = synth.make_track(8, 0.5, rate)
    seg = refload._RefAudioSegment(pcm.tobytes(), 2, rate, 2)
    st = dict(TECHNO)
    x = eng.audio_segment_to_float_array(seg)
    sat = eng.apply_saturation(x, 35)
    eqd = eng.apply_eq_to_samples(sat, rate, st)
    wid = eng.apply_stereo_width(eqd, 1.4)
    q1 = eng.float_array_to_audio_segment(wid, seg)
    mb = eng.apply_multiband_compressor(q1, -25.0, 6.0, -20.0, 3.0, -15.0, 4.0)
    lowshelf = eng.apply_shelf_filter(x[:, 0], rate, 250, 4.0, "low")
    peak = eng.apply_peak_filter(x[:, 1], rate, 4000, -3.0)
    proc = eng.audio_segment_to_float_array(mb)
    thirdparty.Meter.last_loudness = None
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        norm = eng.normalize_to_lufs(proc.copy(), rate, -14.0)
    lim = eng.soft_limiter(norm.copy())
    lim32 = eng.soft_limiter((proc * np.float32(1.7)).copy())
    fin = eng.float_array_to_audio_segment(lim, seg)
    np.savez_compressed(
        os.path.join(GOLDEN, "stages.npz"), pcm=pcm, rate=np.int64(rate), settings=np.array(json.dumps(st)),
        to_float=x, saturation35=sat, eq=eqd, width14=wid, q1=q1.to_numpy(), multiband=mb.to_numpy(),
        lowshelf_L=lowshelf, peak_R=peak, loudness=np.float64(thirdparty.Meter.last_loudness),
        normalized=norm, limited=lim, limited32=lim32, final=fin.to_numpy())
    manifest["stages"] = dict(frames=int(pcm.shape[0]), rate=rate, settings=st)
    write_exciter_tables([st.get("saturation", 0) for _p, _r, st in cases().values()] + [35])
    with open(os.path.join(GOLDEN, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    print("stages.npz written")


if __name__ == "__main__":
    main()
