"""audio_mastering_engine -- B200-native drop-in for the reference module of the same name.

Same import name, presets, entry points and ``settings`` dict as
``/root/reference/worker/audio_mastering_engine.py`` ("ENG"), so ``mastering_gui.py``
(``import audio_mastering_engine as engine``, GUI:18) and ``worker/main.py``
(``from audio_mastering_engine import process_audio_from_gcs``, WRK:11) run unchanged.
Every sample is computed by hand-written sm_100a CUDA kernels behind the C-ABI of
``libb200master.so`` (``include/b200_master.h``); decode / encode / storage stay on the
host.  There is no CPU fallback: importing works anywhere, but the first call needs the
built library and a CUDA device and raises otherwise.

Reference surface mirrored here
    EQ_PRESETS                                     ENG:15-20
    process_audio_from_gcs(gcs_uri, settings)      ENG:24-113
    process_audio_batch_from_gcs(jobs)             the same for several worker jobs in one GPU batch (extension)
    process_audio(settings, status_callback)       called at GUI:204 (absent from the snapshot)
    batch_process_audio(settings, in, out, cb)     called at GUI:220 (absent from the snapshot)
    audio_segment_to_float_array ... soft_limiter  ENG:117-227 (ten helpers)
"""
from __future__ import annotations

import io
import os

import numpy as np

from b200master import lib as _L
from b200master.engine import get_engine, ms_framing
from b200master.plan import (BAND_TIMES, butter4, kweight_biquads, make_band, make_plan,
                             peak_biquad, shelf_biquad)
from b200master.segment import PcmSegment, segment_class

# --- presets (ENG:15-20): same keys and values; the GUI copies them into its sliders ---
EQ_PRESETS = {
    "techno": {"bass_boost": 4.0, "mid_cut": 3.0, "presence_boost": 1.0, "treble_boost": 3.0,
               "description": "Club curve: lifted sub and air, mids scooped."},
    "dubstep": {"bass_boost": 5.0, "mid_cut": 4.0, "presence_boost": 2.0, "treble_boost": 3.5,
                "description": "Heavy low end, crisp top, deep mid scoop."},
    "pop": {"bass_boost": 2.0, "mid_cut": 0.0, "presence_boost": 3.5, "treble_boost": 2.5,
            "description": "Vocal-forward with firm lows and bright highs."},
    "rock": {"bass_boost": 1.5, "mid_cut": -2.0, "presence_boost": 2.5, "treble_boost": 1.0,
             "description": "Warm low-mids for guitars, punchy presence."},
}

AUDIO_EXTENSIONS = (".wav", ".mp3", ".flac", ".aiff", ".aif", ".ogg", ".m4a")


# =========================================================================================
# whole-track path
# =========================================================================================
def _segment_pcm(seg) -> np.ndarray:
    if seg.sample_width != 2:
        # ENG:125 always casts to int16 while scaling by 2^(8*width-1): for any other width the
        # reference emits half-length garbage, so there is nothing to match.
        raise ValueError(f"{8 * seg.sample_width}-bit PCM is not supported by the mastering chain; convert to 16-bit")
    if seg.channels not in (1, 2):
        raise ValueError("only mono and stereo audio are supported")
    pcm = np.frombuffer(seg._data, dtype=np.int16)
    return pcm.reshape(-1, 2) if seg.channels == 2 else pcm


def _job_error(seg, settings):
    """Why the reference could not master this file (None: it can).  The reference handles one file per call, so a
    bad file fails alone; a batch checks every file up front and masters the rest."""
    try:
        _segment_pcm(seg)
    except ValueError as e:
        return str(e)
    if (settings or {}).get("lufs") is not None and seg.frame_count() < 0.4 * seg.frame_rate:
        return "Audio must have length greater than the block size."       # pyloudnorm valid_audio (ENG:218)
    return None


def master_segment(seg, settings, device: int = 0):
    """ENG:46-89 on one decoded segment: chunk loop, loudness, limiter -> new segment."""
    outs, infos = get_engine(device).master([_segment_pcm(seg)], seg.frame_rate, settings)
    if infos[0].get("loudness") is not None:
        loud, target = infos[0]["loudness"], settings.get("lufs")
        print(f"Measured {loud:.2f} LUFS; applied {target - loud:.2f} dB of gain.")
    return seg._spawn(outs[0].tobytes())


def master_segments(segs, settings, device: int = 0):
    """Batch form of ``master_segment``: same-format segments go to the GPU as one launch."""
    segs = list(segs)
    groups, out = {}, [None] * len(segs)
    for i, s in enumerate(segs):
        groups.setdefault((s.frame_rate, s.channels), []).append(i)
    for (rate, _ch), idx in groups.items():
        res, _ = get_engine(device).master([_segment_pcm(segs[i]) for i in idx], rate, settings)
        for i, r in zip(idx, res):
            out[i] = segs[i]._spawn(r.tobytes())
    return out


def master_segments_wav(segs, settings, device: int = 0):
    """``master_segments`` with the WAV export (ENG:96-99) folded into the GPU batch: returns, per
    segment, the bytes ``export(format="wav")`` would write (``b200m_master_batch_wav``: headers written
    on the device ahead of the samples, no host pass over the PCM)."""
    segs = list(segs)
    groups, out = {}, [None] * len(segs)
    for i, s in enumerate(segs):
        groups.setdefault((s.frame_rate, s.channels), []).append(i)
    for (rate, _ch), idx in groups.items():
        images, _ = get_engine(device).master_wav([_segment_pcm(segs[i]) for i in idx], rate, settings)
        for i, img in zip(idx, images):
            out[i] = img
    return out


def process_audio_from_gcs(gcs_uri, settings):
    """ENG:24-113: download from GCS, master on the GPU, upload ``processed/mastered_<name>``
    and its ``.complete`` marker.  Exceptions propagate to the caller (ENG:110-113)."""
    process_audio_batch_from_gcs([(gcs_uri, settings)])


def process_audio_batch_from_gcs(jobs):
    """Several worker jobs in ONE GPU batch: ``jobs`` is a sequence of
    ``(gcs_uri, settings)`` as ``worker/main.py:39`` receives them one by one.  Every object is
    downloaded and decoded on the host (ENG:27-43), all tracks are mastered together with their own
    settings, and each result is uploaded like ENG:91-108: the WAV image comes straight from the GPU
    batch (``b200m_master_batch_wav``), then the ``.complete`` marker.  Exceptions propagate (ENG:110-113)."""
    jobs = list(jobs)
    try:
        from google.cloud import storage
        client = storage.Client()
        decoded = []
        for gcs_uri, _settings in jobs:
            print(f"Fetching {gcs_uri} ...")
            bucket_name, blob_name = gcs_uri.replace("gs://", "").split("/", 1)
            bucket = client.bucket(bucket_name)
            src = io.BytesIO()
            bucket.blob(blob_name).download_to_file(src)
            src.seek(0)
            decoded.append((bucket, blob_name, segment_class().from_file(src)))
        print(f"Decoded {len(decoded)} object(s); mastering on the GPU ...")
        failed = {i: msg for i, (_b, _n, seg) in enumerate(decoded) if (msg := _job_error(seg, jobs[i][1])) is not None}
        groups = {}
        for i, (_b, _n, seg) in enumerate(decoded):
            if i not in failed:
                groups.setdefault((seg.frame_rate, seg.channels), []).append(i)
        images = [None] * len(jobs)
        for (rate, _ch), idx in groups.items():
            res, infos = get_engine().master_wav([_segment_pcm(decoded[i][2]) for i in idx], rate, [jobs[i][1] for i in idx])
            for i, img, info in zip(idx, res, infos):
                images[i] = img
                if info.get("loudness") is not None:
                    print(f"{decoded[i][1]}: measured {info['loudness']:.2f} LUFS; applied "
                          f"{jobs[i][1].get('lufs') - info['loudness']:.2f} dB of gain.")
        for (bucket, blob_name, _seg), img in zip(decoded, images):
            if img is None:
                continue                                     # a job that failed on its own (reported below), like ENG:110-113 for that job
            target = f"processed/mastered_{os.path.basename(blob_name)}"
            print(f"Uploading {target} ...")
            bucket.blob(target).upload_from_file(io.BytesIO(img.tobytes()), content_type="audio/wav")
            bucket.blob(f"{target}.complete").upload_from_string("")
            print(f"Done: {target}.complete written.")
        if failed:
            raise ValueError("; ".join(f"{decoded[i][1]}: {msg}" for i, msg in sorted(failed.items())))
    except Exception as e:
        print(f"FATAL ERROR in mastering engine: {e}")
        raise


def _export(seg, path):
    ext = os.path.splitext(path)[1].lower().lstrip(".") or "wav"
    seg.export(path, format=ext)


def process_audio(settings, status_callback=None):
    """Desktop entry point (GUI:192-206): paths travel inside ``settings``; progress and the
    final "complete" / "error" strings go to ``status_callback`` (GUI:224-232)."""
    say = status_callback or (lambda _m: None)
    try:
        src, dst = settings.get("input_file"), settings.get("output_file")
        if not src or not dst:
            raise ValueError("settings must carry 'input_file' and 'output_file'")
        say("Loading audio...")
        audio = segment_class().from_file(src)
        say("Mastering on the GPU...")
        mastered = master_segment(audio, settings)
        say("Exporting...")
        _export(mastered, dst)
        say(f"Processing complete: {os.path.basename(dst)}")
    except Exception as e:
        say(f"Error: {e}")


def batch_process_audio(settings, input_folder, output_folder, status_callback=None):
    """Desktop batch entry point (GUI:208-222): every audio file of ``input_folder`` is
    mastered with the same settings; same-format files share one GPU launch."""
    say = status_callback or (lambda _m: None)
    try:
        names = sorted(n for n in os.listdir(input_folder) if n.lower().endswith(AUDIO_EXTENSIONS))
        if not names:
            say("No audio files found in the input folder.")
            return
        os.makedirs(output_folder, exist_ok=True)
        say(f"Loading {len(names)} files...")
        cls = segment_class()
        segs = [cls.from_file(os.path.join(input_folder, n)) for n in names]
        bad = [(n, msg) for n, sg in zip(names, segs) if (msg := _job_error(sg, settings)) is not None]
        for n, msg in bad:                                   # the reference would have failed on this file alone
            say(f"Error: {n}: {msg}")
        keep = [i for i, n in enumerate(names) if n not in {b for b, _ in bad}]
        total = len(names)
        names, segs = [names[i] for i in keep], [segs[i] for i in keep]
        say(f"Mastering {len(segs)} files on the GPU...")
        if not segs:
            pass
        elif all(n.lower().endswith(".wav") for n in names):
            # WAV in, WAV out: the GPU returns complete file images
            for n, img in zip(names, master_segments_wav(segs, settings)):
                with open(os.path.join(output_folder, f"mastered_{n}"), "wb") as f:
                    f.write(img)
        else:
            outs = master_segments(segs, settings)
            for n, seg in zip(names, outs):
                _export(seg, os.path.join(output_folder, f"mastered_{n}"))
        say(f"Batch processing complete: {len(names)} of {total} files.")
    except Exception as e:
        say(f"Error: {e}")


# =========================================================================================
# helpers (ENG:117-227), numpy in / numpy out, arithmetic on the GPU
# =========================================================================================
def audio_segment_to_float_array(audio_segment):
    """ENG:117-121."""
    if audio_segment.sample_width != 2:
        raise ValueError("only 16-bit segments are supported")
    pcm = np.frombuffer(audio_segment._data, dtype=np.int16)
    x = get_engine().pcm16_to_float(pcm)
    return x.reshape(-1, 2) if audio_segment.channels == 2 else x


def float_array_to_audio_segment(float_array, audio_segment_template):
    """ENG:123-126 (clip, *2^15, truncate, +FS wrap)."""
    return audio_segment_template._spawn(get_engine().float_to_pcm16(np.asarray(float_array)).tobytes())


def apply_saturation(samples, saturation_percent):
    """ENG:128-134."""
    if saturation_percent == 0:
        return samples
    return get_engine().saturation(samples, saturation_percent)


def apply_stereo_width(samples, width_factor):
    """ENG:136-144."""
    if samples.ndim == 1 or samples.shape[1] != 2:
        return samples
    return get_engine().stereo_width(samples, width_factor)


def apply_eq_to_samples(samples, sample_rate, settings):
    """ENG:146-168: low shelf, two peaks, high shelf per channel; bypassed sections skipped."""
    secs = [shelf_biquad(sample_rate, 250, settings.get("bass_boost", 0.0), "low"),
            peak_biquad(sample_rate, 1000, -settings.get("mid_cut", 0.0)),
            peak_biquad(sample_rate, 4000, settings.get("presence_boost", 0.0)),
            shelf_biquad(sample_rate, 8000, settings.get("treble_boost", 0.0), "high")]
    secs = [s for s in secs if s is not None]
    if not secs:
        return np.array(samples)
    return get_engine().sosfilt(secs, samples)


def apply_shelf_filter(samples, sample_rate, cutoff_hz, gain_db, filter_type, q=0.707):
    """ENG:170-183."""
    sec = shelf_biquad(sample_rate, cutoff_hz, gain_db, filter_type, q)
    return samples if sec is None else get_engine().sosfilt([sec], samples)


def apply_peak_filter(samples, sample_rate, center_hz, gain_db, q=1.0):
    """ENG:185-194."""
    sec = peak_biquad(sample_rate, center_hz, gain_db, q)
    return samples if sec is None else get_engine().sosfilt([sec], samples)


def apply_multiband_compressor(chunk, low_thresh, low_ratio, mid_thresh, mid_ratio, high_thresh, high_ratio,
                               low_crossover=250, high_crossover=4000):
    """ENG:196-210 on one segment (zero state)."""
    st = dict(multiband=True, low_thresh=low_thresh, low_ratio=low_ratio, mid_thresh=mid_thresh,
              mid_ratio=mid_ratio, high_thresh=high_thresh, high_ratio=high_ratio)
    plan = make_plan(st, chunk.frame_rate, chunk.channels, low_crossover, high_crossover)
    pcm = _segment_pcm(chunk)
    out = get_engine().multiband(pcm, plan)
    n = ms_framing(out.shape[0], chunk.frame_rate)      # pydub overlay re-slices by milliseconds
    if n < out.shape[0]:
        out = out[:n]
    elif n > out.shape[0]:
        out = np.concatenate([out, np.zeros((n - out.shape[0],) + out.shape[1:], dtype=np.int16)])
    return chunk._spawn(np.ascontiguousarray(out).tobytes())


def compress_dynamic_range(seg, threshold=-20.0, ratio=4.0, attack=5.0, release=50.0):
    """pydub.effects.compress_dynamic_range (imported by the reference at ENG:8)."""
    band = make_band(seg.frame_rate, threshold, ratio, attack, release)
    return seg._spawn(get_engine().compress_dynamic_range(_segment_pcm(seg), band).tobytes())


def normalize_to_lufs(samples, sample_rate, target_lufs=-14.0):
    """ENG:212-222.  Measured in float32, the dtype the reference chain always passes (ENG:82)."""
    if samples.shape[0] < 0.4 * sample_rate:
        raise ValueError("Audio must have length greater than the block size.")
    out, loud, _gain = get_engine().normalize_to_lufs(samples, sample_rate, target_lufs, kweight_biquads(sample_rate))
    print(f"Measured {loud:.2f} LUFS; applying {target_lufs - loud:.2f} dB of gain...")
    return out


def soft_limiter(samples, threshold=0.98):
    """ENG:224-227: in place, returns its argument."""
    samples[...] = get_engine().soft_limiter(samples, threshold)
    return samples
