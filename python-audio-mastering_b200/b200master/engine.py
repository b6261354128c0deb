"""``Engine``: the Python face of the C-ABI (one handle = one device + one stream).

Host buffers are numpy arrays (or pinned torch CPU tensors); device buffers are torch
CUDA tensors whose ``data_ptr()`` is handed to the library.  torch is used for device
memory and streams only -- every sample is computed by libb200master's kernels.
"""
from __future__ import annotations

import ctypes as C
import threading

import numpy as np

from . import lib as L
from .plan import make_plan

_engines = {}
_lock = threading.Lock()


def ms_framing(n_frames: int, rate: int) -> int:
    """Frames the reference actually processes: pydub derives chunk ends from the track
    length ROUNDED to milliseconds (``len(audio)``, ENG:51-54), so up to half a millisecond of
    tail is dropped or zero padded."""
    len_ms = round(1000 * (float(n_frames) / rate))
    return int(len_ms * (rate / 1000.0))


def _ptr(a):
    return a.ctypes.data if isinstance(a, np.ndarray) else a.data_ptr()


class Engine:
    def __init__(self, device: int = 0):
        self._lib = L.load()
        h = C.c_void_p()
        rc = self._lib.b200m_create(int(device), C.byref(h))
        if rc != L.OK:
            raise RuntimeError("b200m_create failed: " + (self._lib.b200m_last_error(None) or b"").decode())
        self._h = h
        self.device = int(device)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.b200m_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        L.check(self._lib, self._h, rc)

    # -- handle plumbing ---------------------------------------------------------------
    def set_stream(self, cuda_stream_ptr: int):
        self._ck(self._lib.b200m_set_stream(self._h, C.c_void_p(cuda_stream_ptr)))

    def synchronize(self):
        self._ck(self._lib.b200m_synchronize(self._h))

    def set_workspace_limit(self, nbytes: int):
        self._ck(self._lib.b200m_set_workspace_limit(self._h, int(nbytes)))

    def launch_count(self) -> int:
        return int(self._lib.b200m_launch_count(self._h))

    def set_profiling(self, on: bool):
        self._ck(self._lib.b200m_set_profiling(self._h, int(on)))

    def reset_profile(self):
        self._ck(self._lib.b200m_reset_profile(self._h))

    def kernel_time_ms(self, name: str):
        tot, n = C.c_double(), C.c_int64()
        self._ck(self._lib.b200m_kernel_time_ms(self._h, name.encode(), C.byref(tot), C.byref(n)))
        return tot.value, n.value

    def set_recur_tiling(self, tile_frames: int = 0, warm_frames: int = 0, rounds: int = -1):
        self._ck(self._lib.b200m_set_recur_tiling(self._h, int(tile_frames), int(warm_frames), int(rounds)))

    def set_segment_tiles(self, chain_tiles: int = 0, kweight_tiles: int = 0):
        self._ck(self._lib.b200m_set_segment_tiles(self._h, int(chain_tiles), int(kweight_tiles)))

    def set_pipeline(self, on: bool = True):
        self._ck(self._lib.b200m_set_pipeline(self._h, int(on)))

    def set_chain_kernel(self, mode: int = 0):
        self._ck(self._lib.b200m_set_chain_kernel(self._h, int(mode)))

    def set_pipeline_shape(self, groups: int = 0, compute_streams: int = 0):
        self._ck(self._lib.b200m_set_pipeline_shape(self._h, int(groups), int(compute_streams)))

    def recur_stats(self, reset: bool = False):
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        self._ck(self._lib.b200m_recur_stats(self._h, C.byref(a), C.byref(b), C.byref(c), int(reset)))
        return {"wrong_tiles": a.value, "rerun_frames": b.value, "round_repairs": c.value}

    # -- the whole path ------------------------------------------------------------------
    def master_raw(self, pcm_in, in_on_device, in_offsets, in_frames, out_frames, plans, plan_index,
                   pcm_out, out_on_device, want_loudness=True, targets=None, fmt=L.FMT_S16):
        """Thin wrapper over ``b200m_master_batch`` (``b200m_master_batch_targets`` when ``targets`` is
        given: pcm_out then holds ``len(targets)`` copies of the batch output, target-major, and the
        returned gains have shape (len(targets), n)).  pcm_in / pcm_out: numpy int16 arrays or torch
        tensors; offsets / frames / plan_index: sequences.  ``fmt``: format of pcm_in (``lib.FMT_S16``; packed
        s24 and float32 are staged to the reference's 16-bit domain first, see ``b200m_stage_pcm``); the
        output is always int16 (ENG:125)."""
        n = len(in_frames)
        off = np.ascontiguousarray(in_offsets, dtype=np.int64)
        inf = np.ascontiguousarray(in_frames, dtype=np.int64)
        outf = np.ascontiguousarray(out_frames, dtype=np.int64)
        pidx = np.ascontiguousarray(plan_index, dtype=np.int32)
        parr = (L.Plan * len(plans))(*plans)
        loud = np.empty(n, dtype=np.float64) if want_loudness else None
        if targets is not None:
            tg = np.ascontiguousarray(targets, dtype=np.float64)
            gain = np.empty((len(tg), n), dtype=np.float64) if want_loudness else None
            self._ck(self._lib.b200m_master_batch_targets(
                self._h, C.c_void_p(_ptr(pcm_in)), int(in_on_device), int(fmt), n,
                C.c_void_p(off.ctypes.data), C.c_void_p(inf.ctypes.data), C.c_void_p(outf.ctypes.data),
                parr, len(plans), C.c_void_p(pidx.ctypes.data), C.c_void_p(tg.ctypes.data), len(tg),
                C.c_void_p(_ptr(pcm_out)), int(out_on_device),
                C.c_void_p(loud.ctypes.data) if want_loudness else None,
                C.c_void_p(gain.ctypes.data) if want_loudness else None))
            return loud, gain
        gain = np.empty(n, dtype=np.float64) if want_loudness else None
        self._ck(self._lib.b200m_master_batch(
            self._h, C.c_void_p(_ptr(pcm_in)), int(in_on_device), int(fmt), n,
            C.c_void_p(off.ctypes.data), C.c_void_p(inf.ctypes.data), C.c_void_p(outf.ctypes.data),
            parr, len(plans), C.c_void_p(pidx.ctypes.data),
            C.c_void_p(_ptr(pcm_out)), int(out_on_device),
            C.c_void_p(loud.ctypes.data) if want_loudness else None,
            C.c_void_p(gain.ctypes.data) if want_loudness else None))
        return loud, gain

    def master(self, tracks, rate: int, settings):
        """Master a list of int16 numpy tracks ((N, 2) stereo or (N,) mono, same rate and
        channel count).  ``settings``: one dict for all tracks or one per track.
        Returns (list of int16 outputs, list of info dicts)."""
        single = isinstance(tracks, np.ndarray)
        tracks = [tracks] if single else list(tracks)
        if not tracks:
            return [], []
        ch = 1 if tracks[0].ndim == 1 else tracks[0].shape[1]
        sets = [settings] * len(tracks) if isinstance(settings, dict) else list(settings)
        plans, index, keys = [], [], {}
        for s in sets:
            k = repr(sorted((s or {}).items(), key=lambda kv: kv[0]))
            if k not in keys:
                keys[k] = len(plans)
                plans.append(make_plan(s, rate, ch))
            index.append(keys[k])
        in_frames = [int(t.shape[0]) for t in tracks]
        out_frames = [ms_framing(n, rate) for n in in_frames]
        flat = np.ascontiguousarray(np.concatenate([np.ascontiguousarray(t, dtype=np.int16).reshape(-1) for t in tracks]))
        offsets = np.concatenate([[0], np.cumsum(in_frames)[:-1]])
        out = np.empty(sum(out_frames) * ch, dtype=np.int16)
        if out.size == 0:
            return [np.zeros((0, ch) if ch > 1 else (0,), np.int16) for _ in tracks], [{} for _ in tracks]
        loud, gain = self.master_raw(flat, False, offsets, in_frames, out_frames, plans, index, out, False)
        outs, infos, pos = [], [], 0
        for i, n in enumerate(out_frames):
            o = out[pos * ch:(pos + n) * ch]
            outs.append(o.reshape(-1, ch) if ch > 1 else o)
            infos.append({"loudness": float(loud[i]) if plans[index[i]].has_lufs else None,
                          "gain": float(gain[i]) if plans[index[i]].has_lufs else None})
            pos += n
        return outs, infos

    def master_wav(self, tracks, rate: int, settings):
        """``master`` with ENG:96-99 folded in: returns (list of uint8 arrays, each a complete RIFF/WAVE
        file image = what ``export(format="wav")`` would write, views into ONE host buffer; infos).
        The GPU writes the 44-byte headers ahead of the samples; nothing sweeps the PCM on the host."""
        tracks = list(tracks)
        if not tracks:
            return [], []
        ch = 1 if tracks[0].ndim == 1 else tracks[0].shape[1]
        sets = [settings] * len(tracks) if isinstance(settings, dict) else list(settings)
        plans, index, keys = [], [], {}
        for s in sets:
            k = repr(sorted((s or {}).items(), key=lambda kv: kv[0]))
            if k not in keys:
                keys[k] = len(plans)
                plans.append(make_plan(s, rate, ch))
            index.append(keys[k])
        n = len(tracks)
        fw = 2 * ch                                            # bytes per frame
        in_frames = [int(t.shape[0]) for t in tracks]
        out_frames = [ms_framing(f, rate) for f in in_frames]
        # layout: samples start at multiples of 16 bytes, the header sits in the 44 bytes before them
        offs, pos = [], 0
        for f in out_frames:
            start = (pos + 44 + 15) // 16 * 16
            offs.append(start // fw)
            pos = start + f * fw
        buf = np.zeros(pos, dtype=np.uint8)
        if sum(out_frames) == 0:                               # nothing for the GPU to do: header-only files (what export() writes for empty audio)
            for o in offs:
                buf[o * fw - 44:o * fw] = np.frombuffer(self.wav_header(rate, ch, 0), dtype=np.uint8)
            return [buf[o * fw - 44:o * fw] for o in offs], [{"loudness": None, "gain": None} for _ in tracks]
        flat = np.ascontiguousarray(np.concatenate([np.ascontiguousarray(t, dtype=np.int16).reshape(-1) for t in tracks]))
        off = np.ascontiguousarray(np.concatenate([[0], np.cumsum(in_frames)[:-1]]), dtype=np.int64)
        inf = np.ascontiguousarray(in_frames, dtype=np.int64)
        outf = np.ascontiguousarray(out_frames, dtype=np.int64)
        oo = np.ascontiguousarray(offs, dtype=np.int64)
        pidx = np.ascontiguousarray(index, dtype=np.int32)
        loud, gain = np.empty(n, dtype=np.float64), np.empty(n, dtype=np.float64)
        parr = (L.Plan * len(plans))(*plans)
        self._ck(self._lib.b200m_master_batch_wav(
            self._h, C.c_void_p(flat.ctypes.data), 0, L.FMT_S16, n, C.c_void_p(off.ctypes.data), C.c_void_p(inf.ctypes.data),
            C.c_void_p(outf.ctypes.data), parr, len(plans), C.c_void_p(pidx.ctypes.data),
            C.c_void_p(oo.ctypes.data), C.c_void_p(buf.ctypes.data), 0,
            C.c_void_p(loud.ctypes.data), C.c_void_p(gain.ctypes.data)))
        images = [buf[o * fw - 44:(o + f) * fw] for o, f in zip(offs, out_frames)]
        infos = [{"loudness": float(loud[i]) if plans[index[i]].has_lufs else None,
                  "gain": float(gain[i]) if plans[index[i]].has_lufs else None} for i in range(n)]
        return images, infos

    def wav_header(self, rate: int, channels: int, frames: int) -> bytes:
        out = (C.c_ubyte * 44)()
        self._ck(self._lib.b200m_wav_header(int(rate), int(channels), int(frames), out))
        return bytes(out)

    def master_targets(self, tracks, rate: int, settings, targets):
        """``master`` for several loudness targets at once (a preset x loudness sweep): the chain and
        the loudness measurement run once per track, gain / limiter / final cast once per target.
        ``settings["lufs"]`` is ignored.  Returns (outs[target][track], infos[track]) with
        ``infos[t]["gain"]`` a list over targets."""
        tracks = list(tracks)
        targets = [float(t) for t in targets]
        if not targets:
            raise ValueError("master_targets: at least one loudness target")
        if not tracks:
            return [[] for _ in targets], []
        ch = 1 if tracks[0].ndim == 1 else tracks[0].shape[1]
        sets = [settings] * len(tracks) if isinstance(settings, dict) else list(settings)
        plans, index, keys = [], [], {}
        for s in sets:
            s = dict(s or {}, lufs=targets[0])
            k = repr(sorted(s.items(), key=lambda kv: kv[0]))
            if k not in keys:
                keys[k] = len(plans)
                plans.append(make_plan(s, rate, ch))
            index.append(keys[k])
        n = len(tracks)
        in_frames = [int(t.shape[0]) for t in tracks]
        out_frames = [ms_framing(f, rate) for f in in_frames]
        flat = np.ascontiguousarray(np.concatenate([np.ascontiguousarray(t, dtype=np.int16).reshape(-1) for t in tracks]))
        off = np.ascontiguousarray(np.concatenate([[0], np.cumsum(in_frames)[:-1]]), dtype=np.int64)
        inf = np.ascontiguousarray(in_frames, dtype=np.int64)
        outf = np.ascontiguousarray(out_frames, dtype=np.int64)
        pidx = np.ascontiguousarray(index, dtype=np.int32)
        tg = np.ascontiguousarray(targets, dtype=np.float64)
        total = int(sum(out_frames))
        out = np.empty(len(targets) * total * ch, dtype=np.int16)
        loud = np.empty(n, dtype=np.float64)
        gain = np.empty(len(targets) * n, dtype=np.float64)
        parr = (L.Plan * len(plans))(*plans)
        self._ck(self._lib.b200m_master_batch_targets(
            self._h, C.c_void_p(flat.ctypes.data), 0, L.FMT_S16, n, C.c_void_p(off.ctypes.data), C.c_void_p(inf.ctypes.data),
            C.c_void_p(outf.ctypes.data), parr, len(plans), C.c_void_p(pidx.ctypes.data),
            C.c_void_p(tg.ctypes.data), len(targets), C.c_void_p(out.ctypes.data), 0,
            C.c_void_p(loud.ctypes.data), C.c_void_p(gain.ctypes.data)))
        outs = []
        for k in range(len(targets)):
            pos, row = k * total, []
            for f in out_frames:
                o = out[pos * ch:(pos + f) * ch]
                row.append(o.reshape(-1, ch) if ch > 1 else o)
                pos += f
            outs.append(row)
        infos = [{"loudness": float(loud[t]), "gain": [float(gain[k * n + t]) for k in range(len(targets))]} for t in range(n)]
        return outs, infos

    # -- time slices of one long track (device tensors; see longtrack.py) ------------------------
    def stage_pcm(self, pcm_dev, fmt: int, n_samples: int, out_dev):
        self._ck(self._lib.b200m_stage_pcm(self._h, C.c_void_p(_ptr(pcm_dev)), int(fmt), int(n_samples), C.c_void_p(_ptr(out_dev))))

    def slice_halo(self, plan, abs_offset: int):
        a, b = C.c_int64(), C.c_int64()
        self._ck(self._lib.b200m_slice_halo(C.byref(plan), int(abs_offset), C.byref(a), C.byref(b)))
        return a.value, b.value

    def slice_chain(self, pcm_dev, in_frames: int, out_frames: int, plan, proc_dev):
        self._ck(self._lib.b200m_slice_chain(self._h, C.c_void_p(_ptr(pcm_dev)), int(in_frames), int(out_frames),
                                             C.byref(plan), C.c_void_p(_ptr(proc_dev))))

    def slice_energies(self, proc_ext_dev, ext_frames, halo_before, local_frames, abs_offset, track_frames, plan, z_dev):
        j0, nb = C.c_int32(), C.c_int32()
        self._ck(self._lib.b200m_slice_energies(self._h, C.c_void_p(_ptr(proc_ext_dev)), int(ext_frames), int(halo_before),
                                                int(local_frames), int(abs_offset), int(track_frames), C.byref(plan),
                                                C.c_void_p(_ptr(z_dev)), C.byref(j0), C.byref(nb)))
        return j0.value, nb.value

    def track_blocks(self, track_frames: int, rate: int) -> int:
        return int(self._lib.b200m_track_blocks(int(track_frames), int(rate)))

    def gate(self, z_dev, n_blocks: int, plan):
        loud, gain = C.c_double(), C.c_double()
        self._ck(self._lib.b200m_gate(self._h, C.c_void_p(_ptr(z_dev)), int(n_blocks), C.byref(plan), C.byref(loud), C.byref(gain)))
        return loud.value, gain.value

    def slice_final(self, proc_dev, frames: int, plan, gain, out_dev):
        self._ck(self._lib.b200m_slice_final(self._h, C.c_void_p(_ptr(proc_dev)), int(frames), C.byref(plan),
                                             int(gain is not None), float(gain if gain is not None else 1.0),
                                             C.c_void_p(_ptr(out_dev))))

    # -- stage-level helpers (numpy in, numpy out) -----------------------------------------
    def pcm16_to_float(self, pcm: np.ndarray) -> np.ndarray:
        pcm = np.ascontiguousarray(pcm, dtype=np.int16)
        out = np.empty(pcm.shape, dtype=np.float32)
        self._ck(self._lib.b200m_pcm16_to_float(self._h, pcm.ctypes.data, pcm.size, out.ctypes.data))
        return out

    def float_to_pcm16(self, x: np.ndarray) -> np.ndarray:
        x = self._as_float(x)
        out = np.empty(x.shape, dtype=np.int16)
        self._ck(self._lib.b200m_float_to_pcm16(self._h, x.ctypes.data, int(x.dtype == np.float64), x.size, out.ctypes.data))
        return out

    @staticmethod
    def _as_float(x):
        x = np.asarray(x)
        if x.dtype not in (np.float32, np.float64):
            x = x.astype(np.float64)
        return np.ascontiguousarray(x)

    def saturation(self, x: np.ndarray, pct) -> np.ndarray:
        """ENG:128-134.  Samples that came from ENG:117-121 (every value an int16 / 2^15, which is all the
        reference's chain ever feeds it) go through the host-tabulated exciter table and are bit-exact with
        numpy; anything else takes the device's float32 tanhf (<= 1 ulp)."""
        x = np.ascontiguousarray(x, dtype=np.float32)
        out = np.empty_like(x)
        if pct == 0 or x.size == 0:
            out[...] = x
            return out
        k = x * np.float32(32768.0)
        if np.all((k >= -32768.0) & (k <= 32767.0)) and np.array_equal(k, np.rint(k)):
            from .plan import _exciter_lut
            table, key = _exciter_lut(pct)
            pcm = k.astype(np.int16)
            self._ck(self._lib.b200m_saturation_pcm(self._h, pcm.ctypes.data, pcm.size, table.ctypes.data, key, out.ctypes.data))
            return out
        self._ck(self._lib.b200m_saturation(self._h, x.ctypes.data, x.size, float(pct), out.ctypes.data))
        return out

    def stereo_width(self, x: np.ndarray, width) -> np.ndarray:
        x = self._as_float(x)
        out = np.empty_like(x)
        self._ck(self._lib.b200m_stereo_width(self._h, x.ctypes.data, int(x.dtype == np.float64), x.shape[0], float(width), out.ctypes.data))
        return out

    def sosfilt(self, sections, x: np.ndarray) -> np.ndarray:
        """Cascade of ``lib.Biquad`` sections from zero state along axis 0 -> float64."""
        x = self._as_float(x)
        ch = 1 if x.ndim == 1 else x.shape[1]
        out = np.empty(x.shape, dtype=np.float64)
        arr = (L.Biquad * max(len(sections), 1))(*sections)
        self._ck(self._lib.b200m_sosfilt(self._h, arr, len(sections), x.ctypes.data, int(x.dtype == np.float64),
                                         x.shape[0], ch, out.ctypes.data))
        return out

    def multiband(self, pcm: np.ndarray, plan: L.Plan) -> np.ndarray:
        pcm = np.ascontiguousarray(pcm, dtype=np.int16)
        out = np.empty_like(pcm)
        self._ck(self._lib.b200m_multiband(self._h, C.byref(plan), pcm.ctypes.data, pcm.shape[0], out.ctypes.data))
        return out

    def compress_dynamic_range(self, pcm: np.ndarray, band: L.Band, debug=False):
        pcm = np.ascontiguousarray(pcm, dtype=np.int16)
        ch = 1 if pcm.ndim == 1 else pcm.shape[1]
        out = np.empty_like(pcm)
        n = pcm.shape[0]
        att = np.empty(n, dtype=np.float64) if debug else None
        rms = np.empty(n, dtype=np.uint32) if debug else None
        self._ck(self._lib.b200m_compress_dynamic_range(
            self._h, pcm.ctypes.data, n, ch, C.byref(band), out.ctypes.data,
            att.ctypes.data if debug else None, rms.ctypes.data if debug else None))
        return (out, att, rms) if debug else out

    def integrated_loudness(self, mono: np.ndarray, rate: int, kw) -> float:
        mono = np.ascontiguousarray(mono, dtype=np.float32)
        res = C.c_double()
        arr = (L.Biquad * 2)(*kw)
        self._ck(self._lib.b200m_integrated_loudness(self._h, arr, mono.ctypes.data, mono.shape[0], int(rate), C.byref(res)))
        return res.value

    def normalize_to_lufs(self, x: np.ndarray, rate: int, target: float, kw):
        x = np.ascontiguousarray(x, dtype=np.float32)
        ch = 1 if x.ndim == 1 else x.shape[1]
        out = np.empty(x.shape, dtype=np.float64)
        loud, gain = C.c_double(), C.c_double()
        arr = (L.Biquad * 2)(*kw)
        self._ck(self._lib.b200m_normalize_to_lufs(self._h, arr, x.ctypes.data, x.shape[0], ch, int(rate), float(target),
                                                   out.ctypes.data, C.byref(loud), C.byref(gain)))
        return out, loud.value, gain.value

    def soft_limiter(self, x: np.ndarray, threshold=0.98) -> np.ndarray:
        x = self._as_float(x)
        out = np.empty_like(x)
        self._ck(self._lib.b200m_soft_limiter(self._h, x.ctypes.data, int(x.dtype == np.float64), x.size, float(threshold), out.ctypes.data))
        return out


def get_engine(device: int = 0) -> Engine:
    """Process-wide engine per device (the reference's callers hold no state)."""
    with _lock:
        e = _engines.get(device)
        if e is None:
            e = _engines[device] = Engine(device)
        return e
