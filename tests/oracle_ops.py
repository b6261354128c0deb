"""CPU stand-in for ``b200master.longtrack.EngineOps`` built on the oracle (TEST INFRASTRUCTURE):
lets the time-split orchestration (partition, halos, all-reduce) run under gloo without a GPU."""
import numpy as np
import torch

from oracle import port, thirdparty


class OracleOps:
    def __init__(self, rate, channels, settings):
        self.rate, self.ch, self.settings = rate, channels, dict(settings)
        self.has_lufs = settings.get("lufs") is not None

    def empty(self, frames, dtype=torch.int16):
        return torch.zeros((frames, self.ch) if self.ch > 1 else (frames,), dtype=dtype)

    def zeros_f64(self, n):
        return torch.zeros(n, dtype=torch.float64)

    def halo(self, abs_offset):
        # K-weighting settles (below 2^-64) within ~0.5 s; one 400 ms block after the slice
        return (0 if abs_offset == 0 else min(abs_offset, self.rate)), int(0.4 * self.rate) + 1

    def n_blocks(self, track_frames):
        T = track_frames / self.rate
        return int(np.round((T - 0.4) / (0.4 * 0.25)) + 1)          # pyloudnorm numBlocks

    def chain(self, pcm_i16, in_frames, out_frames):
        pcm = pcm_i16.numpy()[:in_frames]
        chunk = int(port.CHUNK_MS * (self.rate / 1000.0))
        outs = []
        for s in range(0, out_frames, chunk):                          # ENG:48-54: every chunk from zero state
            e = min(s + chunk, out_frames)
            outs.append(port.process_chunk(port._take(pcm, s, e, self.rate), self.rate, self.settings))
        return torch.from_numpy(np.concatenate(outs))

    def energies(self, proc_ext, halo_before, local_frames, abs_offset, track_frames, z):
        x = port.pcm_to_float(proc_ext.numpy())
        mono = x.mean(axis=1) if x.ndim == 2 else x                    # ENG:215
        kw = port.k_weight(mono, self.rate)                            # zero state at the buffer start: decayed by the slice start
        abs0 = abs_offset - halo_before
        for j in range(self.n_blocks(track_frames)):
            lo = int(0.4 * (j * 0.25) * self.rate)
            hi = min(int(0.4 * (j * 0.25 + 1) * self.rate), track_frames)
            if abs_offset <= lo < abs_offset + local_frames:
                z[j] = float((1.0 / (0.4 * self.rate)) * np.sum(np.square(kw[lo - abs0:hi - abs0])))

    def gate(self, z, n_blocks):
        """thirdparty.Meter.integrated_loudness from the block energies on (pyloudnorm meter.py)."""
        zz = z.numpy()[:n_blocks]
        with np.errstate(divide="ignore", invalid="ignore"):
            l = -0.691 + 10.0 * np.log10(zz)
            J = [j for j in range(n_blocks) if l[j] >= -70.0]
            gr = -0.691 + 10.0 * np.log10(np.mean([zz[j] for j in J]) if J else np.nan) - 10.0
            J = [j for j in range(n_blocks) if l[j] > gr and l[j] > -70.0]
            zg = np.nan_to_num(np.mean([zz[j] for j in J]) if J else np.nan)
            loud = float(-0.691 + 10.0 * np.log10(zg))
        gain = float(10.0 ** ((self.settings["lufs"] - loud) / 20.0))   # ENG:219-220
        return loud, gain

    def final(self, proc, gain):
        x = port.pcm_to_float(proc.numpy())
        if gain is not None:
            with np.errstate(invalid="ignore", over="ignore"):
                x = x * np.float64(gain)                                # ENG:222 (float64 under NEP 50)
        return torch.from_numpy(port.float_to_pcm16(port.limiter(x)))   # ENG:88-89
