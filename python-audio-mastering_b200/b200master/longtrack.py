"""One long track split along time over several GPUs (BASELINE config 4).

The reference processes a track in 30-s chunks from zero state (ENG:48-54), so a *slice* made
of whole chunks is independent of every other slice up to ``processed_audio`` (ENG:80).  The
only whole-track coupling is the loudness measurement (ENG:82-86, 212-222):

* the K-weighting filter runs through the whole track: a slice needs the filter state at its
  first frame.  It gets a *halo* of the previous slice's processed samples and joins by
  overlap-discard, exactly like the time segments inside one GPU (``b200m_slice_halo``);
* a 400 ms block belongs to the slice that holds its first frame and may reach 400 ms into the
  next slice: a second halo, from the next slice;
* the block energies ``z_j`` of all slices are assembled with ONE ``all_reduce(SUM)``: every
  ``z_j`` is written by exactly one rank and is 0.0 elsewhere, so the sum is exact and
  order-independent; every rank then gates redundantly and obtains the same gain.

Per rank and track that is two point-to-point messages of a few hundred KB to each neighbour
and one all-reduce of ``numBlocks`` doubles (72 k for a 2-hour track): latency-bound on NVLink.
Everything else -- exciter, EQ, width, multiband, limiter -- never leaves the GPU that owns
the slice.

``master_time_split`` is written against two small interfaces so that the same orchestration
runs on NCCL + the CUDA engine (``EngineOps``) and, in the CPU tests, on gloo + the oracle.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import torch

from .engine import ms_framing
from .plan import make_plan

CHUNK_MS = 30 * 1000  # ENG:48


@dataclass
class Slice:
    rank: int
    chunk0: int          # first 30-s chunk of the slice
    chunk1: int          # one past its last chunk
    abs_offset: int      # first frame of the slice in the track
    in_frames: int       # input frames the slice reads (the track may end a little early: zero padded)
    out_frames: int      # frames the slice produces (pydub's millisecond framing applies to the track's tail)


def partition(track_frames: int, rate: int, world: int) -> List[Slice]:
    """Contiguous, chunk-aligned time ranges, as even as the 30-s chunk grid allows.  Ranks
    beyond the number of chunks get an empty slice."""
    out_total = ms_framing(track_frames, rate)

    def chunk_start(k):                                      # pydub's own ms -> frame conversion (ENG:48-54), chunk by chunk
        return int((CHUNK_MS * k) * (rate / 1000.0))

    n_chunks = 0
    while chunk_start(n_chunks) < out_total:
        n_chunks += 1
    base, rem = divmod(n_chunks, world)
    slices, c = [], 0
    for r in range(world):
        n = base + (1 if r < rem else 0)
        c0, c1 = c, c + n
        a, b = min(chunk_start(c0), out_total), min(chunk_start(c1), out_total)
        slices.append(Slice(r, c0, c1, a, max(0, min(b, track_frames) - a) if b > a else 0, b - a))
        c = c1
    return slices


class EngineOps:
    """The arithmetic of one slice on the CUDA engine (device tensors in, device tensors out)."""

    def __init__(self, engine, rate: int, channels: int, settings: dict):
        self.e, self.rate, self.ch = engine, rate, channels
        self.plan = make_plan(settings, rate, channels)
        self.has_lufs = bool(self.plan.has_lufs)
        self.device = torch.device("cuda", engine.device)

    def empty(self, frames: int, dtype=torch.int16):
        shape = (frames, self.ch) if self.ch > 1 and dtype == torch.int16 else (frames,)
        return torch.empty(shape, dtype=dtype, device=self.device)

    def zeros_f64(self, n: int):
        return torch.zeros(n, dtype=torch.float64, device=self.device)

    def halo(self, abs_offset: int):
        return self.e.slice_halo(self.plan, abs_offset)

    def n_blocks(self, track_frames: int) -> int:
        return self.e.track_blocks(track_frames, self.rate)

    def stage(self, pcm, fmt: int, frames: int):
        out = self.empty(frames)
        self.e.stage_pcm(pcm, fmt, frames * self.ch, out)
        return out

    def chain(self, pcm_i16, in_frames: int, out_frames: int):
        proc = self.empty(out_frames)
        self.e.slice_chain(pcm_i16, in_frames, out_frames, self.plan, proc)
        return proc

    def energies(self, proc_ext, halo_before: int, local_frames: int, abs_offset: int, track_frames: int, z):
        self.e.slice_energies(proc_ext, proc_ext.shape[0], halo_before, local_frames, abs_offset, track_frames, self.plan, z)

    def gate(self, z, n_blocks: int):
        return self.e.gate(z, n_blocks, self.plan)

    def final(self, proc, gain: Optional[float]):
        out = torch.empty_like(proc)
        self.e.slice_final(proc, proc.shape[0], self.plan, gain, out)
        return out

    def sync(self):
        self.e.synchronize()


class DistComm:
    """Halo exchange and the block-energy all-reduce over ``torch.distributed`` (NCCL on the GPU
    box, gloo in the CPU tests)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)

    def exchange(self, sends, recvs):
        """sends / recvs: lists of (peer_rank, tensor); tensors are contiguous.  They travel as bytes
        (NCCL has no int16 type)."""
        def raw(t):
            return t.view(torch.uint8) if t.dtype == torch.int16 else t
        ops = [self.dist.P2POp(self.dist.isend, raw(t), p, self.group) for p, t in sends if t.numel()]
        ops += [self.dist.P2POp(self.dist.irecv, raw(t), p, self.group) for p, t in recvs if t.numel()]
        if ops:
            for req in self.dist.batch_isend_irecv(ops):
                req.wait()

    def all_reduce_sum(self, t):
        if t.numel():
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)


def master_time_split(local_pcm, track_frames: int, rate: int, ops, comm, rank: int, world: int, fmt: int = 0):
    """Master this rank's slice of one long track.

    ``local_pcm``: the slice's input frames (``partition(...)[rank]``: ``in_frames`` frames from
    ``abs_offset``), interleaved, in PCM format ``fmt`` (0 = s16, 1 = packed s24, 2 = f32), on
    ``ops``' device.  Returns (int16 output of the slice, info dict with the track's loudness and
    gain -- identical on every rank)."""
    slices = partition(track_frames, rate, world)
    me = slices[rank]
    active = [s for s in slices if s.out_frames > 0]
    out_total = sum(s.out_frames for s in slices)
    # ---- ENG:48-80 on the slice: no communication --------------------------------------------
    if me.out_frames > 0:
        pcm16 = local_pcm if fmt == 0 else ops.stage(local_pcm, fmt, me.in_frames)
        proc = ops.chain(pcm16, me.in_frames, me.out_frames)
    else:
        proc = ops.empty(0)
    info = {"loudness": None, "gain": None, "slice": me}
    if not ops.has_lufs:
        return (ops.final(proc, None) if me.out_frames > 0 else proc), info
    if out_total < 0.4 * rate:
        raise ValueError("Audio must have length greater than the block size.")      # pyloudnorm valid_audio
    # ---- halos: the end of the previous slice (filter warm-up), the start of the next (last blocks) ----
    halos = {s.rank: ops.halo(s.abs_offset) for s in active}
    idx = {s.rank: i for i, s in enumerate(active)}
    hb = ha = 0
    sends, recvs = [], []
    if me.out_frames > 0:
        i = idx[rank]
        prev = active[i - 1] if i > 0 else None
        nxt = active[i + 1] if i + 1 < len(active) else None
        hb = min(halos[rank][0], me.abs_offset)
        ha = min(halos[rank][1], out_total - (me.abs_offset + me.out_frames))
        if prev is not None and hb > prev.out_frames or nxt is not None and ha > nxt.out_frames:
            raise ValueError("slices are shorter than the loudness halos: use fewer ranks for this track")
        ext = ops.empty(hb + me.out_frames + ha)
        ext[hb:hb + me.out_frames] = proc
        if prev is not None:
            recvs.append((prev.rank, ext[:hb]))
            need = min(halos[prev.rank][1], out_total - (prev.abs_offset + prev.out_frames))     # prev's halo_after = my head
            sends.append((prev.rank, proc[:need].contiguous()))
        if nxt is not None:
            recvs.append((nxt.rank, ext[hb + me.out_frames:]))
            need = min(halos[nxt.rank][0], nxt.abs_offset)                                      # next's halo_before = my tail
            sends.append((nxt.rank, proc[me.out_frames - need:].contiguous()))
    comm.exchange(sends, recvs)
    # ---- block energies of the blocks that start in this slice, then one SUM all-reduce ---------------
    nb = ops.n_blocks(out_total)
    z = ops.zeros_f64(nb)
    if me.out_frames > 0:
        ops.energies(ext, hb, me.out_frames, me.abs_offset, out_total, z)
    comm.all_reduce_sum(z)
    loud, gain = ops.gate(z, nb)
    info["loudness"], info["gain"] = loud, gain
    out = ops.final(proc, gain) if me.out_frames > 0 else proc
    return out, info


class ThreadComm:
    """N ranks as N host threads of ONE process (tests and single-GPU emulation): same two meeting
    points as ``DistComm``, implemented with a barrier.  Create one ``ThreadComm.Shared`` per job and
    one ``ThreadComm`` per rank/thread; every rank needs its own ``Engine`` (a handle is not
    re-entrant)."""

    class Shared:
        def __init__(self, world: int):
            import threading
            self.world = world
            self.barrier = threading.Barrier(world)
            self.lock = threading.Lock()
            self.mail = {}
            self.acc = None

    def __init__(self, shared: "ThreadComm.Shared", rank: int):
        self.s, self.rank, self.world = shared, rank, shared.world

    def exchange(self, sends, recvs):
        with self.s.lock:
            for peer, t in sends:
                self.s.mail[(self.rank, peer)] = t
        self.s.barrier.wait()
        for peer, t in recvs:
            if t.numel():
                t.copy_(self.s.mail[(peer, self.rank)])
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        self.s.barrier.wait()
        if self.rank == 0:
            self.s.mail.clear()
        self.s.barrier.wait()

    def all_reduce_sum(self, t):
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        with self.s.lock:
            self.s.acc = t.clone() if self.s.acc is None else self.s.acc + t
        self.s.barrier.wait()
        t.copy_(self.s.acc)
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        self.s.barrier.wait()
        if self.rank == 0:
            self.s.acc = None
        self.s.barrier.wait()


def run_threaded(world: int, fn):
    """Run ``fn(rank, comm)`` on ``world`` threads with a shared ``ThreadComm``; returns the results by rank."""
    import threading
    shared = ThreadComm.Shared(world)
    results, errors = [None] * world, []

    def body(r):
        try:
            results[r] = fn(r, ThreadComm(shared, r))
        except BaseException as e:      # release the others instead of dead-locking the barrier
            errors.append(e)
            shared.barrier.abort()

    threads = [threading.Thread(target=body, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return results
