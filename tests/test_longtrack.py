"""One long track split along time over several ranks (BASELINE config 4): the host-side
orchestration (chunk-aligned partition, loudness halos, SUM all-reduce of the block energies) must
reproduce the single-process result exactly.  CPU tests run it under gloo with the oracle as the
arithmetic; the GPU tests run the CUDA engine with N host threads emulating N ranks on one device."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:
    sys.path.insert(0, HERE)

from b200master import longtrack, synth  # noqa: E402

SETTINGS = dict(bass_boost=4.0, mid_cut=3.0, presence_boost=1.0, treble_boost=3.0, width=1.2, multiband=True, lufs=-14.0)


def test_partition_is_chunk_aligned_and_complete():
    for frames, rate, world in [(7200 * 96000, 96000, 8), (8640000, 48000, 4), (44100 * 61 + 13, 44100, 2),
                                (1000, 48000, 2), (48000 * 31, 48000, 8), (0, 48000, 3)]:
        sl = longtrack.partition(frames, rate, world)
        assert len(sl) == world
        total = sum(s.out_frames for s in sl)
        assert total == longtrack.ms_framing(frames, rate)
        pos = 0
        for s in sl:
            if s.out_frames:
                assert s.abs_offset == pos and s.abs_offset % (30 * rate) == 0      # ENG:48: slices start on chunk boundaries
                assert 0 <= s.in_frames <= s.out_frames or s.in_frames <= frames - s.abs_offset
                pos += s.out_frames
        sizes = [s.chunk1 - s.chunk0 for s in sl]
        assert max(sizes) - min(sizes) <= 1


def _gloo_worker(rank, world, port_no, rate, seconds, settings, ret):
    import torch.distributed as dist
    from oracle_ops import OracleOps
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pcm = synth.make_track(7, seconds, rate)
        me = longtrack.partition(pcm.shape[0], rate, world)[rank]
        local = torch.from_numpy(pcm[me.abs_offset:me.abs_offset + me.in_frames].copy())
        out, info = longtrack.master_time_split(local, pcm.shape[0], rate, OracleOps(rate, 2, settings),
                                                longtrack.DistComm(), rank, world)
        ret[rank] = (out.numpy(), info["loudness"], info["gain"])
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("settings", [SETTINGS, dict(SETTINGS, multiband=False, lufs=None)])
def test_time_split_gloo_world2_matches_single_process(settings):
    """world_size 2 over gloo: slices + halos + all-reduce == the oracle's whole-track result."""
    import torch.multiprocessing as mp
    from oracle import port
    rate, seconds, world = 12000, 61.5, 2         # three 30-s chunks: ranks get 2 + 1
    if settings.get("multiband"):
        settings = dict(settings, presence_boost=0.0, treble_boost=0.0)     # 12 kHz rate: keep the doubled EQ centres below Nyquist
    else:
        settings = dict(settings, presence_boost=0.0, treble_boost=0.0)
    port.build_c()
    pcm = synth.make_track(7, seconds, rate)
    ref, rinfo = port.master(pcm, rate, settings)
    ret = mp.Manager().dict()
    port_no = 29500 + os.getpid() % 2000
    mp.spawn(_gloo_worker, args=(world, port_no, rate, seconds, settings, ret), nprocs=world, join=True)
    got = np.concatenate([ret[r][0] for r in range(world)])
    assert np.array_equal(got, ref)
    if settings.get("lufs") is not None:
        assert ret[0][1] == ret[1][1] == rinfo["loudness"]
        assert ret[0][2] == ret[1][2]


def test_thread_comm_three_ranks_oracle():
    """The same orchestration with 3 ranks as host threads (the emulation the GPU test uses)."""
    from oracle import port
    from oracle_ops import OracleOps
    rate, seconds, world = 8000, 95.0, 3
    st = dict(SETTINGS, presence_boost=0.0, treble_boost=0.0, multiband=False)
    pcm = synth.make_track(9, seconds, rate)
    ref, rinfo = port.master(pcm, rate, st)

    def fn(rank, comm):
        me = longtrack.partition(pcm.shape[0], rate, world)[rank]
        local = torch.from_numpy(pcm[me.abs_offset:me.abs_offset + me.in_frames].copy())
        return longtrack.master_time_split(local, pcm.shape[0], rate, OracleOps(rate, 2, st), comm, rank, world)

    res = longtrack.run_threaded(world, fn)
    assert np.array_equal(np.concatenate([r[0].numpy() for r in res]), ref)
    assert all(r[1]["loudness"] == rinfo["loudness"] for r in res)


@pytest.mark.gpu
@pytest.mark.parametrize("rate,seconds,world", [(48000, 95.0, 3), (44100, 61.3, 2), (96000, 125.0, 4)])
def test_time_split_on_gpu_matches_one_gpu(rate, seconds, world):
    """CUDA engine, N ranks emulated by N host threads (one handle each) on one device: output and
    loudness are bit-identical to b200m_master_batch on the whole track."""
    from b200master import Engine, get_engine
    pcm = synth.make_track(21, seconds, rate)
    st = dict(SETTINGS, saturation=20)
    whole, winfo = get_engine(0).master([pcm], rate, st)
    engines = [Engine(0) for _ in range(world)]

    def fn(rank, comm):
        me = longtrack.partition(pcm.shape[0], rate, world)[rank]
        local = torch.from_numpy(pcm[me.abs_offset:me.abs_offset + me.in_frames].copy()).cuda()
        ops = longtrack.EngineOps(engines[rank], rate, 2, st)
        out, info = longtrack.master_time_split(local, pcm.shape[0], rate, ops, comm, rank, world)
        torch.cuda.synchronize()
        return out.cpu().numpy(), info

    res = longtrack.run_threaded(world, fn)
    got = np.concatenate([r[0] for r in res])
    assert np.array_equal(got, whole[0])
    assert all(r[1]["loudness"] == winfo[0]["loudness"] and r[1]["gain"] == winfo[0]["gain"] for r in res)
    for e in engines:
        e.close()


@pytest.mark.gpu
def test_time_split_s24_input_on_gpu():
    """cfg4's input format: packed 24-bit PCM is staged to the 16-bit domain per slice (declared
    extension) and the split result equals the one-GPU result on the staged track."""
    from b200master import Engine, get_engine
    rate, seconds, world = 96000, 64.0, 2
    pcm16 = synth.make_track(22, seconds, rate)
    low = ((np.arange(pcm16.size, dtype=np.int64) * 37 + 11) & 0xff).astype(np.uint8).reshape(pcm16.shape)
    raw = np.stack([low, (pcm16.view(np.uint16) & 0xff).astype(np.uint8), (pcm16.view(np.uint16) >> 8).astype(np.uint8)], axis=-1)   # (N, 2, 3) little endian
    whole, winfo = get_engine(0).master([pcm16], rate, SETTINGS)
    engines = [Engine(0) for _ in range(world)]

    def fn(rank, comm):
        me = longtrack.partition(pcm16.shape[0], rate, world)[rank]
        local = torch.from_numpy(raw[me.abs_offset:me.abs_offset + me.in_frames].reshape(-1).copy()).cuda()
        ops = longtrack.EngineOps(engines[rank], rate, 2, SETTINGS)
        out, info = longtrack.master_time_split(local, pcm16.shape[0], rate, ops, comm, rank, world, fmt=1)
        torch.cuda.synchronize()
        return out.cpu().numpy(), info

    res = longtrack.run_threaded(world, fn)
    assert np.array_equal(np.concatenate([r[0] for r in res]), whole[0])
    assert all(r[1]["loudness"] == winfo[0]["loudness"] for r in res)
    for e in engines:
        e.close()


@pytest.mark.gpu
def test_stage_pcm_formats():
    """Declared extension: packed s24 keeps the high-order 16 bits (pydub set_sample_width(2) =
    audioop.lin2lin), float32 goes through the reference's quantiser (ENG:123-126)."""
    import audioop
    from b200master import get_engine
    from oracle import port
    eng = get_engine(0)
    rng = np.random.default_rng(5)
    for n in (1, 2, 3, 4, 5, 4096, 100003):
        s32 = rng.integers(-2 ** 23, 2 ** 23, size=n, dtype=np.int64)
        raw = np.zeros((n, 3), dtype=np.uint8)
        raw[:, 0] = s32 & 0xff; raw[:, 1] = (s32 >> 8) & 0xff; raw[:, 2] = (s32 >> 16) & 0xff
        ref = np.frombuffer(audioop.lin2lin(raw.tobytes(), 3, 2), dtype=np.int16)
        d_in = torch.from_numpy(raw.reshape(-1).copy()).cuda()
        d_out = torch.empty(n, dtype=torch.int16, device="cuda")
        eng.stage_pcm(d_in, 1, n, d_out)
        torch.cuda.synchronize()
        assert np.array_equal(d_out.cpu().numpy(), ref)
        f = (rng.standard_normal(n) * 0.7).astype(np.float32)
        f[:1] = 1.0
        d_f = torch.from_numpy(f).cuda()
        eng.stage_pcm(d_f, 2, n, d_out)
        torch.cuda.synchronize()
        assert np.array_equal(d_out.cpu().numpy(), port.float_to_pcm16(f))


def _s24_of(pcm16):
    """Packed little-endian 24-bit PCM whose high-order 16 bits are ``pcm16`` (low byte: deterministic noise)."""
    low = ((np.arange(pcm16.size, dtype=np.int64) * 37 + 11) & 0xff).astype(np.uint8).reshape(pcm16.shape)
    return np.stack([low, (pcm16.view(np.uint16) & 0xff).astype(np.uint8), (pcm16.view(np.uint16) >> 8).astype(np.uint8)], axis=-1)


@pytest.mark.gpu
def test_time_split_s24_vs_the_oracle():
    """cfg4 in miniature against the ORACLE (not against the library): a 95-s 96 kHz s24 track over three ranks
    (host threads, one handle each); the checker is the CPU chain on the staged 16-bit track (audioop.lin2lin =
    the high-order 16 bits, which is what the declared extension promises)."""
    from b200master import Engine
    from oracle import port
    rate, seconds, world = 96000, 95.0, 3
    st = dict(SETTINGS, saturation=25)
    pcm16 = synth.make_track(23, seconds, rate)
    raw = _s24_of(pcm16)
    ref, rinfo = port.master(pcm16, rate, st)
    engines = [Engine(0) for _ in range(world)]

    def fn(rank, comm):
        me = longtrack.partition(pcm16.shape[0], rate, world)[rank]
        local = torch.from_numpy(raw[me.abs_offset:me.abs_offset + me.in_frames].reshape(-1).copy()).cuda()
        out, info = longtrack.master_time_split(local, pcm16.shape[0], rate, longtrack.EngineOps(engines[rank], rate, 2, st), comm, rank, world, fmt=1)
        torch.cuda.synchronize()
        return out.cpu().numpy(), info

    res = longtrack.run_threaded(world, fn)
    assert np.array_equal(np.concatenate([r[0] for r in res]), ref)
    assert all(abs(r[1]["loudness"] - rinfo["loudness"]) <= 1e-12 for r in res)
    for e in engines:
        e.close()


def _nccl_worker(rank, world, port_no, rate, seconds, settings, ret):
    import torch.distributed as dist
    from b200master import Engine
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        pcm16 = synth.make_track(24, seconds, rate)
        raw = _s24_of(pcm16)
        me = longtrack.partition(pcm16.shape[0], rate, world)[rank]
        local = torch.from_numpy(raw[me.abs_offset:me.abs_offset + me.in_frames].reshape(-1).copy()).cuda()
        eng = Engine(rank)
        out, info = longtrack.master_time_split(local, pcm16.shape[0], rate, longtrack.EngineOps(eng, rate, 2, settings),
                                                longtrack.DistComm(), rank, world, fmt=1)
        torch.cuda.synchronize()
        ret[rank] = (out.cpu().numpy(), info["loudness"], info["gain"])
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
def test_time_split_nccl_two_gpus_vs_the_oracle():
    """The same over REAL NCCL ranks, one process per GPU (skipped on a one-GPU box): a 10-minute 96 kHz s24 track
    split over two GPUs, halos and block energies over NVLink, against the oracle on the staged 16-bit track."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (bench.py's cfg4 extra runs the NCCL path at every N >= 2 and checks it against one GPU)")
    import torch.multiprocessing as mp
    from oracle import port
    rate, seconds, world = 96000, 600.0, 2
    st = dict(SETTINGS, saturation=25)
    ref, rinfo = port.master(synth.make_track(24, seconds, rate), rate, st)
    ret = mp.Manager().dict()
    mp.spawn(_nccl_worker, args=(world, 29500 + os.getpid() % 2000, rate, seconds, st, ret), nprocs=world, join=True)
    got = np.concatenate([ret[r][0] for r in range(world)])
    d = np.abs(got.astype(np.int32) - ref.astype(np.int32))
    assert d.max() <= 1 and np.mean(d != 0) <= 1e-5            # full-length tolerance (tests/test_gpu_parity.py docstring)
    assert ret[0][1] == ret[1][1] and abs(ret[0][1] - rinfo["loudness"]) <= 1e-9
