#!/usr/bin/env python
"""BASELINE config 4: ONE long 96 kHz 24-bit stereo track split along time over N GPUs.

    python scripts/bench_longtrack.py [--seconds 7200] [--steps 3] [--warmup 2] [--check]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        scripts/bench_longtrack.py --seconds 7200

Each rank synthesises only its own slice (packed s24, device resident), stages it to the 16-bit
domain, masters it (ENG:48-80 local; loudness through two halo messages per neighbour and one SUM
all-reduce of the 400 ms block energies over NCCL), and rank 0 prints one JSON line: audio-seconds
per wall-second (max over ranks, CUDA events).  --check compares rank 0's slice with the
single-GPU path on the same input (small tracks only)."""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "python-audio-mastering_b200"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

SETTINGS = dict(bass_boost=4.0, mid_cut=3.0, presence_boost=1.0, treble_boost=3.0,
                saturation=25, width=1.2, multiband=True, lufs=-14.0)
RATE = 96000
PIECE_S = 60.0           # the synthetic mix is a sequence of one-minute pieces (keeps the FFT-based generator small)


def make_slice_s24(abs_offset: int, frames: int, device):
    """Packed little-endian s24 stereo for track frames [abs_offset, abs_offset + frames): the 16-bit
    synthetic programme in the high bytes plus a deterministic low byte."""
    from b200master import synth
    piece = int(PIECE_S * RATE)
    out = torch.empty((frames, 2, 3), dtype=torch.uint8, device=device)
    pos = 0
    while pos < frames:
        k = (abs_offset + pos) // piece
        inner = (abs_offset + pos) - k * piece
        n = min(piece - inner, frames - pos)
        t16 = synth.make_tracks_torch(1000 + k, 1, PIECE_S, RATE, device)[0][inner:inner + n]
        v = t16.to(torch.int32)
        out[pos:pos + n, :, 1] = (v & 0xff).to(torch.uint8)
        out[pos:pos + n, :, 2] = ((v >> 8) & 0xff).to(torch.uint8)
        out[pos:pos + n, :, 0] = ((v * 37 + 11) & 0xff).to(torch.uint8)
        pos += n
    return out.reshape(-1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=7200.0)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--profile", action="store_true", help="rank 0 prints the CUDA-event time of every kernel per step (stderr)")
    args = ap.parse_args()
    import torch.distributed as dist
    from b200master import Engine, longtrack

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # keep NCCL's banner off stdout: rank 0 prints ONE JSON line there
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        comm = longtrack.DistComm()
    else:
        comm = longtrack.ThreadComm(longtrack.ThreadComm.Shared(1), 0)
    eng = Engine(local)
    track_frames = int(round(args.seconds * RATE))
    me = longtrack.partition(track_frames, RATE, world)[rank]
    pcm24 = make_slice_s24(me.abs_offset, me.in_frames, f"cuda:{local}")
    ops = longtrack.EngineOps(eng, RATE, 2, SETTINGS)

    def step():
        return longtrack.master_time_split(pcm24, track_frames, RATE, ops, comm, rank, world, fmt=1)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 1)):
        out, info = step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out, info = step()
    eng.synchronize()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    if args.profile:
        eng.set_profiling(True); eng.reset_profile()
        step(); eng.synchronize()
        if rank == 0:
            for k in ["k_stage_s24", "k_chain", "k_detect", "k_comp", "k_comp_repair", "k_comp_fix", "k_kweight", "k_hops",
                      "k_blocks", "k_gate", "k_final"]:
                t, c = eng.kernel_time_ms(k)
                print(f"{k:16s} {t:8.3f} ms ({c} launches)", file=sys.stderr)
        eng.set_profiling(False)
    ok = None
    if args.check:
        # every rank rebuilds the WHOLE track in the 16-bit domain and masters it alone
        full24 = make_slice_s24(0, track_frames, f"cuda:{local}")
        full16 = torch.empty((track_frames, 2), dtype=torch.int16, device=f"cuda:{local}")
        eng.stage_pcm(full24, 1, track_frames * 2, full16)
        torch.cuda.synchronize()
        ref, rinfo = eng.master([full16.cpu().numpy()], RATE, SETTINGS)
        mine = out.cpu().numpy()
        ok = bool(np.array_equal(mine, ref[0][me.abs_offset:me.abs_offset + me.out_frames]) and info["loudness"] == rinfo[0]["loudness"])
        if world > 1:
            t = torch.tensor([1.0 if ok else 0.0], device=f"cuda:{local}")
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            ok = bool(t.item() == 1.0)
    if rank == 0:
        step_ms = ms / args.steps
        frames = longtrack.ms_framing(track_frames, RATE)
        print(json.dumps({
            "metric": "audio-sec mastered/sec (RTF)", "value": args.seconds / (step_ms * 1e-3), "unit": "audio-s/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 1), "ms_per_step": step_ms,
            "higher_is_better": True, "scaling": "strong", "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"cfg4: one {args.seconds:g}-s 96 kHz s24 stereo track split along time over {world} GPU(s); "
                                   "techno preset + exciter 25% + width 1.2 + multiband + -14 LUFS + limiter",
                       "exchange": "2 halo messages per neighbour (processed int16) + 1 SUM all-reduce of block energies, NCCL"},
            "loudness": info["loudness"], "gain": info["gain"],
            "hbm_frac_8B_per_frame": 8.0 * frames / (step_ms * 1e-3) / 1e9 / 6450.6 / world,
            "matches_single_gpu": ok}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
