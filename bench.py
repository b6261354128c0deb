#!/usr/bin/env python
"""bench.py -- audio-seconds mastered per second (RTF) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--tracks B] [--seconds S]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...        # the reference's CPU chain (oracle) on host cores

Workload (BASELINE.json configs[2], the configuration the metric "RTF at 1/2/4/8 B200" is
quoted on; weak scaling: a fixed shard per GPU): B synthetic 3-minute 48 kHz 16-bit stereo
tracks per GPU (B = 64 => the full 512-track batch at 8 GPUs), Techno preset, exciter 25 %,
width 1.2, 3-band multiband compressor, -14 LUFS, limiter: the full chain.  One "step" masters
the whole shard once.  Tracks are independent, so ranks never communicate on the data path;
torch.distributed is used only for the barrier and the max-over-ranks time.

One JSON line is printed by rank 0 (see the keys below).  `value` is timed with the PCM
already resident in HBM; `e2e` is the same metric through the public host-buffer call
(pinned host PCM -> H2D -> kernels -> D2H, every step).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "python-audio-mastering_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

SETTINGS = dict(bass_boost=4.0, mid_cut=3.0, presence_boost=1.0, treble_boost=3.0,   # ENG:16 "techno"
                saturation=25, width=1.2, multiband=True, lufs=-14.0)
RATE = 48000
METRIC = "audio-sec mastered/sec (RTF)"
UNIT = "audio-s/s"
KERNELS = ["k_chain", "k_detect", "k_comp", "k_comp_repair", "k_comp_fix", "k_kweight", "k_hops", "k_blocks", "k_gate", "k_final"]


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--tracks", type=int, default=64, help="tracks per GPU")
    ap.add_argument("--seconds", type=float, default=180.0, help="track length")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-baseline", default="auto", choices=["auto", "off"])
    return ap.parse_args()


def workload_config(args):
    return {
        "workload": f"cfg3 shard: {args.tracks} x {args.seconds:g}-s 48 kHz s16 stereo synthetic tracks per GPU "
                    f"(x{args.gpus} GPUs = {args.tracks * args.gpus} tracks), techno preset + exciter 25% + width 1.2 "
                    f"+ multiband + -14 LUFS + limiter",
        "tracks_per_gpu": args.tracks, "track_seconds": args.seconds, "sample_rate": RATE,
        "settings": SETTINGS, "parallelism": f"track-sharded replicas x{args.gpus}, no data-path collective",
        "l2_policy": "inputs larger than L2 (no flush needed)" if args.tracks * args.seconds * RATE * 4 > 4e8
                     else "L2 flushed between steps",
    }


# ----------------------------------------------------------------------------------------
# clocks sampling (nvidia-smi during the timed region)
# ----------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------
# CPU side: the oracle, timed as the reference's own CPU implementation
# ----------------------------------------------------------------------------------------
def _cpu_chunk_job(job):
    """One worker: master `seconds` of one synthetic track with the oracle; returns wall s."""
    index, seconds, impl = job
    from b200master import synth
    from oracle import port
    import numpy as np  # noqa: F401
    pcm = synth.make_track(index, seconds, RATE)
    t0 = time.perf_counter()
    port.master(pcm, RATE, SETTINGS, impl=impl)
    return time.perf_counter() - t0


def cpu_baseline_single(budget_s=20.0):
    """Rank-0, N=1 leg: the faithful CPU chain (pydub's per-frame Python/audioop loop, as the
    reference runs it) on ONE core over a bounded sample of the same workload."""
    probe = _cpu_chunk_job((0, 1.0, "py"))
    sample = float(min(30.0, max(2.0, budget_s / max(probe, 1e-3))))
    wall = _cpu_chunk_job((0, sample, "py"))
    wall_c = _cpu_chunk_job((0, sample, "c"))
    return {"value": sample / wall, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"first {sample:.1f} s of track 0 of the workload (one 30-s chunk or part of it), full chain, "
                      f"faithful per-frame audioop compressor loop; {wall:.1f} s wall",
            "c_port_value": sample / wall_c,
            "c_port_note": "same sample with the compressor loop restated in C (oracle/compressor.c)"}


def run_reference_arm(args):
    """--impl reference: the oracle port on all host cores, bounded samples per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    from oracle import port
    port.build_c()
    cores = os.cpu_count() or 1
    total_steps = args.steps + args.warmup
    budget = max(4.0, 150.0 / total_steps)                  # seconds of wall per step
    probe = _cpu_chunk_job((0, 1.0, "py"))                  # wall per audio-second on one core
    sample = float(min(30.0, max(1.0, budget / max(probe, 1e-3))))
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        def step(k):
            t0 = time.perf_counter()
            pool.map(_cpu_chunk_job, [(k * cores + i, sample, "py") for i in range(cores)])
            return time.perf_counter() - t0
        for k in range(args.warmup):
            step(k)
        times = [step(args.warmup + k) for k in range(args.steps)]
    wall = sum(times)
    value = cores * sample * args.steps / wall
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"each step: {cores} processes x the first {sample:.1f} s of one workload track "
                                   f"(full chain, faithful per-frame audioop compressor loop)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------
def run_b200_arm(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    try:
        # host side of the e2e leg: run (and pin host memory) on the CPUs next to this rank's GPU, so that the
        # 2 x 2.2 GB per step of every rank stay off the inter-socket link
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local))
    except Exception:
        pass
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # keep NCCL's banner off stdout: rank 0 prints ONE JSON line there
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from b200master import Engine, make_plan, ms_framing, synth

    eng = Engine(local)
    B, n = args.tracks, int(round(args.seconds * RATE))
    d_in = synth.make_tracks_torch(rank * B, B, args.seconds, RATE, f"cuda:{local}")
    d_out = torch.empty_like(d_in)
    h_in = torch.empty(d_in.shape, dtype=torch.int16, pin_memory=True)
    h_out = torch.empty(d_in.shape, dtype=torch.int16, pin_memory=True)
    h_in.copy_(d_in)
    torch.cuda.synchronize()
    plan = make_plan(SETTINGS, RATE, 2)
    offs = [i * n for i in range(B)]
    fr = [n] * B
    of = [ms_framing(n, RATE)] * B
    pidx = [0] * B

    def step_dev():
        eng.master_raw(d_in, True, offs, fr, of, [plan], pidx, d_out, True, want_loudness=False)

    def step_e2e():
        return eng.master_raw(h_in, False, offs, fr, of, [plan], pidx, h_out, False, want_loudness=True)

    def barrier():
        if world > 1:
            dist.barrier()
        eng.synchronize()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        eng.synchronize()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=f"cuda:{local}")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        barrier()
        return ms

    for _ in range(max(args.warmup, 3)):
        step_dev()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    eng.set_profiling(True)
    eng.reset_profile()
    l0 = eng.launch_count()
    ms_dev = timed(step_dev, args.steps)
    launches = eng.launch_count() - l0
    ktimes = {k: eng.kernel_time_ms(k) for k in KERNELS}
    eng.set_profiling(False)

    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    clocks = sampler.stop() if rank == 0 else None      # sampled over both timed regions (100 ms period)

    audio_s = B * args.seconds * world
    value = audio_s * args.steps / (ms_dev * 1e-3)
    e2e_value = audio_s * args.steps / (ms_e2e * 1e-3)

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        dom = max(KERNELS, key=lambda k: ktimes[k][0])
        dom_ms = ktimes[dom][0] / max(ktimes[dom][1], 1)
        frames_per_launch = B * of[0]
        alg_bytes = 8.0 * frames_per_launch                 # 8 B per s16 stereo frame
        achieved = alg_bytes / (dom_ms * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            # the library times k_chain and its warp-per-segment form k_chainw (what runs at this batch size,
            # see b200m_set_chain_kernel) under one name; ncu lists them separately
            key = "k_chainw" if dom == "k_chain" and "k_chainw" in tj else dom
            if key in tj:
                traffic = tj[key]["dram_bytes_per_frame"] * frames_per_launch
        step_ms = ms_dev / args.steps
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args),
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": int(h_in.numel() * 2), "d2h_bytes_per_step": int(h_out.numel() * 2 + B * 16)},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes,
                         "kernel_ms_per_launch": dom_ms,
                         "chain_frac": (alg_bytes / (step_ms * 1e-3) / 1e9) / peak,
                         "note": "8 B per stereo s16 frame (read once + write once) x frames per launch; "
                                 "chain_frac = the same bytes over the whole step"},
            "kernel_ms_per_step": {k: ktimes[k][0] / args.steps for k in KERNELS},
            "clocks": clocks,
        }
        if args.cpu_baseline != "off" and world == 1:
            try:
                line["cpu_baseline"] = cpu_baseline_single()
            except Exception as e:  # the GPU numbers stand even if the host leg fails
                line["cpu_baseline"] = {"error": repr(e)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
