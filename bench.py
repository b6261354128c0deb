#!/usr/bin/env python
"""bench.py -- audio-seconds mastered per second (RTF) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--tracks-total T] [--seconds S]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...        # the reference's CPU chain (oracle) on host cores

Workload (BASELINE.json configs[2], the configuration the metric "RTF at 1/2/4/8 B200" is quoted on):
T = 512 synthetic 3-minute 48 kHz 16-bit stereo tracks, Techno preset, exciter 25 %, width 1.2, 3-band
multiband compressor, -14 LUFS, limiter: the full chain.  STRONG scaling: the 512 tracks are sharded
over the N ranks (512 on one GPU, 64 per GPU on eight).  One "step" masters every track once.  Tracks
are independent, so ranks never communicate on the data path; torch.distributed is used only for the
barrier and the max-over-ranks time.

One JSON line is printed by rank 0.  `value` is timed with the PCM already resident in HBM; `e2e` is the
same metric through the public host-buffer call (pinned host PCM -> H2D -> kernels -> D2H, every step),
with a copy-only leg (the same bytes over the same buffers, no kernels) beside it as its roof.  `extra`
holds the other BASELINE configurations measured the same way: cfg2 (one 3-min 44.1 kHz track), cfg5
(10 000 x 30-s clips, 12 plans, multiband off / on) and cfg4 (one 2-h 96 kHz 24-bit track split along
time over the N ranks, NCCL; checked against the single-GPU result on a 10-minute prefix).
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
for p in (ROOT, PKG, os.path.join(ROOT, "scripts")):
    if p not in sys.path:
        sys.path.insert(0, p)

SETTINGS = dict(bass_boost=4.0, mid_cut=3.0, presence_boost=1.0, treble_boost=3.0,   # ENG:16 "techno"
                saturation=25, width=1.2, multiband=True, lufs=-14.0)
RATE = 48000
METRIC = "audio-sec mastered/sec (RTF)"
UNIT = "audio-s/s"
KERNELS = ["k_chain", "k_detect", "k_comp", "k_comp_sprint", "k_comp_repair", "k_comp_fix", "k_kweight", "k_hops", "k_blocks", "k_gate", "k_final"]
GEN_BATCH = 32          # tracks synthesised per call (bounds the generator's temporaries)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--tracks-total", type=int, default=512, help="tracks of the whole job, sharded over the ranks (strong scaling)")
    ap.add_argument("--tracks", type=int, default=0, help="tracks per GPU (overrides --tracks-total: weak scaling)")
    ap.add_argument("--seconds", type=float, default=180.0, help="track length")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-baseline", default="auto", choices=["auto", "off"])
    ap.add_argument("--extras", default="on", choices=["on", "off"], help="cfg2 / cfg5 / cfg4 after the main timed regions")
    return ap.parse_args()


def shard(args, rank, world):
    """(first track, tracks of this rank, tracks of the job, scaling)."""
    if args.tracks > 0:
        return rank * args.tracks, args.tracks, args.tracks * world, "weak"
    base, rem = divmod(args.tracks_total, world)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0), args.tracks_total, "strong"


def workload_config(args, world):
    _f, mine, total, scaling = shard(args, 0, world)
    return {
        "workload": f"cfg3: {total} x {args.seconds:g}-s 48 kHz s16 stereo synthetic tracks ({scaling} scaling: {mine} per GPU on {world} GPU(s)), "
                    f"techno preset + exciter 25% + width 1.2 + multiband + -14 LUFS + limiter",
        "tracks_total": total, "tracks_per_gpu": mine, "track_seconds": args.seconds, "sample_rate": RATE,
        "settings": SETTINGS, "parallelism": f"track-sharded x{world}, no data-path collective",
        "synthetic": "pink noise + 5 sines + 60 Hz kicks + clicks + dense 6-12 kHz hats (synth.HAT_DENSE), seed 0xB200 + track",
        "l2_policy": "inputs larger than L2 (no flush needed)" if mine * args.seconds * RATE * 4 > 4e8
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
        busy = sorted(x for x in sm if mx and x > 0.5 * mx) or sorted(sm)
        return {"sm_mhz": busy[len(busy) // 2] if busy else None, "sm_max_mhz": mx, "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------
# the workload's bytes: the torch generator on cuda:<local> (the same call in both arms)
# ----------------------------------------------------------------------------------------
def workload_tracks(first, count, seconds, device):
    """(count, N, 2) int16 on `device`: tracks first .. first + count - 1 of the workload."""
    import torch
    from b200master import synth
    n = int(round(seconds * RATE))
    out = torch.empty((count, n, 2), dtype=torch.int16, device=device)
    for k0 in range(0, count, GEN_BATCH):
        k1 = min(count, k0 + GEN_BATCH)
        out[k0:k1] = synth.make_tracks_torch(first + k0, k1 - k0, seconds, RATE, device, hat_cfg=synth.HAT_DENSE)
    return out


def workload_tracks_host(indices, seconds):
    """The same bytes on the host, for the CPU legs: generated on the GPU when there is one (outside every timed
    region), else by the numpy twin of the generator (same recipe, another RNG stream) -- the returned note says which."""
    import numpy as np
    try:
        import torch
        if torch.cuda.is_available():
            dev = f"cuda:{int(os.environ.get('LOCAL_RANK', '0'))}"
            return [workload_tracks(i, 1, seconds, dev)[0].cpu().numpy() for i in indices], \
                "identical bytes to the b200 arm (synth.make_tracks_torch on the GPU, copied to the host before timing)"
    except Exception:
        pass
    from b200master import synth
    return [synth.make_track(i, seconds, RATE, hat_cfg=synth.HAT_DENSE) for i in indices], \
        "numpy twin of the generator (no CUDA device in this process): same recipe, different RNG stream"


# ----------------------------------------------------------------------------------------
# CPU side: the oracle, timed as the reference's own CPU implementation
# ----------------------------------------------------------------------------------------
_CPU_TRACKS = {}


def _cpu_job(job):
    """One worker: master pre-made PCM with the oracle; returns wall s."""
    key, impl = job
    from oracle import port
    pcm = _CPU_TRACKS[key]
    t0 = time.perf_counter()
    port.master(pcm, RATE, SETTINGS, impl=impl)
    return time.perf_counter() - t0


def workload_check(pcm):
    """the synthetic programme must drive the chain the way real material does -- the EQ'd signal
    exceeds full scale somewhere (the +FS wrap of ENG:125 is live) and every compressor band is above its
    threshold for >= 10 % of the frames (the attenuation recurrence has work).  Measured with the oracle on the
    first 30-s chunk of one workload track."""
    import numpy as np
    from oracle import port
    x = port.pcm_to_float(pcm)
    x = port.widen(port.eq(port.exciter(x, SETTINGS["saturation"]), RATE, SETTINGS), SETTINGS["width"])
    peak = float(np.abs(x).max())
    q = port.float_to_pcm16(x)
    act = []
    for b, (thr, ratio), (att, rel) in zip(port.split_bands(q, RATE), port.band_params(SETTINGS), port.BAND_TIMES):
        _o, _a, r = port.compress_band(b, RATE, thr, ratio, att, rel, debug=True)
        act.append(float(np.mean(r > 32768.0 * 10 ** (thr / 20.0))))
    ok = peak > 1.0 and min(act) >= 0.10
    return {"eq_peak_fs": peak, "band_active_frac": act, "ok": bool(ok),
            "note": "first 30-s chunk of track 0: peak of the exciter+EQ+width output (> 1.0 exercises the +FS wrap); "
                    "fraction of frames with window RMS above the band's threshold (>= 0.10 each)"}


def cpu_baseline_single(pcm_track0, budget_s=20.0):
    """Rank-0, N=1 leg: the faithful CPU chain (pydub's per-frame Python/audioop loop, as the
    reference runs it) on ONE core over a bounded sample of the same workload."""
    from oracle import port
    port.build_c()
    _CPU_TRACKS["probe"] = pcm_track0[:RATE]
    probe = _cpu_job(("probe", "py"))
    sample = float(min(30.0, max(2.0, budget_s / max(probe, 1e-3))))
    _CPU_TRACKS["s"] = pcm_track0[:int(sample * RATE)]
    wall = _cpu_job(("s", "py"))
    wall_c = _cpu_job(("s", "c"))
    return {"value": sample / wall, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"first {sample:.1f} s of track 0 of the workload (the bytes the GPU arm mastered, copied back from HBM), full chain, "
                      f"faithful per-frame audioop compressor loop; {wall:.1f} s wall",
            "c_port_value": sample / wall_c,
            "c_port_note": "same sample with the compressor loop restated in C (oracle/compressor.c)"}


def run_reference_arm(args):
    """--impl reference: the oracle port on all host cores, bounded samples per step."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    import multiprocessing as mp
    from oracle import port
    port.build_c()
    cores = os.cpu_count() or 1
    total_steps = args.steps + args.warmup
    budget = max(4.0, 150.0 / total_steps)                  # seconds of wall per step
    tracks, note = workload_tracks_host(list(range(cores)), args.seconds)      # whole tracks: the first seconds of THE workload's tracks
    _CPU_TRACKS["probe"] = tracks[0][:RATE]
    probe = _cpu_job(("probe", "py"))                       # wall per audio-second on one core
    sample = float(min(30.0, max(1.0, budget / max(probe, 1e-3))))
    for i, t in enumerate(tracks):
        _CPU_TRACKS[i] = t[:int(sample * RATE)]
    check = workload_check(tracks[0])
    ctx = mp.get_context("fork")                            # the workers inherit _CPU_TRACKS
    with ctx.Pool(cores) as pool:
        def step():
            t0 = time.perf_counter()
            pool.map(_cpu_job, [(i, "py") for i in range(cores)])
            return time.perf_counter() - t0
        for _ in range(args.warmup):
            step()
        times = [step() for _ in range(args.steps)]
    wall = sum(times)
    value = cores * sample * args.steps / wall
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps,
        "higher_is_better": True, "scaling": shard(args, 0, world)[3], "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, world),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"each step: {cores} processes x the first {sample:.1f} s of workload tracks 0..{cores - 1} "
                                   f"(full chain, faithful per-frame audioop compressor loop); {note}"},
        "workload_check": check,
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
    dev = f"cuda:{local}"
    try:
        # host side of the e2e leg: run (and pin host memory) on the CPUs next to this rank's GPU
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local))
    except Exception:
        pass
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # keep NCCL's banner off stdout: rank 0 prints ONE JSON line there
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from b200master import Engine, make_plan, ms_framing

    eng = Engine(local)
    first, B, total_tracks, scaling = shard(args, rank, world)
    n = int(round(args.seconds * RATE))
    d_in = workload_tracks(first, B, args.seconds, dev)
    d_out = torch.empty_like(d_in)
    torch.cuda.synchronize()
    plan = make_plan(SETTINGS, RATE, 2)
    offs = [i * n for i in range(B)]
    fr = [n] * B
    of = [ms_framing(n, RATE)] * B
    pidx = [0] * B
    # A device-resident batch is cut into as few groups as the library's workspace limit allows (64 GB unless told
    # otherwise); longer groups mean longer compressor tiles behind the same warm-up, so hand it what this GPU has free.
    torch.cuda.empty_cache()
    free_b, _ = torch.cuda.mem_get_info(local)
    ws_limit = min(int(free_b) - (24 << 30), 112 << 30)
    if ws_limit > (64 << 30):
        eng.set_workspace_limit(ws_limit)

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
        rank_ms.clear()
        rank_ms.append(ms)
        if world > 1:
            every = [torch.zeros(1, device=dev) for _ in range(world)]
            dist.all_gather(every, torch.tensor([ms], device=dev))
            rank_ms[:] = [float(x.item()) for x in every]       # every rank's own time (reported beside the max)
            ms = max(rank_ms)
        barrier()
        return ms

    rank_ms = []

    # ---- device-resident: `value` ------------------------------------------------------------------
    def step_dev():
        eng.master_raw(d_in, True, offs, fr, of, [plan], pidx, d_out, True, want_loudness=False)

    warm = max(args.warmup, 3)
    for _ in range(warm):
        step_dev()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    eng.set_profiling(True)
    eng.reset_profile()
    l0 = eng.launch_count()
    ms_dev = timed(step_dev, args.steps)
    rank_ms_dev = [m / args.steps for m in rank_ms]
    launches = eng.launch_count() - l0
    ktimes = {k: eng.kernel_time_ms(k) for k in KERNELS}
    rank_kernel_ms = [sum(v[0] for v in ktimes.values()) / args.steps]
    if world > 1:
        every = [torch.zeros(1, device=dev) for _ in range(world)]
        dist.all_gather(every, torch.tensor([rank_kernel_ms[0]], device=dev))
        rank_kernel_ms = [float(x.item()) for x in every]
    eng.set_profiling(False)
    track0 = d_in[0].cpu().numpy() if rank == 0 else None        # the CPU legs master these very bytes

    # ---- end to end: pinned host buffers through the public call, and the copy-only roof ---------------
    h_in = torch.empty(d_in.shape, dtype=torch.int16, pin_memory=True)
    h_out = torch.empty(d_in.shape, dtype=torch.int16, pin_memory=True)
    h_in.copy_(d_in)
    torch.cuda.synchronize()

    def step_e2e():
        return eng.master_raw(h_in, False, offs, fr, of, [plan], pidx, h_out, False, want_loudness=True)

    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)

    s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()
    pieces = max(8, B // 8)

    def step_copy():
        # the same bytes over the same pinned buffers, both directions at once, no kernels
        s_up.wait_stream(torch.cuda.current_stream()); s_dn.wait_stream(torch.cuda.current_stream())
        for c in range(pieces):
            a, b = c * B // pieces, (c + 1) * B // pieces
            if b > a:
                with torch.cuda.stream(s_up):
                    d_in[a:b].copy_(h_in[a:b], non_blocking=True)
                with torch.cuda.stream(s_dn):
                    h_out[a:b].copy_(d_out[a:b], non_blocking=True)
        torch.cuda.current_stream().wait_stream(s_up); torch.cuda.current_stream().wait_stream(s_dn)

    step_copy()
    ms_copy = timed(step_copy, args.steps)
    clocks = sampler.stop() if rank == 0 else None      # sampled over the timed regions (100 ms period)
    h2d_bytes, d2h_bytes = int(h_in.numel() * 2) * world, int(h_out.numel() * 2 + B * 16) * world
    del h_in, h_out, d_in, d_out
    torch.cuda.empty_cache()

    audio_s = total_tracks * args.seconds
    value = audio_s * args.steps / (ms_dev * 1e-3)
    e2e_value = audio_s * args.steps / (ms_e2e * 1e-3)

    extra = {}
    if args.extras == "on":
        for name, fn in (("cfg2", extra_cfg2), ("cfg5", extra_cfg5), ("cfg4", extra_cfg4)):
            try:
                extra[name] = fn(eng, rank, world, local)
            except Exception as e:              # the headline stands even if an extra fails
                extra[name] = {"error": repr(e)[:300]}
            barrier()

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        dom = max(KERNELS, key=lambda k: ktimes[k][0])
        dom_ms = ktimes[dom][0] / max(ktimes[dom][1], 1)
        # the library cuts a large batch into groups: one launch of the dominant kernel serves one group
        frames_per_launch = B * of[0] * args.steps / max(ktimes[dom][1], 1)
        alg_bytes = 8.0 * frames_per_launch                 # 8 B per s16 stereo frame
        achieved = alg_bytes / (dom_ms * 1e-3) / 1e9
        traffic, traffic_total = None, None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            # the library times a kernel family under one name; ncu lists the variant that ran at this batch size
            key = {"k_chain": "k_chainw", "k_detect": "k_detectw"}.get(dom, dom)
            key = key if key in tj else dom
            if key in tj:
                traffic = tj[key]["dram_bytes_per_frame"] * frames_per_launch
            traffic_total = sum(v["dram_bytes_per_frame"] for v in tj.values() if isinstance(v, dict) and "dram_bytes_per_frame" in v)
        step_ms = ms_dev / args.steps
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": warm, "ms_per_step": step_ms, "higher_is_better": True, "scaling": scaling,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args, world),
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                    "copy_only_ms": ms_copy / args.steps, "frac_of_copy_roof": ms_copy / ms_e2e,
                    "copy_only_note": "the same pinned buffers, H2D and D2H of every rank at once, no kernels: the host-fabric roof of this step"},
            "gpu_launches": int(launches),
            "rank_ms_per_step": rank_ms_dev, "rank_kernel_ms_per_step": rank_kernel_ms,
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes,
                         "kernel_ms_per_launch": dom_ms,
                         "chain_frac": (8.0 * B * of[0] / (step_ms * 1e-3) / 1e9) / peak,
                         "chain_dram_bytes_per_frame": traffic_total,
                         "note": "8 B per stereo s16 frame (read once + write once) x frames per launch (a launch = one group of the "
                                 "batch); chain_frac = the same bytes over the whole step; traffic from profiles/traffic.json (ncu)"},
            "kernel_ms_per_step": {k: ktimes[k][0] / args.steps for k in KERNELS},
            "clocks": clocks,
            "extra": extra,
        }
        if args.cpu_baseline != "off" and world == 1:
            try:
                line["workload_check"] = workload_check(track0[:30 * RATE])
                line["cpu_baseline"] = cpu_baseline_single(track0)
            except Exception as e:  # the GPU numbers stand even if the host leg fails
                line["cpu_baseline"] = {"error": repr(e)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------
# the other BASELINE configurations (after the headline's timed regions; every rank calls them)
# ----------------------------------------------------------------------------------------
def _event_ms(fn, steps, eng):
    import torch
    eng.synchronize(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    eng.synchronize()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def _max_over_ranks(ms, world, dev):
    if world <= 1:
        return ms
    import torch
    import torch.distributed as dist
    t = torch.tensor([ms], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def extra_cfg2(eng, rank, world, local):
    """BASELINE configs[1]: ONE 3-min 44.1 kHz s16 stereo track, full chain, on one GPU (every rank runs a replica;
    rank 0 reports): latency of a single job, device-resident and through host buffers."""
    import torch
    from b200master import make_plan, ms_framing, synth
    rate, seconds = 44100, 180.0
    st = dict(bass_boost=2.0, mid_cut=0.0, presence_boost=3.5, treble_boost=2.5, saturation=25, width=1.2, multiband=True, lufs=-14.0)
    d_in = synth.make_tracks_torch(0, 1, seconds, rate, f"cuda:{local}", hat_cfg=synth.HAT_DENSE)
    d_out = torch.empty_like(d_in)
    n = d_in.shape[1]
    plan = make_plan(st, rate, 2)
    h_in = d_in.cpu().pin_memory(); h_out = torch.empty_like(h_in).pin_memory()
    dev_fn = lambda: eng.master_raw(d_in, True, [0], [n], [ms_framing(n, rate)], [plan], [0], d_out, True, want_loudness=False)
    e2e_fn = lambda: eng.master_raw(h_in, False, [0], [n], [ms_framing(n, rate)], [plan], [0], h_out, False, want_loudness=True)
    for _ in range(3):
        dev_fn(); e2e_fn()
    ms, ms_e = _event_ms(dev_fn, 10, eng), _event_ms(e2e_fn, 10, eng)
    return {"workload": "cfg2: one 180-s 44.1 kHz s16 stereo track, pop preset + exciter 25% + width 1.2 + multiband + -14 LUFS + limiter, one GPU",
            "ms_per_track": ms, "rtf": seconds / (ms * 1e-3), "e2e_ms_per_track": ms_e, "e2e_rtf": seconds / (ms_e * 1e-3),
            "hbm_frac_8B_per_frame": 8.0 * n / (ms * 1e-3) / 1e9 / 6450.6}


def extra_cfg5(eng, rank, world, local):
    """BASELINE configs[4]: 10 000 x 30-s 48 kHz clips, clip k with preset k mod 4 and target {-9, -14, -23}[(k div 4) mod 3]
    (12 plans in one batch), multiband off and on, sharded over the ranks; device-resident.  64 distinct programmes
    repeated; each rank masters its share in calls of <= 2500 clips."""
    import torch
    from b200master import make_plan, ms_framing, synth
    dev = f"cuda:{local}"
    clips_total, seconds = 10000, 30.0
    base, rem = divmod(clips_total, world)
    mine = base + (1 if rank < rem else 0)
    n = int(seconds * RATE)
    per_call = min(2500, mine)
    src = synth.make_tracks_torch(500, 64, seconds, RATE, dev, hat_cfg=synth.HAT_DENSE)
    d_in = src.repeat((per_call + 63) // 64, 1, 1)[:per_call].contiguous()
    del src
    d_out = torch.empty_like(d_in)
    presets = [dict(bass_boost=4.0, mid_cut=3.0, presence_boost=1.0, treble_boost=3.0), dict(bass_boost=5.0, mid_cut=4.0, presence_boost=2.0, treble_boost=3.5),
               dict(bass_boost=2.0, mid_cut=0.0, presence_boost=3.5, treble_boost=2.5), dict(bass_boost=1.5, mid_cut=-2.0, presence_boost=2.5, treble_boost=1.0)]
    targets = [-9.0, -14.0, -23.0]
    res = {"workload": f"cfg5: {clips_total} x 30-s 48 kHz s16 stereo clips over {world} GPU(s), 4 presets x 3 loudness targets (12 plans per batch), "
                       "exciter 25% + width 1.2 + limiter"}
    calls = [(c0, min(per_call, mine - c0)) for c0 in range(0, mine, per_call)]
    for mb in (False, True):
        plans = [make_plan(dict(presets[p], saturation=25, width=1.2, multiband=mb, lufs=t), RATE, 2) for t in targets for p in range(4)]

        def step():
            for c0, cnt in calls:
                eng.master_raw(d_in, True, [i * n for i in range(cnt)], [n] * cnt, [ms_framing(n, RATE)] * cnt, plans,
                               [((c0 + k) % 4) + 4 * (((c0 + k) // 4) % 3) for k in range(cnt)], d_out, True, want_loudness=False)
        step()
        ms = _max_over_ranks(_event_ms(step, 2, eng), world, dev)
        res["multiband_on" if mb else "multiband_off"] = {
            "ms_per_step": ms, "clips_per_s": clips_total / (ms * 1e-3), "rtf": clips_total * seconds / (ms * 1e-3),
            "hbm_frac_8B_per_frame": 8.0 * clips_total * n / (ms * 1e-3) / 1e9 / 6450.6 / world}
    return res


def extra_cfg4(eng, rank, world, local):
    """BASELINE configs[3]: ONE 2-hour 96 kHz 24-bit stereo track split along time over the N ranks (30-s chunk
    aligned slices; halo exchange + SUM all-reduce of the block energies over NCCL), and the same orchestration on a
    10-minute prefix compared bit for bit with the single-GPU result."""
    import numpy as np
    import torch
    import torch.distributed as dist
    import bench_longtrack as bl
    from b200master import longtrack
    dev = f"cuda:{local}"
    comm = longtrack.DistComm() if world > 1 else longtrack.ThreadComm(longtrack.ThreadComm.Shared(1), 0)
    ops = longtrack.EngineOps(eng, bl.RATE, 2, SETTINGS)
    out = {"workload": f"cfg4: one 7200-s 96 kHz s24 stereo track split along time over {world} GPU(s); techno preset + exciter 25% + width 1.2 "
                       "+ multiband + -14 LUFS + limiter",
           "exchange": "2 halo messages per neighbour (processed int16) + 1 SUM all-reduce of block energies" + (", NCCL" if world > 1 else " (one rank: none)")}
    # ---- check on a 10-minute prefix: every rank rebuilds the whole prefix and masters it alone ---------------
    chk_frames = 600 * bl.RATE
    me = longtrack.partition(chk_frames, bl.RATE, world)[rank]
    part = bl.make_slice_s24(me.abs_offset, me.in_frames, dev)
    got, info = longtrack.master_time_split(part, chk_frames, bl.RATE, ops, comm, rank, world, fmt=1)
    full24 = bl.make_slice_s24(0, chk_frames, dev)
    full16 = torch.empty((chk_frames, 2), dtype=torch.int16, device=dev)
    eng.stage_pcm(full24, 1, chk_frames * 2, full16)
    ref = torch.empty_like(full16)
    loud, _gain = eng.master_raw(full16, True, [0], [chk_frames], [longtrack.ms_framing(chk_frames, bl.RATE)], [ops.plan], [0], ref, True)
    torch.cuda.synchronize()
    ok = bool(torch.equal(got, ref[me.abs_offset:me.abs_offset + me.out_frames]) and info["loudness"] == float(loud[0]))
    if world > 1:
        t = torch.tensor([1.0 if ok else 0.0], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        ok = bool(t.item() == 1.0)
    out["matches_single_gpu"] = ok
    out["check"] = "10-minute prefix: every rank's slice equals its rows of the one-GPU result of the same track, loudness equal"
    del part, full24, full16, ref, got
    torch.cuda.empty_cache()
    # ---- the 2-hour track -------------------------------------------------------------------------------------
    seconds = 7200.0
    frames = int(seconds * bl.RATE)
    me = longtrack.partition(frames, bl.RATE, world)[rank]
    pcm24 = bl.make_slice_s24(me.abs_offset, me.in_frames, dev)
    step = lambda: longtrack.master_time_split(pcm24, frames, bl.RATE, ops, comm, rank, world, fmt=1)
    step(); step()
    if world > 1:
        dist.barrier()
    ms = _max_over_ranks(_event_ms(step, 3, eng), world, dev)
    out.update({"ms_per_step": ms, "rtf": seconds / (ms * 1e-3),
                "hbm_frac_12B_per_frame": 12.0 * frames / (ms * 1e-3) / 1e9 / 6450.6 / world})
    return out


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
