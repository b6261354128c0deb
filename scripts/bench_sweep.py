#!/usr/bin/env python
"""BASELINE config 5: preset x loudness sweep over many 30-s clips (small-file throughput).

    python scripts/bench_sweep.py [--clips 1200] [--steps 3] [--warmup 2]

Clip k gets preset k mod 4 (techno, dubstep, pop, rock) and target {-9, -14, -23}[(k div 4) mod 3]
(SURVEY 8d): 12 distinct plans in ONE b200m_master_batch call, every clip exactly one 30-s chunk.
Reported with multiband off and on, PCM resident in HBM.  The clips are 64 distinct synthetic
programmes repeated (throughput does not depend on the content being unique)."""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "python-audio-mastering_b200"))

import torch  # noqa: E402

RATE, SECONDS = 48000, 30.0
PRESETS = [dict(bass_boost=4.0, mid_cut=3.0, presence_boost=1.0, treble_boost=3.0),
           dict(bass_boost=5.0, mid_cut=4.0, presence_boost=2.0, treble_boost=3.5),
           dict(bass_boost=2.0, mid_cut=0.0, presence_boost=3.5, treble_boost=2.5),
           dict(bass_boost=1.5, mid_cut=-2.0, presence_boost=2.5, treble_boost=1.0)]
TARGETS = [-9.0, -14.0, -23.0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=1200)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=2)
    args = ap.parse_args()
    from b200master import Engine, make_plan, ms_framing, synth
    eng = Engine(0)
    n = int(SECONDS * RATE)
    base = synth.make_tracks_torch(500, 64, SECONDS, RATE, "cuda")
    reps = (args.clips + 63) // 64
    d_in = base.repeat(reps, 1, 1)[:args.clips].contiguous()
    d_out = torch.empty_like(d_in)
    offs = [i * n for i in range(args.clips)]
    fr = [n] * args.clips
    of = [ms_framing(n, RATE)] * args.clips
    res = {}
    for mb in (False, True):
        plans = [make_plan(dict(PRESETS[p], saturation=25, width=1.2, multiband=mb, lufs=t), RATE, 2)
                 for t in TARGETS for p in range(4)]
        pidx = [(k % 4) + 4 * ((k // 4) % 3) for k in range(args.clips)]

        def step():
            eng.master_raw(d_in, True, offs, fr, of, plans, pidx, d_out, True, want_loudness=False)

        for _ in range(max(args.warmup, 1)):
            step()
        eng.synchronize(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            step()
        eng.synchronize()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        res["multiband_on" if mb else "multiband_off"] = {
            "ms_per_step": ms, "clips_per_s": args.clips / (ms * 1e-3), "rtf": args.clips * SECONDS / (ms * 1e-3),
            "hbm_frac_8B_per_frame": 8.0 * args.clips * n / (ms * 1e-3) / 1e9 / 6450.6}
    # The full grid on every clip (each clip x 4 presets x 3 targets): 12 outputs per clip.  Without sharing that
    # is 12 whole chains per clip; b200m_master_batch_targets runs ENG:46-82 + the loudness measurement once
    # per (clip, preset) and only gain / limiter / final cast per target.
    gclips = max(1, args.clips // 12)
    g_in = d_in[:gclips].repeat(4, 1, 1).contiguous()                # clip-major within a preset: preset p owns rows p*gclips ..
    g_offs = [i * n for i in range(4 * gclips)]
    g_fr, g_of = [n] * (4 * gclips), [ms_framing(n, RATE)] * (4 * gclips)
    g_out = torch.empty((3,) + tuple(g_in.shape), dtype=torch.int16, device="cuda")
    for mb in (False, True):
        gplans = [make_plan(dict(PRESETS[p], saturation=25, width=1.2, multiband=mb, lufs=TARGETS[0]), RATE, 2) for p in range(4)]
        gidx = [i // gclips for i in range(4 * gclips)]
        g_sep = torch.empty_like(g_in)

        def shared():
            eng.master_raw(g_in, True, g_offs, g_fr, g_of, gplans, gidx, g_out, True, want_loudness=False, targets=TARGETS)

        def separate():
            for t in TARGETS:
                pl = [make_plan(dict(PRESETS[p], saturation=25, width=1.2, multiband=mb, lufs=t), RATE, 2) for p in range(4)]
                eng.master_raw(g_in, True, g_offs, g_fr, g_of, pl, gidx, g_sep, True, want_loudness=False)

        out = {}
        for name, fn in (("shared", shared), ("separate", separate)):
            for _ in range(max(args.warmup, 1)):
                fn()
            eng.synchronize(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.steps):
                fn()
            eng.synchronize()
            e1.record(); torch.cuda.synchronize()
            out[name + "_ms_per_step"] = e0.elapsed_time(e1) / args.steps
        separate()
        eng.synchronize()
        out["last_target_identical"] = bool(torch.equal(g_out[2], g_sep))
        out["outputs"] = 12 * gclips
        out["outputs_per_s_shared"] = 12 * gclips / (out["shared_ms_per_step"] * 1e-3)
        res["grid_multiband_on" if mb else "grid_multiband_off"] = out
    print(json.dumps({"metric": "audio-sec mastered/sec (RTF)", "unit": "audio-s/s", "n_gpus": 1, "steps": args.steps,
                      "config": {"workload": f"cfg5: {args.clips} x 30-s 48 kHz s16 stereo clips, 4 presets x 3 loudness targets "
                                             "(12 plans in one batch), exciter 25% + width 1.2 + limiter"}, **res}))


if __name__ == "__main__":
    main()
