"""Host-buffer pipeline experiments on the bench workload: e2e ms per step against the group size
(B200M_PIPE_MAX_FRAMES is read at handle creation: one engine per setting), the copy-only time of the same
bytes, and the device-resident time of one group."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "python-audio-mastering_b200"))
import torch
from b200master import Engine, synth, make_plan, ms_framing

ntracks = int(sys.argv[1]) if len(sys.argv) > 1 else 256
rate, seconds = 48000, 180.0
st = dict(bass_boost=4.0, mid_cut=3.0, presence_boost=1.0, treble_boost=3.0, saturation=25, width=1.2, multiband=True, lufs=-14.0)
d_in = torch.cat([synth.make_tracks_torch(k, min(32, ntracks - k), seconds, rate, "cuda", hat_cfg=synth.HAT_DENSE) for k in range(0, ntracks, 32)])
n = d_in.shape[1]
h_in = torch.empty(d_in.shape, dtype=torch.int16, pin_memory=True); h_in.copy_(d_in)
h_out = torch.empty(d_in.shape, dtype=torch.int16, pin_memory=True)
plan = make_plan(st, rate, 2)
def ev(fn, K=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K): fn()
    torch.cuda.synchronize(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / K
eng = Engine(0)
for g in (16, 32, 64):
    x = d_in[:g].contiguous(); y = torch.empty_like(x)
    f = lambda: (eng.master_raw(x, True, [i * n for i in range(g)], [n] * g, [ms_framing(n, rate)] * g, [plan], [0] * g, y, True, want_loudness=False), eng.synchronize())
    print(f"device-resident group of {g}: {ev(f):.2f} ms", flush=True)
d_out = torch.empty_like(d_in)
s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()
def copy_only():
    for c in range(16):
        a, b = c * ntracks // 16, (c + 1) * ntracks // 16
        with torch.cuda.stream(s_up): d_in[a:b].copy_(h_in[a:b], non_blocking=True)
        with torch.cuda.stream(s_dn): h_out[a:b].copy_(d_out[a:b], non_blocking=True)
    torch.cuda.synchronize()
print(f"copy only ({ntracks} tracks both ways): {ev(copy_only):.2f} ms", flush=True)
def up_only():
    with torch.cuda.stream(s_up): d_in.copy_(h_in, non_blocking=True)
    torch.cuda.synchronize()
print(f"H2D only: {ev(up_only):.2f} ms", flush=True)
del d_out, eng
offs = [i * n for i in range(ntracks)]; fr = [n] * ntracks; of = [ms_framing(n, rate)] * ntracks
for per_group in (8, 16, 24, 32, 48, 64):
    os.environ["B200M_PIPE_MAX_FRAMES"] = str(per_group * n + 1)
    eng = Engine(0)
    eng.set_pipeline_shape(1, 1)             # groups are bounded by B200M_PIPE_MAX_FRAMES alone
    for streams in (1, 2):
        eng.set_pipeline_shape(1, streams)
        step = lambda: (eng.master_raw(h_in, False, offs, fr, of, [plan], [0] * ntracks, h_out, False, want_loudness=True))
        print(f"tracks per group {per_group:3d}, compute streams {streams}: {ev(step):8.2f} ms/step", flush=True)
    del eng
