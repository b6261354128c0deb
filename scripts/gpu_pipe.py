"""Host-buffer pipeline shape sweep (groups x compute streams) on the bench workload: e2e ms per step."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "python-audio-mastering_b200"))
import torch
from b200master import get_engine, synth, make_plan, ms_framing

ntracks = int(sys.argv[1]) if len(sys.argv) > 1 else 64
rate, seconds = 48000, 180.0
eng = get_engine(0)
st = dict(bass_boost=4.0, mid_cut=3.0, presence_boost=1.0, treble_boost=3.0, saturation=25, width=1.2, multiband=True, lufs=-14.0)
d_in = synth.make_tracks_torch(0, ntracks, seconds, rate, "cuda")
n = d_in.shape[1]
h_in = torch.empty(d_in.shape, dtype=torch.int16, pin_memory=True); h_in.copy_(d_in)
h_out = torch.empty(d_in.shape, dtype=torch.int16, pin_memory=True)
del d_in
plan = make_plan(st, rate, 2)
offs = [i * n for i in range(ntracks)]; fr = [n] * ntracks; of = [ms_framing(n, rate)] * ntracks
def step():
    return eng.master_raw(h_in, False, offs, fr, of, [plan], [0] * ntracks, h_out, False, want_loudness=True)
for groups, streams in [(8, 1), (8, 2), (6, 2), (12, 2), (16, 2), (24, 2), (32, 2), (16, 1)]:
    eng.set_pipeline_shape(groups, streams)
    step(); step(); eng.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 4
    e0.record()
    for _ in range(K): step()
    eng.synchronize(); e1.record(); torch.cuda.synchronize()
    print(f"groups {groups:3d} streams {streams}: {e0.elapsed_time(e1) / K:8.3f} ms/step", flush=True)
