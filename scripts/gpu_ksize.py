"""Device-resident step time against the number of tracks in the batch (small-group efficiency)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "python-audio-mastering_b200"))
import torch
from b200master import get_engine, synth, make_plan, ms_framing
rate, seconds = 48000, 180.0
eng = get_engine(0)
st = dict(bass_boost=4.0, mid_cut=3.0, presence_boost=1.0, treble_boost=3.0, saturation=25, width=1.2, multiband=True, lufs=-14.0)
d_all = synth.make_tracks_torch(0, 32, seconds, rate, "cuda")
n = d_all.shape[1]
plan = make_plan(st, rate, 2)
for nt in [1, 2, 4, 6, 8, 12, 16, 24, 32]:
    d_in = d_all[:nt].contiguous(); d_out = torch.empty_like(d_in)
    offs = [i * n for i in range(nt)]; fr = [n] * nt; of = [ms_framing(n, rate)] * nt
    def step():
        return eng.master_raw(d_in, True, offs, fr, of, [plan], [0] * nt, d_out, True, want_loudness=False)
    step(); step(); eng.synchronize()
    eng.set_profiling(True); eng.reset_profile()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 4
    e0.record()
    for _ in range(K): step()
    eng.synchronize(); e1.record(); torch.cuda.synchronize()
    ks = {k: eng.kernel_time_ms(k)[0] / K for k in ["k_chain", "k_detect", "k_comp", "k_kweight", "k_final"]}
    eng.set_profiling(False)
    ms = e0.elapsed_time(e1) / K
    print(f"tracks {nt:3d}: {ms:8.3f} ms/step = {ms / nt:6.3f} ms/track | " + " ".join(f"{k[2:]} {v:6.2f}" for k, v in ks.items()), flush=True)
