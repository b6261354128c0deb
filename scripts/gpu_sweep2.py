"""Recurrence tile lengths off the power of two (partition camping check) on the bench workload."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "python-audio-mastering_b200"))
import torch
from b200master import get_engine, synth, make_plan, ms_framing
ntracks = int(sys.argv[1]) if len(sys.argv) > 1 else 64
rate, seconds = 48000, 180.0
eng = get_engine(0)
st = dict(bass_boost=4.0, mid_cut=3.0, presence_boost=1.0, treble_boost=3.0, saturation=25, width=1.2, multiband=True, lufs=-14.0)
d_in = synth.make_tracks_torch(0, ntracks, seconds, rate, "cuda")
n = d_in.shape[1]; d_out = torch.empty_like(d_in)
plan = make_plan(st, rate, 2)
offs = [i * n for i in range(ntracks)]; fr = [n] * ntracks; of = [ms_framing(n, rate)] * ntracks
def step():
    return eng.master_raw(d_in, True, offs, fr, of, [plan], [0] * ntracks, d_out, True, want_loudness=False)
KS = ["k_recur_tiles", "k_recur_repair", "k_recur_fix"]
tiles = [int(a) for a in sys.argv[2:]] or [0, 32768, 32800, 33280, 30016, 36000, 24000, 28800, 40000, 48000]
for tile in tiles:
    eng.set_recur_tiling(tile, 0, -1)
    step(); eng.synchronize()
    eng.recur_stats(reset=True)
    eng.set_profiling(True); eng.reset_profile()
    for _ in range(2): step()
    eng.synchronize()
    ms = {k: eng.kernel_time_ms(k)[0] / 2 for k in KS}
    eng.set_profiling(False)
    print(f"tile {tile:6d}: " + " ".join(f"{k[8:]} {v:7.3f}" for k, v in ms.items()) + f" | total {sum(ms.values()):7.3f} ms  {eng.recur_stats()}", flush=True)
