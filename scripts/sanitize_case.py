"""Small batch that takes every kernel and every unaligned path (for compute-sanitizer runs):
two tracks of 35 s and 2.5 s at an odd sample rate, stereo and mono, full chain, both chain kernels."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "python-audio-mastering_b200"))
import numpy as np
from b200master import get_engine, synth

rate = int(sys.argv[1]) if len(sys.argv) > 1 else 44101
eng = get_engine(0)
st = dict(bass_boost=3.0, mid_cut=2.0, presence_boost=1.5, treble_boost=2.0, saturation=18, width=1.15, multiband=True, lufs=-16.0)
for ch in (2, 1):
    tracks = [synth.make_track(80 + i, s, rate, ch) for i, s in enumerate([35.013, 2.507])]
    tracks = [t[: t.shape[0] - (i % 3)] for i, t in enumerate(tracks)]
    for mode in (1, 2, 0):
        eng.set_chain_kernel(mode)
        outs, infos = eng.master(tracks, rate, st)
        eng.synchronize()
        print("channels", ch, "chain kernel", mode, "ok", [o.shape for o in outs], [round(i["loudness"], 3) for i in infos], flush=True)
eng.set_chain_kernel(0)
# the helper entry points
from b200master.plan import make_band
b = synth.make_track(3, 1.0, rate)
eng.compress_dynamic_range(b, make_band(rate, -30.0, 2.0, 3.3, 77.0), debug=True)
eng.saturation(b.astype(np.float32) / 32768, 25)
print("helpers ok")
