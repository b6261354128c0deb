"""k_chain (CTA per segment) against k_chainw (warp per segment) on device-resident batches."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "python-audio-mastering_b200"))
import torch
from b200master import get_engine, synth, make_plan, ms_framing
rate, seconds = 48000, 180.0
eng = get_engine(0)
st = dict(bass_boost=4.0, mid_cut=3.0, presence_boost=1.0, treble_boost=3.0, saturation=25, width=1.2, multiband=True, lufs=-14.0)
plan = make_plan(st, rate, 2)
for nt in [int(a) for a in sys.argv[1:]] or [64, 8]:
    d_in = synth.make_tracks_torch(0, nt, seconds, rate, "cuda"); n = d_in.shape[1]
    outs = {}
    offs = [i * n for i in range(nt)]; fr = [n] * nt; of = [ms_framing(n, rate)] * nt
    for mode, segt in [(1, 0), (2, 0), (0, 0)]:
        eng.set_chain_kernel(mode); eng.set_segment_tiles(segt, 0)
        d_out = torch.empty_like(d_in)
        def step():
            return eng.master_raw(d_in, True, offs, fr, of, [plan], [0] * nt, d_out, True, want_loudness=False)
        step(); eng.synchronize()
        eng.set_profiling(True); eng.reset_profile()
        K = 3
        for _ in range(K): step()
        eng.synchronize()
        ms = eng.kernel_time_ms("k_chain")[0] / K
        msk = eng.kernel_time_ms("k_kweight")[0] / K
        eng.set_profiling(False)
        key = "ref" if mode == 1 else f"w{mode}{segt}"
        outs[key] = d_out
        same = bool(torch.equal(d_out, outs["ref"]))
        print(f"tracks {nt:3d} mode {mode} seg_tiles(2048) {segt:3d}: k_chain {ms:8.3f} k_kweight {msk:7.3f} ms/step   identical to k_chain: {same}", flush=True)
    del outs, d_in
    eng.set_chain_kernel(0); eng.set_segment_tiles(0, 0)
