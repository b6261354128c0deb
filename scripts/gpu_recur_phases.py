"""Per-phase cycles of k_recur_tiles (library built with -DB200M_RECUR_TIMING, loaded via B200M_LIB):

    cd python-audio-mastering_b200 && nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
        -Xcompiler -fPIC -shared -DB200M_RECUR_TIMING -o lib/libb200master_rt.so csrc/b200m_api.cu
    B200M_LIB=$PWD/lib/libb200master_rt.so python ../scripts/gpu_recur_phases.py 1 8 64
"""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "python-audio-mastering_b200"))
import torch
from b200master import get_engine, synth, make_plan, ms_framing
rate, seconds = 48000, 180.0
eng = get_engine(0)
st = dict(bass_boost=4.0, mid_cut=3.0, presence_boost=1.0, treble_boost=3.0, saturation=25, width=1.2, multiband=True, lufs=-14.0)
plan = make_plan(st, rate, 2)
for nt in [int(a) for a in sys.argv[1:]] or [1, 8]:
    d_in = synth.make_tracks_torch(0, nt, seconds, rate, "cuda"); d_out = torch.empty_like(d_in)
    n = d_in.shape[1]
    offs = [i * n for i in range(nt)]; fr = [n] * nt; of = [ms_framing(n, rate)] * nt
    def step():
        return eng.master_raw(d_in, True, offs, fr, of, [plan], [0] * nt, d_out, True, want_loudness=False)
    step(); eng.synchronize()
    buf = (C.c_ulonglong * 16)()
    eng._lib.b200m_debug_counters(eng._h, buf, 1)
    eng.set_profiling(True); eng.reset_profile()
    step(); eng.synchronize()
    ms = eng.kernel_time_ms("k_recur_tiles")[0]
    eng.set_profiling(False)
    eng._lib.b200m_debug_counters(eng._h, buf, 1)
    c = list(buf)
    warps, blocks = max(c[15], 1), max(c[14], 1)
    names = ["wait rms", "B gather", "issue rms", "C steps", "D store"]
    print(f"tracks {nt}: k_recur_tiles {ms:.3f} ms; warps {warps}, block iterations per warp {blocks / warps:.1f}, "
          f"cycles per warp {c[13] / warps:.0f} ({c[13] / warps / 1.965e6:.3f} ms at 1965 MHz)")
    for k, nm in enumerate(names):
        print(f"   {nm:10s} {c[8 + k] / blocks:9.1f} cycles per block iteration  ({100.0 * c[8 + k] / max(c[13], 1):5.1f} %)")
