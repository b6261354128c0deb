"""Per-kernel timing of the batch path with the library's CUDA-event profiler."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "python-audio-mastering_b200"))
import numpy as np, torch
from b200master import get_engine, synth, make_plan, ms_framing

ntracks = int(sys.argv[1]) if len(sys.argv) > 1 else 8
seconds = float(sys.argv[2]) if len(sys.argv) > 2 else 180.0
rate = int(sys.argv[3]) if len(sys.argv) > 3 else 48000
sat = float(sys.argv[4]) if len(sys.argv) > 4 else 25
eng = get_engine(0)
if os.environ.get("WS_GB"):                     # bench.py hands the library what the GPU has free (up to 112 GB): 256-track groups
    eng.set_workspace_limit(int(float(os.environ["WS_GB"]) * (1 << 30)))
if os.environ.get("WARM") or os.environ.get("TILE"):
    eng.set_recur_tiling(int(os.environ.get("TILE", 0)), int(os.environ.get("WARM", 0)), int(os.environ.get("ROUNDS", -1)))
st = dict(bass_boost=4.0, mid_cut=3.0, presence_boost=1.0, treble_boost=3.0, saturation=sat, width=1.2, multiband=True, lufs=-14.0)
if os.environ.get("PRESET") == "pop":           # bench.py's cfg2
    st.update(bass_boost=2.0, mid_cut=0.0, presence_boost=3.5, treble_boost=2.5)
t0 = time.time()
hat = synth.HAT_DENSE if os.environ.get("HAT", "dense") == "dense" else synth.HAT_SPARSE
track0 = int(os.environ.get("TRACK0", 0))       # first track of the synthetic programme (bench.py's rank r starts at r * tracks_per_gpu)
d_in = torch.cat([synth.make_tracks_torch(track0 + k, min(32, ntracks - k), seconds, rate, "cuda", hat_cfg=hat) for k in range(0, ntracks, 32)])
torch.cuda.synchronize(); print("synth", time.time() - t0)
n = d_in.shape[1]
d_out = torch.empty_like(d_in)
plan = make_plan(st, rate, 2)
offs = [i * n for i in range(ntracks)]; fr = [n] * ntracks; of = [ms_framing(n, rate)] * ntracks
def step():
    return eng.master_raw(d_in, True, offs, fr, of, [plan], [0] * ntracks, d_out, True, want_loudness=False)
for _ in range(2): step()
eng.synchronize()
eng.set_profiling(True); eng.reset_profile()
K = 3
t0 = time.time()
for _ in range(K): step()
eng.synchronize(); dt = (time.time() - t0) / K
tot = 0
for k in ["k_chain", "k_detect", "k_comp", "k_comp_sprint", "k_comp_repair", "k_comp_fix", "k_kweight", "k_hops", "k_blocks", "k_gate", "k_final"]:
    ms, cnt = eng.kernel_time_ms(k); tot += ms / K
    print(f"{k:10s} {ms / K:9.3f} ms/step  ({cnt} launches)")
print(f"sum {tot:.3f} ms ; wall {dt * 1e3:.3f} ms/step ; RTF {ntracks * seconds / dt:.0f} ; frames {ntracks * n} ; "
      f"8B/frame roofline frac {ntracks * n * 8 / dt / 6450.6e9:.4f}")
eng.set_profiling(False)
print('recur stats', eng.recur_stats())
