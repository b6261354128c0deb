"""Exploratory GPU parity probe (not a test): per golden case, how far is the CUDA chain
from the reference output."""
import os, sys, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "python-audio-mastering_b200"))
import numpy as np
from b200master import get_engine, synth, make_plan
from b200master.plan import make_band, kweight_biquads
from oracle import port

eng = get_engine(0)
G = os.path.join(ROOT, "tests", "golden")
for f in sorted(os.listdir(G)):
    if not f.endswith(".npz") or f == "stages.npz":
        continue
    z = np.load(os.path.join(G, f))
    st = json.loads(str(z["settings"])); rate = int(z["rate"])
    t0 = time.time()
    try:
        outs, infos = eng.master([z["pcm"]], rate, st)
    except Exception as e:
        print(f"{f}: EXC {e!r}"); continue
    dt = time.time() - t0
    o, ref = outs[0], z["out"]
    if o.shape != ref.shape:
        print(f"{f}: SHAPE {o.shape} vs {ref.shape}"); continue
    d = np.abs(o.astype(np.int32) - ref.astype(np.int32))
    print(f"{f}: n={d.size} max={int(d.max()) if d.size else 0} exact={np.mean(d==0):.6f} >1LSB={int(np.sum(d>1))} "
          f"loud gpu={infos[0]['loudness']} ref={float(z['loudness'])} ({dt*1e3:.1f} ms)")

# stage probes
z = np.load(os.path.join(G, "stages.npz")); rate = int(z["rate"]); st = json.loads(str(z["settings"]))
x = eng.pcm16_to_float(z["pcm"]); print("to_float exact", np.array_equal(x, z["to_float"]))
s = eng.saturation(z["to_float"], 35); d = s.view(np.int32).astype(np.int64) - z["saturation35"].view(np.int32)
print("saturation: exact frac", np.mean(d == 0), "max ulp", np.abs(d).max())
import audio_mastering_engine as ame
e = ame.apply_eq_to_samples(z["saturation35"], rate, st)
print("eq maxabs err", np.abs(e - z["eq"]).max(), "rel", np.abs(e - z["eq"]).max() / np.abs(z["eq"]).max())
w = eng.stereo_width(z["eq"], 1.4); print("width exact", np.array_equal(w, z["width14"]))
q = eng.float_to_pcm16(z["width14"]); print("q1 exact", np.array_equal(q, z["q1"]))
mb = eng.multiband(z["q1"], make_plan(dict(multiband=True), rate, 2)); dd = np.abs(mb.astype(int) - z["multiband"].astype(int))
print("multiband max", dd.max(), "exact", np.mean(dd == 0))
bands = port.split_bands(z["q1"], rate)
for b, (thr, ratio), (att, rel) in zip(bands, port.band_params({}), port.BAND_TIMES):
    ro, ra, rr = port.compress_band(b, rate, thr, ratio, att, rel, debug=True)
    go, ga, gr = eng.compress_dynamic_range(b, make_band(rate, thr, ratio, att, rel), debug=True)
    print(" band: rms exact", np.array_equal(rr, gr), "att exact", np.array_equal(ra, ga), "att maxdiff", np.abs(ra - ga).max(),
          "out exact", np.array_equal(ro, go), "max", np.abs(ro.astype(int) - go.astype(int)).max())
proc = port.pcm_to_float(z["multiband"])
n, loud, gain = eng.normalize_to_lufs(proc, rate, -14.0, kweight_biquads(rate))
print("loudness gpu", loud, "ref", float(z["loudness"]), "diff", loud - float(z["loudness"]), "normalized exact", np.array_equal(n, z["normalized"]),
      "maxrel", np.abs(n - z["normalized"]).max())
l = eng.soft_limiter(z["normalized"]); print("limiter64 exact", np.array_equal(l, z["limited"]))
l32 = eng.soft_limiter((proc * np.float32(1.7))); print("limiter32 exact", np.array_equal(l32, z["limited32"]))
# k-weighting + block energies
kwref = port.k_weight(proc.mean(axis=1), rate)
print("launches", eng.launch_count())
