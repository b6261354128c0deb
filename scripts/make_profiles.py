"""Turn ncu artefacts in gpurun_out/ into the tracked summaries under profiles/.

    python scripts/make_profiles.py <tag> <full.ncu-rep> <frames_per_launch> [launches.csv]

Writes profiles/<tag>_ncu_full.md (one row per captured launch), updates profiles/traffic.json
(DRAM bytes per frame per kernel, from the same capture) and copies the launch list."""
import csv, io, json, os, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, rep, frames = sys.argv[1], sys.argv[2], float(sys.argv[3])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
M = {"time": "gpu__time_duration.sum", "rd": "dram__bytes_read.sum", "wr": "dram__bytes_write.sum",
     "dram": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm": "sm__throughput.avg.pct_of_peak_sustained_elapsed",
     "warps": "sm__warps_active.avg.pct_of_peak_sustained_active", "fp64": "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
     "issue": "smsp__issue_active.avg.pct_of_peak_sustained_active", "regs": "launch__registers_per_thread",
     "grid": "launch__grid_size", "block": "launch__block_size", "inst": "smsp__inst_executed.sum"}
scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}


def val(r, k):
    return float(r[ix[M[k]]]) * scale.get(units[ix[M[k]]], 1.0)


tpath = os.path.join(ROOT, "profiles", "traffic.json")
traffic = json.load(open(tpath)) if os.path.exists(tpath) else {}
out = [f"# {tag}: ncu --set full (--clock-control none), one step of the full chain, {frames / 1e6:.2f} M stereo frames per launch\n",
       f"Source report: `{os.path.basename(rep)}` (gpurun_out/, not tracked). DRAM bytes are per launch; B/frame = DRAM bytes / frames.\n",
       "| kernel | grid x block | time ms | DRAM rd GB | DRAM wr GB | B/frame | DRAM % | SM % | warps % | fp64 pipe % | issue % | regs | warp-inst/frame |",
       "|---|---|---|---|---|---|---|---|---|---|---|---|---|"]
seen = set()
for r in rows[2:]:
    name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "")
    base = name.split("<")[0]
    t, rd, wr = val(r, "time"), val(r, "rd"), val(r, "wr")
    out.append(f"| {name} | {int(val(r, 'grid'))} x {int(val(r, 'block'))} | {t * 1e3:.3f} | {rd / 1e9:.3f} | {wr / 1e9:.3f} | {(rd + wr) / frames:.1f} | "
               f"{val(r, 'dram'):.1f} | {val(r, 'sm'):.1f} | {val(r, 'warps'):.1f} | {val(r, 'fp64'):.1f} | {val(r, 'issue'):.1f} | {int(val(r, 'regs'))} | {val(r, 'inst') / frames:.2f} |")
    if t > 2e-5 and t > traffic.get(base, {}).get("_t", 0.0):          # the longest launch of each kernel (repair rounds re-launch k_comp)
        traffic[base] = {"dram_bytes_per_frame": (rd + wr) / frames, "dram_read_bytes": rd, "dram_write_bytes": wr, "frames": frames,
                         "source": f"profiles/{tag}_ncu_full.md", "_t": t}
open(os.path.join(ROOT, "profiles", f"{tag}_ncu_full.md"), "w").write("\n".join(out) + "\n")
for v in traffic.values():
    v.pop("_t", None)
json.dump(traffic, open(tpath, "w"), indent=1)
if len(sys.argv) > 4:
    shutil.copy(sys.argv[4], os.path.join(ROOT, "profiles", f"{tag}_launches.csv"))
print("\n".join(out))
