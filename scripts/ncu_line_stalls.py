"""Stall reasons per CUDA source line (ncu --page source), for the lines with the most stall samples.
usage: ncu_line_stalls.py report.ncu-rep kernel_regex [top]"""
import csv, subprocess, sys, io, collections
rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 12
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", "regex:" + pat],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, agg, fname = None, collections.defaultdict(lambda: collections.Counter()), None
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if len(r) > 20 and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr) or not r[0].isdigit(): continue
    key = (fname, int(r[0]), r[1].strip()[:70])
    for i, h in enumerate(hdr):
        if h.startswith("stall_") and "Not Issued" not in h and r[i].isdigit(): agg[key][h[6:]] += int(r[i])
    if r[6].isdigit(): agg[key]["_all"] += int(r[6])
tot = sum(v["_all"] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1]["_all"])[:top]:
    rs = ", ".join(f"{n} {100 * c / max(v['_all'], 1):.0f}%" for n, c in v.most_common(5) if n != "_all")
    print(f"{k[0][:16]}:{k[1]:4d} {100 * v['_all'] / max(tot, 1):5.1f}% of samples | {rs} | {k[2]}")
