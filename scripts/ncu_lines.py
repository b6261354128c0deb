"""Aggregate `ncu --page source --print-source cuda,sass --csv` by CUDA source line.
usage: ncu_lines.py report.ncu-rep kernel_regex [top]"""
import csv, subprocess, sys, io
rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", "regex:" + pat],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
agg, tot_i, tot_s, fname, kernels = {}, 0, 0, None, 0
for r in rows:
    if len(r) >= 2 and r[0] == "Function Name": kernels += 1
    if len(r) >= 2 and r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if len(r) < 8 or not r[0].isdigit(): continue
    key = (fname, int(r[0]), r[1].strip()[:100])
    s = int(r[6]) if r[6].isdigit() else 0
    i = int(r[7]) if r[7].isdigit() else 0
    a = agg.setdefault(key, [0, 0]); a[0] += s; a[1] += i
    tot_i += i; tot_s += s
print(f"{pat}: warp-instructions {tot_i}, stall samples {tot_s} (summed over {kernels} file sections)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{k[0][:18]:18s}:{k[1]:4d} inst {100*v[1]/max(tot_i,1):5.2f}% smp {100*v[0]/max(tot_s,1):5.2f}%  {k[2]}")
