"""Aggregate an `ncu --page source --csv` dump by SASS opcode: executed warp instructions,
stall samples.  usage: ncu_opmix.py file.csv [top]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
cnt = collections.Counter(); smp = collections.Counter(); tot = 0; tots = 0
stalls = collections.Counter()
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for r in rows[2:]:
    if len(r) < len(hdr): continue
    src = r[ix["Source"]].strip()
    parts = src.split()
    if not parts: continue
    op = parts[1] if parts[0].startswith("@") and len(parts) > 1 else parts[0]
    op = op.split(".")[0] + ("." + op.split(".")[1] if op.startswith(("LDS", "STS", "LDG", "STG", "SHFL", "BAR")) and "." in op else "")
    if not r[ix["Instructions Executed"]].isdigit(): continue
    n = int(r[ix["Instructions Executed"]] or 0); s = int(r[ix["# Samples"]] or 0)
    cnt[op] += n; smp[op] += s; tot += n; tots += s
    for c in stall_cols: stalls[c] += int(r[ix[c]] or 0)
print(f"total warp-instructions {tot}, samples {tots}")
for op, n in cnt.most_common(top):
    print(f"{op:14s} {n:14d} {100*n/tot:6.2f}%  samples {100*smp[op]/max(tots,1):6.2f}%")
print("stall reasons:", ", ".join(f"{k[6:]} {100*v/max(tots,1):.1f}%" for k, v in stalls.most_common(8)))
