"""Per-kernel headline metrics and top stall reasons of an .ncu-rep (ncu --set full).
usage: ncu_stalls.py report.ncu-rep"""
import csv, sys, subprocess, io
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[0]
def col(n): return hdr.index(n) if n in hdr else -1
base = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'smsp__issue_active.avg.pct', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum']
stall = [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio')]
for r in rows[2:]:
    print(r[col('Kernel Name')][:60])
    print('  ', {b.split('.')[0]: r[col(b)] + ' ' + rows[1][col(b)] for b in base if col(b) >= 0})
    st = sorted(((float(r[col(s)]), s.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')) for s in stall), reverse=True)[:7]
    print('   stalls per issue:', [(n, round(v, 2)) for v, n in st])
