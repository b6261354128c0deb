"""PCIe probe: pinned H2D, D2H, and both at once (what bounds bench.py's e2e number)."""
import torch, time
n = 2211840000
h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True); h_out = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d_a = torch.empty(n, dtype=torch.uint8, device="cuda"); d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(f, reps=3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e3
def h2d():
    with torch.cuda.stream(s1): d_a.copy_(h_in, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h_out.copy_(d_b, non_blocking=True)
def both():
    h2d(); d2h()
for name, f in [("h2d", h2d), ("d2h", d2h), ("both", both)]:
    run(f, 1); ms = run(f)
    print(f"{name}: {ms:.1f} ms for 2.21 GB each way -> {n / ms / 1e6:.1f} GB/s per direction")
