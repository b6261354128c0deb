import os, sys, ctypes, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/python-audio-mastering_b200")
import numpy as np, torch
torch.cuda.init(); torch.zeros(1, device="cuda")
rt = ctypes.CDLL("libcudart.so.12") if os.path.exists("/usr/local/cuda/lib64/libcudart.so.12") else ctypes.CDLL("libcudart.so")
nbytes = 4 * 1024**3
def host_alloc(flags):
    p = ctypes.c_void_p()
    rc = rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(nbytes), ctypes.c_uint(flags))
    assert rc == 0, rc
    return p
d_a = torch.empty(nbytes, dtype=torch.uint8, device="cuda"); d_b = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
rt.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
def run(pin_in, pin_out, both):
    def f():
        rt.cudaMemcpyAsync(d_a.data_ptr(), pin_in, nbytes, 1, s1.cuda_stream)
        if both: rt.cudaMemcpyAsync(pin_out, d_b.data_ptr(), nbytes, 2, s2.cuda_stream)
        torch.cuda.synchronize()
    f()
    t0 = time.perf_counter()
    for _ in range(3): f()
    return (time.perf_counter() - t0) / 3
p_def, p_wc, p_out = host_alloc(0), host_alloc(4), host_alloc(0)
ctypes.memset(p_def, 1, nbytes); ctypes.memset(p_wc, 1, nbytes); ctypes.memset(p_out, 0, nbytes)
for name, pin in (("default", p_def), ("write-combined", p_wc)):
    t1 = run(pin, p_out, False); t2 = run(pin, p_out, True)
    print(f"{name:15s}: H2D only {nbytes / t1 / 1e9:6.1f} GB/s ; H2D + D2H together {nbytes / t2 / 1e9:6.1f} GB/s per direction")
