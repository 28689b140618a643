"""Raw pinned host->device copy bandwidth on this box (context for bench.py's e2e number)."""
import ctypes as C, time, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import denormalized_b200 as d
L = d.lib()
for mb in (0.5, 2, 64, 1024):
    n = int(mb * 2**20)
    h = L.dnz_host_alloc(n); g = L.dnz_device_alloc(0, n)
    C.memset(h, 1, n)
    L.dnz_memcpy(g, h, n, 1)
    reps = max(3, int(2**31 // n // 4))
    t = time.perf_counter()
    for _ in range(reps):
        L.dnz_memcpy(g, h, n, 1)
    dt = time.perf_counter() - t
    print(f"H2D pinned {mb:7.1f} MiB x{reps}: {n * reps / dt / 1e9:6.1f} GB/s")
    L.dnz_host_free(h); L.dnz_device_free(0, g)
