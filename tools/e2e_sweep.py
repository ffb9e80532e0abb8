"""Sweep the host-buffer pipeline's chunk size (option host_chunk_kib) for the headline workload.
usage: python tools/e2e_sweep.py [kib ...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from hehub_b200.binding import Context
Q, LOGN, POLYS = 576460752272228353, 12, 4096
n = 1 << LOGN
ctx = Context()
hx, hy = ctx.pinned((POLYS, n)), ctx.pinned((POLYS, n))
hx[:] = np.random.default_rng(0).integers(0, Q, (POLYS, n), dtype=np.uint64)
for kib in [int(a) for a in sys.argv[1:]] or [512, 1024, 2048, 4096, 8192, 16384, 32768]:
    ctx.set_option("host_chunk_kib", kib)
    for _ in range(2):
        ctx.ntt_host(True, LOGN, [Q], hx, hy)
    reps = 10
    t0 = time.perf_counter()
    for _ in range(reps):
        ctx.ntt_host(True, LOGN, [Q], hx, hy)
    dt = (time.perf_counter() - t0) / reps
    print(f"host_chunk_kib={kib:6d}  {POLYS / dt:.4e} NTT/s  {POLYS * n * 8 / dt / 1e9:.1f} GB/s each way", flush=True)
ctx.close()
