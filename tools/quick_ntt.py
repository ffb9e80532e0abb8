"""Quick A/B timing of the plain transforms for one library build (host-timed over many launches).
usage: python tools/quick_ntt.py <lib.so> [logn ...]"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from hehub_b200.binding import Context, _mod
Q = 576460752272228353
lib = sys.argv[1]
logns = [int(a) for a in sys.argv[2:]] or [12]
ctx = Context(lib_path=lib)
from oracle.binding import Oracle
orc = Oracle()
for logn in logns:
    x = orc.lcg_fill(42, Q, 1 << logn)
    y = orc.ntt_fwd_lazy(logn, Q, x)
    for pipeline in (0, 1):
        ctx.set_option("pipeline", pipeline)
        assert np.array_equal(ctx.ntt_fwd_lazy(logn, Q, x), y), "NTT mismatch"
        assert np.array_equal(ctx.intt_lazy(logn, Q, y), orc.intt_lazy(logn, Q, y)), "INTT mismatch"
for pipeline in (0, 1):
    ctx.set_option("pipeline", pipeline)
    for logn in logns:
        n, polys = 1 << logn, 4096
        m, mp = _mod([Q])
        slabs = [ctx.slab(polys * n) for _ in range(max(2, (512 << 20) // (polys * n * 8)))]
        for i, s in enumerate(slabs):
            ctx._call("lcg_fill", n, mp, 1, s.ptr, polys, 42 + i * polys, 1)
        out = []
        for fwd in (True, False):
            def go(i):
                if fwd: ctx._call("ntt_fwd_lazy", logn, mp, 1, slabs[i % len(slabs)].ptr, polys)
                else: ctx._call("intt_lazy", logn, mp, 1, slabs[i % len(slabs)].ptr, polys, 0)
            for i in range(5): go(i)
            ctx.synchronize()
            reps = 60
            t0 = time.perf_counter()
            for i in range(reps): go(i)
            ctx.synchronize()
            dt = (time.perf_counter() - t0) / reps
            out.append(polys / dt)
        print(f"{os.path.basename(lib):28s} pipeline={pipeline} N={n:6d} ntt {out[0]:.3e}/s intt {out[1]:.3e}/s", flush=True)
        for s in slabs: s.free()
ctx.close()
