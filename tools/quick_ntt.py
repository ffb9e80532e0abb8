"""Quick A/B timing of the plain transforms for one library build (host-timed over many launches).
usage: python tools/quick_ntt.py <lib.so> [--logn 12 ...] [--polys 4096] [--nocheck] [--pipeline 0 1] [--slab-mib 512]
--nocheck: ablation builds (tools/ab_build.sh) compute wrong answers on purpose.
--slab-mib: total footprint the launches rotate over (small => L2-resident inputs)."""
import argparse, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from hehub_b200.binding import Context, _mod
ap = argparse.ArgumentParser()
ap.add_argument("lib")
ap.add_argument("--logn", type=int, nargs="*", default=[12])
ap.add_argument("--polys", type=int, default=4096)
ap.add_argument("--nocheck", action="store_true")
ap.add_argument("--pipeline", type=int, nargs="*", default=[0])  # kept for old logs; the option is gone
ap.add_argument("--slab-mib", type=int, default=512)
ap.add_argument("--reps", type=int, default=60)
ap.add_argument("--persistent", type=int, nargs="*", default=[0])
a = ap.parse_args()
Q = 576460752272228353
ctx = Context(lib_path=a.lib)
if not a.nocheck:
    from oracle.binding import Oracle
    orc = Oracle()
    for logn in a.logn:
        x = orc.lcg_fill(42, Q, 1 << logn)
        y = orc.ntt_fwd_lazy(logn, Q, x)
        for pipeline in a.pipeline:
            assert np.array_equal(ctx.ntt_fwd_lazy(logn, Q, x), y), "NTT mismatch"
            assert np.array_equal(ctx.intt_lazy(logn, Q, y), orc.intt_lazy(logn, Q, y)), "INTT mismatch"
for pipeline in a.persistent:
    for logn in a.logn:
        n, polys = 1 << logn, a.polys
        m, mp = _mod([Q])
        slabs = [ctx.slab(polys * n) for _ in range(max(1, (a.slab_mib << 20) // (polys * n * 8)))]
        for i, s in enumerate(slabs):
            ctx._call("lcg_fill", n, mp, 1, s.ptr, polys, 42 + i * polys, 1)
        out = []
        for fwd in (True, False):
            def go(i):
                if fwd: ctx._call("ntt_fwd_lazy", logn, mp, 1, slabs[i % len(slabs)].ptr, polys)
                else: ctx._call("intt_lazy", logn, mp, 1, slabs[i % len(slabs)].ptr, polys, 0)
            for i in range(5): go(i)
            ctx.synchronize()
            reps = a.reps
            t0 = time.perf_counter()
            for i in range(reps): go(i)
            ctx.synchronize()
            dt = (time.perf_counter() - t0) / reps
            out.append(polys / dt)
        print(f"{os.path.basename(a.lib):28s} persistent={pipeline} N={n:6d} polys={polys} slabs={len(slabs)} ntt {out[0]:.3e}/s intt {out[1]:.3e}/s", flush=True)
        for s in slabs: s.free()
ctx.close()
