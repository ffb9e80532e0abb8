"""Per-call time of the plain transforms on a handful of rows (host-timed over back-to-back calls, device-resident rows).
usage: python tools/quick_few_rows.py <lib.so> [--logn 13] [--rows 1 2 4 8 16 32] [--opt name=value ...]"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hehub_b200.binding import Context, _mod
ap = argparse.ArgumentParser()
ap.add_argument("lib")
ap.add_argument("--logn", type=int, default=13)
ap.add_argument("--rows", type=int, nargs="*", default=[1, 2, 4, 8, 16, 18, 24, 32, 64])
ap.add_argument("--reps", type=int, default=400)
ap.add_argument("--opt", nargs="*", default=[])
a = ap.parse_args()
ctx = Context(lib_path=a.lib)
for kv in a.opt:
    k, v = kv.split("=")
    ctx.set_option(k, int(v))
q = 576460752272228353
m, mp = _mod([q])
n = 1 << a.logn
out = []
for rows in a.rows:
    x = ctx.slab(rows * n)
    ctx._call("lcg_fill", n, mp, 1, x.ptr, rows, 7, 1)
    res = []
    for name in ("ntt_fwd_lazy", "intt_lazy"):
        fn = (lambda: ctx._call("ntt_fwd_lazy", a.logn, mp, 1, x.ptr, rows)) if name == "ntt_fwd_lazy" else (lambda: ctx._call("intt_lazy", a.logn, mp, 1, x.ptr, rows, 0))
        for _ in range(5): fn()
        ctx.synchronize()
        t0 = time.perf_counter()
        for _ in range(a.reps): fn()
        ctx.synchronize()
        res.append((time.perf_counter() - t0) / a.reps * 1e6)
    out.append(f"{rows}: {res[0]:.2f}/{res[1]:.2f}")
    x.free()
print(f"{os.path.basename(a.lib)} N=2^{a.logn} {' '.join(a.opt)} rows: fwd/inv us per call | " + " | ".join(out))
ctx.close()
