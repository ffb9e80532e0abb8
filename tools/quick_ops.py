"""Quick timing of the scheme-level ops for one library build (host-timed, device-resident operands).
usage: python tools/quick_ops.py <lib.so> [--shape c3|c4|c5]"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from hehub_b200.binding import Context, _mod, pick_moduli
ap = argparse.ArgumentParser()
ap.add_argument("lib")
ap.add_argument("--shape", nargs="*", default=["c3"])
ap.add_argument("--only", nargs="*", default=None, help="subset of tensor ext_prod rescale mult_relin")
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--batch", type=int, default=0, help="override the shape's batch (1: single-ciphertext calls)")
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--latency-rows", type=int, default=None, help="context option latency_rows")
ap.add_argument("--scratch-mib", type=int, default=None, help="context option scratch_cap_mib")
ap.add_argument("--single-launch", type=int, default=None, help="context option single_launch")
ap.add_argument("--opt", nargs="*", default=[], help="context options name=value (pair_path=2 pair_tpc=1 ...)")
a = ap.parse_args()
SHAPES = {"ex": (12, [39, 30], 39, 296), "ex3": (12, [39, 30, 30], 39, 296), "c3l6": (13, [40, 30, 30, 30, 30, 30], 40, 296), "c3": (13, [40, 30, 30, 30], 40, 296), "c4": (14, [50] + [40] * 7, 50, 148), "c5": (15, [50] * 12, 55, 74)}
ctx = Context(lib_path=a.lib)
if a.latency_rows is not None: ctx.set_option("latency_rows", a.latency_rows)
if a.scratch_mib is not None: ctx.set_option("scratch_cap_mib", a.scratch_mib)
if a.single_launch is not None: ctx.set_option("single_launch", a.single_launch)
for kv in a.opt:
    k, v = kv.split("=")
    ctx.set_option(k, int(v))
for name in a.shape:
    logn, bits, pbits, batch = SHAPES[name]
    batch = a.batch or batch
    mods, p = pick_moduli(bits, pbits, ctx.lib)
    mods = [int(m) for m in mods]; ext = mods + [int(p)]
    L, n = len(mods), 1 << logn
    em, ep = _mod(ext); mm, mp = _mod(mods)
    key = ctx.slab(L * 2 * (L + 1) * n); ct1 = ctx.slab(batch * 2 * L * n); ct2 = ctx.slab(batch * 2 * L * n)
    res = ctx.slab(batch * 2 * L * n); quad = ctx.slab(batch * 3 * L * n); ext_out = ctx.slab(batch * 2 * (L + 1) * n)
    ctx._call("lcg_fill", n, ep, L + 1, key.ptr, L * 2 * (L + 1), 1000, 1)
    ctx._call("lcg_fill", n, mp, L, ct1.ptr, batch * 2 * L, 100, 1)
    ctx._call("lcg_fill", n, mp, L, ct2.ptr, batch * 2 * L, 200, 1)
    ops = {
        "tensor": lambda: ctx._call("ckks_tensor", logn, mp, L, ct1.ptr, ct2.ptr, quad.ptr, batch),
        "ext_prod": lambda: ctx._call("ext_prod_montgomery", logn, ep, L, ct1.ptr, key.ptr, ext_out.ptr, batch),
        "rescale": lambda: ctx._call("ckks_rescale", logn, mp, L, ct1.ptr, res.ptr, batch),
        "mult_relin": lambda: ctx._call("ckks_mult_relin", logn, ep, L, ct1.ptr, ct2.ptr, key.ptr, res.ptr, batch),
        "relinearize": lambda: ctx._call("ckks_relinearize", logn, ep, L, quad.ptr, key.ptr, res.ptr, batch),
        "rotate": lambda: ctx._call("ckks_rotate", logn, ep, L, ct1.ptr, key.ptr, 5, res.ptr, batch),
    }
    out = []
    for k, fn in ops.items():
        if a.only and k not in a.only: continue
        for _ in range(a.warmup): fn()
        ctx.synchronize()
        reps = a.reps
        t0 = time.perf_counter()
        for _ in range(reps): fn()
        ctx.synchronize()
        dt = (time.perf_counter() - t0) / reps
        out.append(f"{k} {dt * 1e6 / batch:.3f} us/ct")
    print(f"{os.path.basename(a.lib):26s} {name} b={batch} {' '.join(a.opt)}: " + " | ".join(out), flush=True)
    for s in (key, ct1, ct2, res, quad, ext_out): s.free()
ctx.close()
