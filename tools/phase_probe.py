"""Where the time of one two-launch ckks::mult goes (probe build: tools/ab_build.sh phase -DHB_PHASE_CLOCK).
Thread 0 of every CTA logs the SM clock at marked points (ks_pair.cuh, ntt_engine.cuh); this prints, per kernel, the median
cycles between consecutive marks over the CTAs.  usage: python tools/phase_probe.py tools/_variants_phase.so [--shape c3]"""
import argparse, ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from hehub_b200.binding import Context, _mod, pick_moduli
ap = argparse.ArgumentParser()
ap.add_argument("lib")
ap.add_argument("--shape", default="c3")
ap.add_argument("--opt", nargs="*", default=["pair_path=2"])
a = ap.parse_args()
SHAPES = {"c3": (13, [40, 30, 30, 30], 40), "c4": (14, [50] + [40] * 7, 50), "c5": (15, [50] * 12, 55)}
ctx = Context(lib_path=a.lib)
for kv in a.opt:
    k, v = kv.split("=")
    ctx.set_option(k, int(v))
logn, bits, pbits = SHAPES[a.shape]
mods, p = pick_moduli(bits, pbits, ctx.lib)
mods = [int(m) for m in mods]; ext = mods + [int(p)]
L, n = len(mods), 1 << logn
em, ep = _mod(ext); mm, mp = _mod(mods)
key = ctx.slab(L * 2 * (L + 1) * n); ct1 = ctx.slab(2 * L * n); ct2 = ctx.slab(2 * L * n); res = ctx.slab(2 * L * n)
ctx._call("lcg_fill", n, ep, L + 1, key.ptr, L * 2 * (L + 1), 1000, 1)
ctx._call("lcg_fill", n, mp, L, ct1.ptr, 2 * L, 100, 1)
ctx._call("lcg_fill", n, mp, L, ct2.ptr, 2 * L, 200, 1)
NCTA = 4096
log = ctx.slab(NCTA * 32)
setter = ctx.lib.hehub_b200_debug_set_phase_log
setter.argtypes = [ctypes.c_void_p]
def run():
    ctx._call("ckks_mult_relin", logn, ep, L, ct1.ptr, ct2.ptr, key.ptr, res.ptr, 1)
for _ in range(5): run()
ctx.synchronize()
names = {0: "entry", 1: "pdl wait", 2: "inv local + cluster sync", 3: "inv cross (gather, stages, finish)", 4: "fwd cross (levels, wait, scatter)",
         5: "cluster sync", 6: "fwd local passes + store"}
# both kernels write the same slots: run the pair with the log installed, then read per kernel by zeroing in between is not possible
# (one call = two kernels) -> the second kernel (drop) overwrites CTAs [0, grid_drop); the fan kernel's remain above that
z = np.zeros(NCTA * 32, dtype=np.uint64)
log.upload(z)
assert setter(log.ptr) == 0
run()
ctx.synchronize()
got = log.download((NCTA, 32))
setter(None)
used = np.nonzero(got[:, 0])[0]
print(f"{a.shape}: CTAs that logged: {used.size}")
def report(rows, title):
    if rows.size == 0: return
    t = got[rows].astype(np.int64)
    print(f"-- {title} ({rows.size} CTAs), median cycles between marks")
    for s in range(1, 7):
        d = t[:, s] - t[:, s - 1]
        print(f"   {names[s]:40s} {int(np.median(d)):8d}  (min {d.min()}, max {d.max()})")
    print(f"   {'whole CTA':40s} {int(np.median(t[:, 6] - t[:, 0])):8d}")
    for base, label, cnt in ((8, "inverse local pass", 6), (17, "forward pass", 7)):
        prev = None
        for s in range(base, base + cnt):
            if (t[:, s] == 0).all(): continue
            if prev is not None: print(f"   {label} mark {s} - {prev}: {int(np.median(t[:, s] - t[:, prev])):8d}")
            prev = s
pl_cluster = {13: 4, 12: 4, 14: 8, 15: 8}[logn]
n_drop = 2 * L * pl_cluster  # tpc = 1
report(used[used < n_drop], "ks_drop_kernel")
report(used[used >= n_drop], "ks_fan_kernel (CTAs not overwritten by the drop kernel)")
ctx.close()
