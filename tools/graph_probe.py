"""Does a CUDA graph shorten the six dependent launches of one ckks::mult + relinearize?  Captures one call (fixed operand
addresses) on the context's stream and replays it.  usage: python tools/graph_probe.py [c3|c4|c5]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hehub_b200.binding import Context, _mod, pick_moduli
SHAPES = {"c3": (13, [40, 30, 30, 30], 40), "c4": (14, [50] + [40] * 7, 50), "c5": (15, [50] * 12, 55)}
for name in sys.argv[1:] or ["c3"]:
    logn, bits, pbits = SHAPES[name]
    stream = torch.cuda.Stream()
    ctx = Context(device=0, stream=stream.cuda_stream)
    mods, p = pick_moduli(bits, pbits, ctx.lib)
    ext = mods + [p]
    L, n = len(mods), 1 << logn
    em, ep = _mod(ext); mm, mp = _mod(mods)
    key, ct1, ct2, res = (ctx.slab(k) for k in (L * 2 * (L + 1) * n, 2 * L * n, 2 * L * n, 2 * L * n))
    ctx._call("lcg_fill", n, ep, L + 1, key.ptr, L * 2 * (L + 1), 1000, 1)
    ctx._call("lcg_fill", n, mp, L, ct1.ptr, 2 * L, 100, 1)
    ctx._call("lcg_fill", n, mp, L, ct2.ptr, 2 * L, 200, 1)
    call = lambda: ctx._call("ckks_mult_relin", logn, ep, L, ct1.ptr, ct2.ptr, key.ptr, res.ptr, 1)
    for _ in range(5): call()
    ctx.synchronize()
    want = res.download()
    reps = 300
    t0 = time.perf_counter()
    for _ in range(reps): call()
    ctx.synchronize()
    plain = (time.perf_counter() - t0) / reps * 1e6
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=stream):
        call()
    g.replay(); torch.cuda.synchronize()
    import numpy as np
    ok = np.array_equal(res.download(), want)
    t0 = time.perf_counter()
    for _ in range(reps): g.replay()
    torch.cuda.synchronize()
    graph = (time.perf_counter() - t0) / reps * 1e6
    print(f"{name}: stream launches (programmatic dependent launch) {plain:.1f} us per pair, graph replay {graph:.1f} us per pair, same words: {ok}", flush=True)
    ctx.close()
