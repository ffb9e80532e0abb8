"""Parity of the CUDA path (through the C ABI) against the oracle and the reference's golden values.

Every test runs against two back ends:
  * ``gpu``  (marked ``gpu``): hehub_b200/libhehub_b200.so on a real B200 — the parity tests proper;
  * ``sim``  (CPU suite): the same kernel sources compiled for the CTA emulator in tests/kernel_sim —
    checks index arithmetic, table layouts and host logic without a GPU.
The bar is bit-exact raw u64 words (all arithmetic is integer mod q_i).
"""
import numpy as np
import pytest

from conftest import fill_ct, fill_key, fnv

Q59 = 576460752272228353
NTT_MODULI = [65537, 260898817, 35184358850561, 36028796997599233, Q59]


def hx(a):
    return f"{fnv(a):016x}"


@pytest.fixture(scope="module", params=["sim", pytest.param("gpu", marks=pytest.mark.gpu)])
def dev(request):
    from hehub_b200.binding import Context
    if request.param == "sim":
        import __graft_entry__ as ge
        ctx = Context(lib_path=ge.build_sim())
    else:
        ctx = Context(device=0)  # fails loudly without the CUDA library or a device
    ctx.kind = request.param
    yield ctx
    ctx.close()


# ------------------------------------------------------------------ transforms
@pytest.mark.parametrize("logn", [1, 2, 4, 7, 9, 10, 11, 12, 13, 14, 15])
def test_ntt_intt_raw_words_match_oracle(dev, oracle, logn):
    n = 1 << logn
    rng = np.random.default_rng(logn)
    for q in NTT_MODULI:
        x = np.stack([rng.integers(0, q, n, dtype=np.uint64), oracle.lcg_fill(42, q, n),
                      np.eye(1, n, 0, dtype=np.uint64)[0], np.eye(1, n, 1 % n, dtype=np.uint64)[0]])
        want = np.stack([oracle.ntt_fwd_lazy(logn, q, r) for r in x])
        got = dev.poly_ntt_fwd(logn, [q], x)
        assert np.array_equal(got, want), (logn, q)
        assert (got < 2 * q).all()  # tests/ntt_t.cpp:114-118
        back = dev.poly_intt(logn, [q], want)
        want_back = np.stack([oracle.intt_lazy(logn, q, r) for r in want])
        assert np.array_equal(back, want_back), (logn, q)
        # tests/ntt_t.cpp:120-125: one more reduction gives the input back
        strict = dev.poly_intt(logn, [q], want, strict=True)
        assert np.array_equal(strict, x), (logn, q)
        # lazy inputs anywhere in [0, 2^64) are legal for the first butterfly level
        wild = rng.integers(0, 1 << 63, n, dtype=np.uint64) * np.uint64(2) + np.uint64(1)
        if (2 + 2 * logn) * q < (1 << 62):  # headroom so the reference itself does not wrap
            wild = wild >> np.uint64(3)
            assert np.array_equal(dev.poly_ntt_fwd(logn, [q], wild), oracle.ntt_fwd_lazy(logn, q, wild))


@pytest.mark.parametrize("logn", [3, 10, 12, 13, 14, 15])
def test_transforms_wrap_like_the_reference_on_full_range_words(dev, oracle, logn):
    """Words anywhere in [0, 2^64): the reference's growth bound (ntt.cpp:152-175) does not hold, its u64 arithmetic
    wraps — and so must ours, word for word (the kernels run the same u64 operations per butterfly)."""
    n = 1 << logn
    rng = np.random.default_rng(900 + logn)
    for q in (Q59, 36028796997599233, 65537):
        x = rng.integers(0, 1 << 63, (2, n), dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, (2, n), dtype=np.uint64)
        x[1, : n // 2] = np.uint64((1 << 64) - 1)
        want = np.stack([oracle.ntt_fwd_lazy(logn, q, r) for r in x])
        assert np.array_equal(dev.poly_ntt_fwd(logn, [q], x), want), (logn, q)
        want = np.stack([oracle.intt_lazy(logn, q, r) for r in x])
        assert np.array_equal(dev.poly_intt(logn, [q], x), want), (logn, q)


@pytest.mark.parametrize("logn", [12, 13, 14, 15])
def test_latency_and_throughput_plans_agree(dev, oracle, logn):
    """N >= 4096 has two pass plans (ntt_plan.h): the throughput plan, and for launches with few rows a thin plan on a
    4- / 8-CTA cluster.  Both must produce the reference's words, for plain and fused transforms."""
    n = 1 << logn
    mods, ext = _shape(oracle, logn, [40, 30], 40)
    x = np.stack([oracle.lcg_fill(7 + k, Q59, n) for k in range(3)])
    want_f = np.stack([oracle.ntt_fwd_lazy(logn, Q59, r) for r in x])
    want_i = np.stack([oracle.intt_lazy(logn, Q59, r) for r in want_f])
    ct1, ct2, key = fill_ct(oracle, 31, mods, n), fill_ct(oracle, 32, mods, n), fill_key(oracle, 3300, ext, n)
    want_m = oracle.ckks_mult_relin(logn, ext, ct1, ct2, key)
    want_b = oracle.bgv_mod_switch(logn, ext, 65537, fill_ct(oracle, 33, ext, n))
    try:
        dev.set_option("pair_path", 0)  # the wave path, whose launches take these plans (the cluster forms: test_pair_path_...)
        # never the latency plans / always the 4- / 8-CTA thin plans / always the mode-2 plans (N = 4096 / 8192: 8-CTA clusters,
        # tables staged in shared memory)
        for rows, rows2 in ((0, 0), (1 << 30, 0), (1 << 30, 1 << 30))[: 3 if logn <= 13 else 2]:  # (mode 2 exists for N <= 8192)
            dev.set_option("latency_rows", rows)
            dev.set_option("latency2_rows", rows2)
            rows = (rows, rows2)
            assert np.array_equal(dev.poly_ntt_fwd(logn, [Q59], x), want_f), rows
            assert np.array_equal(dev.poly_intt(logn, [Q59], want_f), want_i), rows
            assert np.array_equal(dev.poly_intt(logn, [Q59], want_f, strict=True), x), rows
            assert np.array_equal(dev.ckks_mult_relin(logn, ext, ct1, ct2, key), want_m), rows
            assert np.array_equal(dev.bgv_mod_switch(logn, ext, 65537, fill_ct(oracle, 33, ext, n)), want_b), rows
    finally:
        dev.set_option("latency_rows", -1)
        dev.set_option("latency2_rows", -1)
        dev.set_option("pair_path", 1)


def test_ntt_hashes_match_reference_golden(dev, oracle, kat):
    """SURVEY Appendix B / tests/golden: hashes recorded from the unmodified reference."""
    for row in kat["ntt_hashes"]:
        q, logn = row["q"], row["logn"]
        if logn > 15:
            continue
        x = oracle.lcg_fill(42, q, 1 << logn)
        y = dev.ntt_fwd_lazy(logn, q, x)
        assert hx(y) == row["ntt"], (q, logn)
        assert hx(dev.intt_lazy(logn, q, y)) == row["intt_ntt"], (q, logn)
        assert hx(dev.intt_lazy(logn, q, x)) == row["intt"], (q, logn)


def test_ntt_small_raw_vectors(dev, kat):
    for name, logn in (("ntt_n8", 3), ("ntt_n16", 4)):
        v = kat[name]
        y = dev.ntt_fwd_lazy(logn, v["q"], np.array(v["in"], dtype=np.uint64))
        assert y.tolist() == v["ntt"]
        assert dev.intt_lazy(logn, v["q"], y).tolist() == v["intt"]


@pytest.mark.parametrize("logn", [5, 10, 12, 15])
def test_multi_limb_batched_transform(dev, oracle, logn):
    """[batch][L][N] slabs: limb = row % L, every row independent (ntt.h:41-51, 72-82)."""
    n = 1 << logn
    moduli = [Q59, 36028796997599233, 65537]
    batch = 3
    x = np.stack([np.stack([oracle.lcg_fill(7 + 10 * b + k, q, n) for k, q in enumerate(moduli)]) for b in range(batch)])
    want = np.stack([oracle.poly_ntt_fwd(logn, moduli, x[b]) for b in range(batch)])
    assert np.array_equal(dev.poly_ntt_fwd(logn, moduli, x), want)
    want_i = np.stack([oracle.poly_intt(logn, moduli, want[b]) for b in range(batch)])
    assert np.array_equal(dev.poly_intt(logn, moduli, want), want_i)
    want_s = np.stack([oracle.poly_intt(logn, moduli, want[b], strict=True) for b in range(batch)])
    assert np.array_equal(dev.poly_intt(logn, moduli, want, strict=True), want_s)


@pytest.mark.parametrize("logn", [10, 13, 14])
def test_generic_path_agrees_with_fast_path(dev, oracle, logn):
    n = 1 << logn
    x = oracle.lcg_fill(99, Q59, n)
    fast = dev.ntt_fwd_lazy(logn, Q59, x)
    dev.set_option("force_generic", 1)
    try:
        slow = dev.ntt_fwd_lazy(logn, Q59, x)
        slow_i = dev.intt_lazy(logn, Q59, fast)
    finally:
        dev.set_option("force_generic", 0)
    assert np.array_equal(fast, slow)
    assert np.array_equal(slow_i, dev.intt_lazy(logn, Q59, fast))


@pytest.mark.parametrize("logn", [10, 11, 12, 13])
def test_multi_row_batches_and_unaligned_slabs(dev, oracle, logn):
    """Batches whose row count is not a multiple of anything convenient, mixed moduli per row, and a
    slab that is only 8-byte aligned (rows then move as 64-bit words instead of 128-bit)."""
    n = 1 << logn
    moduli = [Q59, 65537]
    batch = 11
    x = np.stack([np.stack([oracle.lcg_fill(3 + 10 * b + k, q, n) for k, q in enumerate(moduli)]) for b in range(batch)])
    want = np.stack([oracle.poly_ntt_fwd(logn, moduli, x[b]) for b in range(batch)])
    want_i = np.stack([oracle.poly_intt(logn, moduli, want[b]) for b in range(batch)])
    assert np.array_equal(dev.poly_ntt_fwd(logn, moduli, x), want)
    assert np.array_equal(dev.poly_intt(logn, moduli, want), want_i)
    # a slab that starts 8 bytes off a 16-byte boundary
    m = np.ascontiguousarray(np.asarray(moduli, dtype=np.uint64))
    import ctypes as C
    p64 = C.POINTER(C.c_uint64)
    d = dev.slab(x.size + 1)
    try:
        dev._call("slab_h2d", d.ptr + 8, x.ctypes.data, x.size)
        dev._call("ntt_fwd_lazy", logn, m.ctypes.data_as(p64), m.size, d.ptr + 8, batch)
        got = np.empty_like(x)
        dev._call("slab_d2h", got.ctypes.data, d.ptr + 8, x.size)
        dev._call("intt_lazy", logn, m.ctypes.data_as(p64), m.size, d.ptr + 8, batch, 0)
        got_i = np.empty_like(x)
        dev._call("slab_d2h", got_i.ctypes.data, d.ptr + 8, x.size)
        dev.synchronize()
    finally:
        d.free()
    assert np.array_equal(got, want)
    assert np.array_equal(got_i, want_i)


# ------------------------------------------------------------------ coefficient-wise kernels
@pytest.mark.parametrize("n", [1, 2, 5, 1000, 4096, 4099])
def test_coefficient_wise_kernels(dev, oracle, n):
    rng = np.random.default_rng(n)
    moduli = [65537, 33333333, 777777777777777, 1234567890111111111]  # tests/mod_arith_t.cpp:6-32
    full = np.stack([rng.integers(0, 1 << 63, n, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, n, dtype=np.uint64)
                     for _ in moduli])
    assert np.array_equal(dev.barrett_lazy(moduli, full), np.stack([oracle.barrett_lazy(q, full[k]) for k, q in enumerate(moduli)]))
    assert np.array_equal(dev.barrett(moduli, full), np.stack([oracle.barrett(q, full[k]) for k, q in enumerate(moduli)]))
    lazy = np.stack([rng.integers(0, 2 * q, n, dtype=np.uint64) for q in moduli])
    lazy2 = np.stack([rng.integers(0, 2 * q, n, dtype=np.uint64) for q in moduli])
    assert np.array_equal(dev.reduce_strict(moduli, lazy), np.stack([oracle.reduce_strict(q, lazy[k]) for k, q in enumerate(moduli)]))
    assert np.array_equal(dev.add_lazy(moduli, lazy, lazy2), np.stack([oracle.add_lazy(q, lazy[k], lazy2[k]) for k, q in enumerate(moduli)]))
    assert np.array_equal(dev.sub_lazy(moduli, lazy, lazy2), np.stack([oracle.sub_lazy(q, lazy[k], lazy2[k]) for k, q in enumerate(moduli)]))
    scalars = [12345, 1 << 40, 777777777777776, (1 << 64) - 1]
    assert np.array_equal(dev.mul_scalar_lazy(moduli, lazy, scalars),
                          np.stack([oracle.mul_scalar_lazy(q, lazy[k], scalars[k]) for k, q in enumerate(moduli)]))
    odd = [65537, 33333333, 777777777777777, 1234567890111111111]
    assert np.array_equal(dev.mul_hybrid_lazy(odd, lazy, lazy2),
                          np.stack([oracle.mul_hybrid_lazy(q, lazy[k], lazy2[k]) for k, q in enumerate(odd)]))


def test_mod_arith_golden(dev, kat):
    """tests/mod_arith_t.cpp generators, outputs recorded from the unmodified reference."""
    g = kat["mulmod"]
    q, n = g["q"], 1000
    M = (1 << 64) - 1
    seed, f, gg = 42, [], []
    for _ in range(n):  # tests/mod_arith_t.cpp:40-45
        seed = ((seed * 65968279837582827) & M) ^ 3948528936546489545
        f.append(seed % q)
        seed = ((seed * 43534547657678213) & M) ^ 7955436776934235466
        gg.append(seed % q)
    f, gg = np.array(f, dtype=np.uint64), np.array(gg, dtype=np.uint64)
    assert hx(f) == g["f"] and hx(gg) == g["g"]
    out = dev.mul_hybrid_lazy([q], f, gg)
    assert hx(out) == g["hybrid_lazy"]
    assert out[:4].tolist() == g["hybrid_head"]
    m = kat["montgomery128"]
    assert dev.montgomery128_lazy(m["q"], np.array(m["in_lohi"], dtype=np.uint64)).tolist() == m["out"]


def test_montgomery128_matches_oracle(dev, oracle):
    q = 38589379749438777  # tests/mod_arith_t.cpp:61-78
    rng = np.random.default_rng(5)
    lo = rng.integers(0, 1 << 63, 999, dtype=np.uint64) * np.uint64(2) + np.uint64(1)
    hi = rng.integers(0, q, 999, dtype=np.uint64)
    lohi = np.stack([lo, hi], axis=1).ravel()
    assert np.array_equal(dev.montgomery128_lazy(q, lohi), oracle.montgomery128_lazy(q, lohi))


# ------------------------------------------------------------------ scheme-level ops
def _shape(oracle, logn, bits, pbits):
    mods, p = oracle.ckks_pick_moduli(bits, pbits)
    mods = [int(m) for m in mods]
    return mods, mods + [int(p)]


def test_small_golden_vectors(dev, kat):
    """N=16, L=3 full vectors recorded from the unmodified reference (tests/golden)."""
    s = kat["small"]
    logn, mods = s["logn"], s["moduli"]
    ext = mods + [s["P"]]
    L, n = len(mods), 1 << logn
    A = lambda name, shape: np.array(s[name], dtype=np.uint64).reshape(shape)
    ct1, ct2, key = A("ct1", (2, L, n)), A("ct2", (2, L, n)), A("key", (L, 2, L + 1, n))
    quad = dev.ckks_tensor(logn, mods, ct1, ct2)
    assert quad.ravel().tolist() == s["tensor"]
    assert dev.ext_prod(logn, ext, quad[2], key).ravel().tolist() == s["ext_prod"]
    assert dev.ckks_relinearize(logn, ext, quad, key).ravel().tolist() == s["relinearize"]
    assert dev.ckks_mult_relin(logn, ext, ct1, ct2, key).ravel().tolist() == s["mult"]
    assert dev.ckks_rescale(logn, mods, ct1).ravel().tolist() == s["rescale"]
    assert dev.bgv_mod_switch(logn, mods, 65537, ct1).ravel().tolist() == s["mod_switch_t65537"]
    assert dev.bgv_mod_switch(logn, mods, 2, ct1).ravel().tolist() == s["mod_switch_t2"]
    assert dev.bgv_relinearize(logn, ext, 1, quad, key).ravel().tolist() == s["bgv_relinearize_t1"]
    for step, want in s["cycle"].items():
        assert dev.galois_cycle(logn, ct1[0], int(step)).ravel().tolist() == want
    assert dev.galois_involution(logn, ct1[0]).ravel().tolist() == s["involution"]
    for step, want in s["rotate"].items():
        assert dev.ckks_rotate(logn, ext, ct1, key, int(step)).ravel().tolist() == want
    assert dev.ckks_conjugate(logn, ext, ct1, key).ravel().tolist() == s["conjugate"]
    assert dev.add_lazy(mods, ct1[0], ct2[1]).ravel().tolist() == s["poly_add"]
    assert dev.sub_lazy(mods, ct1[0], ct2[1]).ravel().tolist() == s["poly_sub"]
    assert dev.mul_scalar_lazy(mods, ct1[0], [12345] * L).ravel().tolist() == s["poly_mul_scalar_12345"]
    assert dev.poly_intt(logn, mods, ct1[0], strict=True).ravel().tolist() == s["poly_intt_strict"]
    assert dev.poly_ntt_fwd(logn, mods, ct1[0]).ravel().tolist() == s["poly_ntt"]


def test_c3_mult_relin_hashes(dev, oracle, kat):
    """BASELINE config 3: ckks::mult + relinearize, N=8192, L=4 — hashes from the unmodified reference."""
    c3 = kat["c3"]
    logn, mods = c3["logn"], c3["moduli"]
    ext, n = mods + [c3["P"]], 1 << c3["logn"]
    ct1, ct2, key = fill_ct(oracle, 100, mods, n), fill_ct(oracle, 200, mods, n), fill_key(oracle, 1000, ext, n)
    quad = dev.ckks_tensor(logn, mods, ct1, ct2)
    assert [hx(quad[i]) for i in range(3)] == c3["tensor"]
    e = dev.ext_prod(logn, ext, quad[2], key)
    assert [hx(e[i]) for i in range(2)] == c3["ext_prod"]
    r = dev.ckks_relinearize(logn, ext, quad, key)
    assert [hx(r[i]) for i in range(2)] == c3["relinearize"]
    m = dev.ckks_mult_relin(logn, ext, ct1, ct2, key)
    assert [hx(m[i]) for i in range(2)] == c3["mult"]
    rs = dev.ckks_rescale(logn, mods, m)
    assert [hx(rs[i]) for i in range(2)] == c3["rescale"]
    rot = dev.ckks_rotate(logn, ext, ct1, key, 5)
    assert [hx(rot[i]) for i in range(2)] == c3["rotate5"]
    cj = dev.ckks_conjugate(logn, ext, ct1, key)
    assert [hx(cj[i]) for i in range(2)] == c3["conjugate"]
    b = dev.bgv_relinearize(logn, ext, 1, quad, key)
    assert [hx(b[i]) for i in range(2)] == c3["bgv_relinearize_t1"]


def test_c4_rescale_hashes(dev, oracle, kat):
    """BASELINE config 4: rescale / mod-switch N=16384, L=8->7 — hashes from the unmodified reference."""
    c4 = kat["c4"]
    logn, mods, n = c4["logn"], c4["moduli"], 1 << c4["logn"]
    ct = fill_ct(oracle, 300, mods, n)
    r = dev.ckks_rescale(logn, mods, ct)
    assert [hx(r[i]) for i in range(2)] == c4["rescale"]
    m = dev.bgv_mod_switch(logn, mods, 65537, ct)
    assert [hx(m[i]) for i in range(2)] == c4["mod_switch_t65537"]


def test_c5_shape_mult_hash(dev, oracle, kat):
    """BASELINE config 5 shape (one ciphertext): N=32768, L=12."""
    if dev.kind == "sim":
        pytest.skip("182 transforms of 32768 points: GPU only (the sim covers N=32768 with L=2 below)")
    c5 = kat["c5"]
    logn, mods = c5["logn"], c5["moduli"]
    ext, n = mods + [c5["P"]], 1 << c5["logn"]
    ct1, ct2, key = fill_ct(oracle, 100, mods, n), fill_ct(oracle, 200, mods, n), fill_key(oracle, 1000, ext, n)
    m = dev.ckks_mult_relin(logn, ext, ct1, ct2, key)
    assert [hx(m[i]) for i in range(2)] == c5["mult"]


@pytest.mark.parametrize("logn,bits,pbits", [(10, [40, 30, 30], 40), (12, [39, 30], 39), (15, [50, 50], 55), (6, [30, 30, 30, 30], 40)])
def test_scheme_ops_match_oracle(dev, oracle, logn, bits, pbits):
    mods, ext = _shape(oracle, logn, bits, pbits)
    n, L = 1 << logn, len(mods)
    ct1, ct2, key = fill_ct(oracle, 11, mods, n), fill_ct(oracle, 22, mods, n), fill_key(oracle, 3000, ext, n)
    quad = oracle.ckks_tensor(logn, mods, ct1, ct2)
    assert np.array_equal(dev.ckks_tensor(logn, mods, ct1, ct2), quad)
    e = oracle.ext_prod(logn, ext, quad[2], key)
    assert np.array_equal(dev.ext_prod(logn, ext, quad[2], key), e)
    assert np.array_equal(dev.ckks_rescale(logn, ext, e), oracle.ckks_rescale(logn, ext, e))
    for t in (2, 65537):
        assert np.array_equal(dev.bgv_mod_switch(logn, ext, t, e), oracle.bgv_mod_switch(logn, ext, t, e))
    assert np.array_equal(dev.ckks_relinearize(logn, ext, quad, key), oracle.ckks_relinearize(logn, ext, quad, key))
    assert np.array_equal(dev.bgv_relinearize(logn, ext, 1, quad, key), oracle.bgv_relinearize(logn, ext, 1, quad, key))
    assert np.array_equal(dev.ckks_mult_relin(logn, ext, ct1, ct2, key), oracle.ckks_mult_relin(logn, ext, ct1, ct2, key))
    for t in (1, 65537):  # bgv::mult_low_level + bgv::relinearize behind one call
        assert np.array_equal(dev.bgv_mult_relin(logn, ext, t, ct1, ct2, key), oracle.bgv_relinearize(logn, ext, t, quad, key))
    for step in (0, 1, 5, n // 2 - 1):
        assert np.array_equal(dev.galois_cycle(logn, ct1[0], step), oracle.galois_cycle(logn, ct1[0], step))
    assert np.array_equal(dev.galois_involution(logn, ct1[1]), oracle.galois_involution(logn, ct1[1]))
    assert np.array_equal(dev.ckks_rotate(logn, ext, ct1, key, 3), oracle.ckks_rotate(logn, ext, ct1, key, 3))
    assert np.array_equal(dev.ckks_conjugate(logn, ext, ct1, key), oracle.ckks_conjugate(logn, ext, ct1, key))


@pytest.mark.parametrize("logn,bits,pbits", [(5, [30, 30], 40), (10, [40, 30, 30], 40), (13, [59, 59], 59)])
def test_scheme_ops_on_full_range_words(dev, oracle, logn, bits, pbits):
    """Operands anywhere in [0, 2^64) (not residues): every product, reduction and lazy add of the reference is plain
    u64 / u128 arithmetic, so the outputs are still defined — and must still agree word for word."""
    mods, ext = _shape(oracle, logn, bits, pbits)
    n, L = 1 << logn, len(mods)
    rng = np.random.default_rng(4000 + logn)
    wild = lambda *shape: rng.integers(0, 1 << 63, shape, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, shape, dtype=np.uint64)
    ct1, ct2, key = wild(2, L, n), wild(2, L, n), wild(L, 2, L + 1, n)
    quad = oracle.ckks_tensor(logn, mods, ct1, ct2)
    assert np.array_equal(dev.ckks_tensor(logn, mods, ct1, ct2), quad)
    e = oracle.ext_prod(logn, ext, ct1[0], key)
    assert np.array_equal(dev.ext_prod(logn, ext, ct1[0], key), e)
    wide = wild(2, L + 1, n)
    assert np.array_equal(dev.ckks_rescale(logn, ext, wide), oracle.ckks_rescale(logn, ext, wide))
    assert np.array_equal(dev.bgv_mod_switch(logn, ext, 65537, wide), oracle.bgv_mod_switch(logn, ext, 65537, wide))
    assert np.array_equal(dev.ckks_mult_relin(logn, ext, ct1, ct2, key), oracle.ckks_mult_relin(logn, ext, ct1, ct2, key))


# ------------------------------------------------------------------ RLWE cores
def _rlwe_inputs(oracle, logn, mods, seed, batch):
    n = 1 << logn
    rng = np.random.default_rng(seed)
    uni = lambda: np.stack([rng.integers(0, q, n, dtype=np.uint64) for q in mods])
    sk = uni()
    pts, c1s, es, cts = [], [], [], []
    for _ in range(batch):
        small = rng.integers(-19, 20, n)  # the reference's Gaussian is bounded by 6 sigma = 19.2 (sampling.cpp:49-58)
        es.append(np.stack([np.where(small < 0, q + small, small).astype(np.uint64) for q in mods]))
        pts.append(uni())
        c1s.append(uni())
        cts.append(np.stack([uni(), uni()]))
    return sk, np.stack(pts), np.stack(c1s), np.stack(es), np.stack(cts)


@pytest.mark.parametrize("logn,bits", [(5, [30, 30]), (10, [40, 30, 30]), (12, [59]), (13, [40, 30, 30, 30]), (15, [50, 50])])
def test_rlwe_encrypt_decrypt_cores_match_oracle(dev, oracle, logn, bits):
    """decrypt_core / encrypt_core (rlwe.cpp:34-71) with caller-supplied samples: raw words equal the
    oracle's (itself pinned on the reference, tests/test_oracle.py), and decrypt(encrypt(pt)) = pt + e."""
    mods, _ = _shape(oracle, logn, bits, 55)
    if bits == [59]:
        mods = [Q59]
    L, batch = len(mods), 3
    sk, pt, c1, e, ct = _rlwe_inputs(oracle, logn, mods, logn, batch)
    got_d = dev.rlwe_decrypt_core(logn, mods, ct, sk)
    got_e = dev.rlwe_encrypt_core(logn, mods, pt, sk, c1, e)
    for b in range(batch):
        assert np.array_equal(got_d[b], oracle.rlwe_decrypt_core(logn, mods, ct[b], sk))
        assert np.array_equal(got_e[b], oracle.rlwe_encrypt_core(logn, mods, pt[b], sk, c1[b], e[b]))
    back = dev.rlwe_decrypt_core(logn, mods, got_e, sk)
    for k, q in enumerate(mods):
        want = (pt[:, k].astype(object) + e[:, k].astype(object)) % q
        assert np.array_equal(back[:, k], want.astype(np.uint64))


def _rlwe_golden_inputs(oracle, g):
    mods, n = g["moduli"], 1 << g["logn"]
    L = len(mods)
    sk = np.stack([oracle.lcg_fill(4000 + k, mods[k], n) for k in range(L)])
    c1 = np.stack([oracle.lcg_fill(4100 + k, mods[k], n) for k in range(L)])
    pt = np.stack([oracle.lcg_fill(4200 + k, mods[k], n) for k in range(L)])
    small = oracle.lcg_fill(4300, 39, n).astype(np.int64) - 19
    err = np.stack([np.where(small < 0, mods[k] + small, small).astype(np.uint64) for k in range(L)])
    return sk, c1, pt, err, fill_ct(oracle, 4400, mods, n)


def test_rlwe_cores_reference_golden(dev, oracle, kat):
    """Hashes recorded from the unmodified reference (oracle/make_golden.py), N = 8192, L = 4."""
    g = kat["rlwe"]
    sk, c1, pt, err, ct = _rlwe_golden_inputs(oracle, g)
    assert hx(dev.rlwe_decrypt_core(g["logn"], g["moduli"], ct, sk)) == g["decrypt"]
    enc = dev.rlwe_encrypt_core(g["logn"], g["moduli"], pt, sk, c1, err)
    assert [hx(enc[h]) for h in range(2)] == g["encrypt"]


# ------------------------------------------------------------------ base transform + key generation
def _ternary_ntt(oracle, rng, logn, mods):
    t = rng.integers(-1, 2, 1 << logn)
    coeff = np.stack([np.where(t < 0, q + t, t).astype(np.uint64) for q in mods])
    return oracle.poly_ntt_fwd(logn, mods, coeff), coeff, t


def _small_errors(rng, n, mods, rows):
    out = []
    for _ in range(rows):
        sm = rng.integers(-19, 20, n)
        out.append(np.stack([np.where(sm < 0, q + sm, sm).astype(np.uint64) for q in mods]))
    return np.stack(out)


def test_rns_base_transform(dev, oracle):
    """rns_base_transform (rns_transform.cpp:11-126): one modulus -> many on lazy inputs, and the
    small-coefficient many -> one path; large coefficients are refused like the oracle refuses them."""
    rng = np.random.default_rng(9)
    for q_old, news in ((1099507695617, [1073479681, 1099510054913, Q59]), (65537, [260898817]), (Q59, [65537, 1099507695617])):
        x = rng.integers(0, 2 * q_old, (3, 257), dtype=np.uint64)
        got = dev.base_transform_from_single(q_old, x, news)
        for b in range(3):
            assert np.array_equal(got[b], oracle.base_transform_from_single(q_old, x[b], news))
    mods, P = oracle.ckks_pick_moduli([40, 30, 30], 45)
    mods, P = [int(m) for m in mods], int(P)
    small = np.stack([_ternary_ntt(oracle, rng, 8, mods)[1] for _ in range(4)])
    small[1] += np.array(mods, dtype=np.uint64)[:, None]  # lazy representatives: + q per limb
    got = dev.base_transform_to_single(mods, small, P)
    for b in range(4):
        assert np.array_equal(got[b], oracle.base_transform_to_single(mods, small[b], P))
    big = np.stack([rng.integers(0, q, 256, dtype=np.uint64) for q in mods])
    with pytest.raises(ValueError):
        oracle.base_transform_to_single(mods, big, P)
    from hehub_b200.binding import Unsupported
    with pytest.raises(Unsupported):
        dev.base_transform_to_single(mods, big, P)
    # one old modulus: the reference's dispatcher takes the one -> many branch (lazy Barrett only when the new modulus is smaller)
    for q_old, q_new in ((1099507695617, 1073479681), (1073479681, 1099507695617)):
        x = rng.integers(0, 2 * q_old, (2, 129), dtype=np.uint64)
        got = dev.base_transform_to_single([q_old], x[:, None, :], q_new)
        for b in range(2):
            assert np.array_equal(got[b].ravel(), oracle.base_transform_from_single(q_old, x[b], [q_new]).ravel())


def test_key_generation_reports_large_coefficients_at_the_next_synchronize(dev, oracle):
    """hehub_b200_ksk_generate is stream-asynchronous: when the secret's coefficients are not small (the reference then
    composes big integers, rns_transform.cpp:86-105, which the back end does not build) the verdict arrives with the
    next hehub_b200_ctx_synchronize instead of a host round trip inside key generation."""
    from hehub_b200.binding import Unsupported
    logn, n = 6, 64
    mods, ext = _shape(oracle, logn, [30, 30], 40)
    L = len(mods)
    rng = np.random.default_rng(5)
    big_sk = np.stack([rng.integers(0, q, n, dtype=np.uint64) for q in mods])  # INTT of this is not a small polynomial
    masks = np.stack([np.stack([rng.integers(0, q, n, dtype=np.uint64) for q in ext]) for _ in range(L)])
    errs = _small_errors(rng, n, ext, L)
    with pytest.raises(Unsupported):
        dev.ksk_generate(logn, ext, big_sk, big_sk, masks, errs)  # the binding synchronises when it downloads the key
    dev.synchronize()  # the verdict has been consumed: the context is usable again
    sk = _ternary_ntt(oracle, rng, logn, mods)[0]
    assert np.array_equal(dev.ksk_generate(logn, ext, sk, sk, masks, errs), oracle.ksk_generate(logn, ext, sk, sk, masks, errs))


@pytest.mark.parametrize("logn,bits,pbits", [(5, [30, 30], 40), (10, [40, 30, 30], 45), (12, [50], 55), (13, [40, 30, 30, 30], 40),
                                             (6, [50], 45), (10, [40], 30)])  # the last two: one component, P below it
def test_ksk_generate_matches_oracle(dev, oracle, logn, bits, pbits):
    """RlweKsk::RlweKsk (keys.cpp:8-36) on supplied samples: raw key words equal the oracle's."""
    mods, ext = _shape(oracle, logn, bits, pbits)
    L, n = len(mods), 1 << logn
    rng = np.random.default_rng(40 + logn)
    sk_o = _ternary_ntt(oracle, rng, logn, mods)[0]
    sk_c = _ternary_ntt(oracle, rng, logn, mods)[0]
    masks = np.stack([np.stack([rng.integers(0, q, n, dtype=np.uint64) for q in ext]) for _ in range(L)])
    errs = _small_errors(rng, n, ext, L)
    assert np.array_equal(dev.ksk_generate(logn, ext, sk_c, sk_o, masks, errs), oracle.ksk_generate(logn, ext, sk_c, sk_o, masks, errs))


def test_generated_relin_key_relinearizes(dev, oracle):
    """End to end on the device: encrypt two plaintexts, multiply, relinearize with a key generated by
    ksk_generate(s*s, s, P), decrypt: the result is the negacyclic product plus small noise."""
    logn, n = 8, 256
    mods, ext = _shape(oracle, logn, [50, 50], 55)
    L = len(mods)
    rng = np.random.default_rng(77)
    sk, _, s_int = _ternary_ntt(oracle, rng, logn, mods)
    sk2 = dev.mul_hybrid_lazy(mods, sk, sk)  # s*s in NTT form (keys.h:43)
    masks = np.stack([np.stack([rng.integers(0, q, n, dtype=np.uint64) for q in ext]) for _ in range(L)])
    key = dev.ksk_generate(logn, ext, sk2, sk, masks, _small_errors(rng, n, ext, L))
    m = [rng.integers(-50, 51, n) for _ in range(2)]
    cts = []
    for mi in m:
        pt = np.stack([np.where(mi < 0, q + mi, mi).astype(np.uint64) for q in mods])
        c1 = np.stack([rng.integers(0, q, n, dtype=np.uint64) for q in mods])
        cts.append(dev.rlwe_encrypt_core(logn, mods, pt, sk, c1, _small_errors(rng, n, mods, 1)[0]))
    prod = dev.ckks_mult_relin(logn, ext, cts[0], cts[1], key)
    dec = dev.rlwe_decrypt_core(logn, mods, prod, sk)
    # expected: (m0 + e0) * (m1 + e1) in Z[X]/(X^n + 1), up to relinearisation noise
    full = np.convolve(m[0].astype(object), m[1].astype(object))
    want = full[:n].copy()
    want[: n - 1] -= full[n:]
    q0 = mods[0]
    centred = np.array([int(v) if int(v) < q0 // 2 else int(v) - q0 for v in dec[0]], dtype=object)
    assert max(abs(int(a) - int(b)) for a, b in zip(centred, want)) < 2 ** 22  # noise ~ n * 50 * 19 * few


def _keygen_golden_inputs(oracle, g):
    mods, P, logn = g["moduli"], g["P"], g["logn"]
    ext, n, L = mods + [P], 1 << logn, len(mods)

    def ternary(seed):
        t = oracle.lcg_fill(seed, 3, n).astype(np.int64) - 1
        return np.stack([np.where(t < 0, q + t, t).astype(np.uint64) for q in mods])

    so, sc = ternary(5000), ternary(5001)
    masks = np.stack([np.stack([oracle.lcg_fill(5200 + 10 * p + k, ext[k], n) for k in range(L + 1)]) for p in range(L)])
    errs = []
    for p in range(L):
        sm = oracle.lcg_fill(5300 + p, 39, n).astype(np.int64) - 19
        errs.append(np.stack([np.where(sm < 0, q + sm, sm).astype(np.uint64) for q in ext]))
    return ext, so, sc, masks, np.stack(errs), oracle.lcg_fill(5400, 2 * mods[0], n)


def test_keygen_reference_golden(dev, oracle, kat):
    """Hashes recorded from the unmodified reference (oracle/make_golden.py): key-switch key rows and
    both base-transform directions at N = 8192, L = 4."""
    g = kat["keygen"]
    ext, so, sc, masks, errs, lazy = _keygen_golden_inputs(oracle, g)
    mods, logn = g["moduli"], g["logn"]
    sk_o, sk_c = dev.poly_ntt_fwd(logn, mods, so), dev.poly_ntt_fwd(logn, mods, sc)
    ksk = dev.ksk_generate(logn, ext, sk_c, sk_o, masks, errs)
    assert [hx(ksk[p]) for p in range(len(mods))] == g["ksk_rows"]
    assert hx(dev.base_transform_to_single(mods, so, g["P"])) == g["to_single"]
    assert hx(dev.base_transform_from_single(mods[0], lazy, ext[1:])) == g["from_single"]


def test_single_limb_key_switch(dev, oracle):
    """L = 1: the decomposition has one row; rescale of the (q0, P) result leaves one limb."""
    logn = 10
    mods, ext = _shape(oracle, logn, [40], 40)
    n = 1 << logn
    ct1, ct2, key = fill_ct(oracle, 1, mods, n), fill_ct(oracle, 2, mods, n), fill_key(oracle, 3, ext, n)
    assert np.array_equal(dev.ckks_mult_relin(logn, ext, ct1, ct2, key), oracle.ckks_mult_relin(logn, ext, ct1, ct2, key))


def test_batched_ops_and_waves(dev, oracle):
    """Independent ciphertexts in one call; a tiny scratch cap forces wave-by-wave processing."""
    logn = 10
    mods, ext = _shape(oracle, logn, [40, 30, 30], 40)
    n, batch = 1 << logn, 5
    ct1 = np.stack([fill_ct(oracle, 100 + 1000 * b, mods, n) for b in range(batch)])
    ct2 = np.stack([fill_ct(oracle, 200 + 1000 * b, mods, n) for b in range(batch)])
    key = fill_key(oracle, 7000, ext, n)
    want = np.stack([oracle.ckks_mult_relin(logn, ext, ct1[b], ct2[b], key) for b in range(batch)])
    assert np.array_equal(dev.ckks_mult_relin(logn, ext, ct1, ct2, key), want)
    dev.set_option("scratch_cap_mib", 1)
    try:
        assert np.array_equal(dev.ckks_mult_relin(logn, ext, ct1, ct2, key), want)
        want_rot = np.stack([oracle.ckks_rotate(logn, ext, ct1[b], key, 2) for b in range(batch)])
        assert np.array_equal(dev.ckks_rotate(logn, ext, ct1, key, 2), want_rot)
    finally:
        dev.set_option("scratch_cap_mib", 32768)
    want_rs = np.stack([oracle.ckks_rescale(logn, mods, ct1[b]) for b in range(batch)])
    assert np.array_equal(dev.ckks_rescale(logn, mods, ct1), want_rs)
    want_t = np.stack([oracle.ckks_tensor(logn, mods, ct1[b], ct2[b]) for b in range(batch)])
    assert np.array_equal(dev.ckks_tensor(logn, mods, ct1, ct2), want_t)


def test_rescale_is_exact_rounded_division(dev, oracle):
    """tests/ckks_t.cpp:136-175: CRT-composed (x + q_last/2) / q_last == rescaled, per coefficient."""
    logn, n = 3, 8
    mods = [int(m) for m in oracle.prime_row(34, 3)]
    big_q = mods[0] * mods[1] * mods[2]
    rng = np.random.default_rng(3)
    coeffs = [int(rng.integers(0, 1 << 62)) * int(rng.integers(0, 1 << 38)) % big_q for _ in range(n)]
    poly = np.array([[c % q for c in coeffs] for q in mods], dtype=np.uint64)
    ntt = oracle.poly_ntt_fwd(logn, mods, poly)
    ct = np.stack([ntt, ntt])
    out = dev.ckks_rescale(logn, mods, ct)
    back = oracle.poly_intt(logn, mods[:2], out[0], strict=True)
    q01 = mods[0] * mods[1]
    for i, c in enumerate(coeffs):
        want = ((c + mods[2] // 2) // mods[2]) % q01
        r0, r1 = int(back[0][i]), int(back[1][i])
        inv = pow(mods[0], -1, mods[1])
        got = (r0 + mods[0] * (((r1 - r0) * inv) % mods[1])) % q01
        assert got == want, i


# ------------------------------------------------------------------ size-independent properties at full size
@pytest.mark.parametrize("logn", [12, 15])
def test_linearity_and_round_trip_full_batch(dev, oracle, logn):
    """NTT(a) + NTT(b) == NTT(a + b) (mod q) and INTT(NTT(x)) == x over a batch (generated on the device)."""
    n = 1 << logn
    batch = 4096 if dev.kind == "gpu" else 8
    x = dev.lcg_fill([Q59], n, batch, 42, 1)
    if dev.kind == "sim" or logn == 12:
        assert np.array_equal(x[0], oracle.lcg_fill(42, Q59, n)) and np.array_equal(x[-1], oracle.lcg_fill(42 + batch - 1, Q59, n))
    y = dev.poly_ntt_fwd(logn, [Q59], x)
    for b in (0, batch // 2, batch - 1):
        assert np.array_equal(y[b], oracle.ntt_fwd_lazy(logn, Q59, x[b]))
    back = dev.poly_intt(logn, [Q59], y, strict=True)
    assert np.array_equal(back, x)
    q = np.uint64(Q59)
    s = (x[: batch // 2] + x[batch // 2:]) % q
    ys = dev.poly_ntt_fwd(logn, [Q59], s)
    assert np.array_equal(ys % q, (y[: batch // 2] % q + y[batch // 2:] % q) % q)


# ------------------------------------------------------------------ argument errors (reference: exceptions)
def test_argument_errors(dev):
    from hehub_b200.binding import InvalidArgument, Unsupported
    x = np.zeros(1 << 10, dtype=np.uint64)
    with pytest.raises(InvalidArgument):  # ntt.cpp:43-47: primes above 59 bits
        dev.ntt_fwd_lazy(10, (1 << 60) + 33, x)
    with pytest.raises(InvalidArgument):  # ntt.cpp:27-29: 2N must divide q - 1
        dev.ntt_fwd_lazy(10, 1000003, x)
    with pytest.raises((InvalidArgument, Unsupported)):
        dev.poly_ntt_fwd(16, [65537], np.zeros(1 << 16, dtype=np.uint64))
    ct = np.zeros((2, 1, 1 << 10), dtype=np.uint64)
    with pytest.raises(InvalidArgument):  # rescaling.cpp:27-29
        dev.ckks_rescale(10, [1073479681], ct)
    with pytest.raises(InvalidArgument):
        dev.mul_hybrid_lazy([1 << 20], x, x)  # Montgomery needs an odd modulus
    # an empty batch is a no-op, not an error
    assert dev.poly_ntt_fwd(10, [1073479681], np.zeros((0, 1, 1 << 10), dtype=np.uint64)).size == 0


# ------------------------------------------------------------------ host-buffer entry points
def test_host_buffer_pipeline(dev, oracle):
    """ntt_host / ckks_mult_relin_host stream a host batch through device staging slabs in chunks."""
    logn, n = 10, 1 << 10
    batch = 7  # not a multiple of the chunk count
    moduli = [Q59, 65537]
    x = dev.pinned((batch, 2, n))
    for b in range(batch):
        for k, q in enumerate(moduli):
            x[b, k] = oracle.lcg_fill(5 + 10 * b + k, q, n)
    want = np.stack([oracle.poly_ntt_fwd(logn, moduli, x[b]) for b in range(batch)])
    y = dev.pinned((batch, 2, n))
    dev.ntt_host(True, logn, moduli, x, y)
    assert np.array_equal(y, want)
    dev.ntt_host(False, logn, moduli, y, y, strict=True)  # in place on the host buffer
    assert np.array_equal(y, x)
    # a long batch with a small chunk: ramp-up chunks (1, 2, 4 units), full chunks of 8, a ragged one, ramp-down
    big = 41
    xb, yb = dev.pinned((big, 2, n)), dev.pinned((big, 2, n))
    for b in range(big):
        for k, q in enumerate(moduli):
            xb[b, k] = oracle.lcg_fill(1000 + 10 * b + k, q, n)
    dev.set_option("host_chunk_kib", 128)
    try:
        dev.ntt_host(True, logn, moduli, xb, yb)
    finally:
        dev.set_option("host_chunk_kib", 16384)
    assert np.array_equal(yb, np.stack([oracle.poly_ntt_fwd(logn, moduli, xb[b]) for b in range(big)]))
    mods, ext = _shape(oracle, logn, [40, 30], 40)
    ct1, ct2 = dev.pinned((batch, 2, 2, n)), dev.pinned((batch, 2, 2, n))
    for b in range(batch):
        ct1[b] = fill_ct(oracle, 100 + 1000 * b, mods, n)
        ct2[b] = fill_ct(oracle, 200 + 1000 * b, mods, n)
    key = fill_key(oracle, 9000, ext, n)
    dkey = dev.to_device(key)
    out = dev.pinned((batch, 2, 2, n))
    dev.ckks_mult_relin_host(logn, ext, ct1, ct2, dkey, out)
    dkey.free()
    want = np.stack([oracle.ckks_mult_relin(logn, ext, ct1[b], ct2[b], key) for b in range(batch)])
    assert np.array_equal(out, want)


@pytest.mark.parametrize("logn", [4, 10, 12])
def test_scheme_ops_on_unaligned_operands(dev, oracle, logn):
    """ct1 / ct2 / out that are only 8-byte aligned (a caller's sub-slab): the tensor product and the fused mult+relin take
    their one-word-per-thread paths instead of faulting on a misaligned 128-bit access."""
    n = 1 << logn
    mods, ext = _shape(oracle, logn, [40, 30], 40)
    L = len(mods)
    ct1, ct2, key = fill_ct(oracle, 61, mods, n), fill_ct(oracle, 62, mods, n), fill_key(oracle, 6300, ext, n)
    want_t = oracle.ckks_tensor(logn, mods, ct1, ct2)
    want_m = oracle.ckks_mult_relin(logn, ext, ct1, ct2, key)
    import ctypes as C
    p64 = C.POINTER(C.c_uint64)
    m = np.ascontiguousarray(np.asarray(ext, dtype=np.uint64))
    a, b, k = dev.slab(ct1.size + 1), dev.slab(ct2.size + 1), dev.to_device(key)
    quad, out = dev.slab(3 * L * n + 1), dev.slab(2 * L * n + 1)
    try:
        dev._call("slab_h2d", a.ptr + 8, ct1.ctypes.data, ct1.size)
        dev._call("slab_h2d", b.ptr + 8, ct2.ctypes.data, ct2.size)
        dev._call("ckks_tensor", logn, m.ctypes.data_as(p64), L, a.ptr + 8, b.ptr + 8, quad.ptr + 8, 1)
        got_t = np.empty(3 * L * n, dtype=np.uint64)
        dev._call("slab_d2h", got_t.ctypes.data, quad.ptr + 8, got_t.size)
        dev._call("ckks_mult_relin", logn, m.ctypes.data_as(p64), L, a.ptr + 8, b.ptr + 8, k.ptr, out.ptr + 8, 1)
        got_m = np.empty(2 * L * n, dtype=np.uint64)
        dev._call("slab_d2h", got_m.ctypes.data, out.ptr + 8, got_m.size)
        dev.synchronize()
    finally:
        for s in (a, b, k, quad, out):
            s.free()
    assert np.array_equal(got_t.reshape(want_t.shape), want_t)
    assert np.array_equal(got_m.reshape(want_m.shape), want_m)


@pytest.mark.gpu
def test_two_contexts_on_two_devices_in_one_process(oracle):
    """Kernel attributes (dynamic shared memory opt-in) are per device: a second context on another GPU of the same
    process must be able to launch the N >= 8192 transforms (72-147 KB of shared memory) too."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("one GPU on this box")
    from hehub_b200.binding import Context
    ctxs = [Context(device=0), Context(device=1)]
    try:
        for logn in (13, 15, 14):
            n = 1 << logn
            x = np.stack([oracle.lcg_fill(5 + r, Q59, n) for r in range(2)])
            want = np.stack([oracle.ntt_fwd_lazy(logn, Q59, r) for r in x])
            for c in (ctxs[1], ctxs[0], ctxs[1]):
                assert np.array_equal(c.poly_ntt_fwd(logn, [Q59], x), want)
                assert np.array_equal(c.poly_intt(logn, [Q59], want, strict=True), x)
    finally:
        for c in ctxs:
            c.close()


@pytest.mark.parametrize("logn", [3, 10, 12])
def test_rotate_in_place_and_unaligned(dev, oracle, logn):
    """ckks::rotate / conjugate read both polynomials through the Galois permutation inside the key-switch kernels
    (no separate gather pass): in place (out == ct), and on slabs that are only 8-byte aligned (64-bit gathers)."""
    n = 1 << logn
    mods, ext = _shape(oracle, logn, [40, 30, 30], 40)
    L = len(mods)
    ct = np.stack([fill_ct(oracle, 71 + b, mods, n) for b in range(3)])
    key = fill_key(oracle, 7300, ext, n)
    want_r = np.stack([oracle.ckks_rotate(logn, ext, ct[b], key, 7) for b in range(3)])
    want_c = np.stack([oracle.ckks_conjugate(logn, ext, ct[b], key) for b in range(3)])
    import ctypes as C
    p64 = C.POINTER(C.c_uint64)
    m = np.ascontiguousarray(np.asarray(ext, dtype=np.uint64))
    k = dev.to_device(key)
    a, out = dev.slab(ct.size + 1), dev.slab(ct.size + 1)
    try:
        for off in (0, 8):
            dev._call("slab_h2d", a.ptr + off, ct.ctypes.data, ct.size)
            dev._call("ckks_rotate", logn, m.ctypes.data_as(p64), L, a.ptr + off, k.ptr, 7, a.ptr + off, 3)  # in place
            got = np.empty_like(ct)
            dev._call("slab_d2h", got.ctypes.data, a.ptr + off, ct.size)
            dev.synchronize()
            assert np.array_equal(got, want_r), off
            dev._call("slab_h2d", a.ptr + off, ct.ctypes.data, ct.size)
            dev._call("ckks_conjugate", logn, m.ctypes.data_as(p64), L, a.ptr + off, k.ptr, out.ptr + off, 3)
            dev._call("slab_d2h", got.ctypes.data, out.ptr + off, ct.size)
            dev.synchronize()
            assert np.array_equal(got, want_c), off
    finally:
        for s in (a, out, k):
            s.free()


@pytest.mark.parametrize("logn,bits,batch", [(9, [40, 30], 19), (10, [40, 30, 30], 9), (8, [34], 33)])
def test_key_switch_inner_product_with_staged_key(dev, oracle, logn, bits, batch):
    """From 8 ciphertexts per wave on, the key-switch inner product keeps its key slice in shared memory and walks groups of
    ciphertexts per CTA (ragged last group included); aligned and 8-byte-aligned operands, with and without a Galois gather."""
    n = 1 << logn
    mods, ext = _shape(oracle, logn, bits, 40)
    L = len(mods)
    polys = np.stack([np.stack([oracle.lcg_fill(900 + 31 * b + k, mods[k], n) for k in range(L)]) for b in range(batch)])
    key = fill_key(oracle, 9100, ext, n)
    want = np.stack([oracle.ext_prod(logn, ext, polys[b], key) for b in range(batch)])
    assert np.array_equal(dev.ext_prod(logn, ext, polys, key), want)
    ct = np.stack([fill_ct(oracle, 950 + 7 * b, mods, n) for b in range(batch)])
    want_rot = np.stack([oracle.ckks_rotate(logn, ext, ct[b], key, 3) for b in range(batch)])
    assert np.array_equal(dev.ckks_rotate(logn, ext, ct, key, 3), want_rot)
    import ctypes as C
    p64 = C.POINTER(C.c_uint64)
    m = np.ascontiguousarray(np.asarray(ext, dtype=np.uint64))
    a, k, out = dev.slab(polys.size + 1), dev.to_device(key), dev.slab(want.size + 1)
    try:
        dev._call("slab_h2d", a.ptr + 8, polys.ctypes.data, polys.size)
        dev._call("ext_prod_montgomery", logn, m.ctypes.data_as(p64), L, a.ptr + 8, k.ptr, out.ptr + 8, batch)
        got = np.empty_like(want)
        dev._call("slab_d2h", got.ctypes.data, out.ptr + 8, want.size)
        dev.synchronize()
    finally:
        for s in (a, k, out):
            s.free()
    assert np.array_equal(got, want)


@pytest.mark.gpu
@pytest.mark.parametrize("logn,bits", [(12, [39, 30]), (13, [40, 30, 30, 30])])
def test_single_launch_mult_matches_oracle(oracle, logn, bits):
    """Option single_launch: one ckks::mult pair per call as ONE launch (six phases, grid barriers) — same words as the six
    launches and as the oracle, call after call (the barrier counter only grows)."""
    import subprocess, sys, os
    # in a child process with a wall-clock limit: a grid barrier that never completes must not take the test session with it
    code = f"""
import sys, numpy as np
sys.path.insert(0, {os.path.dirname(os.path.dirname(os.path.abspath(__file__)))!r})
sys.path.insert(0, {os.path.dirname(os.path.abspath(__file__))!r})
from hehub_b200.binding import Context, pick_moduli
from oracle.binding import Oracle
from conftest import fill_ct, fill_key
orc, ctx = Oracle(), Context(device=0)
mods, p = pick_moduli({bits!r}, {bits[0]}, ctx.lib)
ext, n = mods + [p], 1 << {logn}
ctx.set_option("single_launch", 1)
for rep in range(3):
    ct1, ct2, key = fill_ct(orc, 11 + rep, mods, n), fill_ct(orc, 21 + rep, mods, n), fill_key(orc, 3100 + rep, ext, n)
    before = ctx.launch_count()
    got = ctx.ckks_mult_relin({logn}, ext, ct1, ct2, key)
    assert ctx.launch_count() - before == 1, ctx.launch_count() - before
    assert np.array_equal(got, orc.ckks_mult_relin({logn}, ext, ct1, ct2, key)), rep
ctx.set_option("single_launch", 0)
assert np.array_equal(ctx.ckks_mult_relin({logn}, ext, ct1, ct2, key), got)
print("ok")
"""
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=180)
    assert res.returncode == 0 and "ok" in res.stdout, res.stdout + res.stderr


@pytest.mark.parametrize("logn,bits,pbits,batch", [(12, [39, 30, 30], 39, 1), (13, [40, 30, 30, 30], 40, 2), (13, [59, 50], 59, 1), (12, [39], 39, 3)])
def test_pair_path_matches_oracle_and_wave_path(dev, oracle, logn, bits, pbits, batch):
    """The two-launch key switch (csrc/ks_pair.cuh: inverse transform handed to the forward transforms in registers, tensor
    product / inner product / drop epilogue inside the loads and stores) produces the words of the six-launch wave path and of
    the oracle — for ckks::mult, relinearize, rotate, conjugate and their bgv:: twins, at every split of the targets."""
    mods, ext = _shape(oracle, logn, bits, pbits)
    n, L = 1 << logn, len(mods)
    cts1 = [fill_ct(oracle, 31 + 7 * b, mods, n) for b in range(batch)]
    cts2 = [fill_ct(oracle, 41 + 7 * b, mods, n) for b in range(batch)]
    key = fill_key(oracle, 3300, ext, n)
    ct1, ct2 = (cts1[0], cts2[0]) if batch == 1 else (np.stack(cts1), np.stack(cts2))
    quads = [oracle.ckks_tensor(logn, mods, a, b) for a, b in zip(cts1, cts2)]
    quad = quads[0] if batch == 1 else np.stack(quads)
    pack = (lambda xs: xs[0]) if batch == 1 else np.stack
    want = {
        "mult": pack([oracle.ckks_mult_relin(logn, ext, a, b, key) for a, b in zip(cts1, cts2)]),
        "relin": pack([oracle.ckks_relinearize(logn, ext, q, key) for q in quads]),
        "bgv": pack([oracle.bgv_relinearize(logn, ext, 65537, q, key) for q in quads]),
        "rot": pack([oracle.ckks_rotate(logn, ext, a, key, 3) for a in cts1]),
        "conj": pack([oracle.ckks_conjugate(logn, ext, a, key) for a in cts1]),
    }
    if L >= 2:  # (one limb: nothing to drop)
        want["rescale"] = pack([oracle.ckks_rescale(logn, mods, a) for a in cts1])
        want["modsw"] = pack([oracle.bgv_mod_switch(logn, mods, 65537, a) for a in cts1])
    try:
        # (pair_path, targets per cluster, launches per call)
        variants = [(0, 0, None), (2, 0, 2), (2, 1, 2), (2, 2, 2), (2, L, 2)]
        if dev.kind == "sim":  # the emulator runs every thread of a cluster as a host thread: fewer variants per shape
            variants = ([(2, 2, 2), (2, 0, 2)] if L > 1 else [(2, 0, 2)]) if logn == 12 else [(2, 1 if L == 2 else 0, 2)]
        for path, tpc, launches in variants:
            dev.set_option("pair_path", path)
            dev.set_option("pair_tpc", tpc)
            before = dev.launch_count()
            assert np.array_equal(dev.ckks_mult_relin(logn, ext, ct1, ct2, key), want["mult"]), (path, tpc)
            if launches:
                assert dev.launch_count() - before == launches
            assert np.array_equal(dev.ckks_rotate(logn, ext, ct1, key, 3), want["rot"]), (path, tpc)
            if not (dev.kind == "sim" and batch > 1):  # (the emulator's largest case keeps to one op of each kernel family)
                assert np.array_equal(dev.ckks_relinearize(logn, ext, quad, key), want["relin"]), (path, tpc)
                assert np.array_equal(dev.bgv_relinearize(logn, ext, 65537, quad, key), want["bgv"]), (path, tpc)
                assert np.array_equal(dev.bgv_mult_relin(logn, ext, 65537, ct1, ct2, key), want["bgv"]), (path, tpc)
                assert np.array_equal(dev.ckks_conjugate(logn, ext, ct1, key), want["conj"]), (path, tpc)
            if L < 2:
                continue
            # rescale / mod-switch of a few ciphertexts: one cluster launch (the same drop kernel, the ciphertext as its source)
            before = dev.launch_count()
            assert np.array_equal(dev.ckks_rescale(logn, mods, ct1), want["rescale"]), (path, tpc)
            if launches:
                assert dev.launch_count() - before == 1
            assert np.array_equal(dev.bgv_mod_switch(logn, mods, 65537, ct1), want["modsw"]), (path, tpc)
    finally:
        dev.set_option("pair_path", 1)
        dev.set_option("pair_tpc", 0)


@pytest.mark.parametrize("logn,bits,pbits", [(14, [45, 40, 40], 45), (15, [50, 50], 55)])
def test_fused_drop_matches_oracle(dev, oracle, logn, bits, pbits):
    """One ciphertext per call at N = 16384 / 32768: the inverse transform of the P limb rides in the key switch's inner-product
    launch (option fused_drop, ext_mac_intt_kernel) — five launches instead of six per relinearize, the same words."""
    mods, ext = _shape(oracle, logn, bits, pbits)
    n = 1 << logn
    ct1, ct2, key = fill_ct(oracle, 51, mods, n), fill_ct(oracle, 52, mods, n), fill_key(oracle, 3500, ext, n)
    quad = oracle.ckks_tensor(logn, mods, ct1, ct2)
    want = {"mult": oracle.ckks_mult_relin(logn, ext, ct1, ct2, key), "bgv": oracle.bgv_relinearize(logn, ext, 65537, quad, key),
            "rot": oracle.ckks_rotate(logn, ext, ct1, key, 7), "conj": oracle.ckks_conjugate(logn, ext, ct1, key)}
    try:
        for fused, launches in ((1, 5), (0, 6)):
            dev.set_option("fused_drop", fused)
            before = dev.launch_count()
            assert np.array_equal(dev.ckks_mult_relin(logn, ext, ct1, ct2, key), want["mult"]), fused
            assert dev.launch_count() - before == launches, (fused, dev.launch_count() - before)
            assert np.array_equal(dev.bgv_relinearize(logn, ext, 65537, quad, key), want["bgv"]), fused
            assert np.array_equal(dev.ckks_rotate(logn, ext, ct1, key, 7), want["rot"]), fused
            assert np.array_equal(dev.ckks_conjugate(logn, ext, ct1, key), want["conj"]), fused
            if fused and dev.kind == "gpu":  # the launch's two counters are put back by the kernel itself: call after call
                for rep in range(10):
                    assert np.array_equal(dev.ckks_mult_relin(logn, ext, ct1, ct2, key), want["mult"]), rep
    finally:
        dev.set_option("fused_drop", 1)
