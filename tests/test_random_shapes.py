"""Randomised shapes (hypothesis): ring size, modulus chain, batch and operand words are drawn at random and the CUDA
path (through the C ABI; the CTA-emulator build in the CPU suite) must reproduce the oracle word for word.  The fixed
cases of test_parity.py pin the BASELINE configs; this file looks for the shapes nobody thought of — ragged batches
against the wave size, chains whose last prime is larger or smaller than the others, one-limb chains, tiny rings."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

BITS = [27, 30, 33, 36, 40, 45, 50, 55, 59]


@pytest.fixture(scope="module", params=["sim", pytest.param("gpu", marks=pytest.mark.gpu)])
def dev(request):
    from hehub_b200.binding import Context
    if request.param == "sim":
        import __graft_entry__ as ge
        ctx = Context(lib_path=ge.build_sim())
    else:
        ctx = Context(device=0)
    ctx.kind = request.param
    yield ctx
    ctx.close()


def _words(rng, moduli, n, lead=()):
    return np.stack([rng.integers(0, 2 * q, lead + (n,), dtype=np.uint64) for q in moduli], axis=len(lead))


@settings(max_examples=25, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])
@given(logn=st.integers(1, 11), bits=st.lists(st.sampled_from(BITS), min_size=1, max_size=4), pbits=st.sampled_from(BITS),
       batch=st.integers(1, 5), seed=st.integers(0, 2**31), cap=st.sampled_from([1, 8192]), t=st.sampled_from([2, 65537, 1032193]))
def test_random_chains_match_oracle(dev, oracle, logn, bits, pbits, batch, seed, cap, t):
    try:
        mods, p = oracle.ckks_pick_moduli(sorted(bits, reverse=(seed & 1) == 0), pbits)
    except ValueError:
        return  # more primes of one size than the reference's table holds
    mods = [int(m) for m in mods]
    ext = mods + [int(p)]
    n, L = 1 << logn, len(mods)
    if any((q - 1) % (2 * n) for q in ext):
        return
    rng = np.random.default_rng(seed)
    ct1, ct2 = _words(rng, mods, n, (batch, 2)), _words(rng, mods, n, (batch, 2))
    key = _words(rng, ext, n, (L, 2))
    wide = _words(rng, ext, n, (batch, 2))
    dev.set_option("scratch_cap_mib", cap)
    try:
        got = dev.ckks_mult_relin(logn, ext, ct1, ct2, key)
        want = np.stack([oracle.ckks_mult_relin(logn, ext, ct1[b], ct2[b], key) for b in range(batch)])
        assert np.array_equal(got, want)
        got = dev.ckks_rescale(logn, ext, wide)
        assert np.array_equal(got, np.stack([oracle.ckks_rescale(logn, ext, wide[b]) for b in range(batch)]))
        got = dev.bgv_mod_switch(logn, ext, t, wide)
        assert np.array_equal(got, np.stack([oracle.bgv_mod_switch(logn, ext, t, wide[b]) for b in range(batch)]))
        x = ct1[:, 0]
        assert np.array_equal(dev.poly_intt(logn, mods, dev.poly_ntt_fwd(logn, mods, x), strict=True) % np.array(mods, dtype=np.uint64)[None, :, None],
                              x % np.array(mods, dtype=np.uint64)[None, :, None])
    finally:
        dev.set_option("scratch_cap_mib", 8192)
