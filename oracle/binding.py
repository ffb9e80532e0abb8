"""ctypes bindings for the CPU oracle and the compiled reference (TEST INFRASTRUCTURE ONLY).

`Oracle()` wraps oracle/_build/libhehub_oracle.so (the C restatement, symbols ``orc_*``);
`Reference()` wraps oracle/_ref/libhehub_ref.so (the unmodified reference behind
oracle/ref_shim.cpp, symbols ``ref_*``).  Both expose the same method names so a parity
test can drive either.  Nothing under hehub_b200/ may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "_build", "libhehub_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libhehub_ref.so")
REF_ROOT = os.environ.get("HEHUB_REFERENCE_ROOT", "/root/reference")

u64 = C.c_uint64
p64 = C.POINTER(C.c_uint64)


def build_oracle(force: bool = False) -> str:
    src = os.path.join(HERE, "hehub_oracle.c")
    if force or not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    return ORACLE_SO


def build_reference(force: bool = False) -> str | None:
    """Compile the real reference when its sources are reachable; otherwise keep any prebuilt .so."""
    if os.path.isdir(os.path.join(REF_ROOT, "src", "fhe")):
        shim = os.path.join(HERE, "ref_shim.cpp")
        if force or not os.path.exists(REF_SO) or os.path.getmtime(REF_SO) < os.path.getmtime(shim):
            subprocess.check_call(["make", "-s", "-C", HERE, "ref", f"REF={REF_ROOT}"])
    return REF_SO if os.path.exists(REF_SO) else None


def _arr(a) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint64)
    return a


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(p64)


class _CpuLib:
    prefix = ""

    def __init__(self, path: str):
        self.path = path
        self.lib = C.CDLL(path)

    def _fn(self, name, restype, *argtypes):
        f = getattr(self.lib, self.prefix + name)
        f.restype = restype
        f.argtypes = list(argtypes)
        return f

    # ---- word level -------------------------------------------------
    def harvey_lazy(self, q, x, w, wh) -> int:
        return self._fn("harvey_lazy", u64, u64, u64, u64, u64)(q, x, w, wh)

    def inverse_mod_prime(self, elem, prime) -> int:
        return self._fn("inverse_mod_prime", u64, u64, u64)(elem, prime)

    def _inplace(self, name, q, x):
        x = _arr(x).copy()
        self._fn(name, None, u64, C.c_size_t, p64)(q, x.size, _ptr(x))
        return x

    def barrett_lazy(self, q, x):
        return self._inplace("barrett_lazy", q, x)

    def barrett(self, q, x):
        return self._inplace("barrett", q, x)

    def reduce_strict(self, q, x):
        return self._inplace("reduce_strict", q, x)

    def _binary(self, name, q, a, b):
        a, b = _arr(a), _arr(b)
        c = np.empty_like(a)
        self._fn(name, None, u64, C.c_size_t, p64, p64, p64)(q, a.size, _ptr(a), _ptr(b), _ptr(c))
        return c

    def mul_hybrid_lazy(self, q, a, b):
        return self._binary("mul_hybrid_lazy", q, a, b)

    def mul_barrett_lazy(self, q, a, b):
        return self._binary("mul_barrett_lazy", q, a, b)

    def montgomery128_lazy(self, q, in_lohi):
        a = _arr(in_lohi)
        out = np.empty(a.size // 2, dtype=np.uint64)
        self._fn("montgomery128_lazy", None, u64, C.c_size_t, p64, p64)(q, out.size, _ptr(a), _ptr(out))
        return out

    # ---- transforms -------------------------------------------------
    def ntt_fwd_lazy(self, logn, q, x):
        x = _arr(x).copy()
        rc = self._fn("ntt_fwd_lazy", C.c_int, C.c_uint, u64, p64)(logn, q, _ptr(x))
        if rc:
            raise ValueError(f"ntt_fwd_lazy rc={rc}")
        return x

    def intt_lazy(self, logn, q, x):
        x = _arr(x).copy()
        rc = self._fn("intt_lazy", C.c_int, C.c_uint, u64, p64)(logn, q, _ptr(x))
        if rc:
            raise ValueError(f"intt_lazy rc={rc}")
        return x

    def bench_ntt(self, logn, q, x, forward=True):
        """In place on the caller's [rows][N] array (timing helper; no copy)."""
        assert x.dtype == np.uint64 and x.flags.c_contiguous
        rows = x.size >> logn
        rc = self._fn("bench_ntt", C.c_int, C.c_uint, u64, p64, C.c_size_t, C.c_int)(logn, q, _ptr(x), rows, int(forward))
        if rc:
            raise ValueError(f"bench_ntt rc={rc}")

    def poly_ntt_fwd(self, logn, moduli, x):
        m, x = _arr(moduli), _arr(x).copy()
        rc = self._fn("poly_ntt_fwd", C.c_int, C.c_uint, C.c_size_t, p64, p64)(logn, m.size, _ptr(m), _ptr(x))
        if rc:
            raise ValueError(f"poly_ntt_fwd rc={rc}")
        return x

    def poly_intt(self, logn, moduli, x, strict=False):
        m, x = _arr(moduli), _arr(x).copy()
        rc = self._fn("poly_intt", C.c_int, C.c_uint, C.c_size_t, p64, p64, C.c_int)(
            logn, m.size, _ptr(m), _ptr(x), int(strict))
        if rc:
            raise ValueError(f"poly_intt rc={rc}")
        return x

    # ---- composite --------------------------------------------------
    def ckks_tensor(self, logn, moduli, ct1, ct2):
        m, a, b = _arr(moduli), _arr(ct1), _arr(ct2)
        L, n = m.size, 1 << logn
        out = np.empty((3, L, n), dtype=np.uint64)
        rc = self._fn("ckks_tensor", C.c_int, C.c_uint, C.c_size_t, p64, p64, p64, p64)(
            logn, L, _ptr(m), _ptr(a), _ptr(b), _ptr(out))
        if rc:
            raise ValueError(f"ckks_tensor rc={rc}")
        return out

    def ext_prod(self, logn, ext_moduli, poly, key):
        m, a, k = _arr(ext_moduli), _arr(poly), _arr(key)
        L, n = m.size - 1, 1 << logn
        out = np.empty((2, L + 1, n), dtype=np.uint64)
        rc = self._fn("ext_prod", C.c_int, C.c_uint, C.c_size_t, p64, p64, p64, p64)(
            logn, L, _ptr(m), _ptr(a), _ptr(k), _ptr(out))
        if rc:
            raise ValueError(f"ext_prod rc={rc}")
        return out

    def ckks_rescale(self, logn, moduli, ct):
        m, a = _arr(moduli), _arr(ct)
        L, n = m.size, 1 << logn
        out = np.empty((2, L - 1, n), dtype=np.uint64)
        rc = self._fn("ckks_rescale", C.c_int, C.c_uint, C.c_size_t, p64, p64, p64)(
            logn, L, _ptr(m), _ptr(a), _ptr(out))
        if rc:
            raise ValueError(f"ckks_rescale rc={rc}")
        return out

    def bgv_mod_switch(self, logn, moduli, t, ct):
        m, a = _arr(moduli), _arr(ct)
        L, n = m.size, 1 << logn
        out = np.empty((2, L - 1, n), dtype=np.uint64)
        rc = self._fn("bgv_mod_switch", C.c_int, C.c_uint, C.c_size_t, p64, u64, p64, p64)(
            logn, L, _ptr(m), t, _ptr(a), _ptr(out))
        if rc:
            raise ValueError(f"bgv_mod_switch rc={rc}")
        return out

    def ckks_relinearize(self, logn, ext_moduli, quad, key):
        m, a, k = _arr(ext_moduli), _arr(quad), _arr(key)
        L, n = m.size - 1, 1 << logn
        out = np.empty((2, L, n), dtype=np.uint64)
        rc = self._fn("ckks_relinearize", C.c_int, C.c_uint, C.c_size_t, p64, p64, p64, p64)(
            logn, L, _ptr(m), _ptr(a), _ptr(k), _ptr(out))
        if rc:
            raise ValueError(f"ckks_relinearize rc={rc}")
        return out

    def bgv_relinearize(self, logn, ext_moduli, t, quad, key):
        m, a, k = _arr(ext_moduli), _arr(quad), _arr(key)
        L, n = m.size - 1, 1 << logn
        out = np.empty((2, L, n), dtype=np.uint64)
        rc = self._fn("bgv_relinearize", C.c_int, C.c_uint, C.c_size_t, p64, u64, p64, p64, p64)(
            logn, L, _ptr(m), t, _ptr(a), _ptr(k), _ptr(out))
        if rc:
            raise ValueError(f"bgv_relinearize rc={rc}")
        return out

    def ckks_mult_relin(self, logn, ext_moduli, ct1, ct2, key):
        m, a, b, k = _arr(ext_moduli), _arr(ct1), _arr(ct2), _arr(key)
        L, n = m.size - 1, 1 << logn
        out = np.empty((2, L, n), dtype=np.uint64)
        rc = self._fn("ckks_mult_relin", C.c_int, C.c_uint, C.c_size_t, p64, p64, p64, p64, p64)(
            logn, L, _ptr(m), _ptr(a), _ptr(b), _ptr(k), _ptr(out))
        if rc:
            raise ValueError(f"ckks_mult_relin rc={rc}")
        return out

    def rlwe_decrypt_core(self, logn, moduli, ct, sk):
        m, ct, sk = _arr(moduli), _arr(ct), _arr(sk)
        out = np.empty((m.size, 1 << logn), dtype=np.uint64)
        rc = self._fn("rlwe_decrypt_core", C.c_int, C.c_uint, C.c_size_t, p64, p64, p64, p64)(
            logn, m.size, _ptr(m), _ptr(ct), _ptr(sk), _ptr(out))
        if rc:
            raise ValueError(f"rlwe_decrypt_core rc={rc}")
        return out

    def rlwe_encrypt_core(self, logn, moduli, pt, sk, c1, e):
        m, pt, sk, c1, e = _arr(moduli), _arr(pt), _arr(sk), _arr(c1), _arr(e)
        out = np.empty((2, m.size, 1 << logn), dtype=np.uint64)
        rc = self._fn("rlwe_encrypt_core", C.c_int, C.c_uint, C.c_size_t, p64, p64, p64, p64, p64, p64)(
            logn, m.size, _ptr(m), _ptr(pt), _ptr(sk), _ptr(c1), _ptr(e), _ptr(out))
        if rc:
            raise ValueError(f"rlwe_encrypt_core rc={rc}")
        return out

    def base_transform_from_single(self, q_old, x, new_moduli):
        x, m = _arr(x), _arr(new_moduli)
        out = np.empty((m.size, x.size), dtype=np.uint64)
        rc = self._fn("base_transform_from_single", C.c_int, u64, C.c_size_t, p64, p64, C.c_size_t, p64)(
            q_old, x.size, _ptr(x), _ptr(m), m.size, _ptr(out))
        if rc:
            raise ValueError(f"base_transform_from_single rc={rc}")
        return out

    def base_transform_to_single(self, old_moduli, x, new_modulus):
        m, x = _arr(old_moduli), _arr(x)
        n = x.size // m.size
        out = np.empty(n, dtype=np.uint64)
        rc = self._fn("base_transform_to_single", C.c_int, C.c_size_t, C.c_size_t, p64, p64, u64, p64)(
            n, m.size, _ptr(m), _ptr(x), new_modulus, _ptr(out))
        if rc:
            raise ValueError(f"base_transform_to_single rc={rc}")
        return out

    def ksk_generate(self, logn, ext_moduli, sk_curr, sk_orig, masks, errors):
        m = _arr(ext_moduli)
        L, n = m.size - 1, 1 << logn
        key = np.empty((L, 2, L + 1, n), dtype=np.uint64)
        rc = self._fn("ksk_generate", C.c_int, C.c_uint, C.c_size_t, p64, p64, p64, p64, p64, p64)(
            logn, L, _ptr(m), _ptr(_arr(sk_curr)), _ptr(_arr(sk_orig)), _ptr(_arr(masks)), _ptr(_arr(errors)), _ptr(key))
        if rc:
            raise ValueError(f"ksk_generate rc={rc}")
        return key

    def galois_cycle(self, logn, poly, step):
        a = _arr(poly)
        L = a.size >> logn
        out = np.empty_like(a)
        rc = self._fn("galois_cycle", C.c_int, C.c_uint, C.c_size_t, p64, p64, C.c_size_t)(
            logn, L, _ptr(a), _ptr(out), step)
        if rc:
            raise ValueError(f"galois_cycle rc={rc}")
        return out

    def galois_involution(self, logn, poly):
        a = _arr(poly)
        L = a.size >> logn
        out = np.empty_like(a)
        rc = self._fn("galois_involution", C.c_int, C.c_uint, C.c_size_t, p64, p64)(
            logn, L, _ptr(a), _ptr(out))
        if rc:
            raise ValueError(f"galois_involution rc={rc}")
        return out

    def ckks_rotate(self, logn, ext_moduli, ct, key, step):
        m, a, k = _arr(ext_moduli), _arr(ct), _arr(key)
        L, n = m.size - 1, 1 << logn
        out = np.empty((2, L, n), dtype=np.uint64)
        rc = self._fn("ckks_rotate", C.c_int, C.c_uint, C.c_size_t, p64, p64, p64, C.c_size_t, p64)(
            logn, L, _ptr(m), _ptr(a), _ptr(k), step, _ptr(out))
        if rc:
            raise ValueError(f"ckks_rotate rc={rc}")
        return out

    def ckks_conjugate(self, logn, ext_moduli, ct, key):
        m, a, k = _arr(ext_moduli), _arr(ct), _arr(key)
        L, n = m.size - 1, 1 << logn
        out = np.empty((2, L, n), dtype=np.uint64)
        rc = self._fn("ckks_conjugate", C.c_int, C.c_uint, C.c_size_t, p64, p64, p64, p64)(
            logn, L, _ptr(m), _ptr(a), _ptr(k), _ptr(out))
        if rc:
            raise ValueError(f"ckks_conjugate rc={rc}")
        return out

    # ---- parameters -------------------------------------------------
    def prime_row(self, bits, count=20):
        out = np.zeros(count, dtype=np.uint64)
        got = self._fn("prime_row", C.c_int, C.c_uint, C.c_size_t, p64)(bits, count, _ptr(out))
        return [int(v) for v in out[:got]]

    def ckks_pick_moduli(self, moduli_bits, additional_bits):
        bits = (C.c_uint * len(moduli_bits))(*moduli_bits)
        out = np.zeros(len(moduli_bits), dtype=np.uint64)
        extra = u64(0)
        rc = self._fn("ckks_pick_moduli", C.c_int, C.POINTER(C.c_uint), C.c_size_t, C.c_uint, p64,
                      C.POINTER(u64))(bits, len(moduli_bits), additional_bits, _ptr(out), C.byref(extra))
        if rc:
            raise ValueError(f"ckks_pick_moduli rc={rc}")
        return [int(v) for v in out], int(extra.value)


class Oracle(_CpuLib):
    """The C restatement (oracle/hehub_oracle.c)."""
    prefix = "orc_"

    def __init__(self):
        super().__init__(build_oracle())

    def intt_lazy_folded(self, logn, q, x):
        x = _arr(x).copy()
        rc = self._fn("intt_lazy_folded", C.c_int, C.c_uint, u64, p64)(logn, q, _ptr(x))
        if rc:
            raise ValueError(f"intt_lazy_folded rc={rc}")
        return x

    def ntt_tables(self, logn, q):
        n = 1 << logn
        fwd, fwd_h = np.empty(n, np.uint64), np.empty(n, np.uint64)
        inv, inv_h = np.empty(2 * n, np.uint64), np.empty(2 * n, np.uint64)
        rc = self._fn("ntt_tables", C.c_int, C.c_uint, u64, p64, p64, p64, p64)(
            logn, q, _ptr(fwd), _ptr(fwd_h), _ptr(inv), _ptr(inv_h))
        if rc:
            raise ValueError(f"ntt_tables rc={rc}")
        return fwd, fwd_h, inv, inv_h

    def mont_consts(self, q):
        a, b, c = u64(0), u64(0), u64(0)
        self._fn("mont_consts", None, u64, C.POINTER(u64), C.POINTER(u64), C.POINTER(u64))(
            q, C.byref(a), C.byref(b), C.byref(c))
        return int(a.value), int(b.value), int(c.value)

    def root_2n(self, q, n):
        return self._fn("root_2n", u64, u64, u64)(q, n)

    def pow_mod(self, q, base, e):
        return self._fn("pow_mod", u64, u64, u64, u64)(q, base, e)

    def add_lazy(self, q, x, y):
        x, y = _arr(x).copy(), _arr(y)
        self._fn("add_lazy", None, u64, C.c_size_t, p64, p64)(q, x.size, _ptr(x), _ptr(y))
        return x

    def sub_lazy(self, q, x, y):
        x, y = _arr(x).copy(), _arr(y)
        self._fn("sub_lazy", None, u64, C.c_size_t, p64, p64)(q, x.size, _ptr(x), _ptr(y))
        return x

    def mul_scalar_lazy(self, q, x, scalar):
        x = _arr(x).copy()
        self._fn("mul_scalar_lazy", None, u64, C.c_size_t, p64, u64)(q, x.size, _ptr(x), scalar)
        return x

    def lcg_fill(self, seed, q, n):
        x = np.empty(n, dtype=np.uint64)
        self._fn("lcg_fill", None, u64, u64, C.c_size_t, p64)(seed, q, n, _ptr(x))
        return x

    def fnv1a(self, x, h=0):
        x = _arr(x)
        return self._fn("fnv1a", u64, p64, C.c_size_t, u64)(_ptr(x), x.size, h)


class Reference(_CpuLib):
    """The unmodified reference compiled into oracle/_ref/ (None-able: see `available`)."""
    prefix = "ref_"

    def __init__(self):
        path = build_reference()
        if path is None:
            raise FileNotFoundError("oracle/_ref/libhehub_ref.so not built and reference sources absent")
        super().__init__(path)

    @staticmethod
    def available() -> bool:
        return build_reference() is not None

    def _poly_binary(self, name, logn, moduli, x, y):
        m, x, y = _arr(moduli), _arr(x).copy(), _arr(y)
        rc = self._fn(name, C.c_int, C.c_uint, C.c_size_t, p64, p64, p64)(logn, m.size, _ptr(m), _ptr(x), _ptr(y))
        if rc:
            raise ValueError(f"{name} rc={rc}")
        return x

    def poly_add(self, logn, moduli, x, y):
        return self._poly_binary("poly_add", logn, moduli, x, y)

    def poly_sub(self, logn, moduli, x, y):
        return self._poly_binary("poly_sub", logn, moduli, x, y)

    def poly_mul_scalar(self, logn, moduli, x, scalar):
        m, x = _arr(moduli), _arr(x).copy()
        rc = self._fn("poly_mul_scalar", C.c_int, C.c_uint, C.c_size_t, p64, p64, u64)(
            logn, m.size, _ptr(m), _ptr(x), scalar)
        if rc:
            raise ValueError(f"poly_mul_scalar rc={rc}")
        return x

    # ---- the sampling-based API under a seeded engine (ref_shim.cpp: ref_rng_*) ----
    def rng_seed(self, seed):
        self._fn("rng_seed", None, u64)(seed)

    def rng_sample(self, kind, logn, moduli):
        """kind: 0 ternary (NTT form), 1 uniform, 2 Gaussian (NTT form); continues the current engine state."""
        m = _arr(moduli)
        out = np.zeros((m.size, 1 << logn), dtype=np.uint64)
        rc = self._fn("rng_sample", C.c_int, C.c_int, C.c_uint, C.c_size_t, p64, p64)(kind, logn, m.size, _ptr(m), _ptr(out))
        if rc:
            raise ValueError(f"rng_sample rc={rc}")
        return out

    def rng_scenario_ckks(self, seed, logn, moduli_bits, additional_bits):
        bits = (C.c_uint * len(moduli_bits))(*moduli_bits)
        out = np.zeros(14, dtype=np.uint64)
        rc = self._fn("rng_scenario_ckks", C.c_int, u64, C.c_uint, C.c_size_t, C.POINTER(C.c_uint), C.c_uint, p64)(
            seed, logn, len(moduli_bits), bits, additional_bits, _ptr(out))
        if rc:
            raise ValueError(f"rng_scenario_ckks rc={rc}")
        return [int(v) for v in out]

    def rng_scenario_bgv(self, seed, logn, moduli_bits, additional_bits, t):
        bits = (C.c_uint * len(moduli_bits))(*moduli_bits)
        out = np.zeros(12, dtype=np.uint64)
        dec = np.zeros(1 << logn, dtype=np.uint64)
        rc = self._fn("rng_scenario_bgv", C.c_int, u64, C.c_uint, C.c_size_t, C.POINTER(C.c_uint), C.c_uint, u64, p64, p64)(
            seed, logn, len(moduli_bits), bits, additional_bits, t, _ptr(out), _ptr(dec))
        if rc:
            raise ValueError(f"rng_scenario_bgv rc={rc}")
        return [int(v) for v in out], dec

    def ckks_codec_scenario(self, logn, moduli_bits, additional_bits, log2_scaling, seed, count):
        bits = (C.c_uint * len(moduli_bits))(*moduli_bits)
        h = np.zeros(1, dtype=np.uint64)
        dec = np.zeros(1 << logn, dtype=np.float64)  # (re, im) per slot, n / 2 slots
        rc = self._fn("ckks_codec_scenario", C.c_int, C.c_uint, C.c_size_t, C.POINTER(C.c_uint), C.c_uint, C.c_double, u64, C.c_size_t,
                      p64, C.POINTER(C.c_double))(logn, len(moduli_bits), bits, additional_bits, float(log2_scaling), seed, count, _ptr(h),
                                                   dec.ctypes.data_as(C.POINTER(C.c_double)))
        if rc:
            raise ValueError(f"ckks_codec_scenario rc={rc}")
        return int(h[0]), dec

    def cache_ntt_factors(self, logn, moduli):
        m = _arr(moduli)
        self._fn("cache_ntt_factors", None, C.c_uint, p64, C.c_size_t)(logn, _ptr(m), m.size)
