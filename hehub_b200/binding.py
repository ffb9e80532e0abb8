"""ctypes binding of the C ABI in include/hehub_b200.h.

This is a thin harness layer (tests, bench, smoke): every method uploads numpy ``uint64``
arrays into device slabs, calls ONE entry point of libhehub_b200.so and downloads the result.
There is no CPU path here: if the CUDA library is missing or no device is present, construction
fails loudly.  Method names mirror the reference functions the entry points replace (and the
oracle binding used to check them).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_SO = os.environ.get("HEHUB_B200_LIB") or os.path.join(HERE, "libhehub_b200.so")  # the variable: A/B builds (tools/ab_build.sh)

u64 = C.c_uint64
p64 = C.POINTER(C.c_uint64)
sz = C.c_size_t
ctxp = C.c_void_p

OK, ERR_INVALID, ERR_CUDA, ERR_UNSUPPORTED, ERR_NOMEM = 0, 1, 2, 3, 4


class HehubB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"hehub_b200 error {code}: {msg}")
        self.code = code


class InvalidArgument(HehubB200Error, ValueError):
    """The reference throws std::invalid_argument for the same condition."""


class Unsupported(HehubB200Error):
    """The reference throws a bare `const char*` for the same condition."""


# name -> argtypes (after the leading ctx), all return int
_SIGS = {
    "tables_prepare": [C.c_uint, p64, sz],
    "slab_alloc": [sz, C.POINTER(C.c_void_p)],
    "slab_free": [C.c_void_p],
    "slab_h2d": [C.c_void_p, C.c_void_p, sz],
    "slab_d2h": [C.c_void_p, C.c_void_p, sz],
    "slab_d2d": [C.c_void_p, C.c_void_p, sz],
    "host_alloc": [sz, C.POINTER(C.c_void_p)],
    "host_free": [C.c_void_p],
    "ntt_fwd_lazy": [C.c_uint, p64, sz, C.c_void_p, sz],
    "intt_lazy": [C.c_uint, p64, sz, C.c_void_p, sz, C.c_int],
    "mulmod_hybrid_lazy": [sz, p64, sz, C.c_void_p, C.c_void_p, C.c_void_p, sz],
    "add_lazy": [sz, p64, sz, C.c_void_p, C.c_void_p, sz],
    "sub_lazy": [sz, p64, sz, C.c_void_p, C.c_void_p, sz],
    "mul_scalar_lazy": [sz, p64, sz, C.c_void_p, p64, sz],
    "reduce_strict": [sz, p64, sz, C.c_void_p, sz],
    "barrett_lazy": [sz, p64, sz, C.c_void_p, sz],
    "barrett": [sz, p64, sz, C.c_void_p, sz],
    "montgomery128_lazy": [u64, sz, C.c_void_p, C.c_void_p],
    "galois_cycle": [C.c_uint, sz, C.c_void_p, C.c_void_p, sz, sz],
    "galois_involution": [C.c_uint, sz, C.c_void_p, C.c_void_p, sz],
    "ckks_tensor": [C.c_uint, p64, sz, C.c_void_p, C.c_void_p, C.c_void_p, sz],
    "ext_prod_montgomery": [C.c_uint, p64, sz, C.c_void_p, C.c_void_p, C.c_void_p, sz],
    "ckks_rescale": [C.c_uint, p64, sz, C.c_void_p, C.c_void_p, sz],
    "bgv_mod_switch": [C.c_uint, p64, sz, u64, C.c_void_p, C.c_void_p, sz],
    "ckks_relinearize": [C.c_uint, p64, sz, C.c_void_p, C.c_void_p, C.c_void_p, sz],
    "bgv_relinearize": [C.c_uint, p64, sz, u64, C.c_void_p, C.c_void_p, C.c_void_p, sz],
    "ckks_mult_relin": [C.c_uint, p64, sz, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, sz],
    "bgv_mult_relin": [C.c_uint, p64, sz, u64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, sz],
    "ckks_rotate": [C.c_uint, p64, sz, C.c_void_p, C.c_void_p, sz, C.c_void_p, sz],
    "ckks_conjugate": [C.c_uint, p64, sz, C.c_void_p, C.c_void_p, C.c_void_p, sz],
    "rlwe_decrypt_core": [C.c_uint, p64, sz, C.c_void_p, C.c_void_p, C.c_void_p, sz],
    "rlwe_encrypt_core": [C.c_uint, p64, sz, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, sz],
    "rns_base_transform_from_single": [u64, p64, sz, C.c_void_p, C.c_void_p, sz, sz],
    "rns_base_transform_to_single": [p64, sz, u64, C.c_void_p, C.c_void_p, sz, sz],
    "ksk_generate": [C.c_uint, p64, sz, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p],
    "lcg_fill": [sz, p64, sz, C.c_void_p, sz, u64, u64],
    "ntt_host": [C.c_int, C.c_uint, p64, sz, C.c_void_p, C.c_void_p, sz, C.c_int],
    "ckks_mult_relin_host": [C.c_uint, p64, sz, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, sz],
    "ct_checksums": [C.c_void_p, sz, sz, C.c_void_p],
    "ctx_set_option": [C.c_char_p, C.c_int64],
    "ctx_set_stream": [C.c_void_p],
    "ctx_synchronize": [],
}

# entry points without a context (host-only parameter selection)
_FREE_SIGS = {
    "prime_row": [C.c_uint, sz, p64],
    "pick_moduli": [C.POINTER(C.c_uint), sz, C.c_uint, p64, p64],
}

EXPORTED = ["hehub_b200_version", "hehub_b200_ctx_create", "hehub_b200_ctx_destroy", "hehub_b200_last_error",
            "hehub_b200_launch_count"] + ["hehub_b200_" + k for k in _SIGS] + ["hehub_b200_" + k for k in _FREE_SIGS] + [
    "hehub_b200_shard_range", "hehub_b200_sweep_create", "hehub_b200_sweep_destroy", "hehub_b200_sweep_make_key", "hehub_b200_sweep_key",
    "hehub_b200_sweep_fill_inputs", "hehub_b200_sweep_run"]


def bind_host_to_gpu(device: int = 0) -> list[int]:
    """Restrict the calling process to the CPUs NVML reports as local to `device` (its NUMA node), so that the
    pinned host buffers allocated afterwards — and the threads that fill them — sit next to the GPU's PCIe root.
    Matters for the host-buffer entry points when several ranks share a two-socket host.  Returns the CPU list
    (empty: nothing changed — NVML missing, or no overlap with the CPUs this process may use)."""
    if os.environ.get("HEHUB_B200_NUMA", "1") == "0":
        return []
    try:
        import pynvml
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(device)
        ncpu = os.cpu_count() or 1
        masks = pynvml.nvmlDeviceGetCpuAffinity(handle, (ncpu + 63) // 64)
        local = {i for i in range(ncpu) if (int(masks[i // 64]) >> (i % 64)) & 1}
        cpus = sorted(local & set(os.sched_getaffinity(0)))
        if cpus:
            os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return []


def load_library(path: str | None = None) -> C.CDLL:
    path = path or DEFAULT_SO
    if not os.path.exists(path):
        raise FileNotFoundError(
            f"{path} not found: build the CUDA library first (python -c 'import __graft_entry__ as g; g.build()'). "
            "hehub_b200 has no CPU fallback.")
    lib = C.CDLL(path)
    lib.hehub_b200_version.restype = C.c_char_p
    lib.hehub_b200_ctx_create.restype = C.c_int
    lib.hehub_b200_ctx_create.argtypes = [C.POINTER(ctxp), C.c_int, C.c_void_p]
    lib.hehub_b200_ctx_destroy.argtypes = [ctxp]
    lib.hehub_b200_last_error.restype = C.c_char_p
    lib.hehub_b200_last_error.argtypes = [ctxp]
    lib.hehub_b200_launch_count.restype = C.c_uint64
    lib.hehub_b200_launch_count.argtypes = [ctxp]
    for name, args in _SIGS.items():
        f = getattr(lib, "hehub_b200_" + name, None)
        if f is None:  # an older A/B build (tools/ab_build.sh): calling the missing entry point raises AttributeError
            continue
        f.restype = C.c_int
        f.argtypes = [ctxp] + args
    for name, args in _FREE_SIGS.items():
        f = getattr(lib, "hehub_b200_" + name, None)
        if f is None:
            continue
        f.restype = C.c_int
        f.argtypes = args
    return lib


def prime_row(bits: int, lib: C.CDLL | None = None) -> list[int]:
    """Row `bits` of the reference's prime table (src/fhe/common/primelists.cpp), host only — no device needed."""
    lib = lib or load_library()
    buf = (C.c_uint64 * 32)()
    n = lib.hehub_b200_prime_row(bits, 32, buf)
    return [int(buf[i]) for i in range(min(n, 32))]


def pick_moduli(moduli_bits, additional_bits: int | None, lib: C.CDLL | None = None):
    """ckks::create_params' prime selection (src/fhe/ckks/basics.cpp:14-38): returns (moduli, additional modulus);
    additional_bits None -> hehub::create_params (rlwe.cpp:9-29), second value None.  Host only."""
    lib = lib or load_library()
    bits = (C.c_uint * len(moduli_bits))(*[int(b) for b in moduli_bits])
    out = (C.c_uint64 * max(1, len(moduli_bits)))()
    add = C.c_uint64(0)
    rc = lib.hehub_b200_pick_moduli(bits, len(moduli_bits), int(additional_bits or 0), out,
                                    C.byref(add) if additional_bits is not None else None)
    if rc == ERR_UNSUPPORTED:
        raise Unsupported(rc, "No suitable primes in the library.")
    if rc != OK:
        raise InvalidArgument(rc, "bad arguments to pick_moduli")
    return [int(out[i]) for i in range(len(moduli_bits))], (int(add.value) if additional_bits is not None else None)


def _arr(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.uint64)


def _mod(moduli):
    m = np.ascontiguousarray(np.asarray(moduli, dtype=np.uint64).ravel())
    return m, m.ctypes.data_as(p64)


class Slab:
    """A pooled device buffer of u64 words (device analogue of the reference's SmartArray)."""

    def __init__(self, ctx: "Context", n_words: int):
        self.ctx, self.n = ctx, int(n_words)
        out = C.c_void_p()
        if self.n:
            ctx._call("slab_alloc", self.n, C.byref(out))
        self.ptr = out.value

    def upload(self, host: np.ndarray):
        host = _arr(host)
        assert host.size == self.n, (host.size, self.n)
        if not self.n:
            return self
        self.ctx._call("slab_h2d", self.ptr, host.ctypes.data, self.n)
        self.ctx.synchronize()  # the numpy buffer may be pageable and short-lived
        return self

    def download(self, shape=None) -> np.ndarray:
        out = np.empty(self.n, dtype=np.uint64)
        if self.n:
            self.ctx._call("slab_d2h", out.ctypes.data, self.ptr, self.n)
            self.ctx.synchronize()
        return out.reshape(shape) if shape is not None else out

    def free(self):
        if self.ptr:
            self.ctx._call("slab_free", self.ptr)
            self.ptr = None


class Context:
    def __init__(self, device: int = 0, stream: int | None = None, lib_path: str | None = None):
        self.lib = load_library(lib_path)
        h = ctxp()
        rc = self.lib.hehub_b200_ctx_create(C.byref(h), device, C.c_void_p(stream))
        if rc != OK:
            raise HehubB200Error(rc, "cannot create a context: no CUDA device / driver (hehub_b200 has no CPU fallback)")
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            for p in getattr(self, "_pinned", []):
                self.lib.hehub_b200_host_free(self.h, p)
            self._pinned = []
            self.lib.hehub_b200_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _call(self, name, *args):
        rc = getattr(self.lib, "hehub_b200_" + name)(self.h, *args)
        if rc != OK:
            msg = self.lib.hehub_b200_last_error(self.h).decode()
            cls = {ERR_INVALID: InvalidArgument, ERR_UNSUPPORTED: Unsupported}.get(rc, HehubB200Error)
            raise cls(rc, msg)

    def version(self) -> str:
        return self.lib.hehub_b200_version().decode()

    def synchronize(self):
        self._call("ctx_synchronize")

    def set_option(self, name: str, value: int):
        self._call("ctx_set_option", name.encode(), int(value))

    def launch_count(self) -> int:
        return int(self.lib.hehub_b200_launch_count(self.h))

    def tables_prepare(self, logn, moduli):
        m, mp = _mod(moduli)
        self._call("tables_prepare", logn, mp, m.size)

    def slab(self, n_words) -> Slab:
        return Slab(self, n_words)

    def to_device(self, host) -> Slab:
        host = _arr(host)
        return Slab(self, host.size).upload(host)

    # ---- host-buffer entry points (pinned numpy views over hehub_b200_host_alloc memory) ----
    def pinned(self, shape) -> np.ndarray:
        n = int(np.prod(shape))
        out = C.c_void_p()
        self._call("host_alloc", max(n, 1), C.byref(out))
        buf = (C.c_uint64 * max(n, 1)).from_address(out.value)
        a = np.frombuffer(buf, dtype=np.uint64, count=n).reshape(shape)
        self._pinned = getattr(self, "_pinned", [])
        self._pinned.append(out.value)
        return a

    def ntt_host(self, forward, logn, moduli, x_in: np.ndarray, x_out: np.ndarray, strict=False):
        m, mp = _mod(moduli)
        batch = x_in.size // (m.size << logn)
        self._call("ntt_host", int(forward), logn, mp, m.size, x_in.ctypes.data, x_out.ctypes.data, batch, int(strict))

    def ckks_mult_relin_host(self, logn, ext_moduli, ct1: np.ndarray, ct2: np.ndarray, key: "Slab", out: np.ndarray):
        m, mp = _mod(ext_moduli)
        L = m.size - 1
        batch = ct1.size // ((2 * L) << logn)
        self._call("ckks_mult_relin_host", logn, mp, L, ct1.ctypes.data, ct2.ctypes.data, key.ptr, out.ctypes.data, batch)

    # ---- raw (device-pointer) entry points used by the bench -----------------------------
    def ntt_fwd_dev(self, logn, moduli, x: Slab, batch):
        m, mp = _mod(moduli)
        self._call("ntt_fwd_lazy", logn, mp, m.size, x.ptr, batch)

    def intt_dev(self, logn, moduli, x: Slab, batch, strict=False):
        m, mp = _mod(moduli)
        self._call("intt_lazy", logn, mp, m.size, x.ptr, batch, int(strict))

    # ---- host-array conveniences (one upload, one call, one download) --------------------
    def _unary(self, name, moduli, x, *extra):
        x = _arr(x)
        m, mp = _mod(moduli)
        n = x.shape[-1]
        batch = x.size // (n * m.size)
        assert batch * n * m.size == x.size, "operand is not [batch][L][n]"
        d = self.to_device(x)
        try:
            self._call(name, n, mp, m.size, d.ptr, *extra, batch)
            return d.download(x.shape)
        finally:
            d.free()

    def reduce_strict(self, moduli, x):
        return self._unary("reduce_strict", moduli, x)

    def barrett_lazy(self, moduli, x):
        return self._unary("barrett_lazy", moduli, x)

    def barrett(self, moduli, x):
        return self._unary("barrett", moduli, x)

    def mul_scalar_lazy(self, moduli, x, scalars):
        s, sp = _mod(scalars)
        return self._unary("mul_scalar_lazy", moduli, x, sp)

    def _binary_inplace(self, name, moduli, x, y):
        x, y = _arr(x), _arr(y)
        assert x.shape == y.shape
        m, mp = _mod(moduli)
        n = x.shape[-1]
        batch = x.size // (n * m.size)
        dx, dy = self.to_device(x), self.to_device(y)
        try:
            self._call(name, n, mp, m.size, dx.ptr, dy.ptr, batch)
            return dx.download(x.shape)
        finally:
            dx.free()
            dy.free()

    def add_lazy(self, moduli, x, y):
        return self._binary_inplace("add_lazy", moduli, x, y)

    def sub_lazy(self, moduli, x, y):
        return self._binary_inplace("sub_lazy", moduli, x, y)

    def mul_hybrid_lazy(self, moduli, a, b):
        a, b = _arr(a), _arr(b)
        assert a.shape == b.shape
        m, mp = _mod(moduli)
        n = a.shape[-1]
        batch = a.size // (n * m.size)
        da, db, dc = self.to_device(a), self.to_device(b), self.slab(a.size)
        try:
            self._call("mulmod_hybrid_lazy", n, mp, m.size, da.ptr, db.ptr, dc.ptr, batch)
            return dc.download(a.shape)
        finally:
            da.free()
            db.free()
            dc.free()

    def montgomery128_lazy(self, q, in_lohi):
        a = _arr(in_lohi)
        n = a.size // 2
        da, do = self.to_device(a), self.slab(max(n, 1))
        try:
            self._call("montgomery128_lazy", q, n, da.ptr, do.ptr)
            return do.download()[:n]
        finally:
            da.free()
            do.free()

    def _transform(self, name, logn, moduli, x, *extra):
        x = _arr(x)
        m, mp = _mod(moduli)
        n = 1 << logn
        assert x.shape[-1] == n
        batch = x.size // (n * m.size)
        assert batch * n * m.size == x.size, "operand is not [batch][L][N]"
        d = self.to_device(x)
        try:
            self._call(name, logn, mp, m.size, d.ptr, batch, *extra)
            return d.download(x.shape)
        finally:
            d.free()

    def poly_ntt_fwd(self, logn, moduli, x):
        return self._transform("ntt_fwd_lazy", logn, moduli, x)

    def poly_intt(self, logn, moduli, x, strict=False):
        return self._transform("intt_lazy", logn, moduli, x, int(strict))

    def ntt_fwd_lazy(self, logn, q, x):
        return self.poly_ntt_fwd(logn, [q], x)

    def intt_lazy(self, logn, q, x):
        return self.poly_intt(logn, [q], x)

    def _op(self, name, logn, moduli, ins, out_shape, *mid, batch):
        m, mp = _mod(moduli)
        slabs = [self.to_device(a) for a in ins]
        out = self.slab(int(np.prod(out_shape)))
        try:
            self._call(name, logn, mp, *mid, *[s.ptr for s in slabs], out.ptr, batch)
            return out.download(out_shape)
        finally:
            for s in slabs:
                s.free()
            out.free()

    @staticmethod
    def _batch_of(a, unit_shape):
        a = _arr(a)
        unit = int(np.prod(unit_shape))
        assert a.size % unit == 0
        return a, a.size // unit

    def ckks_tensor(self, logn, moduli, ct1, ct2):
        L, n = len(moduli), 1 << logn
        ct1, batch = self._batch_of(ct1, (2, L, n))
        lead = ct1.shape[:-3]
        return self._op("ckks_tensor", logn, moduli, [ct1, ct2], lead + (3, L, n), L, batch=batch)

    def ext_prod(self, logn, ext_moduli, poly, key):
        L, n = len(ext_moduli) - 1, 1 << logn
        poly, batch = self._batch_of(poly, (L, n))
        lead = poly.shape[:-2]
        return self._op("ext_prod_montgomery", logn, ext_moduli, [poly, key], lead + (2, L + 1, n), L, batch=batch)

    def ckks_rescale(self, logn, moduli, ct):
        L, n = len(moduli), 1 << logn
        ct, batch = self._batch_of(ct, (2, L, n))
        return self._op("ckks_rescale", logn, moduli, [ct], ct.shape[:-3] + (2, L - 1, n), L, batch=batch)

    def bgv_mod_switch(self, logn, moduli, t, ct):
        L, n = len(moduli), 1 << logn
        ct, batch = self._batch_of(ct, (2, L, n))
        return self._op("bgv_mod_switch", logn, moduli, [ct], ct.shape[:-3] + (2, L - 1, n), L, u64(t), batch=batch)

    def ckks_relinearize(self, logn, ext_moduli, quad, key):
        L, n = len(ext_moduli) - 1, 1 << logn
        quad, batch = self._batch_of(quad, (3, L, n))
        return self._op("ckks_relinearize", logn, ext_moduli, [quad, key], quad.shape[:-3] + (2, L, n), L, batch=batch)

    def bgv_relinearize(self, logn, ext_moduli, t, quad, key):
        L, n = len(ext_moduli) - 1, 1 << logn
        quad, batch = self._batch_of(quad, (3, L, n))
        return self._op("bgv_relinearize", logn, ext_moduli, [quad, key], quad.shape[:-3] + (2, L, n), L, u64(t), batch=batch)

    def ckks_mult_relin(self, logn, ext_moduli, ct1, ct2, key):
        L, n = len(ext_moduli) - 1, 1 << logn
        ct1, batch = self._batch_of(ct1, (2, L, n))
        return self._op("ckks_mult_relin", logn, ext_moduli, [ct1, ct2, key], ct1.shape[:-3] + (2, L, n), L, batch=batch)

    def bgv_mult_relin(self, logn, ext_moduli, t, ct1, ct2, key):
        L, n = len(ext_moduli) - 1, 1 << logn
        ct1, batch = self._batch_of(ct1, (2, L, n))
        return self._op("bgv_mult_relin", logn, ext_moduli, [ct1, ct2, key], ct1.shape[:-3] + (2, L, n), L, u64(t), batch=batch)

    def ckks_rotate(self, logn, ext_moduli, ct, key, step):
        L, n = len(ext_moduli) - 1, 1 << logn
        ct, batch = self._batch_of(ct, (2, L, n))
        m, mp = _mod(ext_moduli)
        dct, dkey, out = self.to_device(ct), self.to_device(key), self.slab(ct.size)
        try:
            self._call("ckks_rotate", logn, mp, L, dct.ptr, dkey.ptr, step, out.ptr, batch)
            return out.download(ct.shape)
        finally:
            dct.free()
            dkey.free()
            out.free()

    def ckks_conjugate(self, logn, ext_moduli, ct, key):
        L, n = len(ext_moduli) - 1, 1 << logn
        ct, batch = self._batch_of(ct, (2, L, n))
        return self._op("ckks_conjugate", logn, ext_moduli, [ct, key], ct.shape, L, batch=batch)

    def rlwe_decrypt_core(self, logn, moduli, ct, sk):
        L, n = len(moduli), 1 << logn
        ct, batch = self._batch_of(ct, (2, L, n))
        m, mp = _mod(moduli)
        dct, dsk, out = self.to_device(ct), self.to_device(sk), self.slab(batch * L * n)
        try:
            self._call("rlwe_decrypt_core", logn, mp, L, dct.ptr, dsk.ptr, out.ptr, batch)
            return out.download(ct.shape[:-3] + (L, n))
        finally:
            dct.free()
            dsk.free()
            out.free()

    def rlwe_encrypt_core(self, logn, moduli, pt, sk, c1, e):
        L, n = len(moduli), 1 << logn
        pt, batch = self._batch_of(pt, (L, n))
        m, mp = _mod(moduli)
        d = [self.to_device(a) for a in (pt, sk, c1, e)]
        out = self.slab(batch * 2 * L * n)
        try:
            self._call("rlwe_encrypt_core", logn, mp, L, d[0].ptr, d[1].ptr, d[2].ptr, d[3].ptr, out.ptr, batch)
            return out.download(pt.shape[:-2] + (2, L, n))
        finally:
            for x in d:
                x.free()
            out.free()

    def base_transform_from_single(self, q_old, x, new_moduli):
        x = _arr(x)
        n = x.shape[-1]
        batch = x.size // n
        m, mp = _mod(new_moduli)
        din, out = self.to_device(x), self.slab(batch * m.size * n)
        try:
            self._call("rns_base_transform_from_single", q_old, mp, m.size, din.ptr, out.ptr, n, batch)
            return out.download(x.shape[:-1] + (m.size, n))
        finally:
            din.free()
            out.free()

    def base_transform_to_single(self, old_moduli, x, new_modulus):
        x = _arr(x)
        m, mp = _mod(old_moduli)
        n = x.shape[-1]
        batch = x.size // (n * m.size)
        din, out = self.to_device(x), self.slab(batch * n)
        try:
            self._call("rns_base_transform_to_single", mp, m.size, new_modulus, din.ptr, out.ptr, n, batch)
            return out.download(x.shape[:-2] + (n,))
        finally:
            din.free()
            out.free()

    def ksk_generate(self, logn, ext_moduli, sk_curr, sk_orig, masks, errors):
        m, mp = _mod(ext_moduli)
        L, n = m.size - 1, 1 << logn
        d = [self.to_device(a) for a in (sk_curr, sk_orig, masks, errors)]
        out = self.slab(L * 2 * (L + 1) * n)
        try:
            self._call("ksk_generate", logn, mp, L, d[0].ptr, d[1].ptr, d[2].ptr, d[3].ptr, out.ptr)
            return out.download((L, 2, L + 1, n))
        finally:
            for x in d:
                x.free()
            out.free()

    def galois_cycle(self, logn, poly, step):
        poly = _arr(poly)
        n = 1 << logn
        rows = poly.size // n
        din, out = self.to_device(poly), self.slab(poly.size)
        try:
            self._call("galois_cycle", logn, rows, din.ptr, out.ptr, step, 1)
            return out.download(poly.shape)
        finally:
            din.free()
            out.free()

    def galois_involution(self, logn, poly):
        poly = _arr(poly)
        n = 1 << logn
        rows = poly.size // n
        din, out = self.to_device(poly), self.slab(poly.size)
        try:
            self._call("galois_involution", logn, rows, din.ptr, out.ptr, 1)
            return out.download(poly.shape)
        finally:
            din.free()
            out.free()

    def lcg_fill(self, moduli, n, rows, seed0, seed_stride=1):
        m, mp = _mod(moduli)
        out = self.slab(rows * n)
        try:
            self._call("lcg_fill", n, mp, m.size, out.ptr, rows, seed0, seed_stride)
            return out.download((rows, n))
        finally:
            out.free()
