"""Batched-ciphertext sweep across GPUs (BASELINE config 5) — a thin caller of the C++ driver.

The driver itself is hehub_b200/csrc/sweep.cu behind `hehub_b200_sweep_*` (include/hehub_b200.h): sharding, waves,
on-device input generation, the mult+relin calls, the checksum kernel and the gather all happen there, in C++ / CUDA.
Collectives (key broadcast from rank 0, all-gather of per-ciphertext checksums — there is no data-path collective)
go through a provider:

  * `nccl_collectives(...)`  — the product: NCCL over NVLink from hehub_b200/libhehub_b200_nccl.so (csrc/nccl_provider.cpp).
    The 128-byte NCCL id is the only thing that needs another channel; `exchange` is that channel (bench.py passes the
    process group torchrun already set up; a C++ job would use MPI_Bcast or a file).
  * `callback_collectives(...)` — the CPU test-suite: the same driver, compiled for the CTA emulator, with the two
    collectives supplied by the caller (world-size-2 gloo in tests/test_sweep.py).

No PyTorch is imported here.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .binding import HERE, Context, _mod

M64 = (1 << 64) - 1
CHECK_MULT = 0x9E3779B97F4A7C15  # odd: position weights w_j = (2j + 1) * CHECK_MULT mod 2^64
NCCL_SO = os.path.join(HERE, "libhehub_b200_nccl.so")

_BCAST = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p)
_GATHER = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p)


class Collectives(C.Structure):
    """struct hehub_b200_collectives (include/hehub_b200.h)"""
    _fields_ = [("self", C.c_void_p), ("broadcast", _BCAST), ("allgather", _GATHER)]


def shard_range(total: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous balanced partition: the first (total % world) ranks take one extra unit (hehub_b200_shard_range)."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError("bad rank / world size")
    base, extra = divmod(total, world)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def ct_checksum_numpy(words: np.ndarray) -> int:
    """The driver's checksum for one ciphertext held as numpy uint64 (oracle side of the parity tests)."""
    w = np.ascontiguousarray(words, dtype=np.uint64).ravel()
    with np.errstate(over="ignore"):
        j = np.arange(w.size, dtype=np.uint64)
        weights = (j * np.uint64(2) + np.uint64(1)) * np.uint64(CHECK_MULT)
        return int((w * weights).sum(dtype=np.uint64)) & M64


def nccl_collectives(rank: int, world: int, device: int, exchange):
    """NCCL provider.  `exchange(id_bytes_or_None) -> id_bytes`: called with the id on rank 0 and None elsewhere, must return
    rank 0's bytes on every rank.  Returns (struct, destroy)."""
    if not os.path.exists(NCCL_SO):
        raise FileNotFoundError(f"{NCCL_SO} not found: python -c 'import __graft_entry__ as g; g.build_nccl()'")
    lib = C.CDLL(NCCL_SO)
    buf = (C.c_uint8 * 128)()
    if rank == 0 and lib.hehub_b200_nccl_unique_id(buf) != 0:
        raise RuntimeError("ncclGetUniqueId failed")
    raw = exchange(bytes(buf) if rank == 0 else None)
    buf = (C.c_uint8 * 128).from_buffer_copy(raw)
    coll = Collectives()
    if lib.hehub_b200_nccl_create(C.byref(coll), rank, world, buf, device) != 0:
        raise RuntimeError("ncclCommInitRank failed")
    return coll, (lambda: lib.hehub_b200_nccl_destroy(C.byref(coll)))


def callback_collectives(broadcast, allgather):
    """Provider from two Python callables (tests): broadcast(address, nbytes, root) and allgather(send_address,
    recv_address, nbytes_per_rank) act on raw addresses of "device" memory (host memory under the CTA emulator)."""
    def _b(_self, buf, nbytes, root, _stream):
        try:
            broadcast(buf, nbytes, root)
            return 0
        except Exception:  # never let an exception cross the C boundary
            return 1

    def _g(_self, send, recv, nbytes, _stream):
        try:
            allgather(send, recv, nbytes)
            return 0
        except Exception:
            return 1
    coll = Collectives(None, _BCAST(_b), _GATHER(_g))
    return coll


class CtSweep:
    """ckks::mult + relinearize over `total` synthetic ciphertext pairs, sharded over `world` ranks (one process per GPU)."""

    def __init__(self, ctx: Context, logn: int, moduli, special: int, seed: int = 42, rank: int = 0, world: int = 1,
                 collectives: Collectives | None = None):
        self.ctx, self.logn = ctx, logn
        self.moduli = [int(m) for m in moduli]
        self.ext = self.moduli + [int(special)]
        self.L, self.n = len(self.moduli), 1 << logn
        self.rank, self.world = rank, world
        self.ct_words = 2 * self.L * self.n
        self._coll = collectives  # keeps the callbacks alive
        lib = ctx.lib
        lib.hehub_b200_sweep_create.argtypes = [C.POINTER(C.c_void_p), C.c_void_p, C.c_uint, C.POINTER(C.c_uint64), C.c_size_t, C.c_uint64,
                                                C.c_int, C.c_int, C.c_void_p]
        lib.hehub_b200_sweep_run.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p, C.POINTER(C.c_size_t),
                                             C.POINTER(C.c_size_t), C.POINTER(C.c_double)]
        lib.hehub_b200_sweep_fill_inputs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t]
        lib.hehub_b200_sweep_make_key.argtypes = [C.c_void_p]
        lib.hehub_b200_sweep_destroy.argtypes = [C.c_void_p]
        lib.hehub_b200_sweep_key.argtypes = [C.c_void_p]
        lib.hehub_b200_sweep_key.restype = C.c_void_p
        _, extp = _mod(self.ext)
        self.h = C.c_void_p()
        self._check(lib.hehub_b200_sweep_create(C.byref(self.h), ctx.h, logn, extp, self.L, seed & M64, rank, world,
                                                C.byref(collectives) if collectives is not None else None))

    def _check(self, rc):
        if rc:
            raise RuntimeError(f"hehub_b200 sweep error {rc}: {self.ctx.lib.hehub_b200_last_error(self.ctx.h).decode()}")

    def close(self):
        if self.h:
            self.ctx.lib.hehub_b200_sweep_destroy(self.h)
            self.h = None

    def make_key(self):
        """Rank 0 fills the key ([L][2][L+1][N], limb L = special modulus); everyone else receives it (broadcast)."""
        self._check(self.ctx.lib.hehub_b200_sweep_make_key(self.h))

    def run(self, total: int, wave: int, time_ops: bool = False):
        """Process this rank's share of `total` pairs in waves of at most `wave`.  Returns {"first", "count", "checksums"
        (this rank, numpy uint64), "all_checksums" (rank 0 only, when world > 1), "op_seconds" (device time of the mult+relin
        calls, CUDA events inside the driver, when time_ops)}."""
        first, count = shard_range(total, self.world, self.rank)
        mine = np.zeros(max(count, 1), dtype=np.uint64)
        everyone = np.zeros(max(total, 1), dtype=np.uint64) if (self.rank == 0 and self.world > 1) else None
        f, n, secs = C.c_size_t(), C.c_size_t(), C.c_double(0.0)
        self._check(self.ctx.lib.hehub_b200_sweep_run(self.h, total, wave, mine.ctypes.data, everyone.ctypes.data if everyone is not None else None,
                                                      C.byref(f), C.byref(n), C.byref(secs) if time_ops else None))
        assert (f.value, n.value) == (first, count)
        out = {"first": first, "count": count, "op_seconds": secs.value, "checksums": mine[:count].copy()}
        if everyone is not None:
            out["all_checksums"] = everyone[:total]
        return out

    def one_ct_inputs_host(self, idx: int):
        """Host copy of the two input ciphertexts of pair `idx` (for checking a sample against the oracle)."""
        a, b = self.ctx.slab(self.ct_words), self.ctx.slab(self.ct_words)
        try:
            self._check(self.ctx.lib.hehub_b200_sweep_fill_inputs(self.h, a.ptr, b.ptr, idx, 1))
            shape = (2, self.L, self.n)
            return a.download(shape), b.download(shape)
        finally:
            a.free()
            b.free()

    def key_host(self):
        L, n = self.L, self.n
        words = L * 2 * (L + 1) * n
        out = np.empty(words, dtype=np.uint64)
        self.ctx._call("slab_d2h", out.ctypes.data, self.ctx.lib.hehub_b200_sweep_key(self.h), words)
        self.ctx.synchronize()
        return out.reshape(L, 2, L + 1, n)
