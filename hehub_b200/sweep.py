"""Batched-ciphertext sweep across GPUs (BASELINE config 5).

Every ciphertext is an independent unit, so the batch is cut into contiguous per-rank ranges with
NO data-path collective.  `torch.distributed` (NCCL over NVLink on the GPU box; gloo in the CPU
test-suite, where the "device" is the CTA emulator build of the same kernels) is used only to

  * broadcast the key-switch key from rank 0 (generated or loaded once, 78 MiB at C5), and
  * gather one checksum per ciphertext so rank 0 can verify the whole sweep.

Inputs are generated on the device from (seed, global row index) by hehub_b200_lcg_fill, so the
words of ciphertext i do not depend on how the batch is sharded.  PyTorch is plumbing here: device
buffers, streams, the process group.  The arithmetic is the C ABI of include/hehub_b200.h.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .binding import Context, _mod

M64 = (1 << 64) - 1
CHECK_MULT = 0x9E3779B97F4A7C15  # odd: position weights w_j = (2j + 1) * CHECK_MULT mod 2^64


def shard_range(total: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous balanced partition: the first (total % world) ranks take one extra unit."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError("bad rank / world size")
    base, extra = divmod(total, world)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def _signed(v: int) -> int:
    v &= M64
    return v - (1 << 64) if v >> 63 else v


def ct_checksums_torch(words, cts: int):
    """Position-weighted wrapping sum per ciphertext, computed on the device that holds `words`
    (int64 view of the u64 words, [cts][words_per_ct])."""
    import torch
    w = words.view(cts, -1)
    j = torch.arange(w.shape[1], dtype=torch.int64, device=w.device)
    weights = (2 * j + 1) * _signed(CHECK_MULT)
    return (w * weights).sum(dim=1)


def ct_checksum_numpy(words: np.ndarray) -> int:
    """The same checksum for one ciphertext held as numpy uint64 (oracle side of the parity test)."""
    w = np.ascontiguousarray(words, dtype=np.uint64).ravel()
    with np.errstate(over="ignore"):
        j = np.arange(w.size, dtype=np.uint64)
        weights = (j * np.uint64(2) + np.uint64(1)) * np.uint64(CHECK_MULT)
        return int((w * weights).sum(dtype=np.uint64)) & M64


class CtSweep:
    """ckks::mult + relinearize over `total` synthetic ciphertext pairs, sharded over the ranks of
    `group_world` processes.  `device` is a torch.device: "cuda:<i>" with the CUDA library, "cpu"
    with the CTA-emulator library (tests only)."""

    def __init__(self, ctx: Context, device, logn: int, moduli, special: int, seed: int = 42):
        import torch
        self.torch = torch
        self.ctx, self.device, self.logn = ctx, torch.device(device), logn
        self.moduli = [int(m) for m in moduli]
        self.ext = self.moduli + [int(special)]
        self.L, self.n = len(self.moduli), 1 << logn
        self.seed = seed
        self._ext, self._extp = _mod(self.ext)
        self._mod, self._modp = _mod(self.moduli)
        self.ct_words = 2 * self.L * self.n
        self.key = None

    # row seeds: ct i of operand a (0/1) has rows [i*2L, (i+1)*2L) of stream a
    def _seed0(self, operand: int, first_ct: int) -> int:
        return (self.seed + operand * 0x5851F42D4C957F2D + first_ct * 2 * self.L) & M64

    def _buf(self, words: int):
        return self.torch.empty(words, dtype=self.torch.int64, device=self.device)

    def make_key(self, dist=None, rank: int = 0):
        """Rank 0 fills the key ([L][2][L+1][N], limb L = special modulus); everyone else receives it."""
        L, n = self.L, self.n
        self.key = self._buf(L * 2 * (L + 1) * n)
        if rank == 0 or dist is None:
            self.ctx._call("lcg_fill", n, self._extp, L + 1, self.key.data_ptr(), L * 2 * (L + 1), (self.seed + 1000) & M64, 1)
            self.ctx.synchronize()
        else:
            self.key.zero_()
        if dist is not None:
            if self.device.type == "cuda":
                self.torch.cuda.synchronize()
            dist.broadcast(self.key, src=0)
            if self.device.type == "cuda":
                self.torch.cuda.synchronize()
        return self.key

    def fill_inputs(self, ct1, ct2, first_ct: int, count: int):
        rows = count * 2 * self.L
        self.ctx._call("lcg_fill", self.n, self._modp, self.L, ct1.data_ptr(), rows, self._seed0(0, first_ct), 1)
        self.ctx._call("lcg_fill", self.n, self._modp, self.L, ct2.data_ptr(), rows, self._seed0(1, first_ct), 1)

    def run(self, total: int, wave: int, rank: int = 0, world: int = 1, dist=None, timer=None):
        """Process this rank's share of `total` ciphertext pairs in waves of at most `wave`.
        Returns {"first", "count", "checksums" (this rank, numpy uint64), "all_checksums" (rank 0
        when dist is given), "op_seconds" (device time of the mult+relin calls if timer given)}.
        `timer(fn)` must run fn() and return its device time in seconds."""
        torch = self.torch
        if self.key is None:
            self.make_key(dist, rank)
        first, count = shard_range(total, world, rank)
        wave = max(1, min(wave, max(count, 1)))
        ct1, ct2, res = (self._buf(wave * self.ct_words) for _ in range(3))
        sums = torch.empty(count, dtype=torch.int64, device=self.device)
        op_seconds = 0.0
        done = 0
        while done < count:
            nb = min(wave, count - done)
            self.fill_inputs(ct1, ct2, first + done, nb)

            def op():
                self.ctx._call("ckks_mult_relin", self.logn, self._extp, self.L, ct1.data_ptr(), ct2.data_ptr(),
                               self.key.data_ptr(), res.data_ptr(), nb)

            if timer is not None:
                op_seconds += timer(op)
            else:
                op()
            self.ctx.synchronize()
            sums[done:done + nb] = ct_checksums_torch(res[: nb * self.ct_words], nb)
            done += nb
        out = {"first": first, "count": count, "op_seconds": op_seconds,
               "checksums": sums.cpu().numpy().view(np.uint64).copy()}
        if dist is not None:
            # ragged gather: pad every rank's vector to the largest share
            cap = shard_range(total, world, 0)[1]
            padded = torch.zeros(cap, dtype=torch.int64, device=self.device)
            padded[:count] = sums
            gathered = [torch.empty_like(padded) for _ in range(world)]
            dist.all_gather(gathered, padded)
            if rank == 0:
                parts = [g.cpu().numpy().view(np.uint64)[: shard_range(total, world, r)[1]] for r, g in enumerate(gathered)]
                out["all_checksums"] = np.concatenate(parts) if parts else np.empty(0, np.uint64)
        return out

    def one_ct_inputs_host(self, idx: int):
        """Host copy of the two input ciphertexts of pair `idx` (for checking a sample against the oracle)."""
        ct1, ct2 = self._buf(self.ct_words), self._buf(self.ct_words)
        self.fill_inputs(ct1, ct2, idx, 1)
        self.ctx.synchronize()
        shape = (2, self.L, self.n)
        return (ct1.cpu().numpy().view(np.uint64).reshape(shape), ct2.cpu().numpy().view(np.uint64).reshape(shape))

    def key_host(self):
        L, n = self.L, self.n
        return self.key.cpu().numpy().view(np.uint64).reshape(L, 2, L + 1, n)
