"""Raw-slab (de)serialisation of RnsPolynomial / ciphertext / key-switch-key words.

The reference has no wire or disk format; this is the trivial `[header][moduli][limb slabs]` dump that
lets parity fixtures travel between machines (dev container <-> GPU box) and between the C++ mirror
(`hehub_b200/cpp/hehub/serialize.h`, same layout) and the Python harness.  Little-endian throughout.

    offset  size  field
    0       8     magic  b"HEHB200\\0"
    8       4     version (1)
    12      4     kind: 1 polynomial [L][N], 2 ciphertext [polys][L][N], 3 key-switch key [rows][2][L][N]
    16      8     N
    24      8     L (limbs per polynomial)
    32      8     polys (1 / 2 or 3 / rows * 2)
    40      4     rep_form: 0 coefficients, 1 NTT values
    44      4     reserved (0)
    48      8*L   moduli
    ...     8*polys*L*N  words, [poly][limb][N]
"""
from __future__ import annotations

import struct

import numpy as np

MAGIC = b"HEHB200\0"
KIND_POLY, KIND_CT, KIND_KSK = 1, 2, 3
_HDR = struct.Struct("<8sIIQQQII")


def dumps(words: np.ndarray, moduli, kind: int, value_form: bool) -> bytes:
    w = np.ascontiguousarray(words, dtype="<u8")
    m = np.ascontiguousarray(np.asarray(moduli, dtype="<u8").ravel())
    if w.ndim < 2 or w.shape[-2] != m.size:
        raise ValueError("words must be [..., L, N] with one modulus per limb")
    n, L = w.shape[-1], m.size
    polys = w.size // (n * L)
    return _HDR.pack(MAGIC, 1, kind, n, L, polys, 1 if value_form else 0, 0) + m.tobytes() + w.tobytes()


def loads(blob: bytes):
    """-> (words [polys][L][N] uint64, moduli list, kind, value_form)"""
    if len(blob) < _HDR.size:
        raise ValueError("truncated header")
    magic, version, kind, n, L, polys, form, _ = _HDR.unpack_from(blob)
    if magic != MAGIC or version != 1:
        raise ValueError("not a hehub_b200 slab file")
    if kind not in (KIND_POLY, KIND_CT, KIND_KSK) or form not in (0, 1):
        raise ValueError("bad kind / representation tag")
    need = _HDR.size + 8 * L + 8 * polys * L * n
    if len(blob) != need:
        raise ValueError(f"size mismatch: {len(blob)} bytes, header says {need}")
    moduli = np.frombuffer(blob, dtype="<u8", count=L, offset=_HDR.size)
    words = np.frombuffer(blob, dtype="<u8", count=polys * L * n, offset=_HDR.size + 8 * L).reshape(polys, L, n)
    return words.astype(np.uint64), [int(q) for q in moduli], kind, bool(form)


def save(path: str, words, moduli, kind: int, value_form: bool) -> None:
    with open(path, "wb") as fh:
        fh.write(dumps(words, moduli, kind, value_form))


def load(path: str):
    with open(path, "rb") as fh:
        return loads(fh.read())
