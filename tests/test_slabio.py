"""Raw-slab file format: Python round trips and rejection of damaged files."""
import numpy as np
import pytest

from hehub_b200 import slabio


def test_round_trip_and_header_checks(tmp_path):
    rng = np.random.default_rng(0)
    mods = [65537, 1099507695617]
    ct = np.stack([np.stack([rng.integers(0, q, 32, dtype=np.uint64) for q in mods]) for _ in range(2)])
    path = tmp_path / "ct.slab"
    slabio.save(str(path), ct, mods, slabio.KIND_CT, True)
    words, moduli, kind, value_form = slabio.load(str(path))
    assert np.array_equal(words, ct) and moduli == mods and kind == slabio.KIND_CT and value_form
    blob = path.read_bytes()
    assert len(blob) == 48 + 8 * 2 + 8 * ct.size
    for damaged in (blob[:-1], blob + b"\0", b"XXXXXXXX" + blob[8:], blob[:10]):
        with pytest.raises(ValueError):
            slabio.loads(damaged)
    with pytest.raises(ValueError):
        slabio.dumps(ct, mods + [3], slabio.KIND_CT, True)
