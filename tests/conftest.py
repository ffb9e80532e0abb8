"""pytest configuration: markers, build-on-demand fixtures, shared input generators."""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

M64 = (1 << 64) - 1


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def lcg(seed, q, n):
    """SURVEY Appendix B input generator (pure python; small n only)."""
    s, out = seed, []
    for _ in range(n):
        s = (s * 6364136223846793005 + 1442695040888963407) & M64
        out.append(s % q)
    return np.array(out, dtype=np.uint64)


def fnv(words, h=1469598103934665603):
    for w in np.asarray(words, dtype=np.uint64).ravel().tolist():
        h = ((h ^ w) * 1099511628211) & M64
    return h


@pytest.fixture(scope="session")
def kat():
    with open(os.path.join(ROOT, "tests", "golden", "reference_kat.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def oracle():
    from oracle.binding import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def reference():
    """The unmodified reference, when oracle/_ref can be built or was shipped prebuilt."""
    from oracle.binding import Reference
    if not Reference.available():
        pytest.skip("reference library not available on this machine")
    return Reference()


def fill_ct(orc, seed0, moduli, n):
    return np.stack([np.stack([orc.lcg_fill(seed0 + 10 * h + k, moduli[k], n) for k in range(len(moduli))])
                     for h in range(2)])


def fill_key(orc, seed0, ext, n):
    L = len(ext) - 1
    return np.stack([np.stack([np.stack([orc.lcg_fill(seed0 + 100 * r + 10 * h + k, ext[k], n)
                                         for k in range(L + 1)]) for h in range(2)]) for r in range(L)])
