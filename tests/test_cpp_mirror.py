"""The C++ host mirror of the hehub:: API (hehub_b200/cpp/hehub) — compiled and run as the reference's
own unit tests would be, against the CUDA library on a GPU box and against the CTA-emulator build of
the same sources in the CPU suite."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "test_hehub_api.cpp")
OUT = os.path.join(ROOT, "tests", "cpp", "_build")


def _build(kind: str) -> str:
    import __graft_entry__ as ge
    from oracle.binding import build_oracle
    oracle_so = build_oracle()
    lib = ge.build_sim() if kind == "sim" else ge.CUDA_SO
    if not os.path.exists(lib):
        raise FileNotFoundError(f"{lib} missing: run __graft_entry__.build() first (there is no CPU fallback)")
    os.makedirs(OUT, exist_ok=True)
    exe = os.path.join(OUT, f"test_hehub_api_{kind}")
    deps = [SRC, lib, oracle_so] + [os.path.join(ROOT, "hehub_b200", "cpp", "hehub", f)
                                    for f in os.listdir(os.path.join(ROOT, "hehub_b200", "cpp", "hehub"))]
    if not os.path.exists(exe) or any(os.path.getmtime(d) > os.path.getmtime(exe) for d in deps):
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Wno-ignored-qualifiers", "-I" + os.path.join(ROOT, "hehub_b200", "cpp"),
                               "-I" + os.path.join(ROOT, "oracle"), SRC, lib, oracle_so,
                               "-Wl,-rpath," + os.path.dirname(lib), "-Wl,-rpath," + os.path.dirname(oracle_so),
                               "-pthread", "-o", exe])
    return exe


@pytest.mark.parametrize("kind", ["sim", pytest.param("gpu", marks=pytest.mark.gpu)])
def test_cpp_mirror(kind, tmp_path_factory):
    exe = _build(kind)
    out_dir = tmp_path_factory.mktemp(f"slabs_{kind}")
    # second argument: terms of the Basel-series example (examples/ckks_example.cpp runs 10^4): a few on the emulator, which
    # starts a host thread per CUDA thread, more on the GPU
    res = subprocess.run([exe, str(out_dir), "3" if kind == "sim" else "200"], capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "0 failures" in res.stdout
    # the sampling-based API (keygen, encrypt, plain ops, BGV encode/decrypt) replayed under seeded engines: every stage's hash
    # must equal what the unmodified reference produced for the same statements (oracle/make_golden.py, kat["rng"])
    import json
    kat = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_kat.json")))["rng"]
    got = {}
    codec = {}
    for line in open(out_dir / "rng_hashes.txt"):
        tag, seed, *hashes = line.split()
        if tag == "codec":
            codec[int(seed)] = hashes
        else:
            got[(tag, int(seed))] = hashes
    for case in kat["samples"]:
        assert got[("samples", case["seed"])] == [case["ternary"], case["uniform"], case["gaussian"]], case
    for case in kat["ckks"]:
        assert got[("ckks", case["seed"])] == case["hashes"], case
    for case in kat["bgv"]:
        assert got[("bgv", case["seed"])] == case["hashes"] + [case["decoded"]], case
    assert len(got) == len(kat["samples"]) + len(kat["ckks"]) + len(kat["bgv"])
    # CKKS encoder / decoder: plaintext words bit-exact with the reference; decoded doubles within 1e-9 relative of its output
    # (the decoder ends in floating point: tolerance instead of identity, as north_star allows for floating-point results)
    all_kat = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_kat.json")))
    for case in all_kat["ckks_codec"]:
        plaintext, abs_sum, *head = codec[case["seed"]]
        assert plaintext == case["plaintext"], case
        assert abs(float(abs_sum) - case["decoded_abs_sum"]) <= 1e-9 * case["decoded_abs_sum"], case
        for a, b in zip(head, case["decoded_head"]):
            assert abs(float(a) - b) <= 1e-9 * max(1.0, abs(b)), (case["seed"], a, b)
    assert len(codec) == len(all_kat["ckks_codec"])
    # the slab files the C++ mirror wrote are read back by the Python side (same layout, hehub_b200/slabio.py)
    import numpy as np
    from hehub_b200 import slabio
    from oracle.binding import Oracle
    orc = Oracle()
    mods = [65537, 260898817, 576460752272228353]
    words, moduli, kind_tag, value_form = slabio.load(str(out_dir / "poly.slab"))
    assert moduli == mods and kind_tag == slabio.KIND_POLY and not value_form
    assert np.array_equal(words[0], np.stack([orc.lcg_fill(6000 + k, q, 64) for k, q in enumerate(mods)]))
    words, moduli, kind_tag, value_form = slabio.load(str(out_dir / "ksk.slab"))
    assert words.shape == (4, 3, 64) and kind_tag == slabio.KIND_KSK and value_form
    assert slabio.dumps(words, moduli, kind_tag, value_form) == open(out_dir / "ksk.slab", "rb").read()
