#!/usr/bin/env python3
"""Generate tests/golden/reference_kat.json by running the UNMODIFIED reference.

TEST INFRASTRUCTURE ONLY.  Run in the dev container (needs /root/reference):

    python oracle/make_golden.py

It builds oracle/_ref/libhehub_ref.so from the reference's own sources (oracle/Makefile),
drives it through oracle/ref_shim.cpp on deterministic inputs, and records

  * FNV-1a hashes of raw u64 outputs at the BASELINE sizes (SURVEY Appendix B layout), and
  * complete raw vectors at small sizes (N = 8 / 16),

so that the oracle and the CUDA path can be pinned on machines where the reference is
absent (the GPU box).  Inputs are the LCG fill of SURVEY Appendix B; the generators of
tests/mod_arith_t.cpp are restated for the word-level KATs.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.binding import Oracle, Reference  # noqa: E402

M64 = (1 << 64) - 1
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                   "reference_kat.json")

NTT_MODULI = [65537, 260898817, 35184358850561, 36028796997599233, 576460752272228353]


def lcg(seed, q, n):
    s, out = seed, []
    for _ in range(n):
        s = (s * 6364136223846793005 + 1442695040888963407) & M64
        out.append(s % q)
    return np.array(out, dtype=np.uint64)


def fnv(words, h=1469598103934665603):
    for w in np.asarray(words, dtype=np.uint64).ravel().tolist():
        h = ((h ^ w) * 1099511628211) & M64
    return h


def hx(words):
    return f"{fnv(words):016x}"


def ints(a):
    return [int(v) for v in np.asarray(a, dtype=np.uint64).ravel()]


def fill_ct(seed0, moduli, n):
    return np.stack([np.stack([lcg(seed0 + 10 * h + k, moduli[k], n) for k in range(len(moduli))])
                     for h in range(2)])


def fill_key(seed0, ext, n):
    L = len(ext) - 1
    return np.stack([np.stack([np.stack([lcg(seed0 + 100 * r + 10 * h + k, ext[k], n)
                                         for k in range(L + 1)]) for h in range(2)]) for r in range(L)])


def main():
    ref = Reference()
    kat = {"_generator": "oracle/make_golden.py", "_reference": "primihub/hehub @ 8d1d4bd (unmodified)"}

    # ---- NTT / INTT (ntt.cpp:145-223) -------------------------------------
    ntt = []
    for q in NTT_MODULI:
        for logn in [1, 2, 3, 4, 5, 7, 9, 10, 11, 12, 13, 14, 15]:
            if (q - 1) % (2 << logn):
                continue
            x = lcg(42, q, 1 << logn)
            y = ref.ntt_fwd_lazy(logn, q, x)
            z = ref.intt_lazy(logn, q, y)
            w = ref.intt_lazy(logn, q, x)  # INTT applied to arbitrary (not NTT-image) data
            ntt.append({"q": q, "logn": logn, "in": hx(x), "ntt": hx(y), "intt_ntt": hx(z), "intt": hx(w)})
    kat["ntt_hashes"] = ntt
    x8 = np.array([1000 * i + 1 for i in range(8)], dtype=np.uint64)
    y8 = ref.ntt_fwd_lazy(3, 65537, x8)
    kat["ntt_n8"] = {"q": 65537, "in": ints(x8), "ntt": ints(y8), "intt": ints(ref.intt_lazy(3, 65537, y8))}
    x16 = lcg(7, 260898817, 16)
    y16 = ref.ntt_fwd_lazy(4, 260898817, x16)
    kat["ntt_n16"] = {"q": 260898817, "in": ints(x16), "ntt": ints(y16),
                      "intt": ints(ref.intt_lazy(4, 260898817, y16))}

    # ---- word-level KATs with the generators of tests/mod_arith_t.cpp ------
    seed, vec = 42, []
    for _ in range(1000):  # mod_arith_t.cpp:12-17
        seed = (((seed ^ 893758435427369) * 65536) + 945738773644543) & M64
        vec.append(seed)
    vec = np.array(vec, dtype=np.uint64)
    kat["barrett_lazy"] = [{"q": q, "in": hx(vec), "out": hx(ref.barrett_lazy(q, vec)),
                            "strict": hx(ref.barrett(q, vec))}
                           for q in [65537, 33333333, 777777777777777, 1234567890111111111]]
    q = 1234567890111111111
    seed, f, g = 42, [], []
    for _ in range(1000):  # mod_arith_t.cpp:40-45
        seed = ((seed * 65968279837582827) & M64) ^ 3948528936546489545
        f.append(seed % q)
        seed = ((seed * 43534547657678213) & M64) ^ 7955436776934235466
        g.append(seed % q)
    f, g = np.array(f, dtype=np.uint64), np.array(g, dtype=np.uint64)
    hyb, bar = ref.mul_hybrid_lazy(q, f, g), ref.mul_barrett_lazy(q, f, g)
    kat["mulmod"] = {"q": q, "f": hx(f), "g": hx(g), "hybrid_lazy": hx(hyb), "barrett_lazy": hx(bar),
                     "hybrid_head": ints(hyb[:4]), "barrett_head": ints(bar[:4])}
    q = 38589379749438777
    seed, lohi = 42, []
    for _ in range(8):  # mod_arith_t.cpp:65-70 (u128 state)
        seed = (seed * 3405898573857435 + 4385837453598385) & ((1 << 128) - 1)
        v = seed % (q << 64)
        lohi += [v & M64, v >> 64]
    kat["montgomery128"] = {"q": q, "in_lohi": lohi,
                            "out": ints(ref.montgomery128_lazy(q, np.array(lohi, dtype=np.uint64)))}

    # ---- C3 chain: ckks::mult + relinearize, N=8192, L=4 -------------------
    mods, P = ref.ckks_pick_moduli([40, 30, 30, 30], 40)
    ext, logn, n = mods + [P], 13, 8192
    ct1, ct2, key = fill_ct(100, ext[:4], n), fill_ct(200, ext[:4], n), fill_key(1000, ext, n)
    quad = ref.ckks_tensor(logn, mods, ct1, ct2)
    e = ref.ext_prod(logn, ext, quad[2], key)
    rl = ref.ckks_relinearize(logn, ext, quad, key)
    mm = ref.ckks_mult_relin(logn, ext, ct1, ct2, key)
    rs = ref.ckks_rescale(logn, mods, rl)
    rot = ref.ckks_rotate(logn, ext, ct1, key, 5)
    cj = ref.ckks_conjugate(logn, ext, ct1, key)
    bre = ref.bgv_relinearize(logn, ext, 1, quad, key)
    kat["c3"] = {"logn": logn, "moduli": mods, "P": P,
                 "tensor": [hx(quad[j]) for j in range(3)], "ext_prod": [hx(e[h]) for h in range(2)],
                 "relinearize": [hx(rl[h]) for h in range(2)], "mult": [hx(mm[h]) for h in range(2)],
                 "rescale": [hx(rs[h]) for h in range(2)], "rotate5": [hx(rot[h]) for h in range(2)],
                 "conjugate": [hx(cj[h]) for h in range(2)],
                 "bgv_relinearize_t1": [hx(bre[h]) for h in range(2)]}

    # ---- C4: rescale / mod-switch N=16384, L=8->7 --------------------------
    mods, P = ref.ckks_pick_moduli([50] + [40] * 7, 50)
    logn, n = 14, 16384
    ct = fill_ct(300, mods, n)
    rs = ref.ckks_rescale(logn, mods, ct)
    ms = ref.bgv_mod_switch(logn, mods, 65537, ct)
    kat["c4"] = {"logn": logn, "moduli": mods, "P": P, "rescale": [hx(rs[h]) for h in range(2)],
                 "mod_switch_t65537": [hx(ms[h]) for h in range(2)]}

    # ---- C5 shape, one ciphertext: N=32768, L=12 ---------------------------
    mods, P = ref.ckks_pick_moduli([50] * 12, 55)
    ext, logn, n = mods + [P], 15, 32768
    ct1, ct2, key = fill_ct(100, mods, n), fill_ct(200, mods, n), fill_key(1000, ext, n)
    mm = ref.ckks_mult_relin(logn, ext, ct1, ct2, key)
    kat["c5"] = {"logn": logn, "moduli": mods, "P": P, "mult": [hx(mm[h]) for h in range(2)]}

    # ---- RLWE cores (rlwe.cpp:34-71), C3 moduli, LCG-filled operands -------------------------
    mods, P = ref.ckks_pick_moduli([40, 30, 30, 30], 40)
    logn, n = 13, 8192
    L = len(mods)
    sk = np.stack([lcg(4000 + k, mods[k], n) for k in range(L)])
    c1 = np.stack([lcg(4100 + k, mods[k], n) for k in range(L)])
    pt = np.stack([lcg(4200 + k, mods[k], n) for k in range(L)])
    small = (lcg(4300, 39, n).astype(np.int64) - 19)       # error coefficients in [-19, 19]
    err = np.stack([np.where(small < 0, mods[k] + small, small).astype(np.uint64) for k in range(L)])
    ct = fill_ct(4400, mods, n)
    dec = ref.rlwe_decrypt_core(logn, mods, ct, sk)
    enc = ref.rlwe_encrypt_core(logn, mods, pt, sk, c1, err)
    kat["rlwe"] = {"logn": logn, "moduli": mods, "decrypt": hx(dec), "encrypt": [hx(enc[h]) for h in range(2)]}

    # ---- base transform + key generation (rns_transform.cpp:11-126, keys.cpp:8-36), C3 chain ----
    mods, P = ref.ckks_pick_moduli([40, 30, 30, 30], 40)
    ext, logn, n = mods + [P], 13, 8192
    L = len(mods)

    def ternary(seed):
        t = lcg(seed, 3, n).astype(np.int64) - 1
        return np.stack([np.where(t < 0, q + t, t).astype(np.uint64) for q in mods])

    sk_o_coeff, sk_c_coeff = ternary(5000), ternary(5001)
    sk_o, sk_c = ref.poly_ntt_fwd(logn, mods, sk_o_coeff), ref.poly_ntt_fwd(logn, mods, sk_c_coeff)
    masks = np.stack([np.stack([lcg(5200 + 10 * p + k, ext[k], n) for k in range(L + 1)]) for p in range(L)])
    errs = []
    for p in range(L):
        sm = lcg(5300 + p, 39, n).astype(np.int64) - 19
        errs.append(np.stack([np.where(sm < 0, q + sm, sm).astype(np.uint64) for q in ext]))
    errs = np.stack(errs)
    ksk = ref.ksk_generate(logn, ext, sk_c, sk_o, masks, errs)
    single = ref.base_transform_to_single(mods, sk_o_coeff, P)
    lazy = lcg(5400, 2 * mods[0], n)
    fan = ref.base_transform_from_single(mods[0], lazy, ext[1:])
    kat["keygen"] = {"logn": logn, "moduli": mods, "P": P, "ksk_rows": [hx(ksk[p]) for p in range(L)],
                     "to_single": hx(single), "from_single": hx(fan)}

    # ---- small raw fixtures: N=16, L=3 ({34,34,34} + P 34) ------------------
    mods, P = ref.ckks_pick_moduli([34, 34, 34], 34)
    ext, logn, n = mods + [P], 4, 16
    ct1, ct2, key = fill_ct(100, mods, n), fill_ct(200, mods, n), fill_key(1000, ext, n)
    quad = ref.ckks_tensor(logn, mods, ct1, ct2)
    small = {"logn": logn, "moduli": mods, "P": P, "ct1": ints(ct1), "ct2": ints(ct2), "key": ints(key),
             "tensor": ints(quad), "ext_prod": ints(ref.ext_prod(logn, ext, quad[2], key)),
             "relinearize": ints(ref.ckks_relinearize(logn, ext, quad, key)),
             "mult": ints(ref.ckks_mult_relin(logn, ext, ct1, ct2, key)),
             "rescale": ints(ref.ckks_rescale(logn, mods, ct1)),
             "mod_switch_t65537": ints(ref.bgv_mod_switch(logn, mods, 65537, ct1)),
             "mod_switch_t2": ints(ref.bgv_mod_switch(logn, mods, 2, ct1)),
             "bgv_relinearize_t1": ints(ref.bgv_relinearize(logn, ext, 1, quad, key)),
             "cycle": {str(s): ints(ref.galois_cycle(logn, ct1[0], s)) for s in [0, 1, 2, 3, 7]},
             "involution": ints(ref.galois_involution(logn, ct1[0])),
             "rotate": {str(s): ints(ref.ckks_rotate(logn, ext, ct1, key, s)) for s in [1, 2, 3]},
             "conjugate": ints(ref.ckks_conjugate(logn, ext, ct1, key)),
             "poly_add": ints(ref.poly_add(logn, mods, ct1[0], ct2[1])),
             "poly_sub": ints(ref.poly_sub(logn, mods, ct1[0], ct2[1])),
             "poly_mul_scalar_12345": ints(ref.poly_mul_scalar(logn, mods, ct1[0], 12345)),
             "poly_intt_strict": ints(ref.poly_intt(logn, mods, ct1[0], True)),
             "poly_ntt": ints(ref.poly_ntt_fwd(logn, mods, ct1[0]))}
    kat["small"] = small

    # ---- the sampling-based API under a seeded engine (sampling.cpp, rlwe.cpp:31-61, keys.cpp, ckks.h:180-197,
    #      bgv/basics.cpp) — hashes of every stage; the mirror's test replays the same statements ----
    rng = {"samples": [], "ckks": [], "bgv": []}
    for seed, logn, moduli in [(1, 6, [65537, 260898817]), (2, 10, [1073479681, 576460752272228353]),
                               (3, 12, [36028796997599233, 576460752272228353, 1099510054913])]:
        ref.rng_seed(seed)
        rng["samples"].append({"seed": seed, "logn": logn, "moduli": moduli,
                               "ternary": hx(ref.rng_sample(0, logn, moduli)), "uniform": hx(ref.rng_sample(1, logn, moduli)),
                               "gaussian": hx(ref.rng_sample(2, logn, moduli))})
    for seed, logn, bits, add in [(5, 10, [40, 30, 30], 40), (6, 12, [39, 30], 39), (7, 13, [40, 30, 30, 30], 40), (8, 11, [55, 50], 55)]:
        rng["ckks"].append({"seed": seed, "logn": logn, "bits": bits, "additional_bits": add,
                            "hashes": [f"{h:016x}" for h in ref.rng_scenario_ckks(seed, logn, bits, add)]})
    # moduli below 2^53: above it the reference's Gaussian conversion (double arithmetic, sampling.cpp:86) yields errors that
    # differ per limb, and bgv::decrypt then needs the big-integer CRT branch, which the back end does not build
    for seed, logn, bits, add, t in [(9, 10, [50, 45, 45], 50, 65537), (10, 12, [52, 50, 48], 52, 65537), (11, 8, [50, 45], 50, 12289)]:
        hs, dec = ref.rng_scenario_bgv(seed, logn, bits, add, t)
        rng["bgv"].append({"seed": seed, "logn": logn, "bits": bits, "additional_bits": add, "t": t,
                           "hashes": [f"{h:016x}" for h in hs], "decoded": hx(dec)})
    kat["rng"] = rng

    # ---- CKKS encoder / decoder (host-side floating point, ckks/basics.cpp:156-369): plaintext hash (bit-exact target) and the
    #      first decoded slots (tolerance target); small and wide-integer branches of both directions ----
    codec = []
    for logn, bits, add, log2s, seed, count in [(10, [40, 30, 30], 40, 30, 3, 512), (12, [39, 30], 39, 30, 4, 2048), (8, [50, 50], 55, 70, 5, 100),
                                                (11, [40, 30], 40, 50, 6, 1024), (6, [30], 40, 20, 7, 32)]:
        # in a separate executable: the reference's stringstream-based big integers crash inside this interpreter (ref_tool.cpp)
        import subprocess
        tool = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "ref_tool")
        line = subprocess.check_output([tool, "codec", str(logn), str(add), str(log2s), str(seed), str(count)] + [str(b) for b in bits], text=True)
        entry = {"logn": logn, "bits": bits, "additional_bits": add, "log2_scaling": log2s, "seed": seed, "count": count}
        entry.update(json.loads(line))
        codec.append(entry)
    kat["ckks_codec"] = codec

    # ---- the prime table (primelists.cpp) ----------------------------------
    kat["prime_rows"] = {str(b): ref.prime_row(b, 32) for b in range(0, 60) if ref.prime_row(b, 32)}
    kat["inverse_mod_prime"] = [[a, p, ref.inverse_mod_prime(a, p)] for a, p in
                                [(65537, 1099510054913), (1, 65537), (2, 1073479681),
                                 (1099502714881, 1125899903827969), (1125899903827969, 1099510054913)]]

    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(OUT, "w") as fh:
        json.dump(kat, fh, indent=1)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")

    # sanity: the oracle agrees with what was just recorded
    orc = Oracle()
    assert ints(orc.ckks_mult_relin(4, ext, ct1, ct2, key)) == small["mult"]
    print("oracle agrees on the small mult fixture")


if __name__ == "__main__":
    main()
