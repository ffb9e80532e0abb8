// test_hehub_api.cpp — the hehub:: host mirror (hehub_b200/cpp/hehub) exercised the way the
// reference's own unit tests exercise the reference (tests/ntt_t.cpp, tests/mod_arith_t.cpp,
// tests/common_t.cpp, tests/ckks_t.cpp:136-175), plus word-for-word comparison of every
// ciphertext op against the CPU oracle (oracle/hehub_oracle.h; test infrastructure).
//
// Linked against hehub_b200/libhehub_b200.so on a GPU box, or against the CTA-emulator build of
// the same sources for the CPU suite (tests/test_cpp_mirror.py drives both).
#include <cstdio>
#include <cstdlib>
#include <random>
#include <string>
#include <vector>

#include "hehub/hehub.h"
#include "hehub_oracle.h"

using namespace hehub;

static int g_fail = 0, g_checks = 0;
#define CHECK(cond)                                                                  \
    do {                                                                             \
        g_checks++;                                                                  \
        if (!(cond)) {                                                               \
            g_fail++;                                                                \
            std::fprintf(stderr, "FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond);     \
        }                                                                            \
    } while (0)
#define CHECK_THROWS(expr, type)                                                     \
    do {                                                                             \
        g_checks++;                                                                  \
        bool thrown = false;                                                         \
        try {                                                                        \
            expr;                                                                    \
        } catch (const type &) {                                                     \
            thrown = true;                                                           \
        } catch (...) {                                                              \
        }                                                                            \
        if (!thrown) {                                                               \
            g_fail++;                                                                \
            std::fprintf(stderr, "FAIL %s:%d: %s did not throw %s\n", __FILE__, __LINE__, #expr, #type); \
        }                                                                            \
    } while (0)

static const u64 Q59 = 576460752272228353ull;

static RnsPolynomial filled(size_t n, const std::vector<u64> &moduli, u64 seed0, PolyRepForm form) {
    RnsPolynomial p(n, moduli.size(), moduli);
    for (size_t k = 0; k < moduli.size(); k++) orc_lcg_fill(seed0 + k, moduli[k], n, p[k].data());
    p.rep_form = form;
    return p;
}
static std::vector<u64> flat(const RnsPolynomial &p) {
    std::vector<u64> out;
    for (size_t k = 0; k < p.component_count(); k++) out.insert(out.end(), p[k].begin(), p[k].end());
    return out;
}
template <size_t K>
static std::vector<u64> flat(const std::array<RnsPolynomial, K> &ct) {
    std::vector<u64> out;
    for (const auto &p : ct) {
        auto f = flat(p);
        out.insert(out.end(), f.begin(), f.end());
    }
    return out;
}

// tests/ntt_t.cpp:91-181
static void test_ntt_round_trip() {
    std::mt19937_64 rng(7);
    for (size_t logn : {4, 7, 10, 12, 13, 15})
        for (u64 q : std::vector<u64>{65537ull, 260898817ull, 35184358850561ull, 36028796997599233ull, Q59}) {
            const size_t n = (size_t)1 << logn;
            std::vector<std::vector<u64>> cases(3, std::vector<u64>(n, 0));
            cases[0][0] = 1;
            cases[1][1] = 1;
            for (auto &c : cases[2]) c = rng() % q;
            for (const auto &x : cases) {
                std::vector<u64> y(x);
                ntt_negacyclic_inplace_lazy(logn, q, y.data());
                bool lazy_ok = true;
                for (u64 v : y) lazy_ok &= v < 2 * q;
                CHECK(lazy_ok);
                std::vector<u64> want(x);
                orc_ntt_fwd_lazy((unsigned)logn, q, want.data());
                CHECK(y == want);
                intt_negacyclic_inplace_lazy(logn, q, y.data());
                for (u64 v : y) lazy_ok &= v < 2 * q;
                CHECK(lazy_ok);
                batched_reduce_strict(q, n, y.data());
                CHECK(y == x);
            }
            // on an RnsPolynomial (ntt_t.cpp:148-180)
            RnsPolynomial poly(n, 2, std::vector<u64>{q, q});
            for (size_t k = 0; k < 2; k++)
                for (auto &c : poly[k]) c = rng() % q;
            RnsPolynomial orig(poly);
            ntt_negacyclic_inplace_lazy(poly);
            CHECK(poly.rep_form == PolyRepForm::value);
            intt_negacyclic_inplace(poly);
            CHECK(poly == orig);
        }
}

// tests/mod_arith_t.cpp:6-78
static void test_mod_arith() {
    const size_t len = 1000;
    for (u64 q : {65537ull, 33333333ull, 777777777777777ull, 1234567890111111111ull}) {
        u64 seed = 42;
        std::vector<u64> vec(len), copy;
        for (auto &v : vec) v = seed = ((seed ^ 893758435427369ull) * 65536) + 945738773644543ull;
        copy = vec;
        batched_barrett_lazy(q, len, vec.data());
        bool ok = true;
        for (size_t i = 0; i < len; i++) ok &= vec[i] < 2 * q && (vec[i] % q) == (copy[i] % q);
        CHECK(ok);
    }
    const u64 q = 1234567890111111111ull;
    u64 seed = 42;
    std::vector<u64> f(len), g(len), h(len);
    for (size_t i = 0; i < len; i++) {
        seed = (seed * 65968279837582827ull) ^ 3948528936546489545ull;
        f[i] = seed % q;
        seed = (seed * 43534547657678213ull) ^ 7955436776934235466ull;
        g[i] = seed % q;
    }
    batched_mul_mod_hybrid(q, len, f.data(), g.data(), h.data());
    bool ok = true;
    for (size_t i = 0; i < len; i++) ok &= h[i] == (u64)((u128)f[i] * g[i] % q);
    CHECK(ok);
    const u64 qm = 38589379749438777ull;
    std::vector<u128> big(len);
    std::vector<u64> red(len);
    std::mt19937_64 rng(3);
    for (auto &b : big) b = ((u128)(rng() % qm) << 64) | rng();
    batched_montgomery_128_lazy(qm, len, big.data(), red.data());
    ok = true;
    for (size_t i = 0; i < len; i++) ok &= red[i] < 2 * qm && (u64)((((u128)red[i]) << 64) % qm) == (u64)(big[i] % qm);
    CHECK(ok);
}

// tests/common_t.cpp:63-166 (container semantics)
static void test_container() {
    std::vector<u64> moduli{1099510054913ull, 1073479681ull, 1072496641ull};
    RnsPolynomial a = filled(64, moduli, 5, PolyRepForm::value);
    RnsPolynomial b(a); // deep copy
    CHECK(a == b);
    b[1][3] ^= 1;
    CHECK(!(a == b));
    CHECK(a.component_count() == 3 && a.dimension() == 64 && a.log_dimension() == 6);
    CHECK_THROWS(RnsPolynomial(48, 1, moduli), std::invalid_argument);
    CHECK_THROWS(RnsPolynomial(64, 4, moduli), std::invalid_argument);
    RnsPolynomial c(a);
    c.remove_components(1);
    CHECK(c.component_count() == 2 && c.modulus_vec().size() == 2);
    const auto fa = flat(a);
    const std::vector<u64> fa2(fa.begin(), fa.begin() + 128);
    CHECK(flat(c) == fa2);
    CHECK_THROWS(c.remove_components(3), std::invalid_argument);
    c.add_components({1072496641ull});
    CHECK(c.component_count() == 3 && c.modulus_at(2) == 1072496641ull);
    {
        const auto fc = flat(c);
        CHECK(std::vector<u64>(fc.begin(), fc.begin() + 128) == fa2);
    }
    c.remove_components(1);
    CHECK_THROWS(a += c, std::invalid_argument);
    RnsPolynomial coeff = filled(64, moduli, 9, PolyRepForm::coeff);
    CHECK_THROWS(a + coeff, std::invalid_argument);
    CHECK_THROWS(coeff * coeff, std::invalid_argument);
    // operators against the oracle, limb by limb
    RnsPolynomial d = filled(64, moduli, 77, PolyRepForm::value);
    auto sum = a + d, diff = a - d, prod = a * d, scaled = a * (u64)12345;
    for (size_t k = 0; k < 3; k++) {
        std::vector<u64> x(a[k].begin(), a[k].end()), y(d[k].begin(), d[k].end()), w(64);
        auto t = x;
        orc_add_lazy(moduli[k], 64, t.data(), y.data());
        CHECK(t == std::vector<u64>(sum[k].begin(), sum[k].end()));
        t = x;
        orc_sub_lazy(moduli[k], 64, t.data(), y.data());
        CHECK(t == std::vector<u64>(diff[k].begin(), diff[k].end()));
        orc_mul_hybrid_lazy(moduli[k], 64, x.data(), y.data(), w.data());
        CHECK(w == std::vector<u64>(prod[k].begin(), prod[k].end()));
        t = x;
        orc_mul_scalar_lazy(moduli[k], 64, t.data(), 12345);
        CHECK(t == std::vector<u64>(scaled[k].begin(), scaled[k].end()));
    }
    // moved-from objects are empty and destructible
    RnsPolynomial m(std::move(sum));
    CHECK(m.component_count() == 3);
}

// ciphertext ops against the oracle on identical inputs
static void test_scheme_ops(size_t logn, const std::vector<unsigned> &bits, unsigned pbits) {
    const size_t n = (size_t)1 << logn, L = bits.size();
    std::vector<u64> mods(L);
    u64 P = 0;
    CHECK(orc_ckks_pick_moduli(bits.data(), L, pbits, mods.data(), &P) == 0);
    std::vector<u64> ext(mods);
    ext.push_back(P);
    ckks::CkksCt ct1(RlweCt{filled(n, mods, 100, PolyRepForm::value), filled(n, mods, 110, PolyRepForm::value)});
    ckks::CkksCt ct2(RlweCt{filled(n, mods, 200, PolyRepForm::value), filled(n, mods, 210, PolyRepForm::value)});
    ct1.scaling_factor = ct2.scaling_factor = 1073741824.0;
    RlweKsk key;
    for (size_t r = 0; r < L; r++)
        key.push_back(RlweCt{filled(n, ext, 1000 + 100 * r, PolyRepForm::value), filled(n, ext, 1010 + 100 * r, PolyRepForm::value)});
    std::vector<u64> fkey;
    for (auto &row : key) {
        auto f = flat(row);
        fkey.insert(fkey.end(), f.begin(), f.end());
    }
    const auto f1 = flat(ct1), f2 = flat(ct2);

    auto quad = ckks::mult_low_level(ct1, ct2);
    std::vector<u64> wq(3 * L * n);
    orc_ckks_tensor((unsigned)logn, L, mods.data(), f1.data(), f2.data(), wq.data());
    CHECK(flat(quad) == wq);
    CHECK(quad.scaling_factor == ct1.scaling_factor * ct2.scaling_factor);

    auto e = ext_prod_montgomery(quad[2], key);
    std::vector<u64> we(2 * (L + 1) * n);
    orc_ext_prod((unsigned)logn, L, ext.data(), wq.data() + 2 * L * n, fkey.data(), we.data());
    CHECK(flat(e) == we);
    CHECK(e[0].component_count() == L + 1 && e[0].rep_form == PolyRepForm::value);

    auto relin = ckks::relinearize(quad, key);
    std::vector<u64> wr(2 * L * n);
    orc_ckks_relinearize((unsigned)logn, L, ext.data(), wq.data(), fkey.data(), wr.data());
    CHECK(flat(relin) == wr);
    auto prod = ckks::mult(ct1, ct2, key);
    CHECK(flat(prod) == wr);
    CHECK(prod.scaling_factor == quad.scaling_factor);
    {
        // the halves of a result share one device slab (no copies between chained calls): value semantics must hold anyway
        CHECK(prod[0].device_adjacent(prod[1]));
        ckks::CkksCt keep(prod);               // deep copy into storage of its own
        CHECK(!keep[0].device_adjacent(keep[1]));
        auto p2 = ckks::mult(ct1, ct2, key);
        p2[0] += p2[1];                        // in place on one half only
        CHECK(flat(keep) == wr);
        CHECK(flat(p2[1]) == std::vector<u64>(wr.begin() + L * n, wr.end()));
        std::vector<u64> sum0(wr.begin(), wr.begin() + L * n); // x += y in place, rns.cpp:78-84
        for (size_t k = 0; k < L; k++) orc_add_lazy(mods[k], n, sum0.data() + k * n, wr.data() + (L + k) * n);
        CHECK(flat(p2[0]) == sum0);
        RnsPolynomial alone(std::move(p2[1])); // one half outlives the other
        p2 = ckks::CkksCt();
        CHECK(flat(alone) == std::vector<u64>(wr.begin() + L * n, wr.end()));
        // a host write to one half un-shares nothing but makes the pair non-contiguous until re-uploaded
        auto p3 = ckks::mult(ct1, ct2, key);
        p3[1][0][0] = p3[1][0][0];
        auto again = ckks::relinearize(ckks::mult_low_level(ct1, ct2), key);
        CHECK(flat(again) == wr);
        CHECK(flat(p3) == wr);
    }

    if (L >= 2) {
        ckks::CkksCt rs(prod);
        rs.scaling_factor = prod.scaling_factor;
        ckks::rescale_inplace(rs);
        std::vector<u64> wrs(2 * (L - 1) * n);
        orc_ckks_rescale((unsigned)logn, L, mods.data(), wr.data(), wrs.data());
        CHECK(flat(rs) == wrs);
        CHECK(rs[0].component_count() == L - 1);
        CHECK(rs.scaling_factor == prod.scaling_factor / (double)mods[L - 1]);
        CHECK_THROWS(ckks::rescale_inplace(rs, 0), std::invalid_argument);
        {
            typedef const char *cstr;
            CHECK_THROWS(ckks::rescale_inplace(rs, 2), cstr);
        }

        bgv::BgvCt bct(RlweCt{ct1[0], ct1[1]});
        bct.plain_modulus = 65537;
        bgv::mod_switch_inplace(bct);
        std::vector<u64> wms(2 * (L - 1) * n);
        orc_bgv_mod_switch((unsigned)logn, L, mods.data(), 65537, f1.data(), wms.data());
        CHECK(flat(bct) == wms);
    }
    {
        bgv::BgvCt b1(RlweCt{ct1[0], ct1[1]}), b2(RlweCt{ct2[0], ct2[1]});
        b1.plain_modulus = b2.plain_modulus = 65537;
        auto bq = bgv::mult_low_level(b1, b2);
        CHECK(flat(bq) == wq);
        auto br = bgv::relinearize(bq, key);
        std::vector<u64> wbr(2 * L * n);
        orc_bgv_relinearize((unsigned)logn, L, ext.data(), 1, wq.data(), fkey.data(), wbr.data());
        CHECK(flat(br) == wbr);
        CHECK(br.plain_modulus == 65537);
        b2.plain_modulus = 3;
        CHECK_THROWS(bgv::mult_low_level(b1, b2), std::invalid_argument);
    }
    auto rot = ckks::rotate(ct1, key, 3);
    std::vector<u64> wrot(2 * L * n);
    orc_ckks_rotate((unsigned)logn, L, ext.data(), f1.data(), fkey.data(), 3, wrot.data());
    CHECK(flat(rot) == wrot);
    auto conj = ckks::conjugate(ct1, key);
    orc_ckks_conjugate((unsigned)logn, L, ext.data(), f1.data(), fkey.data(), wrot.data());
    CHECK(flat(conj) == wrot);
    auto cyc = cycle(ct1[0], 5);
    std::vector<u64> wc(L * n);
    orc_galois_cycle((unsigned)logn, L, f1.data(), wc.data(), 5);
    CHECK(flat(cyc) == wc);
    auto sum = ckks::add(ct1, ct2);
    CHECK(sum[0] == ct1[0] + ct2[0]);
    ct2.scaling_factor *= 2;
    CHECK_THROWS(ckks::add(ct1, ct2), std::invalid_argument);
    // key of the wrong level: exactly L + 1 limbs are required (rgsw.cpp:84-87)
    if (L >= 2) {
        RnsPolynomial shorter(quad[2]);
        shorter.remove_components(1);
        CHECK_THROWS(ext_prod_montgomery(shorter, key), std::invalid_argument);
    }
}

// tests/ckks_t.cpp:136-175 — exact rounded division by the dropped prime
static void test_rescale_exactness() {
    const size_t n = 8;
    std::vector<u64> mods(3);
    orc_prime_row(34, 3, mods.data());
    std::mt19937_64 rng(11);
    const u128 big_q = (u128)mods[0] * mods[1] * mods[2];
    std::vector<u128> coeffs(n);
    for (auto &c : coeffs) c = (((u128)rng() << 64) | rng()) % big_q;
    RnsPolynomial poly(n, 3, mods);
    for (size_t k = 0; k < 3; k++)
        for (size_t i = 0; i < n; i++) poly[k][i] = (u64)(coeffs[i] % mods[k]);
    ntt_negacyclic_inplace_lazy(poly);
    ckks::CkksCt ct(RlweCt{poly, poly});
    ckks::rescale_inplace(ct);
    intt_negacyclic_inplace(ct[0]);
    const u128 q01 = (u128)mods[0] * mods[1];
    const u64 inv = inverse_mod_prime(mods[0] % mods[1], mods[1]);
    bool ok = true;
    for (size_t i = 0; i < n; i++) {
        const u128 want = ((coeffs[i] + mods[2] / 2) / mods[2]) % q01;
        const u64 r0 = ct[0][0][i], r1 = ct[0][1][i];
        const u64 d = (u64)((u128)((r1 + mods[1] - r0 % mods[1]) % mods[1]) * inv % mods[1]);
        ok &= ((u128)r0 + (u128)mods[0] * d) % q01 == want;
    }
    CHECK(ok);
}

// encrypt_core / decrypt_core (rlwe.cpp:34-71) through the hehub:: mirror, word for word against the oracle
static void test_rlwe_cores(size_t logn, const std::vector<unsigned> &bits) {
    const size_t n = (size_t)1 << logn, L = bits.size();
    std::vector<u64> mods(L);
    u64 p = 0;
    orc_ckks_pick_moduli(bits.data(), L, 55, mods.data(), &p);
    auto sk_poly = filled(n, mods, 4000, PolyRepForm::value);
    RlweSk sk(std::move(sk_poly));
    auto mask = filled(n, mods, 4100, PolyRepForm::value);
    auto pt = filled(n, mods, 4200, PolyRepForm::coeff);
    RnsPolynomial err(n, L, mods);
    std::vector<u64> small(n);
    orc_lcg_fill(4300, 39, n, small.data());
    for (size_t k = 0; k < L; k++)
        for (size_t i = 0; i < n; i++) err[k][i] = small[i] < 19 ? mods[k] + small[i] - 19 : small[i] - 19;
    err.rep_form = PolyRepForm::coeff;

    auto ct = encrypt_core(pt, sk, mask, err);
    std::vector<u64> want(2 * L * n);
    orc_rlwe_encrypt_core((unsigned)logn, L, mods.data(), flat(pt).data(), flat(sk).data(), flat(mask).data(), flat(err).data(), want.data());
    CHECK(flat(ct) == want);
    CHECK(ct[0].rep_form == PolyRepForm::value);

    auto back = decrypt_core(ct, sk);
    std::vector<u64> want_pt(L * n);
    orc_rlwe_decrypt_core((unsigned)logn, L, mods.data(), want.data(), flat(sk).data(), want_pt.data());
    CHECK(flat(back) == want_pt);
    CHECK(back.rep_form == PolyRepForm::coeff);
    bool sum_ok = true; // decrypt(encrypt(pt)) == pt + e (mod q)
    for (size_t k = 0; k < L; k++)
        for (size_t i = 0; i < n; i++) sum_ok &= back[k][i] == (u64)(((unsigned __int128)pt[k][i] + err[k][i]) % mods[k]);
    CHECK(sum_ok);
    auto pt_ntt(pt);
    ntt_negacyclic_inplace_lazy(pt_ntt);
    CHECK_THROWS(encrypt_core(pt_ntt, sk, mask, err), std::invalid_argument); // rlwe.cpp:51-53
}

// rns_base_transform (rns_transform.cpp:11-126) and RlweKsk construction (keys.cpp:8-36) through the mirror
static void test_base_transform_and_ksk(size_t logn, const std::vector<unsigned> &bits, unsigned pbits) {
    const size_t n = (size_t)1 << logn, L = bits.size();
    std::vector<u64> mods(L);
    u64 P = 0;
    orc_ckks_pick_moduli(bits.data(), L, pbits, mods.data(), &P);
    std::vector<u64> ext(mods);
    ext.push_back(P);
    // ternary secrets in NTT form
    auto ternary = [&](u64 seed) {
        std::vector<u64> t(n);
        orc_lcg_fill(seed, 3, n, t.data());
        RnsPolynomial s(n, L, mods);
        for (size_t k = 0; k < L; k++)
            for (size_t i = 0; i < n; i++) s[k][i] = t[i] == 2 ? mods[k] - 1 : t[i];
        s.rep_form = PolyRepForm::coeff;
        return s;
    };
    auto so_coeff = ternary(5000), sc_coeff = ternary(5001);
    // many -> one on the small secret, one -> many on a lazy single-modulus polynomial
    auto single = rns_base_transform(so_coeff, {P});
    std::vector<u64> want1(n);
    CHECK(orc_base_transform_to_single(n, L, mods.data(), flat(so_coeff).data(), P, want1.data()) == 0);
    CHECK(flat(single) == want1);
    auto fan = rns_base_transform(single, mods);
    std::vector<u64> want2(L * n);
    orc_base_transform_from_single(P, n, want1.data(), mods.data(), L, want2.data());
    CHECK(flat(fan) == want2);
    auto as_values(so_coeff);
    ntt_negacyclic_inplace_lazy(as_values);
    CHECK_THROWS(rns_base_transform(as_values, {P}), std::logic_error);              // rns_transform.cpp:109-112
    typedef const char *cstr_t;
    CHECK_THROWS(rns_base_transform(so_coeff, std::vector<u64>{P, mods[0]}), cstr_t); // :123
    auto large = filled(n, mods, 5100, PolyRepForm::coeff);
    CHECK_THROWS(rns_base_transform(large, {P}), cstr_t);                      // CRT composition path not built

    RlweSk sk_o(std::move(as_values));
    auto sc_values(sc_coeff);
    ntt_negacyclic_inplace_lazy(sc_values);
    RlweSk sk_c(std::move(sc_values));
    std::vector<RnsPolynomial> masks, errors;
    std::vector<u64> small(n);
    for (size_t p = 0; p < L; p++) {
        masks.push_back(filled(n, ext, 5200 + 10 * p, PolyRepForm::value));
        orc_lcg_fill(5300 + p, 39, n, small.data());
        RnsPolynomial e(n, L + 1, ext);
        for (size_t k = 0; k <= L; k++)
            for (size_t i = 0; i < n; i++) e[k][i] = small[i] < 19 ? ext[k] + small[i] - 19 : small[i] - 19;
        e.rep_form = PolyRepForm::coeff;
        errors.push_back(std::move(e));
    }
    RlweKsk ksk(sk_c, sk_o, P, masks, errors);
    std::vector<u64> fm, fe, got;
    for (size_t p = 0; p < L; p++) {
        auto a = flat(masks[p]), b = flat(errors[p]);
        fm.insert(fm.end(), a.begin(), a.end());
        fe.insert(fe.end(), b.begin(), b.end());
        auto r = flat(ksk[p]);
        got.insert(got.end(), r.begin(), r.end());
    }
    std::vector<u64> want(L * 2 * (L + 1) * n);
    CHECK(orc_ksk_generate((unsigned)logn, L, ext.data(), flat(sk_c).data(), flat(sk_o).data(), fm.data(), fe.data(), want.data()) == 0);
    CHECK(got == want);
    CHECK(ksk.size() == L && ksk[0][0].rep_form == PolyRepForm::value);
}

// raw-slab files (hehub_b200/cpp/hehub/serialize.h): round trip, and a file for the Python side to read
static void test_serialize(const char *dir) {
    const size_t n = 64;
    const std::vector<u64> mods{65537ull, 260898817ull, Q59};
    auto poly = filled(n, mods, 6000, PolyRepForm::coeff);
    RlweCt ct{filled(n, mods, 6100, PolyRepForm::value), filled(n, mods, 6200, PolyRepForm::value)};
    RlweKsk key(RgswCt{ct, RlweCt{ct[1], ct[0]}});
    const std::string base = std::string(dir) + "/";
    b200::save(base + "poly.slab", poly);
    b200::save(base + "ct.slab", ct);
    b200::save(base + "ksk.slab", key);
    auto p2 = b200::load_polynomial(base + "poly.slab");
    CHECK(p2 == poly && p2.rep_form == PolyRepForm::coeff && p2.modulus_vec() == mods);
    auto c2 = b200::load_ciphertext(base + "ct.slab");
    CHECK(flat(c2) == flat(ct) && c2[0].rep_form == PolyRepForm::value);
    auto k2 = b200::load_key_switch_key(base + "ksk.slab");
    CHECK(k2.size() == 2 && flat(k2[1]) == flat(key[1]));
    CHECK_THROWS(b200::load_ciphertext(base + "poly.slab"), std::invalid_argument);
    CHECK_THROWS(b200::load_polynomial(base + "missing.slab"), std::runtime_error);
}

// ---- the sampling-based API under a seeded engine ------------------------------------------------------------
// The same statements oracle/ref_shim.cpp (ref_rng_*) runs through the unmodified reference, here through the mirror.
// Each stage's FNV-1a hash is written to <dir>/rng_hashes.txt; tests/test_cpp_mirror.py compares the file with
// tests/golden/reference_kat.json["rng"] (generated from the real reference by oracle/make_golden.py).
static u64 fnv_words(const std::vector<u64> &w, u64 h = 1469598103934665603ull) {
    for (u64 x : w) {
        h ^= x;
        h *= 1099511628211ull;
    }
    return h;
}
template <class CT>
static u64 fnv_of(const CT &ct) { return fnv_words(flat(ct)); }
static u64 fnv_ksk(const RlweKsk &k) {
    u64 h = 1469598103934665603ull;
    for (const auto &row : k)
        for (const auto &p : row) h = fnv_words(flat(p), h);
    return h;
}
static RnsPolynomial lcg_small_poly(size_t n, const std::vector<u64> &moduli, u64 seed0, u64 bound) {
    RnsPolynomial p(n, moduli.size(), moduli);
    u64 s = seed0;
    std::vector<u64> v(n);
    for (auto &c : v) {
        s = s * 6364136223846793005ull + 1442695040888963407ull;
        c = (s >> 33) % bound;
    }
    for (size_t k = 0; k < moduli.size(); k++)
        for (size_t i = 0; i < n; i++) p[k][i] = v[i] % moduli[k];
    p.rep_form = PolyRepForm::coeff;
    return p;
}
static void put_line(FILE *f, const char *tag, u64 seed, const std::vector<u64> &hashes) {
    std::fprintf(f, "%s %llu", tag, (unsigned long long)seed);
    for (u64 h : hashes) std::fprintf(f, " %016llx", (unsigned long long)h);
    std::fprintf(f, "\n");
}

static void rng_samples(FILE *f, u64 seed, size_t logn, const std::vector<u64> &moduli) {
    rand_engine.seed(seed);
    RnsPolyParams params{(size_t)1 << logn, moduli.size(), moduli};
    std::vector<u64> hs;
    hs.push_back(fnv_of(get_rand_ternary_poly(params)));
    hs.push_back(fnv_of(get_rand_uniform_poly(params)));
    hs.push_back(fnv_of(get_rand_gaussian_poly(params)));
    put_line(f, "samples", seed, hs);
}

static void rng_scenario_ckks(FILE *f, u64 seed, size_t logn, const std::vector<size_t> &bits, size_t additional_bits) {
    rand_engine.seed(seed);
    const size_t n = (size_t)1 << logn;
    auto params = ckks::create_params(n, bits, additional_bits, 1099511627776.0);
    RlweSk sk(params);
    auto relin = get_relin_key(sk, params.additional_mod);
    auto conj = get_conj_key(sk, params.additional_mod);
    auto rot = get_rot_key(sk, params.additional_mod, 3);
    std::vector<u64> hs;
    hs.push_back(fnv_of(static_cast<const RnsPolynomial &>(sk)));
    hs.push_back(fnv_ksk(relin));
    hs.push_back(fnv_ksk(conj));
    hs.push_back(fnv_ksk(rot));
    CkksPt pt1(lcg_small_poly(n, params.moduli, 11, 1 << 20)), pt2(lcg_small_poly(n, params.moduli, 12, 1 << 20));
    pt1.scaling_factor = pt2.scaling_factor = params.initial_scaling_factor;
    auto ct1 = ckks::encrypt(pt1, sk), ct2 = ckks::encrypt(pt2, sk);
    hs.push_back(fnv_of(static_cast<const RlweCt &>(ct1)));
    hs.push_back(fnv_of(static_cast<const RlweCt &>(ct2)));
    hs.push_back(fnv_of(static_cast<const RlweCt &>(ckks::add_plain(ct1, pt2))));
    hs.push_back(fnv_of(static_cast<const RlweCt &>(ckks::sub_plain(ct1, pt2))));
    hs.push_back(fnv_of(static_cast<const RlweCt &>(ckks::mult_plain(ct1, pt2))));
    auto prod = ckks::mult(ct1, ct2, relin);
    hs.push_back(fnv_of(static_cast<const RlweCt &>(prod)));
    ckks::rescale_inplace(prod);
    hs.push_back(fnv_of(static_cast<const RlweCt &>(prod)));
    hs.push_back(fnv_of(static_cast<const RlweCt &>(ckks::rotate(ct1, rot))));
    hs.push_back(fnv_of(static_cast<const RlweCt &>(ckks::conjugate(ct2, conj))));
    // a rescaled ciphertext (L - 1 limbs) against the full-length secret key: the reference's canonical flow, tests/ckks_t.cpp:303
    auto dec = ckks::decrypt(prod, sk);
    CHECK(dec.component_count() == params.component_count - 1);
    CHECK(std::abs(dec.scaling_factor - prod.scaling_factor) == 0.0);
    hs.push_back(fnv_of(static_cast<const RnsPolynomial &>(dec)));
    put_line(f, "ckks", seed, hs);
}

static void rng_scenario_bgv(FILE *f, u64 seed, size_t logn, const std::vector<size_t> &bits, size_t additional_bits, u64 t) {
    rand_engine.seed(seed);
    const size_t n = (size_t)1 << logn;
    auto params = ckks::create_params(n, bits, additional_bits, 1.0);
    RlweSk sk(params);
    auto relin = get_relin_key(sk, params.additional_mod);
    std::vector<u64> d1(n), d2(n);
    u64 s = 77;
    for (size_t i = 0; i < n; i++) {
        s = s * 6364136223846793005ull + 1442695040888963407ull;
        d1[i] = (s >> 20) % t;
        s = s * 6364136223846793005ull + 1442695040888963407ull;
        d2[i] = (s >> 20) % t;
    }
    std::vector<u64> hs;
    auto pt1 = bgv::simd_encode(d1, t, n), pt2 = bgv::simd_encode(d2, t, n);
    hs.push_back(fnv_of(pt1));
    auto ct1 = bgv::encrypt(pt1, sk), ct2 = bgv::encrypt(pt2, sk);
    hs.push_back(fnv_of(static_cast<const RlweCt &>(ct1)));
    hs.push_back(fnv_of(static_cast<const RlweCt &>(ct2)));
    hs.push_back(fnv_of(static_cast<const RlweCt &>(bgv::add_plain(ct1, pt2))));
    hs.push_back(fnv_of(static_cast<const RlweCt &>(bgv::sub_plain(ct1, pt2))));
    auto ct_prod_plain = bgv::mult_plain(ct1, pt2);
    hs.push_back(fnv_of(static_cast<const RlweCt &>(ct_prod_plain)));
    auto ct_res = bgv::add(ct_prod_plain, ct2);
    hs.push_back(fnv_of(static_cast<const RlweCt &>(ct_res)));
    auto prod = bgv::relinearize(bgv::mult_low_level(ct1, ct2), relin);
    hs.push_back(fnv_of(static_cast<const RlweCt &>(prod)));
    bgv::mod_switch_inplace(prod);
    hs.push_back(fnv_of(static_cast<const RlweCt &>(prod)));
    auto dec = bgv::decrypt(ct_res, sk);
    hs.push_back(fnv_of(dec));
    auto out = bgv::simd_decode(dec);
    bool slots_ok = out.size() == n; // tests/bgv_t.cpp:160-190: d1 * d2 + d2 in every slot
    for (size_t i = 0; i < n && slots_ok; i++) slots_ok = out[i] == (u64)(((unsigned __int128)d1[i] * d2[i] + d2[i]) % t);
    // t = 12289 sits far below 2^14, where the approximate reduction of ntt.cpp:171-175 leaves words above 2q and the
    // reference's own decode is wrong: that case only has to equal the reference (hash), not the arithmetic
    if (t == 65537) CHECK(slots_ok);
    auto switched = ct1;
    bgv::mod_switch_inplace(switched);
    hs.push_back(fnv_of(static_cast<const RlweCt &>(switched)));
    auto pt_back = bgv::decrypt(switched, sk);
    hs.push_back(fnv_of(pt_back));
    auto pt1_strict(pt1);
    reduce_strict(pt1_strict);
    CHECK(pt_back == pt1_strict); // tests/bgv_t.cpp:229-259: the plaintext survives the switch
    hs.push_back(fnv_words(out));
    put_line(f, "bgv", seed, hs);
}

// create_params (rlwe.cpp:9-29, ckks/basics.cpp:14-64) against SURVEY §8(d)'s values and the error behaviour
static void test_params() {
    auto c3 = ckks::create_params(8192, {40, 30, 30, 30}, 40, std::pow(2.0, 30));
    CHECK(c3.additional_mod == 1099510054913ull);
    CHECK((c3.moduli == std::vector<u64>{1099507695617ull, 1073479681ull, 1072496641ull, 1071513601ull}));
    CHECK(c3.component_count == 4 && c3.dimension == 8192 && c3.initial_scaling_factor == std::pow(2.0, 30));
    auto ex = ckks::create_params(4096, 30); // examples/ckks_example.cpp: L = 2 {39, 30} + P 39
    CHECK(ex.component_count == 2 && ex.moduli.size() == 2 && ex.moduli[1] == 1073479681ull);
    CHECK(ex.additional_mod == prime_lists[39][0] && ex.moduli[0] == prime_lists[39][1]);
    auto six = ckks::create_params(8192, 30); // the two-argument form yields L = 6 at N = 8192
    CHECK(six.component_count == 6);
    auto plain = create_params(4096, {30, 30, 45});
    CHECK((plain.moduli == std::vector<u64>{prime_lists[30][0], prime_lists[30][1], prime_lists[45][0]}));
    CHECK(prime_lists.size() == 60 && prime_lists[26].empty() && prime_lists[45].size() == 19 && prime_lists[59].size() == 20);
    typedef const char *cstr_t;
    CHECK_THROWS(ckks::create_params(3000, 30), cstr_t);
    CHECK_THROWS(ckks::create_params(1024, 20), cstr_t);
    CHECK_THROWS(create_params(4096, std::vector<int>(21, 30)), cstr_t); // a row holds 20 primes
}

// operands of different lengths (rns.cpp:58-140): the ADVICE case — rescale, then decrypt with the full-length key
static void test_mismatched_component_counts() {
    const size_t n = 256;
    const std::vector<u64> mods{1099507695617ull, 1073479681ull, 1072496641ull};
    auto longer = filled(n, mods, 7000, PolyRepForm::value);
    auto shorter = filled(n, {mods[0], mods[1]}, 7100, PolyRepForm::value);
    auto prod = shorter * longer; // min(a, b) components
    CHECK(prod.component_count() == 2);
    auto prod2 = longer * shorter;
    CHECK(prod2.component_count() == 2 && prod2 == prod);
    std::vector<u64> want(2 * n);
    for (size_t k = 0; k < 2; k++) orc_mul_hybrid_lazy(mods[k], n, shorter[k].data(), longer[k].data(), want.data() + k * n);
    CHECK(flat(prod) == want);
    auto sum(shorter);
    sum += longer; // b may be longer than self
    for (size_t k = 0; k < 2; k++)
        for (size_t i = 0; i < n; i++) {
            u64 x = shorter[k][i] + longer[k][i];
            want[k * n + i] = x - (x >= 2 * mods[k] ? 2 * mods[k] : 0);
        }
    CHECK(flat(sum) == want);
    auto bad(longer);
    CHECK_THROWS(bad += shorter, std::invalid_argument); // "Operand b contains less components than self."
    CHECK_THROWS(bad -= shorter, std::invalid_argument);
    auto other = filled(n, {mods[0], mods[2]}, 7200, PolyRepForm::value);
    CHECK_THROWS(shorter * other, std::invalid_argument); // moduli mismatch in the common prefix
    auto self_mul(longer);
    self_mul *= shorter; // rns.h:255-259: self = self * b, so self shrinks
    CHECK(self_mul.component_count() == 2 && self_mul == prod);
    // decrypt_core of a 2-limb ciphertext with the 3-limb key
    RlweSk sk(filled(n, mods, 7300, PolyRepForm::value));
    RlweCt ct{filled(n, {mods[0], mods[1]}, 7400, PolyRepForm::value), filled(n, {mods[0], mods[1]}, 7500, PolyRepForm::value)};
    auto pt = decrypt_core(ct, sk);
    std::vector<u64> want_pt(2 * n), ctw = flat(ct), skw = flat(sk);
    skw.resize(2 * n);
    std::vector<u64> two{mods[0], mods[1]};
    orc_rlwe_decrypt_core(8, 2, two.data(), ctw.data(), skw.data(), want_pt.data());
    CHECK(flat(pt) == want_pt);
}

// CKKS encoder / decoder (ckks_encoding.h): the same data as oracle/ref_shim.cpp::ref_ckks_codec_scenario.  The plaintext hash
// goes to the file (bit-exact target), the decoded slots as %.17g (tolerance target, checked by tests/test_cpp_mirror.py).
static void codec_scenario(FILE *f, size_t logn, const std::vector<size_t> &bits, size_t additional_bits, double log2_scaling, u64 seed,
                           size_t count) {
    const size_t n = (size_t)1 << logn;
    auto params = ckks::create_params(n, bits, additional_bits, std::pow(2.0, log2_scaling));
    std::vector<cc_double> data(count);
    u64 s = seed;
    for (auto &d : data) {
        s = s * 6364136223846793005ull + 1442695040888963407ull;
        const double re = (double)(int64_t)(s >> 11) / 4503599627370496.0 - 1.0;
        s = s * 6364136223846793005ull + 1442695040888963407ull;
        const double im = (double)(int64_t)(s >> 11) / 4503599627370496.0 - 1.0;
        d = cc_double(re, im);
    }
    auto pt = ckks::simd_encode(data, params);
    CHECK(pt.rep_form == PolyRepForm::coeff && pt.scaling_factor == std::pow(2.0, log2_scaling));
    auto back = ckks::simd_decode<cc_double>(pt);
    CHECK(back.size() == n / 2);
    double worst = 0; // decode(encode(x)) == x up to the rounding of one coefficient unit
    for (size_t i = 0; i < count; i++) worst = std::max(worst, std::abs(back[i] - data[i]));
    CHECK(worst < (double)n / std::pow(2.0, log2_scaling) + 1e-9);
    std::fprintf(f, "codec %llu %016llx", (unsigned long long)seed, (unsigned long long)fnv_of(static_cast<const RnsPolynomial &>(pt)));
    double abs_sum = 0;
    for (const auto &c : back) abs_sum += std::abs(c.real()) + std::abs(c.imag());
    std::fprintf(f, " %.17g", abs_sum);
    for (size_t i = 0; i < 8 && i < back.size(); i++) std::fprintf(f, " %.17g %.17g", back[i].real(), back[i].imag());
    std::fprintf(f, "\n");
}

// examples/ckks_example.cpp with fewer terms: the Basel series through encode -> encrypt -> mult -> add -> decrypt -> decode
static void test_ckks_example(int terms) {
    rand_engine.seed(2024);
    auto params = ckks::create_params(4096, 30);
    CkksSk sk(params);
    auto relin_key = get_relin_key(sk, params.additional_mod);
    CkksCt ct_sum;
    double want = 0;
    for (int i = 1; i <= terms; i++) {
        auto pt = ckks::encode(1.0 / i, params);
        auto ct = ckks::encrypt(pt, sk);
        auto ct_squared = ckks::mult(ct, ct, relin_key);
        ct_sum = (i == 1) ? ct_squared : ckks::add(ct_sum, ct_squared);
        want += 1.0 / i / i;
    }
    const double sum = ckks::decode(ckks::decrypt(ct_sum, sk));
    CHECK(std::abs(sum - want) < 1e-4); // the reference prints (1.64483, 1.64493) after 10^4 terms at this precision
    if (std::abs(sum - want) >= 1e-4) std::fprintf(stderr, "ckks example: got %.9f want %.9f\n", sum, want);
}

static void test_rng_api(const char *dir) {
    const std::string path = std::string(dir) + "/rng_hashes.txt";
    FILE *f = std::fopen(path.c_str(), "w");
    CHECK(f != nullptr);
    if (!f) return;
    // the cases of oracle/make_golden.py ("rng")
    rng_samples(f, 1, 6, {65537ull, 260898817ull});
    rng_samples(f, 2, 10, {1073479681ull, Q59});
    rng_samples(f, 3, 12, {36028796997599233ull, Q59, 1099510054913ull});
    rng_scenario_ckks(f, 5, 10, {40, 30, 30}, 40);
    rng_scenario_ckks(f, 6, 12, {39, 30}, 39);
    rng_scenario_ckks(f, 7, 13, {40, 30, 30, 30}, 40);
    rng_scenario_ckks(f, 8, 11, {55, 50}, 55);
    rng_scenario_bgv(f, 9, 10, {50, 45, 45}, 50, 65537);
    rng_scenario_bgv(f, 10, 12, {52, 50, 48}, 52, 65537);
    rng_scenario_bgv(f, 11, 8, {50, 45}, 50, 12289);
    // the cases of oracle/make_golden.py ("ckks_codec")
    codec_scenario(f, 10, {40, 30, 30}, 40, 30, 3, 512);
    codec_scenario(f, 12, {39, 30}, 39, 30, 4, 2048);
    codec_scenario(f, 8, {50, 50}, 55, 70, 5, 100);
    codec_scenario(f, 11, {40, 30}, 40, 50, 6, 1024);
    codec_scenario(f, 6, {30}, 40, 20, 7, 32);
    std::fclose(f);
}

int main(int argc, char **argv) {
    try {
        test_params();
        test_mismatched_component_counts();
        test_ntt_round_trip();
        test_mod_arith();
        test_container();
        test_scheme_ops(10, {40, 30, 30}, 40);
        test_scheme_ops(12, {39}, 39);
        test_scheme_ops(13, {40, 30, 30, 30}, 40);
        test_rescale_exactness();
        test_rlwe_cores(10, {40, 30});
        test_rlwe_cores(13, {40, 30, 30, 30});
        test_base_transform_and_ksk(8, {40, 30}, 45);
        test_base_transform_and_ksk(12, {40, 30, 30}, 45);
        if (argc > 1) test_serialize(argv[1]);
        if (argc > 1) test_rng_api(argv[1]);
        test_ckks_example(argc > 2 ? std::atoi(argv[2]) : 12);
    } catch (const std::exception &e) {
        std::fprintf(stderr, "unexpected exception: %s\n", e.what());
        return 2;
    } catch (const char *msg) {
        std::fprintf(stderr, "unexpected exception: %s\n", msg);
        return 2;
    }
    std::printf("%d checks, %d failures\n", g_checks, g_fail);
    return g_fail ? 1 : 0;
}
