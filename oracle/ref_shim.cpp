/*
 * ref_shim.cpp — C entry points onto the UNMODIFIED reference (primihub/hehub).
 *
 * TEST INFRASTRUCTURE ONLY.  oracle/Makefile compiles this file together with the
 * reference's own sources where they lie under $(REF) (default /root/reference)
 * into oracle/_ref/libhehub_ref.so.  No reference source is copied into this
 * repository.  The entry points mirror oracle/hehub_oracle.h one-for-one
 * (prefix ref_ instead of orc_) so the same harness can drive either; they
 * call the reference's public functions on its own containers.
 */
#include "fhe/bgv/bgv.h"
#include "fhe/ckks/ckks.h"
#include "fhe/common/mod_arith.h"
#include "fhe/common/ntt.h"
#include "fhe/common/permutation.h"
#include "fhe/common/primelists.h"
#include "fhe/common/rns_transform.h"
#include "fhe/primitives/keys.h"
#include "fhe/primitives/rgsw.h"

#include <cstring>
#include <random>
#include <chrono>
#include <vector>

#include "fhe/common/sampling.h"

namespace hehub {
extern std::default_random_engine rand_engine; // src/fhe/common/sampling.cpp:13 (a plain global of the reference)
}

using namespace hehub;

namespace {

RnsPolynomial load_poly(size_t n, size_t L, const u64 *moduli, const u64 *src, bool value_form) {
    RnsPolynomial p(n, L, std::vector<u64>(moduli, moduli + L));
    for (size_t k = 0; k < L; k++) std::memcpy(p[k].data(), src + k * n, n * sizeof(u64));
    p.rep_form = value_form ? PolyRepForm::value : PolyRepForm::coeff;
    return p;
}

void store_poly(const RnsPolynomial &p, u64 *dst) {
    const size_t n = p.dimension();
    for (size_t k = 0; k < p.component_count(); k++)
        std::memcpy(dst + k * n, p[k].data(), n * sizeof(u64));
}

RlweKsk load_key(size_t n, size_t L, const u64 *ext_moduli, const u64 *key) {
    RlweKsk ksk;
    for (size_t p = 0; p < L; p++) {
        RlweCt row{load_poly(n, L + 1, ext_moduli, key + (p * 2 + 0) * (L + 1) * n, true),
                   load_poly(n, L + 1, ext_moduli, key + (p * 2 + 1) * (L + 1) * n, true)};
        ksk.push_back(std::move(row));
    }
    return ksk;
}

static u64 fnv_poly(const RnsPolynomial &p, u64 h = 1469598103934665603ull) {
    for (size_t k = 0; k < p.component_count(); k++)
        for (size_t i = 0; i < p.dimension(); i++) {
            h ^= p[k][i];
            h *= 1099511628211ull;
        }
    return h;
}
template <size_t K> static u64 fnv_ct(const std::array<RnsPolynomial, K> &ct) {
    u64 h = 1469598103934665603ull;
    for (const auto &p : ct) h = fnv_poly(p, h);
    return h;
}
static u64 fnv_ksk(const RlweKsk &k) {
    u64 h = 1469598103934665603ull;
    for (const auto &row : k)
        for (const auto &p : row) h = fnv_poly(p, h);
    return h;
}
static RnsPolynomial lcg_poly(size_t n, const std::vector<u64> &moduli, u64 seed0, u64 bound) {
    RnsPolynomial p(n, moduli.size(), moduli);
    u64 s = seed0; // one small signed-free coefficient stream shared by the limbs (a valid small plaintext)
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

template <class F> int guarded(F &&f) {
    try {
        f();
        return 0;
    } catch (const std::invalid_argument &) {
        return 1;
    } catch (const std::logic_error &) {
        return 3;
    } catch (const char *) {
        return 4;
    } catch (...) {
        return 5;
    }
}

} // namespace

extern "C" {

u64 ref_harvey_lazy(u64 q, u64 x, u64 w, u64 wh) { return mul_mod_harvey_lazy(q, x, w, wh); }
u64 ref_inverse_mod_prime(u64 elem, u64 prime) { return inverse_mod_prime(elem, prime); }
void ref_barrett_lazy(u64 q, size_t n, u64 *x) { batched_barrett_lazy(q, n, x); }
void ref_barrett(u64 q, size_t n, u64 *x) { batched_barrett(q, n, x); }
void ref_reduce_strict(u64 q, size_t n, u64 *x) { batched_reduce_strict(q, n, x); }
void ref_mul_hybrid_lazy(u64 q, size_t n, const u64 *a, const u64 *b, u64 *c) {
    batched_mul_mod_hybrid_lazy(q, n, a, b, c);
}
void ref_mul_barrett_lazy(u64 q, size_t n, const u64 *a, const u64 *b, u64 *c) {
    batched_mul_mod_barrett_lazy(q, n, a, b, c);
}
void ref_montgomery128_lazy(u64 q, size_t n, const u64 *in_lohi, u64 *out) {
    std::vector<u128> in(n);
    for (size_t i = 0; i < n; i++) in[i] = ((u128)in_lohi[2 * i + 1] << 64) | in_lohi[2 * i];
    batched_montgomery_128_lazy(q, n, in.data(), out);
}

int ref_ntt_fwd_lazy(unsigned logn, u64 q, u64 *x) {
    return guarded([&] { ntt_negacyclic_inplace_lazy(logn, q, x); });
}
int ref_intt_lazy(unsigned logn, u64 q, u64 *x) {
    return guarded([&] { intt_negacyclic_inplace_lazy(logn, q, x); });
}

/* RnsPolynomial-level elementwise ops: rns.cpp:58-171 */
int ref_poly_add(unsigned logn, size_t L, const u64 *moduli, u64 *x, const u64 *y) {
    return guarded([&] {
        size_t n = (size_t)1 << logn;
        auto a = load_poly(n, L, moduli, x, true), b = load_poly(n, L, moduli, y, true);
        a += b;
        store_poly(a, x);
    });
}
int ref_poly_sub(unsigned logn, size_t L, const u64 *moduli, u64 *x, const u64 *y) {
    return guarded([&] {
        size_t n = (size_t)1 << logn;
        auto a = load_poly(n, L, moduli, x, true), b = load_poly(n, L, moduli, y, true);
        a -= b;
        store_poly(a, x);
    });
}
int ref_poly_mul_scalar(unsigned logn, size_t L, const u64 *moduli, u64 *x, u64 scalar) {
    return guarded([&] {
        size_t n = (size_t)1 << logn;
        auto a = load_poly(n, L, moduli, x, true);
        a *= scalar;
        store_poly(a, x);
    });
}
int ref_poly_ntt_fwd(unsigned logn, size_t L, const u64 *moduli, u64 *x) {
    return guarded([&] {
        size_t n = (size_t)1 << logn;
        auto a = load_poly(n, L, moduli, x, false);
        ntt_negacyclic_inplace_lazy(a);
        store_poly(a, x);
    });
}
int ref_poly_intt(unsigned logn, size_t L, const u64 *moduli, u64 *x, int strict) {
    return guarded([&] {
        size_t n = (size_t)1 << logn;
        auto a = load_poly(n, L, moduli, x, true);
        if (strict)
            intt_negacyclic_inplace(a);
        else
            intt_negacyclic_inplace_lazy(a);
        store_poly(a, x);
    });
}

int ref_ckks_tensor(unsigned logn, size_t L, const u64 *moduli, const u64 *ct1, const u64 *ct2,
                    u64 *quad) {
    return guarded([&] {
        size_t n = (size_t)1 << logn;
        CkksCt a, b;
        for (int h = 0; h < 2; h++) {
            a[h] = load_poly(n, L, moduli, ct1 + h * L * n, true);
            b[h] = load_poly(n, L, moduli, ct2 + h * L * n, true);
        }
        auto prod = ckks::mult_low_level(a, b);
        for (int j = 0; j < 3; j++) store_poly(prod[j], quad + j * L * n);
    });
}

int ref_ext_prod(unsigned logn, size_t L, const u64 *ext_moduli, const u64 *in, const u64 *key,
                 u64 *out) {
    return guarded([&] {
        size_t n = (size_t)1 << logn;
        auto pt = load_poly(n, L, ext_moduli, in, true);
        auto ksk = load_key(n, L, ext_moduli, key);
        auto ct = ext_prod_montgomery(pt, ksk);
        for (int h = 0; h < 2; h++) store_poly(ct[h], out + h * (L + 1) * n);
    });
}

int ref_ckks_rescale(unsigned logn, size_t L, const u64 *moduli, const u64 *ct, u64 *out) {
    return guarded([&] {
        size_t n = (size_t)1 << logn;
        CkksCt c;
        for (int h = 0; h < 2; h++) c[h] = load_poly(n, L, moduli, ct + h * L * n, true);
        ckks::rescale_inplace(c);
        for (int h = 0; h < 2; h++) store_poly(c[h], out + h * (L - 1) * n);
    });
}

int ref_bgv_mod_switch(unsigned logn, size_t L, const u64 *moduli, u64 t, const u64 *ct,
                       u64 *out) {
    return guarded([&] {
        size_t n = (size_t)1 << logn;
        BgvCt c;
        for (int h = 0; h < 2; h++) c[h] = load_poly(n, L, moduli, ct + h * L * n, true);
        c.plain_modulus = t;
        bgv::mod_switch_inplace(c);
        for (int h = 0; h < 2; h++) store_poly(c[h], out + h * (L - 1) * n);
    });
}

int ref_ckks_relinearize(unsigned logn, size_t L, const u64 *ext_moduli, const u64 *quad,
                         const u64 *key, u64 *out) {
    return guarded([&] {
        size_t n = (size_t)1 << logn;
        ckks::CkksQuadraticCt q3;
        for (int j = 0; j < 3; j++) q3[j] = load_poly(n, L, ext_moduli, quad + j * L * n, true);
        auto ksk = load_key(n, L, ext_moduli, key);
        auto ct = ckks::relinearize(q3, ksk);
        for (int h = 0; h < 2; h++) store_poly(ct[h], out + h * L * n);
    });
}

/* bgv::relinearize always mod-switches with the default plain_modulus 1
 * (bgv/arith.cpp:72-73), so `t` is accepted only for signature symmetry. */
int ref_bgv_relinearize(unsigned logn, size_t L, const u64 *ext_moduli, u64 t, const u64 *quad,
                        const u64 *key, u64 *out) {
    return guarded([&] {
        size_t n = (size_t)1 << logn;
        bgv::BgvQuadraticCt q3;
        for (int j = 0; j < 3; j++) q3[j] = load_poly(n, L, ext_moduli, quad + j * L * n, true);
        q3.plain_modulus = t;
        auto ksk = load_key(n, L, ext_moduli, key);
        auto ct = bgv::relinearize(q3, ksk);
        for (int h = 0; h < 2; h++) store_poly(ct[h], out + h * L * n);
    });
}

int ref_ckks_mult_relin(unsigned logn, size_t L, const u64 *ext_moduli, const u64 *ct1,
                        const u64 *ct2, const u64 *key, u64 *out) {
    return guarded([&] {
        size_t n = (size_t)1 << logn;
        CkksCt a, b;
        for (int h = 0; h < 2; h++) {
            a[h] = load_poly(n, L, ext_moduli, ct1 + h * L * n, true);
            b[h] = load_poly(n, L, ext_moduli, ct2 + h * L * n, true);
        }
        auto ksk = load_key(n, L, ext_moduli, key);
        auto ct = ckks::mult(a, b, ksk);
        for (int h = 0; h < 2; h++) store_poly(ct[h], out + h * L * n);
    });
}

int ref_galois_cycle(unsigned logn, size_t L, const u64 *in, u64 *out, size_t step) {
    return guarded([&] {
        size_t n = (size_t)1 << logn;
        std::vector<u64> moduli(L, 65537);
        auto p = load_poly(n, L, moduli.data(), in, true);
        store_poly(cycle(p, step), out);
    });
}

int ref_galois_involution(unsigned logn, size_t L, const u64 *in, u64 *out) {
    return guarded([&] {
        size_t n = (size_t)1 << logn;
        std::vector<u64> moduli(L, 65537);
        auto p = load_poly(n, L, moduli.data(), in, true);
        store_poly(involution(p), out);
    });
}

int ref_ckks_rotate(unsigned logn, size_t L, const u64 *ext_moduli, const u64 *ct, const u64 *key,
                    size_t step, u64 *out) {
    return guarded([&] {
        size_t n = (size_t)1 << logn;
        CkksCt c;
        for (int h = 0; h < 2; h++) c[h] = load_poly(n, L, ext_moduli, ct + h * L * n, true);
        auto ksk = load_key(n, L, ext_moduli, key);
        auto r = ckks::rotate(c, ksk, step);
        for (int h = 0; h < 2; h++) store_poly(r[h], out + h * L * n);
    });
}

int ref_ckks_conjugate(unsigned logn, size_t L, const u64 *ext_moduli, const u64 *ct,
                       const u64 *key, u64 *out) {
    return guarded([&] {
        size_t n = (size_t)1 << logn;
        CkksCt c;
        for (int h = 0; h < 2; h++) c[h] = load_poly(n, L, ext_moduli, ct + h * L * n, true);
        auto ksk = load_key(n, L, ext_moduli, key);
        auto r = ckks::conjugate(c, ksk);
        for (int h = 0; h < 2; h++) store_poly(r[h], out + h * L * n);
    });
}

/* decrypt_core is the reference's own function; encrypt_core draws its samples from a global RNG,
 * so the shim repeats its three statements (rlwe.cpp:50, 54-58) with the reference's own operators
 * on caller-supplied samples. */
int ref_rlwe_decrypt_core(unsigned logn, size_t L, const u64 *moduli, const u64 *ct, const u64 *sk, u64 *pt) {
    return guarded([&] {
        size_t n = (size_t)1 << logn;
        RlweCt c;
        for (int h = 0; h < 2; h++) c[h] = load_poly(n, L, moduli, ct + h * L * n, true);
        RlweSk s(load_poly(n, L, moduli, sk, true));
        auto r = decrypt_core(c, s);
        store_poly(r, pt);
    });
}

int ref_rlwe_encrypt_core(unsigned logn, size_t L, const u64 *moduli, const u64 *pt, const u64 *sk, const u64 *c1,
                          const u64 *e, u64 *out) {
    return guarded([&] {
        size_t n = (size_t)1 << logn;
        RlweSk s(load_poly(n, L, moduli, sk, true));
        auto mask = load_poly(n, L, moduli, c1, true);
        auto ex = load_poly(n, L, moduli, e, false);
        ntt_negacyclic_inplace_lazy(ex);          // sampling.cpp:66
        auto c0 = ex - mask * s;                  // rlwe.cpp:50
        auto pt_ntt = load_poly(n, L, moduli, pt, false);
        ntt_negacyclic_inplace_lazy(pt_ntt);      // rlwe.cpp:54-55
        c0 += pt_ntt;                             // rlwe.cpp:58
        store_poly(c0, out);
        store_poly(mask, out + L * n);
    });
}

/* rns_base_transform is the reference's own function (both directions). */
int ref_base_transform_from_single(u64 q_old, size_t n, const u64 *in, const u64 *new_moduli, size_t Lnew, u64 *out) {
    return guarded([&] {
        auto p = load_poly(n, 1, &q_old, in, false);
        auto r = rns_base_transform(p, std::vector<u64>(new_moduli, new_moduli + Lnew));
        store_poly(r, out);
    });
}

int ref_base_transform_to_single(size_t n, size_t L, const u64 *old_moduli, const u64 *in, u64 new_modulus, u64 *out) {
    return guarded([&] {
        auto p = load_poly(n, L, old_moduli, in, false);
        auto r = rns_base_transform(p, std::vector<u64>{new_modulus});
        store_poly(r, out);
    });
}

/* RlweKsk::RlweKsk (keys.cpp:8-36) draws its RLWE samples from the global RNG inside rgsw_encrypt; the shim
 * repeats the constructor's statements with the reference's own operators, taking the samples from the caller
 * (mask = c1 of get_rlwe_sample, error = the Gaussian coefficients before their NTT). */
int ref_ksk_generate(unsigned logn, size_t L, const u64 *ext_moduli, const u64 *sk_curr, const u64 *sk_orig,
                     const u64 *masks, const u64 *errors, u64 *key) {
    return guarded([&] {
        const size_t n = (size_t)1 << logn, L1 = L + 1;
        const u64 P = ext_moduli[L];
        RlweSk sk_c(load_poly(n, L, ext_moduli, sk_curr, true)), sk_o(load_poly(n, L, ext_moduli, sk_orig, true));
        auto sk_curr_extended = sk_c;                                   // keys.cpp:10-11
        sk_curr_extended.add_components({P});
        std::fill(sk_curr_extended.last()->begin(), sk_curr_extended.last()->end(), (u64)0); // keys.cpp:13-16 (x0 anyway)
        auto sk_orig_extended(sk_o);                                    // keys.cpp:21-26
        intt_negacyclic_inplace_lazy(sk_orig_extended);
        auto extended_part = rns_base_transform(sk_orig_extended, {P});
        sk_orig_extended.add_components({P});
        *sk_orig_extended.last() = std::move(extended_part[0]);
        ntt_negacyclic_inplace_lazy(sk_orig_extended);
        std::vector<u64> mont_consts;                                   // rgsw.cpp:36-45
        for (size_t k = 0; k < L1; k++) mont_consts.push_back(((u64)(-1LL) % ext_moduli[k]) + 1);
        for (size_t p = 0; p < L; p++) {
            std::vector<u64> basis(L1, 0);                              // keys.cpp:28-33
            basis[p] = P % ext_moduli[p];
            auto c1 = load_poly(n, L1, ext_moduli, masks + p * L1 * n, true);      // rlwe.cpp:41
            auto ex = load_poly(n, L1, ext_moduli, errors + p * L1 * n, false);    // sampling.cpp:47-66
            ntt_negacyclic_inplace_lazy(ex);
            auto c0 = ex - c1 * sk_orig_extended;                       // rlwe.cpp:50
            c0 += sk_curr_extended * basis;                             // rgsw.cpp:27
            c0 *= mont_consts;                                          // rgsw.cpp:47-51
            c1 *= mont_consts;
            store_poly(c0, key + (p * 2 + 0) * L1 * n);
            store_poly(c1, key + (p * 2 + 1) * L1 * n);
        }
    });
}

/* the raw prime table, for checking the restated selection rule */
int ref_prime_row(unsigned bits, size_t count, u64 *out) {
    if (bits >= prime_lists.size()) return 0;
    const auto &row = prime_lists[bits];
    size_t m = row.size() < count ? row.size() : count;
    for (size_t i = 0; i < m; i++) out[i] = row[i];
    return (int)m;
}

int ref_ckks_pick_moduli(const unsigned *moduli_bits, size_t L, unsigned additional_bits,
                         u64 *moduli_out, u64 *additional_out) {
    return guarded([&] {
        std::vector<size_t> bits(moduli_bits, moduli_bits + L);
        auto params = ckks::create_params(8, bits, additional_bits, 1.0);
        for (size_t k = 0; k < L; k++) moduli_out[k] = params.moduli[k];
        *additional_out = params.additional_mod;
    });
}

/* timing helper: `rows` independent single-limb transforms back to back, exactly the calls
 * bench/ntt_bm.cpp:17-25 times (twiddle cache already warm after the first row) */
int ref_bench_ntt(unsigned logn, u64 q, u64 *x, size_t rows, int forward) {
    return guarded([&] {
        const size_t n = (size_t)1 << logn;
        for (size_t r = 0; r < rows; r++) {
            if (forward)
                ntt_negacyclic_inplace_lazy(logn, q, x + r * n);
            else
                intt_negacyclic_inplace_lazy(logn, q, x + r * n);
        }
    });
}

/* ---- the sampling-based API, made deterministic by seeding the reference's own global engine ----------------
 * Each scenario runs the reference's public functions end to end and records the FNV-1a hash of
 * every intermediate's raw words, limbs in order.  tests/cpp/test_hehub_api.cpp runs the same statements through
 * the hehub_b200 mirror after seeding ITS engine with the same value. */
void ref_rng_seed(u64 seed) { rand_engine.seed(seed); }

/* 0 ternary (NTT form), 1 uniform, 2 Gaussian (NTT form) — sampling.cpp:16-93 */
int ref_rng_sample(int kind, unsigned logn, size_t L, const u64 *moduli, u64 *out) {
    return guarded([&] {
        RnsPolyParams params{(size_t)1 << logn, L, std::vector<u64>(moduli, moduli + L)};
        auto p = kind == 0 ? get_rand_ternary_poly(params) : kind == 1 ? get_rand_uniform_poly(params) : get_rand_gaussian_poly(params);
        store_poly(p, out);
    });
}

/* CKKS flow: keygen, relin / conj / rot keys, encrypt, mult, rescale, rotate, conjugate, add/sub/mult_plain, decrypt.
 * hashes[14]. */
int ref_rng_scenario_ckks(u64 seed, unsigned logn, size_t L, const unsigned *moduli_bits, unsigned additional_bits, u64 *hashes) {
    return guarded([&] {
        rand_engine.seed(seed);
        const size_t n = (size_t)1 << logn;
        auto params = ckks::create_params(n, std::vector<size_t>(moduli_bits, moduli_bits + L), additional_bits, 1099511627776.0);
        RlweSk sk(params);
        auto relin = get_relin_key(sk, params.additional_mod);
        auto conj = get_conj_key(sk, params.additional_mod);
        auto rot = get_rot_key(sk, params.additional_mod, 3);
        size_t j = 0;
        hashes[j++] = fnv_poly(sk);
        hashes[j++] = fnv_ksk(relin);
        hashes[j++] = fnv_ksk(conj);
        hashes[j++] = fnv_ksk(rot);
        CkksPt pt1(lcg_poly(n, params.moduli, 11, 1 << 20)), pt2(lcg_poly(n, params.moduli, 12, 1 << 20));
        pt1.scaling_factor = pt2.scaling_factor = params.initial_scaling_factor;
        auto ct1 = ckks::encrypt(pt1, sk), ct2 = ckks::encrypt(pt2, sk);
        hashes[j++] = fnv_ct(ct1);
        hashes[j++] = fnv_ct(ct2);
        hashes[j++] = fnv_ct(ckks::add_plain(ct1, pt2));
        hashes[j++] = fnv_ct(ckks::sub_plain(ct1, pt2));
        hashes[j++] = fnv_ct(ckks::mult_plain(ct1, pt2));
        auto prod = ckks::mult(ct1, ct2, relin);
        hashes[j++] = fnv_ct(prod);
        ckks::rescale_inplace(prod);
        hashes[j++] = fnv_ct(prod);
        hashes[j++] = fnv_ct(ckks::rotate(ct1, rot));
        hashes[j++] = fnv_ct(ckks::conjugate(ct2, conj));
        hashes[j++] = fnv_poly(ckks::decrypt(prod, sk)); // rescaled (L-1 limbs) ciphertext against the full-length key
    });
}

/* BGV flow: slot encoding, encrypt, plain ops, mult + relinearize, mod-switch, decrypt, decode.  hashes[12];
 * decoded[n] receives simd_decode(decrypt(ct1 * pt2 + ct2)) (the shape of tests/bgv_t.cpp:160-190).  Decryption is
 * only asked of ciphertexts whose noise keeps rns_base_transform on its small-coefficient path (rns_transform.cpp:47-84). */
int ref_rng_scenario_bgv(u64 seed, unsigned logn, size_t L, const unsigned *moduli_bits, unsigned additional_bits, u64 t, u64 *hashes,
                         u64 *decoded) {
    return guarded([&] {
        rand_engine.seed(seed);
        const size_t n = (size_t)1 << logn;
        auto params = ckks::create_params(n, std::vector<size_t>(moduli_bits, moduli_bits + L), additional_bits, 1.0);
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
        size_t j = 0;
        auto pt1 = bgv::simd_encode(d1, t, n), pt2 = bgv::simd_encode(d2, t, n);
        hashes[j++] = fnv_poly(pt1);
        auto ct1 = bgv::encrypt(pt1, sk), ct2 = bgv::encrypt(pt2, sk);
        hashes[j++] = fnv_ct(ct1);
        hashes[j++] = fnv_ct(ct2);
        hashes[j++] = fnv_ct(bgv::add_plain(ct1, pt2));
        hashes[j++] = fnv_ct(bgv::sub_plain(ct1, pt2));
        auto ct_prod_plain = bgv::mult_plain(ct1, pt2);
        hashes[j++] = fnv_ct(ct_prod_plain);
        auto ct_res = bgv::add(ct_prod_plain, ct2);
        hashes[j++] = fnv_ct(ct_res);
        auto prod = bgv::relinearize(bgv::mult_low_level(ct1, ct2), relin);
        hashes[j++] = fnv_ct(prod);
        bgv::mod_switch_inplace(prod);
        hashes[j++] = fnv_ct(prod);
        auto dec = bgv::decrypt(ct_res, sk);
        hashes[j++] = fnv_poly(dec);
        auto out = bgv::simd_decode(dec);
        std::memcpy(decoded, out.data(), n * sizeof(u64));
        auto switched = ct1; // tests/bgv_t.cpp:229-259: the plaintext survives the switch
        bgv::mod_switch_inplace(switched);
        hashes[j++] = fnv_ct(switched);
        hashes[j++] = fnv_poly(bgv::decrypt(switched, sk));
    });
}

/* CKKS encoder / decoder (host-side floating point, ckks/basics.cpp:156-369): `count` complex data from an LCG in
 * [-1, 1) x [-1, 1), encoded at scaling 2^log2_scaling; hash[0] = FNV of the plaintext words; decoded[2 * slots] = (re, im) of
 * simd_decode of that plaintext. */
int ref_ckks_codec_scenario(unsigned logn, size_t L, const unsigned *moduli_bits, unsigned additional_bits, double log2_scaling,
                            u64 seed, size_t count, u64 *hash, double *decoded) {
    return guarded([&] {
        const size_t n = (size_t)1 << logn;
        auto params = ckks::create_params(n, std::vector<size_t>(moduli_bits, moduli_bits + L), additional_bits, std::pow(2.0, log2_scaling));
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
        hash[0] = fnv_poly(pt);
        auto back = ckks::simd_decode<cc_double>(pt);
        for (size_t i = 0; i < back.size(); i++) {
            decoded[2 * i] = back[i].real();
            decoded[2 * i + 1] = back[i].imag();
        }
    });
}

/* The reference's own application benchmark (bench/benchmarks.cpp, bench/ckks_bm.cpp; README table) timed on this host:
 * ckks::create_params(N, scaling_bits); out_us[0..3] = microseconds per call of encode + encrypt, decrypt + decode, rotate by one
 * slot, mult (tensor + relinearize); out_us[4] = RNS components chosen.  One thread, like the reference. */
int ref_api_bench(unsigned logn, unsigned scaling_bits, int reps, double *out_us) {
    return guarded([&] {
        const size_t n = (size_t)1 << logn;
        auto params = ckks::create_params(n, (size_t)scaling_bits);
        cache_ntt_factors_strict(logn, params.moduli);
        CkksSk sk(params);
        auto rot_key = get_rot_key(sk, params.additional_mod, 1);
        auto relin_key = get_relin_key(sk, params.additional_mod);
        std::vector<cc_double> data(n / 2);
        for (size_t i = 0; i < data.size(); i++) data[i] = cc_double(0.001 * (double)(i % 997), -0.002 * (double)(i % 499));
        auto ct = ckks::encrypt(ckks::simd_encode(data, params), sk);
        auto time_us = [&](auto f) {
            f();
            const auto t0 = std::chrono::steady_clock::now();
            for (int i = 0; i < reps; i++) f();
            return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / reps;
        };
        out_us[0] = time_us([&] { auto c = ckks::encrypt(ckks::simd_encode(data, params), sk); });
        out_us[1] = time_us([&] { auto d = ckks::simd_decode<cc_double>(ckks::decrypt(ct, sk)); });
        out_us[2] = time_us([&] { auto r = ckks::rotate(ct, rot_key, 1); });
        out_us[3] = time_us([&] { auto m = ckks::mult(ct, ct, relin_key); });
        out_us[4] = (double)params.moduli.size();
    });
}

void ref_cache_ntt_factors(unsigned logn, const u64 *moduli, size_t count) {
    cache_ntt_factors_strict(logn, std::vector<u64>(moduli, moduli + count));
}

} // extern "C"
