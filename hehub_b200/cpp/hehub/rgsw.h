// rgsw.h / keys.h — key-switch key container and the gadget-decomposition inner product
// (src/fhe/primitives/rgsw.{h,cpp}:57-156, keys.h:19-32).
#pragma once
#include <memory>
#include <vector>

#include "permutation.h"
#include "rlwe.h"

namespace hehub {

using RgswCt = std::vector<RlweCt>;

/// rgsw.cpp:11-33 — one fresh RLWE sample per basis row, plus pt_ntt scaled by that row's RNS constants
inline RgswCt rgsw_encrypt(const RlwePt &pt_ntt, const RlweSk &sk, const std::vector<std::vector<u64>> &decomp_basis) {
    if (pt_ntt.rep_form == PolyRepForm::coeff) throw std::invalid_argument("Plaintext is expected in NTT form.");
    RgswCt rgsw(decomp_basis.size());
    for (auto &rlwe_sample : rgsw) rlwe_sample = get_rlwe_sample(sk);
    for (size_t r = 0; r < rgsw.size(); r++) rgsw[r][0] += pt_ntt * decomp_basis[r];
    return rgsw;
}

/// rgsw.cpp:35-55 — the same, every polynomial then multiplied by 2^64 mod q_k (Montgomery form for ext_prod_montgomery)
inline RgswCt rgsw_encrypt_montgomery(const RlwePt &pt_ntt, const RlweSk &sk, const std::vector<std::vector<u64>> &decomp_basis) {
    auto rgsw = rgsw_encrypt(pt_ntt, sk, decomp_basis);
    if (rgsw.empty()) return rgsw;
    std::vector<u64> mont_consts;
    for (u64 modulus : rgsw[0][0].modulus_vec()) mont_consts.push_back((u64)(-1LL) % modulus + 1);
    for (auto &rlwe_sample : rgsw)
        for (auto &poly : rlwe_sample) poly *= mont_consts;
    return rgsw;
}

/// RlweKsk = vector<RlweCt> (keys.h:19-32).  The device kernels read the key as ONE slab laid out
/// [row p][half][limb k <= L][N]; it is packed from the rows on first use and cached in the object
/// (keys are reused by every relinearize / rotate).
struct RlweKsk : public RgswCt {
    using RgswCt::RgswCt;
    RlweKsk() {}
    RlweKsk(RgswCt &&rgsw) : RgswCt(std::move(rgsw)) {}

    /// keys.cpp:8-36 on the device, with the RLWE samples supplied by the caller: the reference draws
    /// them from a process-global RNG inside rgsw_encrypt (rgsw.cpp:21-23), which this back end does
    /// not replicate.  masks[p]: uniform, NTT form; errors[p]: small coefficients, coefficient form;
    /// both over (q_0..q_{L-1}, additional_mod), one per row p < L.
    RlweKsk(const RlweSk &sk_curr, const RlweSk &sk_orig, const u64 additional_mod, const std::vector<RnsPolynomial> &masks,
            const std::vector<RnsPolynomial> &errors) {
        const size_t L = sk_orig.component_count(), n = sk_orig.dimension();
        if (masks.size() != L || errors.size() != L) throw std::invalid_argument("One RLWE sample per RNS component is needed.");
        auto ext = sk_orig.modulus_vec();
        ext.push_back(additional_mod);
        const size_t poly_words = (L + 1) * n;
        detail::Staged m(L * poly_words), e(L * poly_words), key(L * 2 * poly_words);
        for (size_t p = 0; p < L; p++) {
            if (masks[p].modulus_vec() != ext || errors[p].modulus_vec() != ext || masks[p].dimension() != n || errors[p].dimension() != n)
                throw std::invalid_argument("Samples do not match the extended modulus chain.");
            b200::check(hehub_b200_slab_d2d(b200::context(), m.dev + p * poly_words, masks[p].dev(), poly_words));
            b200::check(hehub_b200_slab_d2d(b200::context(), e.dev + p * poly_words, errors[p].dev(), poly_words));
        }
        b200::check(hehub_b200_ksk_generate(b200::context(), (unsigned)sk_orig.log_dimension(), ext.data(), L, sk_curr.dev(), sk_orig.dev(),
                                            m.dev, e.dev, key.dev));
        RnsPolyParams params{n, L + 1, ext};
        for (size_t p = 0; p < L; p++) {
            RlweCt row{RnsPolynomial(params), RnsPolynomial(params)};
            for (size_t h = 0; h < 2; h++) {
                b200::check(hehub_b200_slab_d2d(b200::context(), row[h].dev_mut(), key.dev + (p * 2 + h) * poly_words, poly_words));
                row[h].rep_form = PolyRepForm::value;
            }
            push_back(std::move(row));
        }
    }

    /// keys.cpp:8-36, the reference's signature: the L RLWE samples are drawn here on the host in the reference's order
    /// (per row: uniform mask over (q_0..q_{L-1}, P), then Gaussian error), then one fused device call builds the key.
    RlweKsk(const RlweSk &sk_curr, const RlweSk &sk_orig, const u64 additional_mod)
        : RlweKsk(sk_curr, sk_orig, additional_mod, Samples(sk_orig, additional_mod)) {}

    struct Packed {
        u64 *dev = nullptr;
        size_t rows = 0, limbs = 0, dimension = 0;
        std::vector<u64> ext_moduli;
        ~Packed() {
            if (dev) hehub_b200_slab_free(b200::context(), dev);
        }
    };
    /// validates the shape like rgsw.cpp:59-89 and returns the packed device copy
    const Packed &packed(const RlwePt &pt) const {
        if (empty()) throw std::invalid_argument("Empty RGSW ciphertext.");
        const size_t L = pt.component_count(), n = pt.dimension();
        auto ext = (*this)[0][0].modulus_vec();
        if (ext.size() < L + 1) throw std::invalid_argument("Invalid component number in RGSW ciphertext.");
        const u64 last = ext.back();
        ext.resize(L + 1);
        ext.back() = last;
        for (size_t k = 0; k < L; k++)
            if (ext[k] != pt.modulus_at((int)k)) throw std::invalid_argument("Moduli mismatch.");
        for (const auto &sample : *this)
            for (const auto &poly : sample) {
                if (poly.dimension() != n) throw std::invalid_argument("Polynomial lengths mismatch.");
                if (poly.component_count() != L + 1 || poly.modulus_vec() != ext) throw std::invalid_argument("Inconsistent RGSW ciphertext.");
            }
        if (size() < L) throw std::invalid_argument("Inconsistent RGSW ciphertext."); // one row per digit
        if (cache_ && cache_->rows == L && cache_->dimension == n && cache_->ext_moduli == ext) return *cache_;
        auto p = std::make_shared<Packed>();
        p->rows = L;
        p->limbs = L + 1;
        p->dimension = n;
        p->ext_moduli = ext;
        const size_t poly_words = (L + 1) * n;
        b200::check(hehub_b200_slab_alloc(b200::context(), L * 2 * poly_words, &p->dev));
        for (size_t r = 0; r < L; r++)
            for (size_t h = 0; h < 2; h++)
                b200::check(hehub_b200_slab_d2d(b200::context(), p->dev + (r * 2 + h) * poly_words, (*this)[r][h].dev(), poly_words));
        cache_ = p;
        return *cache_;
    }
    /// call after mutating the rows in place
    void invalidate_packed() { cache_.reset(); }

private:
    mutable std::shared_ptr<Packed> cache_;
    struct Samples {
        std::vector<RnsPolynomial> masks, errors;
        Samples(const RlweSk &sk_orig, const u64 additional_mod) {
            auto ext = sk_orig.modulus_vec();
            ext.push_back(additional_mod);
            RlweParams params{sk_orig.dimension(), ext.size(), ext};
            for (size_t p = 0; p + 1 < ext.size(); p++) { // rgsw.cpp:21-23 through rlwe.cpp:34-50
#ifdef HEHUB_DEBUG_RLWE_ZERO_C1
                masks.push_back(get_zero_poly(params));
#else
                masks.push_back(get_rand_uniform_poly(params, PolyRepForm::value));
#endif
#ifdef HEHUB_DEBUG_RLWE_ZERO_E
                errors.push_back(get_zero_poly(params, PolyRepForm::coeff));
#else
                errors.push_back(detail::gaussian_coeffs(params, 3.2));
#endif
            }
        }
    };
    RlweKsk(const RlweSk &sk_curr, const RlweSk &sk_orig, const u64 additional_mod, const Samples &s)
        : RlweKsk(sk_curr, sk_orig, additional_mod, s.masks, s.errors) {}
};

/// keys.h:42-44 — RGSW encryption of sk^2 under sk
inline RlweKsk get_relin_key(const RlweSk &sk, const u64 additional_mod) {
    return RlweKsk(RlweSk(static_cast<const RnsPolynomial &>(sk) * static_cast<const RnsPolynomial &>(sk)), sk, additional_mod);
}
/// keys.h:54-56 — RGSW encryption of sk(X^-1) under sk
inline RlweKsk get_conj_key(const RlweSk &sk, const u64 additional_mod) { return RlweKsk(RlweSk(involution(sk)), sk, additional_mod); }

/// keys.h:63-67
struct RotKey : public RlweKsk {
    using RlweKsk::RlweKsk;
    RotKey() {}
    RotKey(RlweKsk &&ksk) : RlweKsk(std::move(ksk)) {}
    size_t step = 0;
};
/// keys.h:78-83 — RGSW encryption of sk(X^(3^step)) under sk
inline RotKey get_rot_key(const RlweSk &sk, const u64 additional_mod, const size_t step) {
    RotKey rot_key(RlweKsk(RlweSk(cycle(sk, step)), sk, additional_mod));
    rot_key.step = step;
    return rot_key;
}

/// rgsw.cpp:57-156 — RNS-digit decomposition of `pt` (NTT form) and inner product with the key,
/// accumulated in 128 bits and Montgomery-reduced once; result over (q_0..q_{L-1}, P), value form.
inline RlweCt ext_prod_montgomery(const RlwePt &pt, const RlweKsk &rgsw) {
    const auto &key = rgsw.packed(pt);
    const size_t L = pt.component_count();
    RnsPolyParams ext{pt.dimension(), L + 1, key.ext_moduli};
    // one slab [2][L+1][N] from the kernel, split into the two polynomials of the result
    detail::Staged out(2 * (L + 1) * pt.dimension());
    b200::check(hehub_b200_ext_prod_montgomery(b200::context(), (unsigned)pt.log_dimension(), key.ext_moduli.data(), L, pt.dev(),
                                               key.dev, out.dev, 1));
    RlweCt ct; // the two halves adopt the slab: no copy, and they stay contiguous for the rescale that follows
    for (size_t h = 0; h < 2; h++) {
        ct[h] = RnsPolynomial(RnsIntVec::adopt(out.block, out.dev + h * (L + 1) * pt.dimension(), ext));
        ct[h].rep_form = PolyRepForm::value;
    }
    return ct;
}
inline RlweCt ext_prod_montgomery(const RlwePt &pt, const RgswCt &rgsw) { return ext_prod_montgomery(pt, RlweKsk(RgswCt(rgsw))); }

} // namespace hehub
