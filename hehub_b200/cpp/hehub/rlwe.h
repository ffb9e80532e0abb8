// rlwe.h — RLWE value types, parameter construction, sampling-based encryption and the coefficient-wise
// ciphertext operators of hehub (src/fhe/primitives/rlwe.{h,cpp}).  Random draws happen on the host with the
// reference's engine and distributions (sampling.h); every transform and product runs on the device.
#pragma once
#include <array>

#include "ntt.h"
#include "primelists.h"
#include "sampling.h"

namespace hehub {

using RlweParams = RnsPolynomial::Params;
using RlwePt = RnsPolynomial;
using RlweCt = std::array<RnsPolynomial, 2>;

/// rlwe.cpp:9-29 — one prime per requested bit size, drawn in order from prime_lists with one cursor per row
inline RlweParams create_params(size_t dimension, std::vector<int> moduli_bits) {
    RlweParams params;
    params.dimension = dimension;
    std::vector<unsigned> bits(moduli_bits.begin(), moduli_bits.end());
    params.moduli.assign(bits.size(), 0);
    if (hehub_b200_pick_moduli(bits.data(), bits.size(), 0, params.moduli.data(), nullptr) != HEHUB_B200_OK)
        throw "No suitable primes in the library.";
    params.component_count = params.moduli.size();
    return params;
}

/// Secret key (rlwe.h:34-50): ternary coefficients, kept in NTT form.
struct RlweSk : public RnsPolynomial {
    using RnsPolynomial::RnsPolynomial;
    RlweSk() {}
    RlweSk(RnsPolynomial &&rns_poly) : RnsPolynomial(std::move(rns_poly)) {}
    RlweSk(const RnsPolynomial &rns_poly) : RnsPolynomial(rns_poly) {}
    /// rlwe.cpp:31-32 — sample the key
    RlweSk(const RlweParams &params) : RnsPolynomial(get_rand_ternary_poly(params)) {}
};

/// rlwe.cpp:34-50 — (e - c1 * sk, c1) over the first `components` moduli of the key (0: all of them), NTT form.
/// HEHUB_DEBUG_RLWE_ZERO_C1 / HEHUB_DEBUG_RLWE_ZERO_E switch a draw off, like the reference's macros.
inline RlweCt get_rlwe_sample(const RlweSk &sk, size_t components = 0) {
    if (components == 0) components = sk.component_count();
    RlweParams params{sk.dimension(), components, sk.modulus_vec()};
#ifdef HEHUB_DEBUG_RLWE_ZERO_C1
    auto c1 = get_zero_poly(params);
#else
    auto c1 = get_rand_uniform_poly(params, PolyRepForm::value);
#endif
#ifdef HEHUB_DEBUG_RLWE_ZERO_E
    auto ex = get_zero_poly(params);
#else
    auto ex = get_rand_gaussian_poly(params);
#endif
    auto c0 = ex - c1 * static_cast<const RnsPolynomial &>(sk);
    return RlweCt{std::move(c0), std::move(c1)};
}

inline RlweCt add(const RlweCt &ct1, const RlweCt &ct2) { return RlweCt{ct1[0] + ct2[0], ct1[1] + ct2[1]}; }        // rlwe.cpp:83-85
inline RlweCt add_plain_core(const RlweCt &ct, const RlwePt &pt) { return RlweCt{ct[0] + pt, ct[1]}; }               // rlwe.cpp:87-89
inline RlweCt sub(const RlweCt &ct1, const RlweCt &ct2) { return RlweCt{ct1[0] - ct2[0], ct1[1] - ct2[1]}; }        // rlwe.cpp:91-93
inline RlweCt sub_plain_core(const RlweCt &ct, const RlwePt &pt) { return RlweCt{ct[0] - pt, ct[1]}; }               // rlwe.cpp:95-97
inline RlweCt mult_plain_core(const RlweCt &ct, const RlwePt &pt) { return RlweCt{ct[0] * pt, ct[1] * pt}; }         // rlwe.cpp:99-101

/// rlwe.cpp:72-81 — c0 + c1 * sk, INTT, strict reduction (all on the device)
inline RlwePt decrypt_core(const RlweCt &ct, const RlweSk &sk) {
    auto pt = ct[0] + ct[1] * static_cast<const RnsPolynomial &>(sk);
    intt_negacyclic_inplace(pt);
    return pt;
}

/// rlwe.cpp:52-71 — the reference's signature: mask and error are drawn here (host), in the reference's order (mask
/// first), then ONE fused device call does NTT(e) - c1 * sk + NTT(pt).
inline RlweCt encrypt_core(const RlwePt &pt, const RlweSk &sk) {
    if (pt.rep_form == PolyRepForm::value) throw std::invalid_argument("Plaintext not in coeff representation.");
    const size_t L = pt.component_count(), n = pt.dimension();
    RlweParams params{n, L, pt.modulus_vec()};
    if (sk.dimension() != n) throw std::invalid_argument("Operands' poly len mismatch.");
    if (sk.component_count() < L) throw std::invalid_argument("Operand b contains less components than self."); // via rns.cpp:58-72
    for (size_t k = 0; k < L; k++)
        if (sk.modulus_at((int)k) != pt.modulus_at((int)k)) throw std::invalid_argument("Operands' moduli mismatch.");
#ifdef HEHUB_DEBUG_RLWE_ZERO_C1
    auto c1 = get_zero_poly(params);
#else
    auto c1 = get_rand_uniform_poly(params, PolyRepForm::value);
#endif
#ifdef HEHUB_DEBUG_RLWE_ZERO_E
    auto e = get_zero_poly(params, PolyRepForm::coeff);
#else
    auto e = detail::gaussian_coeffs(params, 3.2);
#endif
    detail::Staged out(2 * L * n);
    b200::check(hehub_b200_rlwe_encrypt_core(b200::context(), (unsigned)pt.log_dimension(), params.moduli.data(), L, pt.dev(), sk.dev(),
                                             c1.dev(), e.dev(), out.dev, 1));
    RlweCt ct;
    for (size_t h = 0; h < 2; h++) {
        ct[h] = RnsPolynomial(RnsIntVec::adopt(out.block, out.dev + h * L * n, params));
        ct[h].rep_form = PolyRepForm::value;
    }
    return ct;
}

/// rlwe.cpp:34-61 with the samples supplied by the caller: the reference draws `mask` (uniform, NTT
/// form) and `error` (small coefficients, coefficient form) from a process-global RNG
/// (sampling.cpp:12-69), which this back end does not replicate.  Same statements, same operators.
inline RlweCt encrypt_core(const RlwePt &pt, const RlweSk &sk, const RnsPolynomial &mask, RnsPolynomial error) {
    if (pt.rep_form == PolyRepForm::value) throw std::invalid_argument("Plaintext not in coeff representation."); // rlwe.cpp:51-53
    ntt_negacyclic_inplace_lazy(error);                                   // sampling.cpp:66
    auto c0 = error - mask * static_cast<const RnsPolynomial &>(sk);      // rlwe.cpp:50
    auto pt_ntt(pt);
    ntt_negacyclic_inplace_lazy(pt_ntt);                                  // rlwe.cpp:54-55
    c0 += pt_ntt;                                                         // rlwe.cpp:58
    return RlweCt{std::move(c0), mask};
}

} // namespace hehub
