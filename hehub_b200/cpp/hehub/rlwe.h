// rlwe.h — RLWE value types and the coefficient-wise ciphertext operators of hehub
// (src/fhe/primitives/rlwe.{h,cpp}).  Sampling / encryption stay on the host side of the reference
// (std::default_random_engine, sampling.cpp:12-14) and are outside this back end's path.
#pragma once
#include <array>

#include "ntt.h"

namespace hehub {

using RlweParams = RnsPolynomial::Params;
using RlwePt = RnsPolynomial;
using RlweCt = std::array<RnsPolynomial, 2>;

/// Secret key container (rlwe.h:34-50); key sampling itself is host-side reference code.
struct RlweSk : public RnsPolynomial {
    using RnsPolynomial::RnsPolynomial;
    RlweSk() {}
    RlweSk(RnsPolynomial &&rns_poly) : RnsPolynomial(std::move(rns_poly)) {}
    RlweSk(const RnsPolynomial &rns_poly) : RnsPolynomial(rns_poly) {}
};

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
