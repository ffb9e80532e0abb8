// bgv.h — the ciphertext-op part of hehub::bgv (src/fhe/bgv/bgv.h:24-177, arith.cpp, mod_switch.cpp).
#pragma once
#include "ckks.h"

namespace hehub {
namespace bgv {

using BgvPt = RlwePt;

struct BgvCt : public RlweCt {
    using RlweCt::RlweCt;
    BgvCt() {}
    BgvCt(RlweCt &&other) : RlweCt(std::move(other)) {}
    u64 plain_modulus = 1;
};

struct BgvQuadraticCt : public std::array<RnsPolynomial, 3> {
    using std::array<RnsPolynomial, 3>::array;
    u64 plain_modulus = 1;
};

inline BgvCt add(const BgvCt &ct1, const BgvCt &ct2) { // bgv/arith.cpp:9-15
    if (ct1.plain_modulus != ct2.plain_modulus) throw std::invalid_argument("Plain moduli mismatch.");
    BgvCt sum_ct = ::hehub::add(ct1, ct2);
    sum_ct.plain_modulus = ct1.plain_modulus;
    return sum_ct;
}
inline BgvCt sub(const BgvCt &ct1, const BgvCt &ct2) { // bgv/arith.cpp:29-35
    if (ct1.plain_modulus != ct2.plain_modulus) throw std::invalid_argument("Plain moduli mismatch.");
    BgvCt diff_ct = ::hehub::sub(ct1, ct2);
    diff_ct.plain_modulus = ct1.plain_modulus;
    return diff_ct;
}

/// bgv/arith.cpp:59-69 — same tensor product kernel as CKKS
inline BgvQuadraticCt mult_low_level(const BgvCt &ct1, const BgvCt &ct2) {
    if (ct1.plain_modulus != ct2.plain_modulus) throw std::invalid_argument("Plain moduli mismatch.");
    ckks::CkksCt a(RlweCt{ct1[0], ct1[1]}), b(RlweCt{ct2[0], ct2[1]});
    auto q = ckks::mult_low_level(a, b);
    BgvQuadraticCt prod;
    for (size_t h = 0; h < 3; h++) prod[h] = std::move(q[h]);
    prod.plain_modulus = ct1.plain_modulus;
    return prod;
}

/// mod_switch.cpp:80-90 → :13-78
inline void mod_switch_inplace(BgvCt &ct, size_t dropping_primes = 1) {
    if (dropping_primes >= 2) throw "under development";
    if (dropping_primes != 1) throw std::invalid_argument("The number of primes to be dropped is not positive.");
    ckks::detail::check_ct(ct);
    if (ct[0].component_count() == 1) throw std::invalid_argument("Unable to drop the only one prime.");
    const auto params = ct[0].params();
    const size_t L = params.component_count, n = params.dimension;
    ::hehub::detail::Staged out(2 * (L - 1) * n);
    const auto in = ckks::detail::gather<2>(ct);
    b200::check(hehub_b200_bgv_mod_switch(b200::context(), (unsigned)ct[0].log_dimension(), params.moduli.data(), L, ct.plain_modulus,
                                          in.dev, out.dev, 1));
    RnsPolyParams dropped{n, L - 1, std::vector<u64>(params.moduli.begin(), params.moduli.end() - 1)};
    static_cast<RlweCt &>(ct) = ckks::detail::scatter<2>(out, dropped);
}

/// bgv/arith.cpp:71-79 — the internal mod-switch runs with the default plain modulus 1, as in the reference
inline BgvCt relinearize(const BgvQuadraticCt &ct, const RlweKsk &relin_key) {
    const auto &key = relin_key.packed(ct[2]);
    const auto params = ct[0].params();
    const size_t L = params.component_count, words = L * params.dimension;
    ::hehub::detail::Staged out(2 * words);
    const auto q = ckks::detail::gather<3>(ct);
    b200::check(hehub_b200_bgv_relinearize(b200::context(), (unsigned)ct[0].log_dimension(), key.ext_moduli.data(), L, 1, q.dev, key.dev,
                                           out.dev, 1));
    BgvCt ct_new(ckks::detail::scatter<2>(out, params));
    ct_new.plain_modulus = ct.plain_modulus;
    return ct_new;
}

} // namespace bgv

using BgvPt = bgv::BgvPt;
using BgvCt = bgv::BgvCt;

} // namespace hehub
