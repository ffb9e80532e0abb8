// bgv.h — hehub::bgv on the B200 back end (src/fhe/bgv/bgv.h:18-177, basics.cpp, arith.cpp, mod_switch.cpp):
// slot encoding over the plain modulus, encrypt / decrypt and every ciphertext operator.
#pragma once
#include <algorithm>
#include <cmath>
#include <string>

#include "ckks.h"
#include "rns_transform.h"

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

/// basics.cpp:11-42 — data into the slots (NTT values modulo the plain modulus), then to coefficient form
inline RlwePt simd_encode(const std::vector<u64> &data, const u64 modulus, size_t slot_count = 0) {
    for (auto datum : data)
        if (datum >= modulus) throw std::invalid_argument("Data not being valid Z_p elements with p = " + std::to_string(modulus) + ".");
    if (slot_count == 0) slot_count = (size_t)1 << (size_t)std::ceil(std::log2(data.size()));
    const auto data_size = data.size();
    if (data_size > slot_count)
        throw std::invalid_argument("Cannot encode " + std::to_string(data_size) + " data into " + std::to_string(slot_count) + " slots.");
    RlwePt pt(RnsPolyParams{slot_count, 1, std::vector<u64>{modulus}});
    pt.rep_form = PolyRepForm::value;
    u64 *words = pt[0].data();
    std::copy(data.begin(), data.end(), words);
    std::fill(words + data_size, words + slot_count, (u64)0);
    intt_negacyclic_inplace_lazy(pt);
    return pt;
}

/// basics.cpp:44-61
inline std::vector<u64> simd_decode(const RlwePt &pt, size_t data_size = 0) {
    if (data_size == 0) data_size = pt.dimension();
    if (pt.component_count() != 1)
        throw std::invalid_argument("Plaintext is in RNS with multiple moduli. Use big int version of decoding.");
    auto pt_copy(pt);
    ntt_negacyclic_inplace_lazy(pt_copy);
    reduce_strict(pt_copy);
    std::vector<u64> data(pt_copy[0].begin(), pt_copy[0].end());
    data.resize(data_size);
    return data;
}

/// basics.cpp:63-79 — an RLWE sample whose both halves are multiplied by the plain modulus
inline RlweCt get_rlwe_sample_lift_noise(const RlweSk &sk, const u64 lifting_factor, size_t components = 0) {
    if (components == 0) components = sk.component_count();
    auto rlwe_sample = get_rlwe_sample(sk, components);
    for (auto &rns_poly : rlwe_sample) rns_poly *= lifting_factor;
    return rlwe_sample;
}

inline BgvCt add(const BgvCt &ct1, const BgvCt &ct2) { // bgv/arith.cpp:9-15
    if (ct1.plain_modulus != ct2.plain_modulus) throw std::invalid_argument("Plain moduli mismatch.");
    BgvCt sum_ct = ::hehub::add(ct1, ct2);
    sum_ct.plain_modulus = ct1.plain_modulus;
    return sum_ct;
}
namespace detail {
/// the plaintext under the ciphertext's moduli, in NTT form (bgv/arith.cpp:21-22, 41-42, 52-53)
inline RnsPolynomial plain_under_ct_moduli(const BgvCt &ct, const BgvPt &pt) {
    if (pt.component_count() != 1 || pt.modulus_at(0) != ct.plain_modulus) throw std::invalid_argument("plain moduli mismatch.");
    auto pt_under_ct_mod = rns_base_transform(pt, ct[0].modulus_vec());
    ntt_negacyclic_inplace_lazy(pt_under_ct_mod);
    return pt_under_ct_mod;
}
} // namespace detail

inline BgvCt add_plain(const BgvCt &ct, const BgvPt &pt) { // bgv/arith.cpp:17-27
    BgvCt sum_ct = ::hehub::add_plain_core(ct, detail::plain_under_ct_moduli(ct, pt));
    sum_ct.plain_modulus = ct.plain_modulus;
    return sum_ct;
}
inline BgvCt sub_plain(const BgvCt &ct, const BgvPt &pt) { // bgv/arith.cpp:37-47
    BgvCt diff_ct = ::hehub::sub_plain_core(ct, detail::plain_under_ct_moduli(ct, pt));
    diff_ct.plain_modulus = ct.plain_modulus;
    return diff_ct;
}
inline BgvCt mult_plain(const BgvCt &ct, const BgvPt &pt) { // bgv/arith.cpp:49-57
    BgvCt prod_ct = ::hehub::mult_plain_core(ct, detail::plain_under_ct_moduli(ct, pt));
    prod_ct.plain_modulus = ct.plain_modulus;
    return prod_ct;
}

inline BgvCt sub(const BgvCt &ct1, const BgvCt &ct2) { // bgv/arith.cpp:29-35
    if (ct1.plain_modulus != ct2.plain_modulus) throw std::invalid_argument("Plain moduli mismatch.");
    BgvCt diff_ct = ::hehub::sub(ct1, ct2);
    diff_ct.plain_modulus = ct1.plain_modulus;
    return diff_ct;
}

/// basics.cpp:81-108 — noise lifted by the plain modulus, plaintext migrated under the ciphertext moduli
inline BgvCt encrypt(const RlwePt &pt, const RlweSk &rlwe_sk, std::vector<u64> ct_moduli = std::vector<u64>{}) {
    const auto pt_modulus = pt.modulus_at(0);
    if (ct_moduli.empty()) ct_moduli = rlwe_sk.modulus_vec();
    if (std::find(ct_moduli.begin(), ct_moduli.end(), pt_modulus) != ct_moduli.end())
        throw std::logic_error("Plaintext modulus needs to be coprime with ciphertext modulus (i.e. cannot belong to ct_moduli).");
    auto sample = get_rlwe_sample_lift_noise(rlwe_sk, pt_modulus, ct_moduli.size());
    auto pt_under_ct_mod = rns_base_transform(pt, ct_moduli);
    ntt_negacyclic_inplace_lazy(pt_under_ct_mod);
    sample[0] += pt_under_ct_mod;
    BgvCt ct(std::move(sample));
    ct.plain_modulus = pt_modulus;
    return ct;
}

/// basics.cpp:110-117 — RLWE decryption, then back under the plain modulus
inline BgvPt decrypt(const BgvCt &ct, const RlweSk &rlwe_sk) {
    auto pt_under_ct_mod = ::hehub::decrypt_core(ct, rlwe_sk);
    return rns_base_transform(pt_under_ct_mod, std::vector<u64>{ct.plain_modulus});
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
