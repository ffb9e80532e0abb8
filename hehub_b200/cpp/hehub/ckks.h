// ckks.h — hehub::ckks on the B200 back end (src/fhe/ckks/ckks.h:19-329, basics.cpp:14-64, arith.cpp,
// rescaling.cpp): parameter selection, encrypt / decrypt and every ciphertext operator.  The encoder (complex fp64
// FFT, basics.cpp:68-369) is host-side code outside the integer hot path; it lives in
// ckks_encoding.h, restated so that it hands the integer path the same words as the reference.
#pragma once
#include <cmath>
#include <map>
#include <memory>

#include "permutation.h"
#include "rgsw.h"

namespace hehub {
namespace ckks {

/// ckks.h:19-28
struct CkksParams : public RlweParams {
    CkksParams() {}
    CkksParams(RlweParams &&other) : RlweParams(std::move(other)) {}
    u64 additional_mod = 1;
    double initial_scaling_factor = 1.0;
};

/// basics.cpp:14-38 — the additional modulus is drawn FIRST, then the chain, one cursor per bit-size row
inline CkksParams create_params(size_t dimension, std::vector<size_t> moduli_bits, size_t additional_mod_bits,
                                double initial_scaling_factor) {
    CkksParams params;
    params.dimension = dimension;
    std::vector<unsigned> bits(moduli_bits.begin(), moduli_bits.end());
    params.moduli.assign(bits.size(), 0);
    if (hehub_b200_pick_moduli(bits.data(), bits.size(), (unsigned)additional_mod_bits, params.moduli.data(), &params.additional_mod) !=
        HEHUB_B200_OK)
        throw "No suitable primes in the library.";
    params.component_count = params.moduli.size();
    params.initial_scaling_factor = initial_scaling_factor;
    return params;
}

/// basics.cpp:40-64 — the modulus budget for 128-bit security at `dimension`, cut into primes of the scaling size
inline CkksParams create_params(size_t dimension, size_t initial_scaling_bits) {
    static const std::map<size_t, size_t> std_log_q_size{{1024, 27}, {2048, 54}, {4096, 109}, {8192, 218}, {16384, 438}, {32768, 881}};
    const auto it = std_log_q_size.find(dimension);
    if (it == std_log_q_size.end()) throw "No suitable primes for this dimension.";
    const size_t log_q_size = it->second;
    if (log_q_size < 2 * initial_scaling_bits) throw "Initial scaling bits too big.";
    std::vector<size_t> mod_bits((log_q_size + 1) / initial_scaling_bits - 1, initial_scaling_bits);
    const size_t rest_bits = log_q_size - (log_q_size + 1) / initial_scaling_bits * initial_scaling_bits;
    mod_bits[0] += rest_bits / 2;
    return create_params(dimension, mod_bits, initial_scaling_bits + rest_bits / 2, std::pow(2.0, (double)initial_scaling_bits));
}

struct CkksPt : public RlwePt {
    using RlwePt::RlwePt;
    CkksPt() {}
    CkksPt(RlwePt &&other) : RlwePt(std::move(other)) {}
    double scaling_factor = 1.0;
};

struct CkksCt : public RlweCt {
    using RlweCt::RlweCt;
    CkksCt() {}
    CkksCt(RlweCt &&other) : RlweCt(std::move(other)) {}
    double scaling_factor = 1.0;
};

struct CkksQuadraticCt : public std::array<RnsPolynomial, 3> {
    using std::array<RnsPolynomial, 3>::array;
    double scaling_factor = 1.0;
};

namespace detail {
inline void check_scaling_factor(double a, double b) {
    if (std::abs(a - b) > std::pow(2.0, -50)) throw std::invalid_argument("The scaling factors mismatch"); // arith.cpp:7-13
}
inline void check_ct(const RlweCt &ct) { // rescaling.cpp:15-29
    if (ct[0].modulus_vec() != ct[1].modulus_vec()) throw std::invalid_argument("Ill-formed ciphertext: modulus sets mismatch.");
    if (ct[0].dimension() != ct[1].dimension()) throw std::invalid_argument("Ill-formed ciphertext: polynomial lengths mismatch.");
    if (ct[0].component_count() != ct[1].component_count()) throw std::invalid_argument("Ill-formed ciphertext: component numbers mismatch.");
}
/// The K polynomials as ONE contiguous [K][L][N] device operand (what the C ABI consumes): their own storage when
/// they already sit back to back in one slab — results of earlier calls do, see scatter() — else a gathered copy.
struct Operand {
    const u64 *dev = nullptr;
    std::unique_ptr<::hehub::detail::Staged> copy;
};
template <size_t K>
inline Operand gather(const std::array<RnsPolynomial, K> &polys) {
    Operand op;
    bool adjacent = true;
    for (size_t h = 0; h + 1 < K; h++) adjacent = adjacent && polys[h].device_adjacent(polys[h + 1]);
    if (adjacent) {
        op.dev = polys[0].dev();
        return op;
    }
    const size_t words = polys[0].component_count() * polys[0].dimension();
    op.copy.reset(new ::hehub::detail::Staged(K * words));
    for (size_t h = 0; h < K; h++) b200::check(hehub_b200_slab_d2d(b200::context(), op.copy->dev + h * words, polys[h].dev(), words));
    op.dev = op.copy->dev;
    return op;
}
/// the K polynomials a kernel wrote back to back into `out`, adopted without a copy (they share the slab)
template <size_t K>
inline std::array<RnsPolynomial, K> scatter(const ::hehub::detail::Staged &out, const RnsPolyParams &params) {
    std::array<RnsPolynomial, K> polys;
    const size_t words = params.component_count * params.dimension;
    for (size_t h = 0; h < K; h++) {
        polys[h] = RnsPolynomial(RnsIntVec::adopt(out.block, out.dev + h * words, params));
        polys[h].rep_form = PolyRepForm::value;
    }
    return polys;
}
} // namespace detail

/// ckks.h:180-184
inline CkksCt encrypt(const CkksPt &pt, const RlweSk &sk) {
    CkksCt ct = encrypt_core(pt, sk);
    ct.scaling_factor = pt.scaling_factor;
    return ct;
}
/// ckks.h:193-197
inline CkksPt decrypt(const CkksCt &ct, const RlweSk &sk) {
    CkksPt pt = decrypt_core(ct, sk);
    pt.scaling_factor = ct.scaling_factor;
    return pt;
}

inline CkksCt add(const CkksCt &ct1, const CkksCt &ct2) { // arith.cpp:15-20
    detail::check_scaling_factor(ct1.scaling_factor, ct2.scaling_factor);
    CkksCt sum_ct = ::hehub::add(ct1, ct2);
    sum_ct.scaling_factor = ct1.scaling_factor;
    return sum_ct;
}
inline CkksCt sub(const CkksCt &ct1, const CkksCt &ct2) { // arith.cpp:31-36
    detail::check_scaling_factor(ct1.scaling_factor, ct2.scaling_factor);
    CkksCt diff_ct = ::hehub::sub(ct1, ct2);
    diff_ct.scaling_factor = ct1.scaling_factor;
    return diff_ct;
}
inline CkksCt add_plain(const CkksCt &ct, const CkksPt &pt) { // arith.cpp:22-29
    detail::check_scaling_factor(ct.scaling_factor, pt.scaling_factor);
    RnsPolynomial pt_ntt(pt);
    ntt_negacyclic_inplace_lazy(pt_ntt);
    CkksCt sum_ct = add_plain_core(ct, pt_ntt);
    sum_ct.scaling_factor = ct.scaling_factor;
    return sum_ct;
}
inline CkksCt sub_plain(const CkksCt &ct, const CkksPt &pt) { // arith.cpp:38-45
    detail::check_scaling_factor(ct.scaling_factor, pt.scaling_factor);
    RnsPolynomial pt_ntt(pt);
    ntt_negacyclic_inplace_lazy(pt_ntt);
    CkksCt diff_ct = sub_plain_core(ct, pt_ntt);
    diff_ct.scaling_factor = ct.scaling_factor;
    return diff_ct;
}
inline CkksCt mult_plain(const CkksCt &ct, const CkksPt &pt) { // arith.cpp:47-53
    RnsPolynomial pt_ntt(pt);
    ntt_negacyclic_inplace_lazy(pt_ntt);
    CkksCt prod_ct = mult_plain_core(ct, pt_ntt);
    prod_ct.scaling_factor = ct.scaling_factor * pt.scaling_factor;
    return prod_ct;
}

/// arith.cpp:55-62 — one fused tensor-product kernel (4 multiplications + 1 lazy addition per coefficient)
inline CkksQuadraticCt mult_low_level(const CkksCt &ct1, const CkksCt &ct2) {
    detail::check_ct(ct1);
    detail::check_ct(ct2);
    ::hehub::detail::check_same_shape(ct1[0], ct2[0]);
    for (auto *ct : {&ct1, &ct2})
        for (const auto &p : *ct)
            if (p.rep_form != PolyRepForm::value)
                throw std::invalid_argument("Polynomial multiplication requires NTT form (value representation).");
    const auto params = ct1[0].params();
    const size_t words = params.component_count * params.dimension;
    ::hehub::detail::Staged q(3 * words);
    const auto a = detail::gather<2>(ct1), b = detail::gather<2>(ct2);
    b200::check(hehub_b200_ckks_tensor(b200::context(), (unsigned)ct1[0].log_dimension(), params.moduli.data(), params.component_count,
                                       a.dev, b.dev, q.dev, 1));
    CkksQuadraticCt prod;
    static_cast<std::array<RnsPolynomial, 3> &>(prod) = detail::scatter<3>(q, params);
    prod.scaling_factor = ct1.scaling_factor * ct2.scaling_factor;
    return prod;
}

/// rescaling.cpp:80-90 → :14-78
inline void rescale_inplace(CkksCt &ct, size_t dropping_primes = 1) {
    if (dropping_primes >= 2) throw "under development";
    if (dropping_primes != 1) throw std::invalid_argument("The number of primes to be dropped is not positive.");
    detail::check_ct(ct);
    if (ct[0].component_count() == 1) throw std::invalid_argument("Unable to drop the only one prime.");
    const auto params = ct[0].params();
    const size_t L = params.component_count, n = params.dimension;
    ::hehub::detail::Staged out(2 * (L - 1) * n);
    const auto in = detail::gather<2>(ct);
    b200::check(hehub_b200_ckks_rescale(b200::context(), (unsigned)ct[0].log_dimension(), params.moduli.data(), L, in.dev, out.dev, 1));
    RnsPolyParams dropped{n, L - 1, std::vector<u64>(params.moduli.begin(), params.moduli.end() - 1)};
    const double sf = ct.scaling_factor / (double)params.moduli[L - 1]; // rescaling.cpp:77
    static_cast<RlweCt &>(ct) = detail::scatter<2>(out, dropped);
    ct.scaling_factor = sf;
}

/// arith.cpp:64-73 — key switch of d2, drop of the special prime and the final additions in one call
inline CkksCt relinearize(const CkksQuadraticCt &ct, const RlweKsk &relin_key) {
    const auto &key = relin_key.packed(ct[2]);
    const auto params = ct[0].params();
    const size_t L = params.component_count, words = L * params.dimension;
    ::hehub::detail::Staged out(2 * words);
    const auto q = detail::gather<3>(ct);
    b200::check(hehub_b200_ckks_relinearize(b200::context(), (unsigned)ct[0].log_dimension(), key.ext_moduli.data(), L, q.dev, key.dev,
                                            out.dev, 1));
    CkksCt ct_new(detail::scatter<2>(out, params));
    ct_new.scaling_factor = ct.scaling_factor;
    return ct_new;
}

/// ckks.h:270-274 — tensor product + relinearize fused behind one C-ABI call
inline CkksCt mult(const CkksCt &ct1, const CkksCt &ct2, const RlweKsk &relin_key) {
    detail::check_ct(ct1);
    detail::check_ct(ct2);
    ::hehub::detail::check_same_shape(ct1[0], ct2[0]);
    const auto &key = relin_key.packed(ct1[1]);
    const auto params = ct1[0].params();
    const size_t L = params.component_count, words = L * params.dimension;
    ::hehub::detail::Staged out(2 * words);
    const auto a = detail::gather<2>(ct1), b = detail::gather<2>(ct2);
    b200::check(hehub_b200_ckks_mult_relin(b200::context(), (unsigned)ct1[0].log_dimension(), key.ext_moduli.data(), L, a.dev, b.dev,
                                           key.dev, out.dev, 1));
    CkksCt prod(detail::scatter<2>(out, params));
    prod.scaling_factor = ct1.scaling_factor * ct2.scaling_factor;
    return prod;
}

/// arith.cpp:75-83
inline CkksCt conjugate(const CkksCt &ct, const RlweKsk &conj_key) {
    detail::check_ct(ct);
    const auto &key = conj_key.packed(ct[1]);
    const auto params = ct[0].params();
    const size_t L = params.component_count, words = L * params.dimension;
    ::hehub::detail::Staged out(2 * words);
    const auto a = detail::gather<2>(ct);
    b200::check(hehub_b200_ckks_conjugate(b200::context(), (unsigned)ct[0].log_dimension(), key.ext_moduli.data(), L, a.dev, key.dev,
                                          out.dev, 1));
    CkksCt res(detail::scatter<2>(out, params));
    res.scaling_factor = ct.scaling_factor;
    return res;
}

/// arith.cpp:85-93
inline CkksCt rotate(const CkksCt &ct, const RlweKsk &rot_key, const size_t step) {
    detail::check_ct(ct);
    const auto &key = rot_key.packed(ct[1]);
    const auto params = ct[0].params();
    const size_t L = params.component_count, words = L * params.dimension;
    ::hehub::detail::Staged out(2 * words);
    const auto a = detail::gather<2>(ct);
    b200::check(hehub_b200_ckks_rotate(b200::context(), (unsigned)ct[0].log_dimension(), key.ext_moduli.data(), L, a.dev, key.dev, step,
                                       out.dev, 1));
    CkksCt res(detail::scatter<2>(out, params));
    res.scaling_factor = ct.scaling_factor;
    return res;
}

/// ckks.h:303-305
inline CkksCt rotate(const CkksCt &ct, const RotKey &rot_key) { return rotate(ct, rot_key, rot_key.step); }

} // namespace ckks

} // namespace hehub

#include "ckks_encoding.h" // simd_encode / simd_decode / encode / decode: host-side, see the header

namespace hehub {

using CkksParams = ckks::CkksParams;
using CkksPt = ckks::CkksPt;
using CkksCt = ckks::CkksCt;
using CkksSk = RlweSk;

} // namespace hehub
