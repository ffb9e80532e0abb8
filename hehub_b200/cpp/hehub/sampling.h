// sampling.h — random polynomials (src/fhe/common/sampling.{h,cpp}) for the B200 back end.
//
// The draws stay on the host, as SURVEY §8(f).3 keeps them: the same engine type, default seed and
// distributions as the reference (sampling.cpp:12-14), consumed in the same order — so a program that
// seeds nothing sees the reference's stream under the same standard library, and the parity tests can
// compare whole encryptions and keys with the reference after `rand_engine.seed(s)` on both sides.
// Everything after the draw (reduction into each RNS component, NTT) runs on the device.
#pragma once
#include <cmath>
#include <random>

#include "ntt.h"

namespace hehub {

inline std::default_random_engine rand_engine;                                  // sampling.cpp:13 (default-seeded)
inline std::uniform_int_distribution<signed char> rand_ternary((signed char)-1, (signed char)1); // sampling.cpp:14

namespace detail {
/// the draws of get_rand_gaussian_poly (sampling.cpp:60-90) before its NTT, clipped at 6 sigma.  Two quirks of the
/// reference are part of its output and reproduced: `std::round<i64>(g)` converts g to an integer FIRST (so the value
/// is truncated toward zero, not rounded), and `modulus + <that double>` is evaluated in double precision, which is
/// inexact for moduli above 2^53 (the sum is rounded to the spacing of doubles near q before it becomes a u64 again).
inline RnsPolynomial gaussian_coeffs(const RnsPolyParams &params, double std_dev) {
    const double bound = std_dev * 6;
    std::normal_distribution<double> rand_gaussian(0, std_dev);
    std::vector<double> gaussians(params.dimension);
    for (auto &g : gaussians) {
        do {
            g = rand_gaussian(rand_engine);
        } while (std::abs(g) > bound);
    }
    RnsPolynomial poly(params);
    for (size_t k = 0; k < poly.component_count(); k++) {
        const u64 q = poly.modulus_at((int)k);
        u64 *dst = poly[(int)k].data();
        for (size_t i = 0; i < gaussians.size(); i++) {
            u64 c = (u64)((double)q + (double)(int64_t)gaussians[i]);
            c -= (c >= q) ? q : 0;
            dst[i] = c;
        }
    }
    return poly;
}
} // namespace detail

/// sampling.cpp:16-37 — coefficients uniform in {-1, 0, 1}, result in NTT form
inline RnsPolynomial get_rand_ternary_poly(const RnsPolyParams &params) {
    std::vector<signed char> ternary(params.dimension);
    for (auto &t : ternary) t = rand_ternary(rand_engine);
    RnsPolynomial poly(params);
    for (size_t k = 0; k < poly.component_count(); k++) { // sampling.cpp:27-32
        const u64 q = poly.modulus_at((int)k);
        u64 *dst = poly[(int)k].data();
        for (size_t i = 0; i < ternary.size(); i++) {
            u64 c = q + (u64)ternary[i];
            c -= (c >= q) ? q : 0;
            dst[i] = c;
        }
    }
    ntt_negacyclic_inplace_lazy(poly);
    return poly;
}

/// sampling.cpp:39-58 — component k uniform in [0, q_k); `form` only labels the result
inline RnsPolynomial get_rand_uniform_poly(const RnsPolyParams &params, PolyRepForm form = PolyRepForm::coeff) {
    RnsPolynomial poly(params);
    for (size_t k = 0; k < poly.component_count(); k++) {
        std::uniform_int_distribution<u64> uni_mod((u64)0, poly.modulus_at((int)k) - 1);
        u64 *dst = poly[(int)k].data();
        for (size_t i = 0; i < poly.dimension(); i++) dst[i] = uni_mod(rand_engine);
    }
    poly.rep_form = form;
    return poly;
}

/// sampling.cpp:60-93 — rounded Gaussian coefficients (|g| <= 6 sigma), result in NTT form
inline RnsPolynomial get_rand_gaussian_poly(const RnsPolyParams &params, double std_dev = 3.2) {
    auto poly = detail::gaussian_coeffs(params, std_dev);
    ntt_negacyclic_inplace_lazy(poly);
    return poly;
}

/// sampling.cpp:95-97
inline RnsPolynomial get_zero_poly(const RnsPolyParams &params, PolyRepForm form = PolyRepForm::value) {
    RnsPolynomial poly(params);
    poly.rep_form = form;
    for (size_t k = 0; k < poly.component_count(); k++) std::fill(poly[(int)k].begin(), poly[(int)k].end(), (u64)0);
    return poly;
}

} // namespace hehub
