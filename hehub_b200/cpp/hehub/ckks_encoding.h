// ckks_encoding.h — ckks::simd_encode / simd_decode / encode / decode (src/fhe/ckks/ckks.h:103-172, basics.cpp:66-369).
//
// Host-side by nature (complex doubles, and integers wider than a machine word): SURVEY §2 row 12 keeps this off the
// GPU path.  It is here so that an application written against the reference — e.g. examples/ckks_example.cpp — builds
// against the mirror unchanged.  Encoding must hand the integer path the SAME words as the reference, so the floating-point
// recipe is restated operation for operation (same libm / libstdc++ calls in the same order, same truncations); decoding
// ends in doubles and is held to a relative tolerance instead.  The reference's decimal-string big integers are replaced
// by little-endian machine-word integers with the same exact results.
#pragma once
#include <cmath>
#include <complex>
#include <map>
#include <numeric>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "mod_arith.h"
#include "rns_transform.h"

namespace hehub {

using cc_double = std::complex<double>;

namespace ckks {
namespace detail {

/// reverses the low `bits` bits of x (bits <= 16), like permutation.h:41-55
inline u64 reverse_bits16(u64 x, int bits) {
    u64 r = 0;
    for (int b = 0; b < bits; b++) r |= ((x >> b) & 1) << (bits - 1 - b);
    return r;
}

/// 3^i modulo 2^32 for i < 2^17: the root index of slot i is that value modulo 2N (permutation.cpp:11-26)
inline const std::vector<uint32_t> &root_index_factors() {
    static const std::vector<uint32_t> table = [] {
        std::vector<uint32_t> t((size_t)1 << 17);
        t[0] = 1;
        for (size_t i = 1; i < t.size(); i++) t[i] = t[i - 1] * 3u;
        return t;
    }();
    return table;
}

/// The negacyclic FFT of basics.cpp:66-154 over complex doubles, natural order in and out.  Twiddles per (size, direction)
/// are cached like the reference's static map; every product and sum below is the one the reference performs.
class NegacyclicFft {
public:
    static void run(cc_double *values, size_t log_dimension, bool inverse) {
        const Factors &f = factors(log_dimension, inverse);
        const size_t dimension = (size_t)1 << log_dimension;
        std::vector<cc_double> work(dimension);
        for (size_t i = 0; i < dimension; i++) work[i] = inverse ? values[i] : values[i] * f.twist[i];
        size_t idx = 0;
        for (size_t level = 1, gap = dimension / 2; level <= log_dimension; level++, gap >>= 1)
            for (size_t start = 0; start < dimension; start += 2 * gap, idx++)
                for (size_t l = start; l < start + gap; l++) {
                    const size_t h = l + gap;
                    const cc_double t = work[h] * f.butterfly[idx];
                    work[h] = work[l] - t;
                    work[l] = work[l] + t;
                }
        for (size_t i = 0; i < dimension; i++) values[i] = work[reverse_bits16(i, (int)log_dimension)];
        if (inverse)
            for (size_t i = 0; i < dimension; i++) values[i] *= f.twist[i];
    }

private:
    struct Factors {
        std::vector<cc_double> twist, butterfly;
    };
    static const Factors &factors(size_t log_dimension, bool inverse) {
        static std::map<std::pair<size_t, bool>, Factors> cache;
        auto it = cache.find({log_dimension, inverse});
        if (it != cache.end()) return it->second;
        Factors f;
        const auto dimension = 1ULL << log_dimension;
        cc_double zeta = std::polar(1.0, 2 * M_PI / dimension);
        if (inverse) zeta = std::conj(zeta);
        for (size_t i = 0; i < dimension; i++)
            f.twist.push_back(inverse ? std::polar(1.0 / dimension, i * M_PI / dimension * -1.0) : std::polar(1.0, i * M_PI / dimension));
        for (size_t level = 1, gap = dimension / 2; level <= log_dimension; level++, gap >>= 1)
            for (size_t local = 0; local < dimension / gap / 2; local++)
                f.butterfly.push_back(std::pow(zeta, (reverse_bits16(local, (int)level - 1) << (log_dimension - level))));
        return cache.emplace(std::make_pair(log_dimension, inverse), std::move(f)).first->second;
    }
};

/// Non-negative integers of any size, little-endian 64-bit words: what the reference's decimal UBInt is used for here.
struct WideUInt {
    std::vector<u64> w; // no leading zero words; empty = 0
    WideUInt() {}
    explicit WideUInt(u64 v) {
        if (v) w.push_back(v);
    }
    /// floor(d) for a finite d >= 0, exactly (UBInt::from_double, bigint.cpp:35-46, prints the exact digits)
    static WideUInt floor_of(double d) {
        WideUInt r;
        if (!(d >= 1.0)) return r;
        int e;
        const double m = std::frexp(d, &e);                 // d = m * 2^e, 0.5 <= m < 1
        u64 mant = (u64)std::ldexp(m, 53);                  // 53-bit integer mantissa
        int shift = e - 53;                                 // value = mant * 2^shift
        if (shift < 0) {
            mant >>= -shift;
            shift = 0;
        }
        r.w.assign((size_t)shift / 64 + 2, 0);
        r.w[shift / 64] = mant << (shift % 64);
        if (shift % 64) r.w[shift / 64 + 1] = mant >> (64 - shift % 64);
        r.trim();
        return r;
    }
    void trim() {
        while (!w.empty() && w.back() == 0) w.pop_back();
    }
    u64 mod(u64 q) const {
        u128 r = 0;
        for (size_t i = w.size(); i-- > 0;) r = ((r << 64) | w[i]) % q;
        return (u64)r;
    }
    void mul_add(u64 m, u64 a) { // this = this * m + a
        u128 carry = a;
        for (auto &x : w) {
            const u128 t = (u128)x * m + carry;
            x = (u64)t;
            carry = t >> 64;
        }
        if (carry) w.push_back((u64)carry);
    }
    void halve() {
        u64 carry = 0;
        for (size_t i = w.size(); i-- > 0;) {
            const u64 next = w[i] & 1;
            w[i] = (w[i] >> 1) | (carry << 63);
            carry = next;
        }
        trim();
    }
    bool less_than(const WideUInt &o) const {
        if (w.size() != o.w.size()) return w.size() < o.w.size();
        for (size_t i = w.size(); i-- > 0;)
            if (w[i] != o.w[i]) return w[i] < o.w[i];
        return false;
    }
    WideUInt minus(const WideUInt &o) const { // this >= o
        WideUInt r(*this);
        u64 borrow = 0;
        for (size_t i = 0; i < r.w.size(); i++) {
            const u64 sub = i < o.w.size() ? o.w[i] : 0;
            const u64 before = r.w[i];
            r.w[i] = before - sub - borrow;
            borrow = (before < sub || (before == sub && borrow)) ? 1 : 0;
        }
        r.trim();
        return r;
    }
    /// nearest double, ties to even (to_double, bigint.cpp:295-302, reads the decimal digits back)
    double to_double() const {
        if (w.empty()) return 0.0;
        const size_t top = w.size() - 1;
        const int lead = 63 - __builtin_clzll(w[top]);
        const long bits = (long)top * 64 + lead + 1;
        if (bits <= 64) return (double)w[0]; // the conversion rounds to nearest even
        // the top 64 bits, with everything below folded into a sticky bit so that one rounding decides
        const long drop = bits - 64;
        u64 head = 0;
        bool sticky = false;
        for (long b = 0; b < 64; b++) {
            const long pos = drop + b;
            head |= ((w[pos / 64] >> (pos % 64)) & 1) << b;
        }
        for (long pos = 0; pos < drop && !sticky; pos++) sticky = (w[pos / 64] >> (pos % 64)) & 1;
        // 64 -> 53 bits by hand (the low 11 bits and the sticky bit decide), then scale
        u64 mant = head >> 11;
        const u64 rest = head & 0x7FF;
        if (rest > 0x400 || (rest == 0x400 && (sticky || (mant & 1)))) mant++;
        return std::ldexp((double)mant, (int)(drop + 11));
    }
};

inline void check_scaling(double scaling_factor) {
    if (scaling_factor <= 0) throw std::invalid_argument("Scaling factor should be positive.");
}

} // namespace detail

/// basics.cpp:156-255 — slots to the two conjugate halves, inverse negacyclic FFT, scale, truncate, reduce into every limb
inline CkksPt simd_encode_cc(const std::vector<cc_double> &data, const double scaling_factor, const CkksParams &pt_params) {
    detail::check_scaling(scaling_factor);
    const auto dimension = pt_params.dimension;
    const size_t log_dimension = std::round(std::log2(dimension));
    const auto slot_count = dimension / 2;
    if (data.size() > slot_count)
        throw std::invalid_argument("Cannot encode " + std::to_string(data.size()) + " data into " + std::to_string(slot_count) + " slots.");
    std::vector<cc_double> interpolated(dimension, 0.0);
    const auto &root_indices = detail::root_index_factors();
    const auto mask = (1 << (log_dimension + 1)) - 1;
    for (size_t i = 0; i < data.size(); i++) {
        const auto position = ((root_indices[i] & mask) - 1) / 2;
        interpolated[position] = data[i];
        interpolated[dimension - 1 - position] = std::conj(data[i]);
    }
    detail::NegacyclicFft::run(interpolated.data(), size_t(std::log2(interpolated.size())), /*inverse=*/true);

    CkksPt pt(RnsPolynomial(pt_params.dimension, pt_params.component_count, pt_params.moduli));
    const size_t L = pt_params.component_count;
    std::vector<u64 *> limb(L);
    for (size_t k = 0; k < L; k++) limb[k] = pt[(int)k].data(); // host words; uploaded when a device operator first needs them
    bool small_coeff = true; // every scaled coefficient below 2^64?
    const double small_bound = std::pow(2.0, 64) / scaling_factor;
    for (const auto &d : interpolated)
        if (std::abs(d.real()) > small_bound) {
            small_coeff = false;
            break;
        }
    for (size_t i = 0; i < dimension; i++) {
        double coeff = interpolated[i].real();
        coeff *= scaling_factor;
        const bool negative = coeff <= 0;
        if (small_coeff) {
            const u64 magnitude = u64(std::abs(coeff));
            for (size_t k = 0; k < L; k++) {
                const u64 q = pt_params.moduli[k];
                // batched_barrett_lazy (mod_arith.cpp:9-17): a representative below 2q, then the lazy negation of :215-221
                u64 x = magnitude - q * (u64)(((u128)magnitude * ((u64)(-1) / q)) >> 64);
                if (negative) x = (x == 0) ? 0 : (2 * q - x);
                limb[k][i] = x;
            }
        } else {
            const auto magnitude = detail::WideUInt::floor_of(std::abs(coeff));
            for (size_t k = 0; k < L; k++) {
                const u64 q = pt_params.moduli[k];
                const u64 x = magnitude.mod(q);
                limb[k][i] = negative ? q - x : x; // basics.cpp:243-247: q itself when x == 0, like the reference
            }
        }
    }
    pt.scaling_factor = scaling_factor;
    return pt;
}

inline CkksPt simd_encode(const std::vector<cc_double> &data, const CkksParams &pt_params) {
    return simd_encode_cc(data, pt_params.initial_scaling_factor, pt_params);
}
inline CkksPt simd_encode(const std::vector<double> &data, const CkksParams &pt_params) {
    std::vector<cc_double> data_cc;
    for (auto d : data) data_cc.push_back(cc_double(d));
    return simd_encode_cc(data_cc, pt_params.initial_scaling_factor, pt_params);
}
/// ckks.h:123-139 — one datum in every slot
inline CkksPt encode(const cc_double datum, const CkksParams &pt_params) {
    return simd_encode(std::vector<cc_double>(pt_params.dimension / 2, datum), pt_params);
}
inline CkksPt encode(const double datum, const CkksParams &pt_params) {
    return simd_encode(std::vector<double>(pt_params.dimension / 2, datum), pt_params);
}

/// basics.cpp:274-352 — centre every coefficient (composing the limbs when it does not fit the first one), unscale, FFT
inline std::vector<cc_double> simd_decode_cc(const CkksPt &pt, size_t data_size) {
    const auto scaling_factor = pt.scaling_factor;
    detail::check_scaling(scaling_factor);
    const auto slot_count = pt.dimension() / 2;
    if (data_size == 0) data_size = slot_count;
    if (data_size > slot_count)
        throw std::invalid_argument("Cannot decode " + std::to_string(data_size) + " items from " + std::to_string(slot_count) + " slots.");
    RnsPolynomial reduced(pt);
    reduce_strict(reduced);
    const auto dimension = pt.dimension();
    const size_t log_dimension = std::round(std::log2(dimension));
    const size_t L = pt.component_count();
    const auto &moduli = pt.modulus_vec();
    std::vector<const u64 *> limb(L);
    for (size_t k = 0; k < L; k++) limb[k] = reduced[(int)k].data();
    // small: every coefficient, lifted from the first modulus to the others as a centred value, reproduces the other limbs
    bool small_coeff = true;
    const u64 q0 = moduli[0], half_q0 = q0 / 2;
    for (size_t k = 1; k < L && small_coeff; k++) {
        const u64 q = moduli[k], lift = (q0 / q + 1) * q - q0; // rns_transform.cpp:11-37
        for (size_t i = 0; i < dimension; i++) {
            u64 x = limb[0][i];
            if (x >= half_q0) x += lift;
            if (q < q0) x -= q * (u64)(((u128)x * ((u64)(-1) / q)) >> 64); // lazy Barrett: compare raw words, like the reference
            if (x != limb[k][i]) {
                small_coeff = false;
                break;
            }
        }
    }
    std::vector<cc_double> interpolated(dimension);
    if (small_coeff) {
        for (size_t i = 0; i < dimension; i++)
            interpolated[i] = limb[0][i] < half_q0 ? (double)limb[0][i] : -(double)(q0 - limb[0][i]);
    } else {
        // compose by mixed radix (Garner): X = v_0 + v_1 q_0 + v_2 q_0 q_1 + ..., exact
        detail::WideUInt whole(1), half;
        for (u64 q : moduli) whole.mul_add(q, 0);
        half = whole;
        half.halve();
        std::vector<std::vector<u64>> prefix_mod(L, std::vector<u64>(L, 1)); // (q_0 .. q_{j-1}) mod q_k
        std::vector<u64> prefix_inv(L, 1);                                   // (q_0 .. q_{k-1})^{-1} mod q_k
        for (size_t k = 0; k < L; k++) {
            for (size_t j = 1; j <= k; j++) prefix_mod[j][k] = (u64)((u128)prefix_mod[j - 1][k] * (moduli[j - 1] % moduli[k]) % moduli[k]);
            if (k) prefix_inv[k] = inverse_mod_prime(prefix_mod[k][k], moduli[k]);
        }
        std::vector<u64> v(L);
        for (size_t i = 0; i < dimension; i++) {
            for (size_t k = 0; k < L; k++) {
                const u64 q = moduli[k];
                u128 acc = 0;
                for (size_t j = 0; j < k; j++) acc = (acc + (u128)(v[j] % q) * prefix_mod[j][k]) % q;
                const u64 diff = (u64)(((u128)limb[k][i] + q - (u64)acc) % q);
                v[k] = k ? (u64)((u128)diff * prefix_inv[k] % q) : limb[0][i];
            }
            detail::WideUInt x(v[L - 1]); // Horner: v_0 + q_0 (v_1 + q_1 (v_2 + ...))
            for (size_t k = L - 1; k-- > 0;) x.mul_add(moduli[k], v[k]);
            interpolated[i] = x.less_than(half) ? x.to_double() : -whole.minus(x).to_double();
        }
    }
    for (auto &c : interpolated) c /= scaling_factor;
    detail::NegacyclicFft::run(interpolated.data(), log_dimension, /*inverse=*/false);
    std::vector<cc_double> data(slot_count);
    const auto &root_indices = detail::root_index_factors();
    const auto mask = (1 << (log_dimension + 1)) - 1;
    for (size_t i = 0; i < slot_count; i++) data[i] = interpolated[((root_indices[i] & mask) - 1) / 2];
    (void)data_size; // like the reference, every slot is returned (basics.cpp:343-351 ignores data_size after the check)
    return data;
}

template <typename T = double, typename std::enable_if<std::is_same<T, double>::value || std::is_same<T, cc_double>::value>::type * = nullptr>
inline std::vector<T> simd_decode(const CkksPt &pt, size_t data_size = 0) {
    auto data_cc = simd_decode_cc(pt, data_size);
    if constexpr (std::is_same<T, cc_double>::value) {
        return data_cc;
    } else {
        std::vector<double> data;
        for (const auto &d : data_cc) data.push_back(d.real());
        return data;
    }
}
/// ckks.h:167-172 — the mean of the slots
template <typename T = double, typename std::enable_if<std::is_same<T, double>::value || std::is_same<T, cc_double>::value>::type * = nullptr>
inline T decode(const CkksPt &pt) {
    std::vector<T> decoded = simd_decode<T>(pt);
    T sum = std::accumulate(decoded.begin(), decoded.end(), (T)0);
    return sum / decoded.size();
}

} // namespace ckks
} // namespace hehub
