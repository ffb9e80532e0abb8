// mod_arith.h — word-level modular kernels of hehub (src/fhe/common/mod_arith.{h,cpp}) on the B200
// back end.  The batched_* functions keep the reference's signatures (host pointers in, host
// pointers out): each call stages the vector through a device slab and runs ONE coefficient-wise
// CUDA kernel.  They exist for drop-in completeness and for the parity tests; code that cares about
// speed keeps its data in RnsPolynomial objects, whose operators never leave the device.
#pragma once
#include <memory>
#include <vector>

#include "rns.h"

namespace hehub {

namespace detail {
// RAII device staging of a host vector
struct Staged {
    std::shared_ptr<SlabBlock> block;
    u64 *dev = nullptr;
    size_t words;
    explicit Staged(size_t n) : block(std::make_shared<SlabBlock>(n)), dev(block->p), words(n) {}
    Staged(const u64 *host, size_t n) : Staged(n) {
        if (n) {
            b200::check(hehub_b200_slab_h2d(b200::context(), dev, host, n));
            b200::synchronize();
        }
    }
    void download(u64 *host) {
        if (!words) return;
        b200::check(hehub_b200_slab_d2h(b200::context(), host, dev, words));
        b200::synchronize();
    }
};
} // namespace detail

/// mod_arith.cpp:9-17 — vec[i] -= q * floor(vec[i] * floor((2^64-1)/q) / 2^64), result in [0, 2q)
inline void batched_barrett_lazy(const u64 modulus, const size_t vec_len, u64 vec[]) {
    detail::Staged s(vec, vec_len);
    b200::check(hehub_b200_barrett_lazy(b200::context(), vec_len, &modulus, 1, s.dev, 1));
    s.download(vec);
}
/// mod_arith.h:18-25
inline void batched_barrett(const u64 modulus, const size_t vec_len, u64 vec[]) {
    detail::Staged s(vec, vec_len);
    b200::check(hehub_b200_barrett(b200::context(), vec_len, &modulus, 1, s.dev, 1));
    s.download(vec);
}
/// mod_arith.h:58-63
inline void batched_reduce_strict(const u64 modulus, const size_t vec_len, u64 vec[]) {
    detail::Staged s(vec, vec_len);
    b200::check(hehub_b200_reduce_strict(b200::context(), vec_len, &modulus, 1, s.dev, 1));
    s.download(vec);
}
/// mod_arith.cpp:64-92 — Montgomery reduce then Harvey multiply by 2^64 mod q; result in [0, 2q)
inline void batched_mul_mod_hybrid_lazy(const u64 modulus, const size_t vec_len, const u64 in_vec1[], const u64 in_vec2[],
                                        u64 out_vec[]) {
    detail::Staged a(in_vec1, vec_len), b(in_vec2, vec_len), c(vec_len);
    b200::check(hehub_b200_mulmod_hybrid_lazy(b200::context(), vec_len, &modulus, 1, a.dev, b.dev, c.dev, 1));
    c.download(out_vec);
}
/// mod_arith.h:31-39
inline void batched_mul_mod_hybrid(const u64 modulus, const size_t vec_len, const u64 in_vec1[], const u64 in_vec2[],
                                   u64 out_vec[]) {
    detail::Staged a(in_vec1, vec_len), b(in_vec2, vec_len), c(vec_len);
    b200::check(hehub_b200_mulmod_hybrid_lazy(b200::context(), vec_len, &modulus, 1, a.dev, b.dev, c.dev, 1));
    b200::check(hehub_b200_reduce_strict(b200::context(), vec_len, &modulus, 1, c.dev, 1));
    c.download(out_vec);
}
/// mod_arith.cpp:113-134 — (a + (lo64(a) * (-q^-1) mod 2^64) * q) >> 64
inline void batched_montgomery_128_lazy(const u64 modulus, const size_t len, const u128 in[], u64 out[]) {
    detail::Staged a(reinterpret_cast<const u64 *>(in), 2 * len), c(len); // little-endian (lo, hi) pairs
    b200::check(hehub_b200_montgomery128_lazy(b200::context(), modulus, len, a.dev, c.dev));
    c.download(out);
}

/// mod_arith.h:65-72 — on the device, no host round trip
inline void reduce_strict(RnsPolynomial &rns_poly) {
    b200::check(hehub_b200_reduce_strict(b200::context(), rns_poly.dimension(), rns_poly.modulus_vec().data(),
                                         rns_poly.component_count(), rns_poly.dev_mut(), 1));
}

/// mod_arith.h:74-78 (scalar; host)
inline u64 mul_mod_harvey_lazy(const u64 modulus, const u64 in1, const u64 in2, const u64 in2_harvey) {
    const u64 approx_quotient = (u64)(((u128)in1 * in2_harvey) >> 64);
    return (u64)((u128)in1 * in2 - (u128)approx_quotient * modulus);
}

/// mod_arith.cpp:138-149 (host; Fermat)
inline u64 inverse_mod_prime(const u64 elem, const u64 prime) {
    u64 result = 1 % prime, base = elem % prime, e = prime - 2;
    while (e) {
        if (e & 1) result = (u64)((u128)result * base % prime);
        base = (u64)((u128)base * base % prime);
        e >>= 1;
    }
    return result;
}

} // namespace hehub
