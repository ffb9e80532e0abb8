// modarith.cuh — word-level modular arithmetic on the device (sm_100a).
//
// Every function reproduces the *raw lazy representative* the reference computes
// (primihub/hehub, src/fhe/common/mod_arith.{h,cpp}); file:line cited per function.
// All products are exact 64x64->128 pieces built from 32-bit IMADs: this is 64-bit
// integer modular arithmetic, tensor cores are not applicable.
#pragma once
#include <cstdint>

#include "compat.h"

namespace hb {

typedef unsigned long long u64;
typedef unsigned int u32;

// Per-modulus constants and table pointers, resident in device memory.
struct LimbConst {
    u64 q;          // modulus
    u64 q2;         // 2q
    u64 nq;         // -q mod 2^64
    u64 minus_qinv; // -q^{-1} mod 2^64          (mod_arith.cpp:49-52)
    u64 r;          // 2^64 mod q                 (mod_arith.cpp:54-57)
    u64 r_h;        // floor(r * 2^64 / q)        (mod_arith.cpp:59-62)
    u64 barrett_c;  // floor((2^64-1)/q)          (mod_arith.cpp:10)
    u32 logq;       // (u64)(log2(q)+0.5)         (ntt.cpp:171)
    u32 fix;        // q >= 2^logq                (ntt.cpp:172)
    const ulonglong2 *fwd;       // forward twiddles (w, w'), pass/slot-major layout (tables.cu)
    const ulonglong2 *inv;       // inverse twiddles, pass/slot-major layout
    const ulonglong2 *inv_scale; // psi^{-i}/N (strict) and its Harvey companion, natural order
    const ulonglong2 *fwd_nat;   // reference order T[i] = psi^{bitrev(i)}      (ntt.cpp:54-58)
    const ulonglong2 *inv_nat;   // reference order, per-level inverse twiddles (ntt.cpp:64-74)
    const ulonglong2 *fwd_lat;   // forward / inverse twiddles in the layout of the latency plan (ntt_plan.h), or null
    const ulonglong2 *inv_lat;
    const ulonglong2 *fwd_lat2;  // the same for the plans of the two-launch key switch of ONE ciphertext (mode 2), or null
    const ulonglong2 *inv_lat2;
};

// lo64(x*w + h*n): one accumulation chain of 2 wide + 4 narrow IMADs, no carries needed.
HB_D u64 mul2_lo64(u64 x, u64 w, u64 h, u64 n) {
#if defined(HB_KERNEL_SIM)
    return x * w + h * n;
#else
    u64 t;
    asm("{\n\t"
        ".reg .u64 acc;\n\t"
        ".reg .u32 x0, x1, w0, w1, h0, h1, n0, n1, lo, hi;\n\t"
        "mov.b64 {x0, x1}, %1;\n\t"
        "mov.b64 {w0, w1}, %2;\n\t"
        "mov.b64 {h0, h1}, %3;\n\t"
        "mov.b64 {n0, n1}, %4;\n\t"
        "mul.wide.u32 acc, x0, w0;\n\t"
        "mad.wide.u32 acc, h0, n0, acc;\n\t"
        "mov.b64 {lo, hi}, acc;\n\t"
        "mad.lo.u32 hi, x0, w1, hi;\n\t"
        "mad.lo.u32 hi, x1, w0, hi;\n\t"
        "mad.lo.u32 hi, h0, n1, hi;\n\t"
        "mad.lo.u32 hi, h1, n0, hi;\n\t"
        "mov.b64 %0, {lo, hi};\n\t"
        "}"
        : "=l"(t)
        : "l"(x), "l"(w), "l"(h), "l"(n));
    return t;
#endif
}

// hi64(x * w), exact.  Same value as __umul64hi; the x0*w0 partial product only contributes its
// high word, so it is asked for as a 32-bit high multiply (IMAD.HI) instead of a full IMAD.WIDE,
// which occupies the multiplier pipe for twice as long on sm_100 (profiles/r1*_int_peak.json).
HB_D u64 umul64hi_split(u64 x, u64 w) {
#if defined(HB_KERNEL_SIM)
    return __umul64hi(x, w);
#else
    u64 r;
    asm("{\n\t"
        ".reg .u32 x0, x1, w0, w1, p00h, ml, mh, m2h, unused;\n\t"
        ".reg .u64 mid, mid2, acc;\n\t"
        "mov.b64 {x0, x1}, %1;\n\t"
        "mov.b64 {w0, w1}, %2;\n\t"
        "mul.hi.u32 p00h, x0, w0;\n\t"
        "cvt.u64.u32 mid, p00h;\n\t"
        "mad.wide.u32 mid, x0, w1, mid;\n\t"   // x0*w1 + hi32(x0*w0)            < 2^64
        "mov.b64 {ml, mh}, mid;\n\t"
        "cvt.u64.u32 mid2, ml;\n\t"
        "mad.wide.u32 mid2, x1, w0, mid2;\n\t" // x1*w0 + lo32(mid)              < 2^64
        "mov.b64 {unused, m2h}, mid2;\n\t"
        "cvt.u64.u32 acc, mh;\n\t"
        "mad.wide.u32 acc, x1, w1, acc;\n\t"   // x1*w1 + hi32(mid)
        "cvt.u64.u32 mid, m2h;\n\t"
        "add.u64 %0, acc, mid;\n\t"            // + hi32(mid2): the true high word, no wrap
        "}"
        : "=l"(r)
        : "l"(x), "l"(w));
    return r;
#endif
}

// mul_mod_harvey_lazy — mod_arith.h:74-78.  r = lo64(x*w) - lo64(hi64(x*w')*q), any x < 2^64,
// w < q  ->  r in [0, 2q).  `nq` is -q mod 2^64 so the subtraction folds into the IMAD chain.
HB_D u64 harvey_lazy(u64 x, u64 w, u64 wh, u64 nq) {
#if defined(HB_HI64_SPLIT)
    u64 qhat = umul64hi_split(x, wh);
#else
    u64 qhat = __umul64hi(x, wh); // exact: the low partial product's carry is kept
#endif
    return mul2_lo64(x, w, qhat, nq);
}

// lo64(a*b + c): one wide + two narrow IMADs, the addition rides on the accumulator
HB_D u64 mad_lo64(u64 a, u64 b, u64 c) {
#if defined(HB_KERNEL_SIM)
    return a * b + c;
#else
    u64 t;
    asm("{\n\t"
        ".reg .u64 acc;\n\t"
        ".reg .u32 a0, a1, b0, b1, lo, hi;\n\t"
        "mov.b64 {a0, a1}, %1;\n\t"
        "mov.b64 {b0, b1}, %2;\n\t"
        "mad.wide.u32 acc, a0, b0, %3;\n\t"
        "mov.b64 {lo, hi}, acc;\n\t"
        "mad.lo.u32 hi, a0, b1, hi;\n\t"
        "mad.lo.u32 hi, a1, b0, hi;\n\t"
        "mov.b64 %0, {lo, hi};\n\t"
        "}"
        : "=l"(t)
        : "l"(a), "l"(b), "l"(c));
    return t;
#endif
}

// the sweep at ntt.cpp:171-175: x -= ((x >> logq) - fix) * q, computed as x + m * (-q) mod 2^64
HB_D u64 approx_reduce(u64 x, const LimbConst &c) {
    return mad_lo64((x >> c.logq) - (u64)c.fix, c.nq, x);
}

// batched_reduce_strict — mod_arith.h:58-63
HB_D u64 reduce_strict(u64 x, u64 q) { return x - ((x >= q) ? q : 0ull); }

// batched_barrett_lazy — mod_arith.cpp:9-17
HB_D u64 barrett_lazy(u64 x, const LimbConst &c) {
    return x - c.q * __umul64hi(x, c.barrett_c);
}

// lazy add / sub — rns.cpp:78-84, 109-115
HB_D u64 add_lazy(u64 x, u64 y, u64 q2) {
    x += y;
    return x - ((x >= q2) ? q2 : 0ull);
}
HB_D u64 sub_lazy(u64 x, u64 y, u64 q2) {
    x += q2 - y;
    return x - ((x >= q2) ? q2 : 0ull);
}

// 128-bit accumulate acc += a*b (exact; rgsw.cpp:131-134).  Written as one carry chain over the four
// 32x32 partial products so that ptxas emits exactly four IMAD.WIDE (separate lo64 / hi64 multiplies
// cost five wide and two narrow IMADs on the multiplier pipe that bounds these kernels).
HB_D void mac128(u64 &lo, u64 &hi, u64 a, u64 b) {
#if defined(HB_KERNEL_SIM)
    u64 plo = a * b;
    u64 phi = __umul64hi(a, b);
    lo += plo;
    hi += phi + ((lo < plo) ? 1ull : 0ull);
#else
    u32 l0 = (u32)lo, l1 = (u32)(lo >> 32), h0 = (u32)hi, h1 = (u32)(hi >> 32);
    const u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32);
    asm("mad.lo.cc.u32 %0, %4, %6, %0;\n\t"
        "madc.hi.cc.u32 %1, %4, %6, %1;\n\t"
        "madc.lo.cc.u32 %2, %5, %7, %2;\n\t"
        "madc.hi.u32 %3, %5, %7, %3;\n\t"
        "mad.lo.cc.u32 %1, %4, %7, %1;\n\t"
        "madc.hi.cc.u32 %2, %4, %7, %2;\n\t"
        "addc.u32 %3, %3, 0;\n\t"
        "mad.lo.cc.u32 %1, %5, %6, %1;\n\t"
        "madc.hi.cc.u32 %2, %5, %6, %2;\n\t"
        "addc.u32 %3, %3, 0;\n\t"
        : "+r"(l0), "+r"(l1), "+r"(h0), "+r"(h1)
        : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
    lo = ((u64)l1 << 32) | l0;
    hi = ((u64)h1 << 32) | h0;
#endif
}

// The same accumulation for inner loops that run it many times on the same accumulator (key-switch inner product).
// mac128 keeps its four 32-bit words in ONE chain, so the cross products a0*b1, a1*b0 land on the register pair
// (word 1, word 2) while a0*b0 and a1*b1 want (0, 1) and (2, 3): ptxas cannot give both pairings an even register and
// pays ~9 register moves per call, half of them on the multiplier pipe (profiles/r3_mac_census.md).  Here the cross
// products go to a separate 96-bit accumulator (m0, m1, m2), every IMAD.WIDE works on an aligned pair, and the two are
// added once at the end: (r3 r2 r1 r0) + ((m2 m1 m0) << 32), the same value mod 2^128.
struct Acc128 {
    u64 r01, r23, m01; // register pairs, as IMAD.WIDE wants them
    u32 m2;
};
HB_D void acc128_clear(Acc128 &s) {
    s.r01 = s.r23 = s.m01 = 0;
    s.m2 = 0;
}
HB_D void acc128_mac(Acc128 &s, u64 a, u64 b) {
#if defined(HB_KERNEL_SIM)
    const unsigned __int128 prod = (unsigned __int128)a * b; // the emulator keeps everything in (r23, r01); m stays 0
    const unsigned __int128 acc = (((unsigned __int128)s.r23 << 64) | s.r01) + prod;
    s.r01 = (u64)acc;
    s.r23 = (u64)(acc >> 64);
#else
    asm("{\n\t"
        ".reg .u32 a0, a1, b0, b1, r0, r1, r2, r3, m0, m1;\n\t"
        "mov.b64 {a0, a1}, %4;\n\t"
        "mov.b64 {b0, b1}, %5;\n\t"
        "mov.b64 {r0, r1}, %0;\n\t"
        "mov.b64 {r2, r3}, %1;\n\t"
        "mov.b64 {m0, m1}, %2;\n\t"
        "mad.lo.cc.u32 r0, a0, b0, r0;\n\t"
        "madc.hi.cc.u32 r1, a0, b0, r1;\n\t"
        "madc.lo.cc.u32 r2, a1, b1, r2;\n\t"
        "madc.hi.u32 r3, a1, b1, r3;\n\t"
        "mad.lo.cc.u32 m0, a0, b1, m0;\n\t"
        "madc.hi.cc.u32 m1, a0, b1, m1;\n\t"
        "addc.u32 %3, %3, 0;\n\t"
        "mad.lo.cc.u32 m0, a1, b0, m0;\n\t"
        "madc.hi.cc.u32 m1, a1, b0, m1;\n\t"
        "addc.u32 %3, %3, 0;\n\t"
        "mov.b64 %0, {r0, r1};\n\t"
        "mov.b64 %1, {r2, r3};\n\t"
        "mov.b64 %2, {m0, m1};\n\t"
        "}"
        : "+l"(s.r01), "+l"(s.r23), "+l"(s.m01), "+r"(s.m2)
        : "l"(a), "l"(b));
#endif
}
// (r3 r2 r1 r0) + ((m2 m1 m0) << 32) mod 2^128
HB_D void acc128_fold(const Acc128 &s, u64 &lo, u64 &hi) {
    const u64 shifted_lo = s.m01 << 32, shifted_hi = (s.m01 >> 32) | ((u64)s.m2 << 32);
    lo = s.r01 + shifted_lo;
    hi = s.r23 + shifted_hi + (lo < shifted_lo ? 1ull : 0ull);
}

// full 128-bit product (lo, hi) = a * b with four IMAD.WIDE
HB_D void mul_full(u64 a, u64 b, u64 &lo, u64 &hi) {
    lo = 0;
    hi = 0;
    mac128(lo, hi, a, b);
}

// 128-bit a = (hi, lo);  Montgomery reduce: (a + (lo*minus_qinv mod 2^64)*q) >> 64
// — mod_arith.cpp:126-133 (also the first half of the hybrid mulmod, :80-87).
HB_D u64 montgomery128(u64 lo, u64 hi, const LimbConst &c) {
    u64 u = lo * c.minus_qinv;
    u64 phi = __umul64hi(u, c.q);
    // lo + lo64(u*q) == 0 mod 2^64 by construction, so that addition carries exactly when lo != 0:
    // the low product itself is never needed.  (Clearing 32 bits per round instead — 4 wide + 2 narrow
    // multiplies for the identical quotient — measured 3-4 % slower in the tensor kernel: more
    // instructions, and that kernel is as much issue- as multiplier-bound.)
    u64 carry = (lo != 0) ? 1ull : 0ull;
    return hi + phi + carry;
}

// batched_mul_mod_hybrid_lazy — mod_arith.cpp:64-92
HB_D u64 mul_hybrid_lazy(u64 a, u64 b, const LimbConst &c) {
    u64 lo, hi;
    mul_full(a, b, lo, hi);
    u64 t = montgomery128(lo, hi, c);
    return harvey_lazy(t, c.r, c.r_h, c.nq);
}

} // namespace hb
