// ntt_engine.cuh — negacyclic NTT / INTT kernels for sm_100a.
//
// Replaces ntt_negacyclic_inplace_lazy / intt_negacyclic_inplace_lazy of the reference
// (src/fhe/common/ntt.cpp:145-223).  The butterfly dataflow graph, the absence of reduction
// between levels, the final approximate reduction and the psi^{-i}/N scaling are reproduced
// operation for operation, so outputs are the same raw u64 words as the CPU path.  Only the
// *order* in which independent butterflies run is changed:
//
//   * one CTA owns a whole row (N <= 16384) in shared memory; a row of N = 32768 (256 KiB, more
//     than one SM holds) is owned by a 2-CTA thread-block cluster, half a row each: the single
//     cross-CTA level (gap N/2) is computed from global/L2 reads on the way in (forward) or
//     exchanged through the row's own global words between two cluster barriers (inverse);
//   * each thread keeps 8 or 16 coefficients in registers and runs 3-4 levels on them
//     ("pass"), so the row crosses shared memory 2-3 times instead of log2(N) times;
//   * the inverse uses the folded form (no bit-reversal permutations).
//
// Kernels are parameterised by an IO policy that supplies the row's input words and consumes
// its output words; the fused rescale / key-switch kernels (ops.cu) are instantiations with
// prologue/epilogue arithmetic inside the policy, so the coefficient slab is read from and
// written to HBM exactly once per transform.
//
// IO policy interface (all __device__):
//   int  limb(int row)                                   -> index into the LimbConst array
//   u64  load(int row, int i, const LimbConst&)          -> input word i of the row
//   void store(int row, int i, u64 v, const LimbConst&)  -> output word i of the row
//   u64 *raw(int row)                                    -> row-sized exchange area in global memory
//                                                           (N = 32768 inverse only; may be the output row)
#pragma once
#include "modarith.cuh"
#include "ntt_plan.h"

namespace hb {

// ------------------------------------------------------------------------------------------
// butterfly — ntt.cpp:161-167 (identical for both directions)
// ------------------------------------------------------------------------------------------
HB_D void bfly(u64 &lo, u64 &hi, const ulonglong2 tw, u64 nq, u64 q2) {
    u64 t = harvey_lazy(hi, tw.x, tw.y, nq);
    hi = lo + q2 - t;
    lo = lo + t;
}

// K forward levels on 2^K registers.  Level m pairs registers 2^(K-m) apart; twiddle slot
// (2^(m-1) - 1 + blk) is read at tw[slot * stride].
template <int K, int M = 1>
HB_D void fwd_levels(u64 (&v)[1 << K], const ulonglong2 *__restrict__ tw,
                                           int stride, u64 nq, u64 q2) {
    constexpr int half = 1 << (K - M);
#pragma unroll
    for (int blk = 0; blk < (1 << (M - 1)); blk++) {
        const ulonglong2 z = __ldg(tw + ((1 << (M - 1)) - 1 + blk) * stride);
#pragma unroll
        for (int jj = 0; jj < half; jj++) bfly(v[blk * 2 * half + jj], v[blk * 2 * half + jj + half], z, nq, q2);
    }
    if constexpr (M < K) fwd_levels<K, M + 1>(v, tw, stride, nq, q2);
}

// K inverse (folded) stages on 2^K registers.  Stage m pairs registers 2^(m-1) apart; twiddle
// slot (2^(m-1) - 1 + jj) depends on the register's index modulo 2^(m-1).
template <int K, int M = 1>
HB_D void inv_levels(u64 (&v)[1 << K], const ulonglong2 *__restrict__ tw,
                                           int stride, u64 nq, u64 q2) {
    constexpr int d = 1 << (M - 1);
#pragma unroll
    for (int jj = 0; jj < d; jj++) {
        const ulonglong2 z = __ldg(tw + (d - 1 + jj) * stride);
#pragma unroll
        for (int blk = 0; blk < (1 << (K - M)); blk++) bfly(v[blk * 2 * d + jj], v[blk * 2 * d + jj + d], z, nq, q2);
    }
    if constexpr (M < K) inv_levels<K, M + 1>(v, tw, stride, nq, q2);
}

HB_D int sphys(int i) { return i + ((i >> 4) << 1); }

// contiguous 2^K-word group at logical index base (aligned to 2^K <= 16): 128-bit accesses
template <int K>
HB_D void lds_contig(const u64 *sm, int base, u64 (&v)[1 << K]) {
    const ulonglong2 *p = reinterpret_cast<const ulonglong2 *>(sm + sphys(base));
#pragma unroll
    for (int c = 0; c < (1 << K) / 2; c++) {
        ulonglong2 t = p[c];
        v[2 * c] = t.x;
        v[2 * c + 1] = t.y;
    }
}
template <int K>
HB_D void sts_contig(u64 *sm, int base, const u64 (&v)[1 << K]) {
    ulonglong2 *p = reinterpret_cast<ulonglong2 *>(sm + sphys(base));
#pragma unroll
    for (int c = 0; c < (1 << K) / 2; c++) p[c] = make_ulonglong2(v[2 * c], v[2 * c + 1]);
}

// ------------------------------------------------------------------------------------------
// forward, fast path
// ------------------------------------------------------------------------------------------
template <int LOGN, int P, class IO>
HB_D void fwd_pass(u64 *sm, const IO &io, const LimbConst &lc, int row, int B) {
    constexpr NttPlan pl = plan_for(LOGN);
    constexpr int LOGNC = LOGN - pl.lpre, NC = 1 << LOGNC, T = pl.threads;
    constexpr int K = pl.k[P], L0 = fwd_lambda0(pl, P), GSL = LOGNC - L0 - K; // log2(smallest gap)
    constexpr int NG = NC >> K;
    constexpr bool first = (P == 0), last = (P == pl.npass - 1);
    static_assert(!last || GSL == 0, "last forward pass must be contiguous");
    const ulonglong2 *tw_pass = lc.fwd + fwd_pass_offset(pl, P);
    constexpr int stride = 1 << (pl.lpre + L0);

#pragma unroll 1
    for (int g = threadIdx.x; g < NG; g += T) {
        const int lo = g & ((1 << GSL) - 1), hb = g >> GSL;
        const int base = (hb << (LOGNC - L0)) + lo;
        u64 v[1 << K];
        if constexpr (first) {
            if constexpr (pl.lpre == 0) {
#pragma unroll
                for (int j = 0; j < (1 << K); j++) v[j] = io.load(row, base + (j << GSL), lc);
            } else {
                // level 1 of the full row (gap N/2) is computed here by both CTAs of the row;
                // CTA B keeps the low (B = 0) or high (B = 1) output — ntt.cpp:161-167
                const ulonglong2 z = __ldg(lc.fwd);
#pragma unroll
                for (int j = 0; j < (1 << K); j++) {
                    const int i = base + (j << GSL);
                    u64 a = io.load(row, i, lc), b = io.load(row, i + NC, lc);
                    u64 t = harvey_lazy(b, z.x, z.y, lc.nq);
                    v[j] = B ? (a + lc.q2 - t) : (a + t);
                }
            }
        } else if constexpr (last) {
            lds_contig<K>(sm, base, v);
        } else {
#pragma unroll
            for (int j = 0; j < (1 << K); j++) v[j] = sm[sphys(base + (j << GSL))];
        }

        fwd_levels<K>(v, tw_pass + ((B << L0) + hb), stride, lc.nq, lc.q2);

        if constexpr (last) {
#pragma unroll
            for (int j = 0; j < (1 << K); j++) v[j] = approx_reduce(v[j], lc); // ntt.cpp:171-175
            sts_contig<K>(sm, base, v);
        } else {
#pragma unroll
            for (int j = 0; j < (1 << K); j++) sm[sphys(base + (j << GSL))] = v[j];
        }
    }
}

template <int LOGN, int P, class IO>
HB_D void fwd_passes(u64 *sm, const IO &io, const LimbConst &lc, int row, int B) {
    constexpr NttPlan pl = plan_for(LOGN);
    fwd_pass<LOGN, P>(sm, io, lc, row, B);
    if constexpr (P == 0 && pl.lpre == 1) {
        // both CTAs of the row have read all of it: from here on either may overwrite it
        // (in-place transforms store into the words the sibling CTA has just read)
        hb_cluster_sync();
    } else {
        __syncthreads();
    }
    if constexpr (P + 1 < pl.npass) fwd_passes<LOGN, P + 1>(sm, io, lc, row, B);
}

template <int LOGN, class IO>
HB_GLOBAL(plan_for(LOGN).threads, plan_for(LOGN).min_blocks)
ntt_fwd_fast_kernel(const IO io, const LimbConst *__restrict__ limbs) {
    constexpr NttPlan pl = plan_for(LOGN);
    constexpr int NC = 1 << (LOGN - pl.lpre), T = pl.threads;
    HB_SHARED_U64(sm);
    const int row = blockIdx.x >> pl.lpre, B = blockIdx.x & ((1 << pl.lpre) - 1);
    const LimbConst lc = limbs[io.limb(row)];
    fwd_passes<LOGN, 0>(sm, io, lc, row, B);
    for (int i = threadIdx.x; i < NC; i += T) io.store(row, B * NC + i, sm[sphys(i)], lc);
}

// ------------------------------------------------------------------------------------------
// inverse, fast path
// ------------------------------------------------------------------------------------------
template <int LOGN, int P, class IO>
HB_D void inv_pass(u64 *sm, const IO &io, const LimbConst &lc, int row, int B) {
    constexpr NttPlan pl = plan_for(LOGN);
    constexpr int LOGNC = LOGN - pl.lpre, NC = 1 << LOGNC, T = pl.threads;
    constexpr int K = inv_k(pl, P), S0 = inv_s0(pl, P);
    constexpr int NG = NC >> K;
    constexpr bool first = (P == 0), last = (P == pl.npass - 1);
    static_assert(!first || S0 == 0, "first inverse pass must be contiguous");
    const ulonglong2 *tw_pass = lc.inv + inv_pass_offset(pl, P);
    constexpr int stride = 1 << S0;

#pragma unroll 1
    for (int g = threadIdx.x; g < NG; g += T) {
        const int lo = g & ((1 << S0) - 1), hi = g >> S0;
        const int base = (hi << (S0 + K)) + lo;
        u64 v[1 << K];
        if constexpr (first) {
            lds_contig<K>(sm, base, v);
        } else {
#pragma unroll
            for (int j = 0; j < (1 << K); j++) v[j] = sm[sphys(base + (j << S0))];
        }

        inv_levels<K>(v, tw_pass + lo, stride, lc.nq, lc.q2);

        if constexpr (last) {
            if constexpr (pl.lpre == 0) {
#pragma unroll
                for (int j = 0; j < (1 << K); j++) {
                    const int i = base + (j << S0);
                    u64 x = approx_reduce(v[j], lc);                 // ntt.cpp:218
                    const ulonglong2 s = __ldg(lc.inv_scale + i);    // psi^{-i}/N, ntt.cpp:219-221
                    io.store(row, i, harvey_lazy(x, s.x, s.y, lc.nq), lc);
                }
            } else {
#pragma unroll
                for (int j = 0; j < (1 << K); j++) sm[sphys(base + (j << S0))] = v[j];
            }
        } else if constexpr (first) {
            sts_contig<K>(sm, base, v);
        } else {
#pragma unroll
            for (int j = 0; j < (1 << K); j++) sm[sphys(base + (j << S0))] = v[j];
        }
    }
}

template <int LOGN, int P, class IO>
HB_D void inv_passes(u64 *sm, const IO &io, const LimbConst &lc, int row, int B) {
    constexpr NttPlan pl = plan_for(LOGN);
    inv_pass<LOGN, P>(sm, io, lc, row, B);
    if constexpr (P + 1 < pl.npass) {
        __syncthreads();
        inv_passes<LOGN, P + 1>(sm, io, lc, row, B);
    }
}

template <int LOGN, class IO>
HB_GLOBAL(plan_for(LOGN).threads, plan_for(LOGN).min_blocks)
intt_fast_kernel(const IO io, const LimbConst *__restrict__ limbs) {
    constexpr NttPlan pl = plan_for(LOGN);
    constexpr int NC = 1 << (LOGN - pl.lpre), T = pl.threads;
    HB_SHARED_U64(sm);
    const int row = blockIdx.x >> pl.lpre, B = blockIdx.x & ((1 << pl.lpre) - 1);
    const LimbConst lc = limbs[io.limb(row)];
    for (int i = threadIdx.x; i < NC; i += T) sm[sphys(i)] = io.load(row, B * NC + i, lc);
    __syncthreads();
    inv_passes<LOGN, 0>(sm, io, lc, row, B);
    if constexpr (pl.lpre == 1) {
        // last stage (gap N/2) pairs word i of CTA 0 with word i of CTA 1.  Exchange the halves
        // through the row's exchange area (L2), combine into shared memory, and only after the
        // sibling has also finished reading write the final words (the area may be the output).
        __syncthreads();
        u64 *raw = io.raw(row);
        for (int i = threadIdx.x; i < NC; i += T) raw[B * NC + i] = sm[sphys(i)];
        hb_cluster_sync();
        const ulonglong2 *tw = lc.inv + inv_pass_offset(pl, pl.npass);
        for (int i = threadIdx.x; i < NC; i += T) {
            const u64 mine = sm[sphys(i)], other = hb_ldcg(raw + (1 - B) * NC + i);
            const u64 lo = B ? other : mine, hi = B ? mine : other;
            const ulonglong2 z = __ldg(tw + i);
            const u64 t = harvey_lazy(hi, z.x, z.y, lc.nq);
            sm[sphys(i)] = B ? (lo + lc.q2 - t) : (lo + t); // ntt.cpp:199-206, own half only
        }
        hb_cluster_sync();
        for (int i = threadIdx.x; i < NC; i += T) {
            const int gi = B * NC + i;
            const ulonglong2 s = __ldg(lc.inv_scale + gi);
            io.store(row, gi, harvey_lazy(approx_reduce(sm[sphys(i)], lc), s.x, s.y, lc.nq), lc);
        }
    }
}

// ------------------------------------------------------------------------------------------
// generic path: any logn <= 14, one level per shared-memory sweep, reference table order.
// Used for small rings (N < 1024) and as an on-device cross-check of the fast path.
// ------------------------------------------------------------------------------------------
template <class IO>
HB_GLOBAL(256, 1) ntt_fwd_generic_kernel(const IO io, const LimbConst *__restrict__ limbs, int logn) {
    HB_SHARED_U64(sm);
    const int n = 1 << logn, row = blockIdx.x;
    const LimbConst lc = limbs[io.limb(row)];
    for (int i = threadIdx.x; i < n; i += blockDim.x) sm[i] = io.load(row, i, lc);
    __syncthreads();
    for (int level = 1; level <= logn; level++) { // ntt.cpp:155-169
        const int gl = logn - level;              // log2(gap)
        for (int b = threadIdx.x; b < n / 2; b += blockDim.x) {
            const int blk = b >> gl, l = (blk << (gl + 1)) + (b & ((1 << gl) - 1));
            bfly(sm[l], sm[l + (1 << gl)], __ldg(lc.fwd_nat + (1 << (level - 1)) + blk), lc.nq, lc.q2);
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x) io.store(row, i, approx_reduce(sm[i], lc), lc);
}

template <class IO>
HB_GLOBAL(256, 1) intt_generic_kernel(const IO io, const LimbConst *__restrict__ limbs, int logn) {
    HB_SHARED_U64(sm);
    const int n = 1 << logn, row = blockIdx.x;
    const LimbConst lc = limbs[io.limb(row)];
    for (int i = threadIdx.x; i < n; i += blockDim.x) sm[i] = io.load(row, i, lc);
    __syncthreads();
    for (int s = 1; s <= logn; s++) { // folded form of ntt.cpp:185-212
        const int gl = s - 1;
        for (int b = threadIdx.x; b < n / 2; b += blockDim.x) {
            const int blk = b >> gl, j = b & ((1 << gl) - 1), l = (blk << (gl + 1)) + j;
            const int rev = gl ? (int)(__brev((unsigned)j) >> (32 - gl)) : 0;
            bfly(sm[l], sm[l + (1 << gl)], __ldg(lc.inv_nat + (1 << gl) - 1 + rev), lc.nq, lc.q2);
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const ulonglong2 sc = __ldg(lc.inv_scale + i);
        io.store(row, i, harvey_lazy(approx_reduce(sm[i], lc), sc.x, sc.y, lc.nq), lc);
    }
}

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
struct LaunchStats {
    unsigned long long launches = 0;
};

template <int LOGN, class IO>
inline cudaError_t launch_fwd_fast(cudaStream_t st, const IO &io, const LimbConst *limbs, int rows, LaunchStats &ls) {
    constexpr NttPlan pl = plan_for(LOGN);
    constexpr int smem = smem_words(1 << (LOGN - pl.lpre)) * 8;
    auto kern = ntt_fwd_fast_kernel<LOGN, IO>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    ls.launches++;
    if constexpr (pl.lpre == 1) {
        return HB_LAUNCH_CLUSTER(kern, (unsigned)(rows << 1), pl.threads, smem, st, 2, io, limbs);
    } else {
        HB_LAUNCH(kern, rows, pl.threads, smem, st, 1, io, limbs);
        return cudaGetLastError();
    }
}

template <int LOGN, class IO>
inline cudaError_t launch_inv_fast(cudaStream_t st, const IO &io, const LimbConst *limbs, int rows, LaunchStats &ls) {
    constexpr NttPlan pl = plan_for(LOGN);
    constexpr int smem = smem_words(1 << (LOGN - pl.lpre)) * 8;
    auto kern = intt_fast_kernel<LOGN, IO>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    ls.launches++;
    if constexpr (pl.lpre == 1) {
        return HB_LAUNCH_CLUSTER(kern, (unsigned)(rows << 1), pl.threads, smem, st, 2, io, limbs);
    } else {
        HB_LAUNCH(kern, rows, pl.threads, smem, st, 1, io, limbs);
        return cudaGetLastError();
    }
}

template <class IO>
inline cudaError_t launch_generic(bool forward, cudaStream_t st, unsigned logn, const IO &io,
                                  const LimbConst *limbs, int rows, LaunchStats &ls) {
    const int smem = (1 << logn) * 8;
    int threads = (1 << logn) / 2;
    threads = threads < 32 ? 32 : (threads > 256 ? 256 : threads);
    if (forward) {
        auto kern = ntt_fwd_generic_kernel<IO>;
        static int configured = 0;
        if (smem > configured) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e != cudaSuccess) return e;
            configured = smem;
        }
        HB_LAUNCH(kern, rows, threads, smem, st, 1, io, limbs, (int)logn);
    } else {
        auto kern = intt_generic_kernel<IO>;
        static int configured = 0;
        if (smem > configured) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e != cudaSuccess) return e;
            configured = smem;
        }
        HB_LAUNCH(kern, rows, threads, smem, st, 1, io, limbs, (int)logn);
    }
    ls.launches++;
    return cudaGetLastError();
}

// Dispatch on ring size.  `force_generic` routes sizes the generic kernel supports through it
// (parity cross-check).  Returns cudaErrorInvalidValue for unsupported sizes.
template <class IO>
inline cudaError_t launch_ntt(bool forward, cudaStream_t st, unsigned logn, const IO &io, const LimbConst *limbs,
                              int rows, bool force_generic, LaunchStats &ls) {
    if (rows <= 0) return cudaSuccess;
    if (logn < 1 || logn > kFastLogMax) return cudaErrorInvalidValue;
    if (logn < kFastLogMin || (force_generic && logn <= kGenericLogMax)) return launch_generic(forward, st, logn, io, limbs, rows, ls);
#define HB_CASE(LN)                                                                       \
    case LN:                                                                              \
        return forward ? launch_fwd_fast<LN>(st, io, limbs, rows, ls) : launch_inv_fast<LN>(st, io, limbs, rows, ls);
    switch (logn) {
        HB_CASE(10) HB_CASE(11) HB_CASE(12) HB_CASE(13) HB_CASE(14) HB_CASE(15)
    }
#undef HB_CASE
    return cudaErrorInvalidValue;
}

} // namespace hb
