// ntt_engine.cuh — negacyclic NTT / INTT kernels for sm_100a.
//
// Replaces ntt_negacyclic_inplace_lazy / intt_negacyclic_inplace_lazy of the reference
// (src/fhe/common/ntt.cpp:145-223).  The butterfly dataflow graph, the absence of reduction
// between levels, the final approximate reduction and the psi^{-i}/N scaling are reproduced
// operation for operation, so outputs are the same raw u64 words as the CPU path.  Only the
// *order* in which independent butterflies run is changed:
//
//   * a row (one limb of one polynomial, N words) lives in shared memory; each thread keeps 8 or
//     16 coefficients in registers and runs 3-4 levels on them ("pass"), so the row crosses
//     shared memory 2-3 times instead of log2(N) times;
//   * the inverse uses the folded form (no bit-reversal permutations);
//   * N <= 8192: PERSISTENT kernel, one CTA per resident slot looping over rows; the next row's
//     words are streamed into a second shared-memory buffer with cp.async while the current row
//     is being transformed, so HBM latency and the copy-in never stall the integer pipes;
//   * N = 16384: one CTA per row; N = 32768 (256 KiB, more than one SM holds): a 2-CTA
//     thread-block cluster per row, half a row each; the single cross-CTA level (gap N/2) is
//     computed from global/L2 reads on the way in (forward) or exchanged through the row's own
//     global words between two cluster barriers (inverse).
//
// Kernels are parameterised by an IO policy; the fused rescale / key-switch kernels (ops.cu)
// are instantiations with prologue/epilogue arithmetic inside the policy, so those values are
// read from and written to HBM exactly once per transform.
//
// IO policy interface (all __device__):
//   int        limb(int row)                              -> index into the LimbConst array
//   const u64 *src(int row)                               -> the row's N contiguous input words
//   u64        pre(int row, int i, u64 raw, LimbConst&)   -> input word i from the raw word read at src(row)[i]
//   void       store(int row, int i, u64 v, LimbConst&)   -> consumes output word i
//   u64       *raw(int row)                               -> row-sized exchange area in global memory
//                                                            (N = 32768 inverse only; may be the output row)
//   bool       vec                                        -> every row pointer is 16-byte aligned
//   void       store2(int row, int i, u64 v0, u64 v1, LimbConst&) -> words i (even) and i+1, used when vec
#pragma once
#include <type_traits>

#include "modarith.cuh"
#include "ntt_plan.h"

namespace hb {

template <class IO>
HB_D u64 io_load(const IO &io, int row, int i, const LimbConst &lc) {
#if defined(HB_ABL_NOLOAD) // ablation builds only: synthesise the input words, no global reads
    return io.pre(row, i, (u64)i * 0x9E3779B97F4A7C15ull + (u64)row, lc);
#else
    return io.pre(row, i, io.src(row)[i], lc);
#endif
}

// ------------------------------------------------------------------------------------------
// butterfly — ntt.cpp:161-167 (identical for both directions)
// ------------------------------------------------------------------------------------------
HB_D void bfly(u64 &lo, u64 &hi, const ulonglong2 tw, u64 nq, u64 q2) {
    u64 t = harvey_lazy(hi, tw.x, tw.y, nq);
    hi = lo + q2 - t;
    lo = lo + t;
}

// Where a pass takes its 2^K - 1 twiddle pairs from: the table in global memory (L1/L2), or a
// register copy the persistent kernels keep across rows for the contiguous pass, whose twiddles
// are private to the thread and identical for every row of the same limb.
struct TwTable {
    const ulonglong2 *p;
    int stride;
#if defined(HB_ABL_TWCONST) // ablation builds only (tools/ab_build.sh): one twiddle per pass, no table traffic
    HB_D ulonglong2 get(int) const { return __ldg(p); }
#else
    HB_D ulonglong2 get(int slot) const { return __ldg(p + slot * stride); }
#endif
};
template <int K>
struct TwRegs {
    ulonglong2 r[(1 << K) - 1];
    HB_D ulonglong2 get(int slot) const { return r[slot]; }
    HB_D void fill(const ulonglong2 *__restrict__ p, int stride) {
#pragma unroll
        for (int s = 0; s < (1 << K) - 1; s++) r[s] = __ldg(p + s * stride);
    }
};

// K forward levels on 2^K registers.  Level m pairs registers 2^(K-m) apart and uses twiddle
// slot (2^(m-1) - 1 + blk).
template <int K, int M = 1, class TW>
HB_D void fwd_levels(u64 (&v)[1 << K], const TW &tw, u64 nq, u64 q2) {
    constexpr int half = 1 << (K - M);
#pragma unroll
    for (int blk = 0; blk < (1 << (M - 1)); blk++) {
        const ulonglong2 z = tw.get((1 << (M - 1)) - 1 + blk);
#pragma unroll
        for (int jj = 0; jj < half; jj++) bfly(v[blk * 2 * half + jj], v[blk * 2 * half + jj + half], z, nq, q2);
    }
    if constexpr (M < K) fwd_levels<K, M + 1>(v, tw, nq, q2);
}

// K inverse (folded) stages on 2^K registers.  Stage m pairs registers 2^(m-1) apart; twiddle
// slot (2^(m-1) - 1 + jj) depends on the register's index modulo 2^(m-1).
template <int K, int M = 1, class TW>
HB_D void inv_levels(u64 (&v)[1 << K], const TW &tw, u64 nq, u64 q2) {
    constexpr int d = 1 << (M - 1);
#pragma unroll
    for (int jj = 0; jj < d; jj++) {
        const ulonglong2 z = tw.get(d - 1 + jj);
#pragma unroll
        for (int blk = 0; blk < (1 << (K - M)); blk++) bfly(v[blk * 2 * d + jj], v[blk * 2 * d + jj + d], z, nq, q2);
    }
    if constexpr (M < K) inv_levels<K, M + 1>(v, tw, nq, q2);
}

HB_D int sphys(int i) { return i + ((i >> 4) << 1); }

// padded offset of a stride that is a multiple of 16 words: sphys(b + j * s) == sphys(b) + j * sstride(s),
// so strided passes address shared memory as one base register plus immediates
HB_D constexpr int sstride(int s) { return s + (s >> 3); }

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

// stream a row of NC raw words into the padded shared-memory layout, 16 bytes per cp.async
template <int NC, int T>
HB_D void prefetch_row(u64 *sm, const u64 *__restrict__ src) {
    static_assert((NC / 2) % T == 0 && (2 * T) % 16 == 0, "whole 16-byte chunks per thread");
    u64 *const d = sm + sphys(2 * threadIdx.x);
    const u64 *const g = src + 2 * threadIdx.x;
#pragma unroll
    for (int k = 0; k < NC / 2 / T; k++) hb_cp_async16(d + k * sstride(2 * T), g + k * 2 * T);
    hb_cp_async_commit();
}

// the CTA's NC words between shared memory (padded) and the IO policy.  io.vec (uniform): every row
// pointer of the policy is 16-byte aligned, so rows move as 128-bit words (two coefficients).
template <int NC, int T, class IO>
HB_D void store_row(const u64 *sm, const IO &io, const LimbConst &lc, int row, int first_word) {
    static_assert(NC % (2 * T) == 0 && T % 16 == 0, "whole steps");
    if (io.vec) {
        const u64 *const s = sm + sphys(2 * threadIdx.x);
#pragma unroll
        for (int k = 0; k < NC / 2 / T; k++) {
            const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(s + k * sstride(2 * T));
            io.store2(row, first_word + 2 * threadIdx.x + k * 2 * T, v.x, v.y, lc);
        }
    } else {
        const u64 *const s = sm + sphys(threadIdx.x);
#pragma unroll
        for (int k = 0; k < NC / T; k++) io.store(row, first_word + threadIdx.x + k * T, s[k * sstride(T)], lc);
    }
}
template <int NC, int T, class IO>
HB_D void load_row(u64 *sm, const IO &io, const LimbConst &lc, int row, int first_word) {
    static_assert(NC % (2 * T) == 0 && T % 16 == 0, "whole steps");
    if (io.vec) {
        u64 *const d = sm + sphys(2 * threadIdx.x);
        const u64 *const g = io.src(row) + first_word + 2 * threadIdx.x;
#pragma unroll
        for (int k = 0; k < NC / 2 / T; k++) {
            const int i = first_word + 2 * threadIdx.x + k * 2 * T;
#if defined(HB_ABL_NOLOAD)
            ulonglong2 v = make_ulonglong2((u64)i * 0x9E3779B97F4A7C15ull + (u64)row, (u64)i);
#else
            ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(g + k * 2 * T);
#endif
            v.x = io.pre(row, i, v.x, lc);
            v.y = io.pre(row, i + 1, v.y, lc);
            *reinterpret_cast<ulonglong2 *>(d + k * sstride(2 * T)) = v;
        }
    } else {
        u64 *const d = sm + sphys(threadIdx.x);
#pragma unroll
        for (int k = 0; k < NC / T; k++) d[k * sstride(T)] = io_load(io, row, first_word + threadIdx.x + k * T, lc);
    }
}

// ------------------------------------------------------------------------------------------
// forward passes.  SRC: 0 = first pass reads global memory, 1 = first pass reads the raw words
// a prefetch left in shared memory.
// ------------------------------------------------------------------------------------------
// TWC: register twiddles for the contiguous pass (TwRegs<K>) or NoTw (read the table)
struct NoTw {};
template <int LOGN, int T, int P, int SRC, class IO, class TWC>
HB_D void fwd_pass(u64 *sm, const IO &io, const LimbConst &lc, int row, int B, const TWC &twc) {
    constexpr NttPlan pl = plan_for(LOGN);
    constexpr int LOGNC = LOGN - pl.lpre, NC = 1 << LOGNC;
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
        static_assert(last || GSL >= 4, "strided passes step by multiples of 16 words");
        u64 *const smb = sm + sphys(base); // strided passes: element j lives at smb[j * SJ]
        constexpr int SJ = sstride(1 << GSL);
        u64 v[1 << K];
        if constexpr (first && SRC == 1) {
            static_assert(pl.lpre == 0, "prefetched rows are whole rows");
#pragma unroll
            for (int j = 0; j < (1 << K); j++) v[j] = io.pre(row, base + (j << GSL), smb[j * SJ], lc);
        } else if constexpr (first) {
            if constexpr (pl.lpre == 0) {
#pragma unroll
                for (int j = 0; j < (1 << K); j++) v[j] = io_load(io, row, base + (j << GSL), lc);
            } else {
                // level 1 of the full row (gap N/2) is computed here by both CTAs of the row;
                // CTA B keeps the low (B = 0) or high (B = 1) output — ntt.cpp:161-167
                const ulonglong2 z = __ldg(lc.fwd);
#pragma unroll
                for (int j = 0; j < (1 << K); j++) {
                    const int i = base + (j << GSL);
                    u64 a = io_load(io, row, i, lc), b = io_load(io, row, i + NC, lc);
                    u64 t = harvey_lazy(b, z.x, z.y, lc.nq);
                    v[j] = B ? (a + lc.q2 - t) : (a + t);
                }
            }
        } else if constexpr (last) {
            lds_contig<K>(sm, base, v);
        } else {
#pragma unroll
            for (int j = 0; j < (1 << K); j++) v[j] = smb[j * SJ];
        }

        if constexpr (last && !std::is_same<TWC, NoTw>::value) {
            static_assert(NG == T, "register twiddles need one contiguous group per thread");
            fwd_levels<K>(v, twc, lc.nq, lc.q2);
        } else {
            fwd_levels<K>(v, TwTable{tw_pass + ((B << L0) + hb), stride}, lc.nq, lc.q2);
        }

        if constexpr (last) {
#pragma unroll
            for (int j = 0; j < (1 << K); j++) v[j] = approx_reduce(v[j], lc); // ntt.cpp:171-175
            sts_contig<K>(sm, base, v);
        } else {
#pragma unroll
            for (int j = 0; j < (1 << K); j++) smb[j * SJ] = v[j];
        }
    }
}

template <int LOGN, int T, int P, int SRC, class IO, class TWC>
HB_D void fwd_passes(u64 *sm, const IO &io, const LimbConst &lc, int row, int B, const TWC &twc) {
    constexpr NttPlan pl = plan_for(LOGN);
    fwd_pass<LOGN, T, P, SRC>(sm, io, lc, row, B, twc);
    if constexpr (P == 0 && pl.lpre == 1) {
        // both CTAs of the row have read all of it: from here on either may overwrite it
        // (in-place transforms store into the words the sibling CTA has just read)
        hb_cluster_sync();
    } else {
        __syncthreads();
    }
    if constexpr (P + 1 < pl.npass) fwd_passes<LOGN, T, P + 1, SRC>(sm, io, lc, row, B, twc);
}

// one CTA (or one CTA of a 2-CTA cluster) per row
template <int LOGN, class IO>
HB_GLOBAL(plan_for(LOGN).threads, plan_for(LOGN).min_blocks)
ntt_fwd_fast_kernel(const IO io, const LimbConst *__restrict__ limbs) {
    constexpr NttPlan pl = plan_for(LOGN);
    constexpr int NC = 1 << (LOGN - pl.lpre), T = pl.threads;
    HB_SHARED_U64(sm);
    const int row = blockIdx.x >> pl.lpre, B = blockIdx.x & ((1 << pl.lpre) - 1);
    const LimbConst lc = limbs[io.limb(row)];
    fwd_passes<LOGN, T, 0, 0>(sm, io, lc, row, B, NoTw{});
    store_row<NC, T>(sm, io, lc, row, B * NC);
}

// persistent, double-buffered: grid = resident CTA slots, rows visited with stride gridDim.x
template <int LOGN, class IO>
HB_GLOBAL(pipe_plan_for(LOGN).threads, pipe_plan_for(LOGN).min_blocks)
ntt_fwd_pipe_kernel(const IO io, const LimbConst *__restrict__ limbs, int rows) {
    constexpr PipePlan pp = pipe_plan_for(LOGN);
    constexpr int NC = 1 << LOGN, T = pp.threads, BUF = smem_words(NC);
    HB_SHARED_U64(smem);
    int row = blockIdx.x;
    if (row >= rows) return;
    prefetch_row<NC, T>(smem, io.src(row));
    // the last (contiguous) pass uses 15 twiddle pairs private to this thread and identical for
    // every row of a limb: they stay in registers for as long as the CTA keeps seeing that limb
    constexpr NttPlan pl = plan_for(LOGN);
    constexpr int KL = pl.k[pl.npass - 1], L0L = fwd_lambda0(pl, pl.npass - 1);
    constexpr bool kCache = (NC >> KL) == T;
    typename std::conditional<kCache, TwRegs<KL>, NoTw>::type twc;
    int cached_limb = -1;
    int cur = 0;
    for (; row < rows; row += gridDim.x) {
        u64 *sm = smem + cur * BUF;
        hb_cp_async_wait_all();
        __syncthreads(); // this row has landed; every thread is done with the other buffer
        const int next = row + gridDim.x;
        if (next < rows) prefetch_row<NC, T>(smem + (cur ^ 1) * BUF, io.src(next));
        const int limb = io.limb(row);
        const LimbConst lc = limbs[limb];
        if constexpr (kCache) {
            if (limb != cached_limb) {
                twc.fill(lc.fwd + fwd_pass_offset(pl, pl.npass - 1) + threadIdx.x, 1 << L0L);
                cached_limb = limb;
            }
        }
        fwd_passes<LOGN, T, 0, 1>(sm, io, lc, row, 0, twc);
        store_row<NC, T>(sm, io, lc, row, 0);
        cur ^= 1;
    }
}

// ------------------------------------------------------------------------------------------
// inverse passes
// ------------------------------------------------------------------------------------------
template <int LOGN, int T, int P, int SRC, class IO, class TWC>
HB_D void inv_pass(u64 *sm, const IO &io, const LimbConst &lc, int row, int B, const TWC &twc) {
    constexpr NttPlan pl = plan_for(LOGN);
    constexpr int LOGNC = LOGN - pl.lpre, NC = 1 << LOGNC;
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
        static_assert(first || S0 >= 4, "strided passes step by multiples of 16 words");
        u64 *const smb = sm + sphys(base);
        constexpr int SJ = sstride(1 << S0);
        u64 v[1 << K];
        if constexpr (first) {
            lds_contig<K>(sm, base, v);
            if constexpr (SRC == 1) { // raw prefetched words: apply the policy's input map
#pragma unroll
                for (int j = 0; j < (1 << K); j++) v[j] = io.pre(row, base + j, v[j], lc);
            }
        } else {
#pragma unroll
            for (int j = 0; j < (1 << K); j++) v[j] = smb[j * SJ];
        }

        if constexpr (last && !std::is_same<TWC, NoTw>::value) {
            static_assert(NG == T, "register twiddles need one group per thread");
            inv_levels<K>(v, twc, lc.nq, lc.q2);
        } else {
            inv_levels<K>(v, TwTable{tw_pass + lo, stride}, lc.nq, lc.q2);
        }

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
                for (int j = 0; j < (1 << K); j++) smb[j * SJ] = v[j];
            }
        } else if constexpr (first) {
            sts_contig<K>(sm, base, v);
        } else {
#pragma unroll
            for (int j = 0; j < (1 << K); j++) smb[j * SJ] = v[j];
        }
    }
}

template <int LOGN, int T, int P, int SRC, class IO, class TWC>
HB_D void inv_passes(u64 *sm, const IO &io, const LimbConst &lc, int row, int B, const TWC &twc) {
    constexpr NttPlan pl = plan_for(LOGN);
    inv_pass<LOGN, T, P, SRC>(sm, io, lc, row, B, twc);
    if constexpr (P + 1 < pl.npass) {
        __syncthreads();
        inv_passes<LOGN, T, P + 1, SRC>(sm, io, lc, row, B, twc);
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
    load_row<NC, T>(sm, io, lc, row, B * NC);
    __syncthreads();
    inv_passes<LOGN, T, 0, 0>(sm, io, lc, row, B, NoTw{});
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

template <int LOGN, class IO>
HB_GLOBAL(pipe_plan_for(LOGN).threads, pipe_plan_for(LOGN).min_blocks)
intt_pipe_kernel(const IO io, const LimbConst *__restrict__ limbs, int rows) {
    constexpr PipePlan pp = pipe_plan_for(LOGN);
    constexpr int NC = 1 << LOGN, T = pp.threads, BUF = smem_words(NC);
    HB_SHARED_U64(smem);
    int row = blockIdx.x;
    if (row >= rows) return;
    prefetch_row<NC, T>(smem, io.src(row));
    // the last inverse pass (largest gaps) uses twiddle pairs private to this thread and identical
    // for every row of a limb: kept in registers while the CTA keeps seeing that limb
    constexpr NttPlan pl = plan_for(LOGN);
    constexpr int KL = inv_k(pl, pl.npass - 1), S0L = inv_s0(pl, pl.npass - 1);
    constexpr bool kCache = (NC >> KL) == T;
    typename std::conditional<kCache, TwRegs<KL>, NoTw>::type twc;
    int cached_limb = -1;
    int cur = 0;
    for (; row < rows; row += gridDim.x) {
        u64 *sm = smem + cur * BUF;
        hb_cp_async_wait_all();
        __syncthreads();
        const int next = row + gridDim.x;
        if (next < rows) prefetch_row<NC, T>(smem + (cur ^ 1) * BUF, io.src(next));
        const int limb = io.limb(row);
        const LimbConst lc = limbs[limb];
        if constexpr (kCache) {
            if (limb != cached_limb) {
                twc.fill(lc.inv + inv_pass_offset(pl, pl.npass - 1) + threadIdx.x, 1 << S0L);
                cached_limb = limb;
            }
        }
        inv_passes<LOGN, T, 0, 1>(sm, io, lc, row, 0, twc); // the last pass stores straight from registers
        cur ^= 1;
    }
}

// ------------------------------------------------------------------------------------------
// generic path: any logn <= 14, one level per shared-memory sweep, reference table order.
// Used for small rings (N < 1024) and as an on-device cross-check of the fast paths.
// ------------------------------------------------------------------------------------------
template <class IO>
HB_GLOBAL(256, 1) ntt_fwd_generic_kernel(const IO io, const LimbConst *__restrict__ limbs, int logn) {
    HB_SHARED_U64(sm);
    const int n = 1 << logn, row = blockIdx.x;
    const LimbConst lc = limbs[io.limb(row)];
    for (int i = threadIdx.x; i < n; i += blockDim.x) sm[i] = io_load(io, row, i, lc);
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
    for (int i = threadIdx.x; i < n; i += blockDim.x) sm[i] = io_load(io, row, i, lc);
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

struct LaunchEnv {
    cudaStream_t stream;
    int sm_count;       // persistent grids are sized from it
    bool force_generic; // parity cross-check path
    bool pipeline;      // use the persistent double-buffered kernels where they exist
    LaunchStats *stats;
};

template <int LOGN, bool FWD, class IO>
constexpr auto fast_kernel() {
    if constexpr (FWD) return &ntt_fwd_fast_kernel<LOGN, IO>;
    else return &intt_fast_kernel<LOGN, IO>;
}
template <int LOGN, bool FWD, class IO>
constexpr auto pipe_kernel() {
    if constexpr (FWD) return &ntt_fwd_pipe_kernel<LOGN, IO>;
    else return &intt_pipe_kernel<LOGN, IO>;
}

// Opt in to `smem` bytes of dynamic shared memory and ask for just enough carveout for `blocks`
// resident CTAs: whatever is left of the SM's 228 KB stays L1, which the twiddle tables live in.
template <class K>
inline cudaError_t configure_smem(K kern, int smem, int blocks) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    int pct = (int)(((long long)blocks * (smem + 1024) * 100 + 228 * 1024 - 1) / (228 * 1024));
    if (pct > 100) pct = 100;
    if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    return e;
}

template <int LOGN, bool FWD, class IO>
inline cudaError_t launch_fast(const LaunchEnv &env, const IO &io, const LimbConst *limbs, int rows) {
    constexpr NttPlan pl = plan_for(LOGN);
    constexpr int smem = smem_words(1 << (LOGN - pl.lpre)) * 8;
    auto kern = fast_kernel<LOGN, FWD, IO>();
    static bool configured = false;
    if (!configured) {
        cudaError_t e = configure_smem(kern, smem, pl.min_blocks);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    env.stats->launches++;
    if constexpr (pl.lpre == 1) {
        return HB_LAUNCH_CLUSTER(kern, (unsigned)(rows << 1), pl.threads, smem, env.stream, 2, io, limbs);
    } else {
        HB_LAUNCH(kern, rows, pl.threads, smem, env.stream, 1, io, limbs);
        return cudaGetLastError();
    }
}

template <int LOGN, bool FWD, class IO>
inline cudaError_t launch_pipe(const LaunchEnv &env, const IO &io, const LimbConst *limbs, int rows) {
    constexpr PipePlan pp = pipe_plan_for(LOGN);
    constexpr int smem = 2 * smem_words(1 << LOGN) * 8;
    auto kern = pipe_kernel<LOGN, FWD, IO>();
    static bool configured = false;
    if (!configured) {
        cudaError_t e = configure_smem(kern, smem, pp.min_blocks);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    // balanced persistent grid: every CTA visits ceil(rows / slots) rows or one fewer
    const int slots = env.sm_count * pp.min_blocks;
    const int per = (rows + slots - 1) / slots;
    const int grid = (rows + per - 1) / per;
    env.stats->launches++;
    HB_LAUNCH(kern, grid, pp.threads, smem, env.stream, 1, io, limbs, rows);
    return cudaGetLastError();
}

template <class IO>
inline cudaError_t launch_generic(bool forward, const LaunchEnv &env, unsigned logn, const IO &io, const LimbConst *limbs,
                                  int rows) {
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
        HB_LAUNCH(kern, rows, threads, smem, env.stream, 1, io, limbs, (int)logn);
    } else {
        auto kern = intt_generic_kernel<IO>;
        static int configured = 0;
        if (smem > configured) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e != cudaSuccess) return e;
            configured = smem;
        }
        HB_LAUNCH(kern, rows, threads, smem, env.stream, 1, io, limbs, (int)logn);
    }
    env.stats->launches++;
    return cudaGetLastError();
}

// Dispatch on ring size.  `aligned16`: every src(row) is 16-byte aligned (cp.async granularity).
template <class IO>
inline cudaError_t launch_ntt(bool forward, const LaunchEnv &env, unsigned logn, const IO &io, const LimbConst *limbs, int rows,
                              bool aligned16) {
    if (rows <= 0) return cudaSuccess;
    if (logn < 1 || logn > kFastLogMax) return cudaErrorInvalidValue;
    if (logn < kFastLogMin || (env.force_generic && logn <= kGenericLogMax)) return launch_generic(forward, env, logn, io, limbs, rows);
    const bool pipe = env.pipeline && aligned16 && logn <= (unsigned)kPipeLogMax;
#define HB_CASE_PIPE(LN)                                                                                     \
    case LN:                                                                                                 \
        if (pipe) return forward ? launch_pipe<LN, true>(env, io, limbs, rows) : launch_pipe<LN, false>(env, io, limbs, rows); \
        return forward ? launch_fast<LN, true>(env, io, limbs, rows) : launch_fast<LN, false>(env, io, limbs, rows);
#define HB_CASE_FAST(LN)                                                                                     \
    case LN:                                                                                                 \
        return forward ? launch_fast<LN, true>(env, io, limbs, rows) : launch_fast<LN, false>(env, io, limbs, rows);
    switch (logn) {
        HB_CASE_PIPE(10) HB_CASE_PIPE(11) HB_CASE_PIPE(12) HB_CASE_PIPE(13) HB_CASE_FAST(14) HB_CASE_FAST(15)
    }
#undef HB_CASE_PIPE
#undef HB_CASE_FAST
    return cudaErrorInvalidValue;
}

} // namespace hb
