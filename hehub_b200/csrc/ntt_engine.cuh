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
//   * the contiguous pass moves the row between global and shared memory warp by warp (each warp
//     owns 32 << K adjacent words), and when the neighbouring pass keeps the same number of words
//     per thread the exchange between the two stays inside a warp: N = 4096 needs ONE CTA barrier
//     per transform;
//   * N <= 8192: one CTA per row (a 2-CTA cluster for launches with few rows: latency plans, ntt_plan.h);
//     N = 16384 and N = 32768 (256 KiB, more than one SM holds): a 2-CTA
//     thread-block cluster per row, half a row each; the single cross-CTA level (gap N/2) is
//     computed from global/L2 reads on the way in (forward) or finished from the sibling's shared
//     memory (distributed shared memory) on the way out (inverse).
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
//   void       prefetch(int row, int first, int nwords)   -> optional: called once per CTA before the first pass
//   u64 fetch(int row, int i) / ulonglong2 fetch2(row, i) -> optional: the raw word(s) of src(row) the transform reads as word i
//                                                            (i, i + 1), for policies that gather through a permutation
//   bool       vec                                        -> every row pointer is 16-byte aligned
//   void       store2(int row, int i, u64 v0, u64 v1, LimbConst&) -> words i (even) and i+1, used when vec
#pragma once
#include <atomic>
#include <type_traits>
#include <utility>

#include "modarith.cuh"
#include "ntt_plan.h"

namespace hb {

// Probe builds only (tools/ab_build.sh phase -DHB_PHASE_CLOCK; tools/phase_probe.py): thread 0 of every CTA records the SM
// clock at marked points into a log the probe installs, 32 slots per CTA.
#if defined(HB_PHASE_CLOCK) && !defined(HB_KERNEL_SIM)
static __device__ unsigned long long *hb_phase_ptr;
#define HB_PHASE(slot)                                                                         \
    do {                                                                                       \
        if (threadIdx.x == 0 && hb_phase_ptr) hb_phase_ptr[(size_t)blockIdx.x * 32 + (slot)] = clock64(); \
    } while (0)
#else
#define HB_PHASE(slot) ((void)0)
#endif

// optional policy hooks: fetch(row, i) / fetch2(row, i) replace the streaming load of word i (words i, i + 1; i even) of
// src(row) — a policy that gathers its input through a permutation (Galois automorphisms fused into the key switch)
template <class IO, class = void>
struct io_has_fetch : std::false_type {};
template <class IO>
struct io_has_fetch<IO, std::void_t<decltype(std::declval<const IO &>().fetch(0, 0))>> : std::true_type {};

template <class IO>
HB_D u64 io_load(const IO &io, int row, int i, const LimbConst &lc) {
#if defined(HB_ABL_NOLOAD) // ablation builds only: synthesise the input words, no global reads
    return io.pre(row, i, (u64)i * 0x9E3779B97F4A7C15ull + (u64)row, lc);
#else
    if constexpr (io_has_fetch<IO>::value) return io.pre(row, i, io.fetch(row, i), lc);
    else return io.pre(row, i, hb_ld_stream(io.src(row) + i), lc);
#endif
}

// optional: a policy whose kernel should be compiled for MORE resident CTAs per SM than its plan asks for (fewer registers), where
// that was measured to pay: specialise IoResidency<IO>
template <class IO>
struct IoResidency {
    static constexpr int extra(int /*logn*/, int /*mode*/, bool /*forward*/) { return 0; }
};

// optional policy hook: prefetch(row, first_word, nwords) is called once per CTA before the first pass
template <class IO, class = void>
struct io_has_prefetch : std::false_type {};
template <class IO>
struct io_has_prefetch<IO, std::void_t<decltype(std::declval<const IO &>().prefetch(0, 0, 0))>> : std::true_type {};

// the twiddle tables of a plan family (ntt_plan.h: 0 throughput, 1 latency, 2 single-ciphertext key switch)
template <int MODE>
HB_D const ulonglong2 *tw_fwd(const LimbConst &lc) { return MODE == 0 ? lc.fwd : (MODE == 1 ? lc.fwd_lat : lc.fwd_lat2); }
template <int MODE>
HB_D const ulonglong2 *tw_inv(const LimbConst &lc) { return MODE == 0 ? lc.inv : (MODE == 1 ? lc.inv_lat : lc.inv_lat2); }

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
#if defined(HB_ABL_TWNOLOAD) // ablation builds only (tools/ab_build.sh): no table load at all (wrong results, timing only)
    HB_D ulonglong2 get(int slot) const { return make_ulonglong2((u64)(size_t)p + slot, (u64)stride); }
#elif defined(HB_ABL_TWCONST) // ablation builds only: one twiddle per pass, no table traffic
    HB_D ulonglong2 get(int) const { return __ldg(p); }
#else
    HB_D ulonglong2 get(int slot) const { return __ldg(p + slot * stride); }
#endif
};
// the same from a table staged in shared memory (mode-2 plans, ntt_plan.h)
struct TwSmem {
    const ulonglong2 *p;
    int stride;
    HB_D ulonglong2 get(int slot) const { return p[slot * stride]; }
};
// Staged tables live behind the CTA's row in shared memory: forward block, inverse local block, inverse cross block.
template <int LOGN, int MODE>
HB_D const ulonglong2 *stage_fwd(const u64 *sm) {
    return reinterpret_cast<const ulonglong2 *>(sm + smem_words(1 << (LOGN - plan_for(LOGN, true, MODE).lpre)));
}
template <int LOGN, int MODE>
HB_D const ulonglong2 *stage_inv(const u64 *sm) { return stage_fwd<LOGN, MODE>(sm) + fwd_stage_block(plan_for(LOGN, true, MODE)); }
template <int LOGN, int MODE>
constexpr int kStagedSmemBytes = smem_words(1 << (LOGN - plan_for(LOGN, true, MODE).lpre)) * 8 +
                                 16 * (fwd_stage_block(plan_for(LOGN, true, MODE)) + inv_stage_local(plan_for(LOGN, false, MODE)) +
                                       inv_stage_cross_block(plan_for(LOGN, false, MODE)));
HB_CX bool staged_mode(int mode) { return mode == 2; }
// copy `entries` (a multiple of T) 16-byte entries; every thread moves entries tid, tid + T, ...
template <int T>
HB_D void stage_copy(const ulonglong2 *dst, const ulonglong2 *src, int entries_const) {
    ulonglong2 *d = const_cast<ulonglong2 *>(dst) + threadIdx.x;
    const ulonglong2 *s = src + threadIdx.x;
#pragma unroll
    for (int e = 0; e < entries_const; e += T) hb_cp_async16(d + e, s + e);
}
// the forward tables of limb `lc` for CTA B of its cluster / the inverse tables
template <int LOGN, int T, int MODE>
HB_D void stage_fwd_tables(const u64 *sm, const LimbConst &lc, int B) {
    constexpr NttPlan pl = plan_for(LOGN, true, MODE);
    static_assert(kStagePad % T == 0, "blocks are whole multiples of the CTA size");
    stage_copy<T>(stage_fwd<LOGN, MODE>(sm), tw_fwd<MODE>(lc) + B * fwd_stage_block(pl), fwd_stage_block(pl));
}
template <int LOGN, int T, int MODE>
HB_D void stage_inv_tables(const u64 *sm, const LimbConst &lc, int B) {
    constexpr NttPlan pl = plan_for(LOGN, false, MODE);
    stage_copy<T>(stage_inv<LOGN, MODE>(sm), tw_inv<MODE>(lc), inv_stage_local(pl));
    stage_copy<T>(stage_inv<LOGN, MODE>(sm) + inv_stage_local(pl), tw_inv<MODE>(lc) + inv_stage_local(pl) + B * inv_stage_cross_block(pl),
                  inv_stage_cross_block(pl));
}

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

// ------------------------------------------------------------------------------------------
// warp-local row movement.  In the contiguous pass every thread owns 2^K adjacent words, so a warp
// owns the 32 << K adjacent words of its lanes: those move between global and shared memory with
// coalesced 128-bit accesses of the warp itself and only a warp barrier, no CTA barrier.
// ------------------------------------------------------------------------------------------
template <int K, class IO>
HB_D void warp_store(const u64 *sm, const IO &io, const LimbConst &lc, int row, int first_word, int wbase) {
    const int lane = threadIdx.x & 31;
    if (io.vec) {
        const u64 *const s = sm + sphys(wbase + 2 * lane);
#pragma unroll
        for (int k = 0; k < (1 << K) / 2; k++) {
            const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(s + k * sstride(64));
            io.store2(row, first_word + wbase + 2 * lane + 64 * k, v.x, v.y, lc);
        }
    } else {
        const u64 *const s = sm + sphys(wbase + lane);
#pragma unroll
        for (int k = 0; k < (1 << K); k++) io.store(row, first_word + wbase + lane + 32 * k, s[k * sstride(32)], lc);
    }
}
template <int K, class IO>
HB_D void warp_load(u64 *sm, const IO &io, const LimbConst &lc, int row, int first_word, int wbase) {
    const int lane = threadIdx.x & 31;
    if (io.vec) {
        u64 *const d = sm + sphys(wbase + 2 * lane);
        const u64 *const g = io.src(row) + first_word + wbase + 2 * lane;
#pragma unroll
        for (int k = 0; k < (1 << K) / 2; k++) {
            const int i = first_word + wbase + 2 * lane + 64 * k;
#if defined(HB_ABL_NOLOAD)
            ulonglong2 v = make_ulonglong2((u64)i * 0x9E3779B97F4A7C15ull + (u64)row, (u64)i);
#else
            ulonglong2 v;
            if constexpr (io_has_fetch<IO>::value) v = io.fetch2(row, i);
            else v = hb_ld_stream2(g + 64 * k);
#endif
            v.x = io.pre(row, i, v.x, lc);
            v.y = io.pre(row, i + 1, v.y, lc);
            *reinterpret_cast<ulonglong2 *>(d + k * sstride(64)) = v;
        }
    } else {
        u64 *const d = sm + sphys(wbase + lane);
#pragma unroll
        for (int k = 0; k < (1 << K); k++) d[k * sstride(32)] = io_load(io, row, first_word + wbase + lane + 32 * k, lc);
    }
}

// Two consecutive passes that keep the same number of words per thread, one of them the contiguous
// pass, exchange data only inside groups of 2^K consecutive threads (K <= 5: inside a warp), so
// the barrier between them is a warp barrier.
HB_CX bool warp_local_pair(int ka, int kb) { return ka == kb && ka <= 5; }

// ------------------------------------------------------------------------------------------
// forward passes: gaps shrink, the last pass is contiguous and stores the row
// ------------------------------------------------------------------------------------------
template <int LOGN, int T, int P, int MODE, class IO>
HB_D void fwd_pass(u64 *sm, const IO &io, const LimbConst &lc, int row, int B) {
    constexpr NttPlan pl = plan_for(LOGN, true, MODE);
    constexpr int LOGNC = LOGN - pl.lpre, NC = 1 << LOGNC;
    constexpr int K = pl.k[P], L0 = fwd_lambda0(pl, P), GSL = LOGNC - L0 - K; // log2(smallest gap)
    constexpr int NG = NC >> K;
    constexpr bool first = (P == 0), last = (P == pl.npass - 1);
    static_assert(!(first && pl.xchg), "pass 0 of an exchanging plan is fwd_cross_pass");
    static_assert(!last || GSL == 0, "last forward pass must be contiguous");
    // strides of at least 16 words address shared memory as one base register plus immediates (sstride); the thin passes of
    // the latency plans may step by less and pay the padding arithmetic per access
    constexpr bool imm = GSL >= 4;
    static_assert(T % 32 == 0 && (NG % T == 0 || (T % NG == 0 && NG % 32 == 0)), "whole warps in every step (idle warps allowed)");
    const ulonglong2 *tw_pass = tw_fwd<MODE>(lc) + fwd_pass_offset(pl, P);
    constexpr int stride = 1 << (pl.lpre + L0);
    constexpr int SJ = imm ? sstride(1 << (imm ? GSL : 4)) : 0;

#pragma unroll 1
    for (int g = threadIdx.x; g < NG; g += T) {
        const int lo = g & ((1 << GSL) - 1), hb = g >> GSL;
        const int base = (hb << (LOGNC - L0)) + lo;
        u64 *const smb = sm + sphys(base); // strided passes: element j lives at smb[j * SJ]
        u64 v[1 << K];
        if constexpr (first) {
            if constexpr (pl.lpre == 0) {
#pragma unroll
                for (int j = 0; j < (1 << K); j++) v[j] = io_load(io, row, base + (j << GSL), lc);
            } else {
                // level 1 of the full row (gap N/2) is computed here by both CTAs of the row;
                // CTA B keeps the low (B = 0) or high (B = 1) output — ntt.cpp:161-167
                const ulonglong2 z = __ldg(tw_fwd<MODE>(lc));
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
        } else if constexpr (imm) {
#pragma unroll
            for (int j = 0; j < (1 << K); j++) v[j] = smb[j * SJ];
        } else {
#pragma unroll
            for (int j = 0; j < (1 << K); j++) v[j] = sm[sphys(base + (j << GSL))];
        }

        if constexpr (staged_mode(MODE)) fwd_levels<K>(v, TwSmem{stage_fwd<LOGN, MODE>(sm) + fwd_stage_offset(pl, P) + hb, 1 << L0}, lc.nq, lc.q2);
        else fwd_levels<K>(v, TwTable{tw_pass + ((B << L0) + hb), stride}, lc.nq, lc.q2);

        if constexpr (last) {
#pragma unroll
            for (int j = 0; j < (1 << K); j++) v[j] = approx_reduce(v[j], lc); // ntt.cpp:171-175
            sts_contig<K>(sm, base, v); // own words only
            hb_syncwarp();
            warp_store<K>(sm, io, lc, row, B * NC, (g - (int)(threadIdx.x & 31)) << K);
        } else if constexpr (imm) {
#pragma unroll
            for (int j = 0; j < (1 << K); j++) smb[j * SJ] = v[j];
        } else {
#pragma unroll
            for (int j = 0; j < (1 << K); j++) sm[sphys(base + (j << GSL))] = v[j];
        }
    }
}

// Exchanging plans (xchg = 1): pass 0 is a pass of the FULL row shared by the 2^lpre CTAs of the cluster.  Group t holds
// the words t + j * G (G = N >> K); CTA B takes the groups [B, B + 1) * G / C, reads them from global memory, runs
// the K levels (the first lpre of them pair words of different CTAs) and hands every result to the CTA that owns its
// position: word i belongs to CTA i / NC, at offset i % NC of that CTA's shared memory (st.shared::cluster).
template <int LOGN, int T, int MODE, class IO>
HB_D void fwd_cross_pass(u64 *sm, const IO &io, const LimbConst &lc, int row, int B) {
    constexpr NttPlan pl = plan_for(LOGN, true, MODE);
    constexpr int K = pl.k[0], C = 1 << pl.lpre, KL = K - pl.lpre; // KL: levels of this pass that stay inside the owner
    constexpr int LOGG = LOGN - K, GPC = (1 << LOGG) / C;
    static_assert(K >= pl.lpre && LOGG >= 4, "the cross pass covers every cross-CTA level; strides are multiples of 16 words");
    static_assert(GPC % T == 0 && T % 32 == 0, "whole warps in every step");
    constexpr int SJ = sstride(1 << LOGG);
    const ulonglong2 *tw_pass = tw_fwd<MODE>(lc);
#pragma unroll 1
    for (int g = threadIdx.x; g < GPC; g += T) {
        const int t = B * GPC + g;
        u64 v[1 << K];
#pragma unroll
        for (int j = 0; j < (1 << K); j++) v[j] = io_load(io, row, t + (j << LOGG), lc);
        if constexpr (staged_mode(MODE)) fwd_levels<K>(v, TwSmem{stage_fwd<LOGN, MODE>(sm), 1}, lc.nq, lc.q2);
        else fwd_levels<K>(v, TwTable{tw_pass, 1}, lc.nq, lc.q2);
        // every CTA of the cluster is running before its shared memory is written (the arrive is at kernel start)
        if (g == (int)threadIdx.x) hb_cluster_wait();
        u64 *const smb = sm + sphys(t);
#pragma unroll
        for (int o = 0; o < C; o++) {
            if (o == B) {
#pragma unroll
                for (int jj = 0; jj < (1 << KL); jj++) smb[jj * SJ] = v[(o << KL) + jj];
            } else {
#pragma unroll
                for (int jj = 0; jj < (1 << KL); jj++) hb_st_dsmem(smb + jj * SJ, o, v[(o << KL) + jj]);
            }
        }
    }
}

#if defined(HB_ABL_SHFL) && !defined(HB_KERNEL_SIM)
// A/B build only (tools/ab_build.sh shfl -DHB_ABL_SHFL; profiles/r3_headline_ab.md): the exchange between the last two forward
// passes (both 4 levels wide, so it stays inside groups of 16 lanes) done with warp shuffles instead of shared memory: a
// 16 x 16 transpose of 64-bit words in four butterfly stages.  new v[i] of lane l = old v[l] of lane i.
HB_D void shfl_transpose16(u64 (&v)[16]) {
    const unsigned lane = threadIdx.x & 31;
#pragma unroll
    for (int b = 8; b; b >>= 1) {
        const bool up = lane & b;
#pragma unroll
        for (int r = 0; r < 16; r++) {
            if (r & b) continue;
            const u64 send = up ? v[r] : v[r | b];
            const u64 got = __shfl_xor_sync(0xffffffffu, send, b);
            if (up) v[r] = got;
            else v[r | b] = got;
        }
    }
}
// the last two passes of a {.., 4, 4} forward plan fused over registers
template <int LOGN, int T, int MODE, class IO>
HB_D void fwd_last_two_shfl(u64 *sm, const IO &io, const LimbConst &lc, int row, int B) {
    constexpr NttPlan pl = plan_for(LOGN, true, MODE);
    constexpr int LOGNC = LOGN - pl.lpre, NC = 1 << LOGNC, P = pl.npass - 2;
    constexpr int L0 = fwd_lambda0(pl, P), NG = NC >> 4;
    static_assert(pl.k[P] == 4 && pl.k[P + 1] == 4 && LOGNC - L0 - 4 == 4, "two 4-level passes at the end");
    const ulonglong2 *twa = tw_fwd<MODE>(lc) + fwd_pass_offset(pl, P), *twb = tw_fwd<MODE>(lc) + fwd_pass_offset(pl, P + 1);
#pragma unroll 1
    for (int g = threadIdx.x; g < NG; g += T) {
        const int lo = g & 15, hb = g >> 4, base = (hb << 8) + lo;
        const u64 *const smb = sm + sphys(base);
        u64 v[16];
#pragma unroll
        for (int j = 0; j < 16; j++) v[j] = smb[j * sstride(16)];
        fwd_levels<4>(v, TwTable{twa + ((B << L0) + hb), 1 << (pl.lpre + L0)}, lc.nq, lc.q2);
        shfl_transpose16(v); // now v[i] = word g * 16 + i
        fwd_levels<4>(v, TwTable{twb + ((B << (L0 + 4)) + g), 1 << (pl.lpre + L0 + 4)}, lc.nq, lc.q2);
#pragma unroll
        for (int j = 0; j < 16; j++) v[j] = approx_reduce(v[j], lc);
        hb_syncwarp(); // every lane has read its strided words before the contiguous ones are written
        sts_contig<4>(sm, g << 4, v);
        hb_syncwarp();
        warp_store<4>(sm, io, lc, row, B * NC, (g - (int)(threadIdx.x & 31)) << 4);
    }
}
#endif

template <int LOGN, int T, int P, int MODE, class IO>
HB_D void fwd_passes(u64 *sm, const IO &io, const LimbConst &lc, int row, int B) {
    constexpr NttPlan pl = plan_for(LOGN, true, MODE);
#if defined(HB_ABL_SHFL) && !defined(HB_KERNEL_SIM)
    if constexpr (P == pl.npass - 2 && P > 0 && pl.k[P] == 4 && pl.k[P + 1] == 4 && LOGN == 12 && MODE == 0) {
        fwd_last_two_shfl<LOGN, T, MODE>(sm, io, lc, row, B);
        return;
    } else
#endif
    if constexpr (P == 0 && pl.xchg) {
        fwd_cross_pass<LOGN, T, MODE>(sm, io, lc, row, B);
        HB_PHASE(16);
        hb_cluster_sync(); // every word has reached its owner (and all of the row has been read: in-place stores may follow)
        HB_PHASE(17);
        fwd_passes<LOGN, T, 1, MODE>(sm, io, lc, row, B);
    } else {
        fwd_pass<LOGN, T, P, MODE>(sm, io, lc, row, B);
        HB_PHASE(17 + P);
        if constexpr (P + 1 < pl.npass) {
            if constexpr (P == 0 && pl.lpre == 1) {
                // both CTAs of the row have read all of it: from here on either may overwrite it
                // (in-place transforms store into the words the sibling CTA has just read)
                hb_cluster_sync();
            } else if constexpr (P + 2 == pl.npass && warp_local_pair(pl.k[P], pl.k[P + 1])) {
                hb_syncwarp();
            } else {
                __syncthreads();
            }
            fwd_passes<LOGN, T, P + 1, MODE>(sm, io, lc, row, B);
        }
    }
}

// Latency plans: every twiddle this thread will use, requested into L1 before the kernel waits for its predecessor (the
// tables are written once, when the modulus is first seen).  A row alone on its SMs has nothing to hide a table load
// behind: with five thin passes the L2 latency of each pass's twiddles was a third of the kernel (profiles/r3_latency_plans.md).
template <int LOGN, int T, int MODE, int P = 0>
HB_D void prefetch_fwd_twiddles(const LimbConst &lc, int B) {
    constexpr NttPlan pl = plan_for(LOGN, true, MODE);
    const ulonglong2 *tab = tw_fwd<MODE>(lc);
    if constexpr (P == 0 && pl.xchg) {
        for (int s = threadIdx.x; s < (1 << pl.k[0]) - 1; s += T) hb_prefetch_l1(tab + s); // CTA-uniform: a few lines
    } else {
        constexpr int LOGNC = LOGN - pl.lpre, K = pl.k[P], L0 = fwd_lambda0(pl, P), GSL = LOGNC - L0 - K, NG = (1 << LOGNC) >> K;
        const ulonglong2 *tw = tab + fwd_pass_offset(pl, P);
        for (int g = threadIdx.x; g < NG; g += T)
#pragma unroll
            for (int s = 0; s < (1 << K) - 1; s++) hb_prefetch_l1(tw + ((B << L0) + (g >> GSL)) + ((size_t)s << (pl.lpre + L0)));
    }
    if constexpr (P + 1 < pl.npass) prefetch_fwd_twiddles<LOGN, T, MODE, P + 1>(lc, B);
}
template <int LOGN, int T, int MODE, int P = 0>
HB_D void prefetch_inv_twiddles(const LimbConst &lc, int B) {
    constexpr NttPlan pl = plan_for(LOGN, false, MODE);
    const ulonglong2 *tab = tw_inv<MODE>(lc);
    if constexpr (pl.xchg && P == pl.npass - 1) {
        constexpr int K = pl.k[0], LOGG = LOGN - K, GPC = (1 << LOGG) >> pl.lpre;
        const ulonglong2 *tw = tab + inv_pass_offset(pl, P);
        for (int g = threadIdx.x; g < GPC; g += T) {
            const int t = B * GPC + g;
#pragma unroll
            for (int s = 0; s < (1 << K) - 1; s++) hb_prefetch_l1(tw + t + ((size_t)s << LOGG));
#pragma unroll
            for (int j = 0; j < (1 << K); j++) hb_prefetch_l1(lc.inv_scale + t + (j << LOGG));
        }
    } else {
        constexpr int LOGNC = LOGN - pl.lpre, K = inv_k(pl, P), S0 = inv_s0(pl, P), NG = (1 << LOGNC) >> K;
        const ulonglong2 *tw = tab + inv_pass_offset(pl, P);
        for (int g = threadIdx.x; g < NG; g += T)
#pragma unroll
            for (int s = 0; s < (1 << K) - 1; s++) hb_prefetch_l1(tw + (g & ((1 << S0) - 1)) + ((size_t)s << S0));
    }
    if constexpr (P + 1 < pl.npass) prefetch_inv_twiddles<LOGN, T, MODE, P + 1>(lc, B);
}

#if defined(HB_ABL_TMA) && !defined(HB_KERNEL_SIM)
// A/B build only (tools/ab_build.sh tma -DHB_ABL_TMA; profiles/r3_headline_ab.md): the row is brought into an unpadded
// staging area of shared memory by ONE bulk asynchronous copy (cp.async.bulk, the 1-D TMA path; SASS: UBLKCP) issued by one
// thread and awaited on an mbarrier; pass 0 then reads shared memory instead of global memory.
template <class IO>
struct StagedIO : IO {
    const u64 *stage;
    HB_D u64 fetch(int, int i) const { return stage[i]; }
    HB_D ulonglong2 fetch2(int, int i) const { return *reinterpret_cast<const ulonglong2 *>(stage + i); }
};
HB_D void hb_tma_row_load(u64 *stage, const u64 *src, unsigned bytes) {
    __shared__ __align__(8) unsigned long long mbar;
    const unsigned bar = (unsigned)__cvta_generic_to_shared(&mbar), dst = (unsigned)__cvta_generic_to_shared(stage);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
                     "r"(bar)
                     : "memory");
    }
    unsigned done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar) : "memory");
    }
}
#endif

// one CTA (or one CTA of a 2-CTA cluster) per row
template <int LOGN, class IO, int MODE = 0>
HB_GLOBAL(plan_for(LOGN, true, MODE).threads, plan_for(LOGN, true, MODE).min_blocks + IoResidency<IO>::extra(LOGN, MODE, true))
ntt_fwd_fast_kernel(const IO io, const LimbConst *__restrict__ limbs) {
    // One row per CTA on purpose: resident CTAs walking several rows each (with or without a start
    // skew between the CTAs of an SM) measured 15-20 % slower than letting the block scheduler hand
    // out rows (profiles/r1_plan_sweep.md).
    constexpr NttPlan pl = plan_for(LOGN, true, MODE);
    constexpr int T = pl.threads;
    HB_SHARED_U64(sm);
    const int row = blockIdx.x >> pl.lpre, B = blockIdx.x & ((1 << pl.lpre) - 1);
    const LimbConst lc = limbs[io.limb(row)]; // uploaded when the chain was built: safe to read before the wait
    if constexpr (pl.xchg) hb_cluster_arrive(); // matched by the wait in front of the first remote store
#if HB_LAT_PREFETCH
    if constexpr (MODE == 1 && pl.xchg) prefetch_fwd_twiddles<LOGN, T, MODE>(lc, B);
#endif
    if constexpr (staged_mode(MODE)) stage_fwd_tables<LOGN, T, MODE>(sm, lc, B); // static data: on its way before the row may be read
    hb_pdl_wait();
    if constexpr (staged_mode(MODE)) {
        hb_cp_async_wait_all();
        __syncthreads();
    }
    if constexpr (io_has_prefetch<IO>::value) io.prefetch(row, B << (LOGN - pl.lpre), 1 << (LOGN - pl.lpre));
#if defined(HB_ABL_TMA) && !defined(HB_KERNEL_SIM)
    if constexpr (pl.lpre == 0 && !io_has_fetch<IO>::value) {
        if (io.vec) { // 16-byte aligned rows only
            u64 *stage = sm + smem_words(1 << LOGN);
            hb_tma_row_load(stage, io.src(row), (unsigned)(8u << LOGN));
            fwd_passes<LOGN, T, 0, MODE>(sm, StagedIO<IO>{io, stage}, lc, row, B);
            return;
        }
    }
#endif
    fwd_passes<LOGN, T, 0, MODE>(sm, io, lc, row, B);
}

// ------------------------------------------------------------------------------------------
// inverse passes: the first pass is contiguous and loads the row, gaps grow
// ------------------------------------------------------------------------------------------
template <int LOGN, int T, int P, int MODE, class IO>
HB_D void inv_pass(u64 *sm, const IO &io, const LimbConst &lc, int row, int B) {
    constexpr NttPlan pl = plan_for(LOGN, false, MODE);
    constexpr int LOGNC = LOGN - pl.lpre, NC = 1 << LOGNC;
    constexpr int K = inv_k(pl, P), S0 = inv_s0(pl, P);
    constexpr int NG = NC >> K;
    constexpr bool first = (P == 0), last = (P == pl.npass - 1);
    static_assert(!first || S0 == 0, "first inverse pass must be contiguous");
    constexpr bool imm = S0 >= 4; // see fwd_pass
    static_assert(T % 32 == 0 && (NG % T == 0 || (T % NG == 0 && NG % 32 == 0)), "whole warps in every step (idle warps allowed)");
    const ulonglong2 *tw_pass = tw_inv<MODE>(lc) + inv_pass_offset(pl, P);
    constexpr int stride = 1 << S0;
    constexpr int SJ = imm ? sstride(1 << (imm ? S0 : 4)) : 0;

#pragma unroll 1
    for (int g = threadIdx.x; g < NG; g += T) {
        const int lo = g & ((1 << S0) - 1), hi = g >> S0;
        const int base = (hi << (S0 + K)) + lo;
        u64 *const smb = sm + sphys(base);
        u64 v[1 << K];
        if constexpr (first) {
            warp_load<K>(sm, io, lc, row, B * NC, (g - (int)(threadIdx.x & 31)) << K);
            hb_syncwarp();
            lds_contig<K>(sm, base, v);
        } else if constexpr (imm) {
#pragma unroll
            for (int j = 0; j < (1 << K); j++) v[j] = smb[j * SJ];
        } else {
#pragma unroll
            for (int j = 0; j < (1 << K); j++) v[j] = sm[sphys(base + (j << S0))];
        }

        if constexpr (staged_mode(MODE)) inv_levels<K>(v, TwSmem{stage_inv<LOGN, MODE>(sm) + inv_pass_offset(pl, P) + lo, stride}, lc.nq, lc.q2);
        else inv_levels<K>(v, TwTable{tw_pass + lo, stride}, lc.nq, lc.q2);

        if constexpr (last && pl.lpre == 0) {
#pragma unroll
            for (int j = 0; j < (1 << K); j++) {
                const int i = base + (j << S0);
                u64 x = approx_reduce(v[j], lc);                 // ntt.cpp:218
                const ulonglong2 s = __ldg(lc.inv_scale + i);    // psi^{-i}/N, ntt.cpp:219-221
                io.store(row, i, harvey_lazy(x, s.x, s.y, lc.nq), lc);
            }
        } else if constexpr (first) {
            sts_contig<K>(sm, base, v);
        } else if constexpr (imm) {
#pragma unroll
            for (int j = 0; j < (1 << K); j++) smb[j * SJ] = v[j];
        } else {
#pragma unroll
            for (int j = 0; j < (1 << K); j++) sm[sphys(base + (j << S0))] = v[j];
        }
    }
}

// Exchanging plans: the last inverse pass is a pass of the FULL row (stage gaps G ... N/2, the last lpre of them pair words
// of different CTAs).  CTA B takes the groups [B, B + 1) * G / C, gathers word t + j * G from the shared memory of the
// CTA that owns it (ld.shared::cluster), runs the K stages and finishes every word: approximate reduction and the
// psi^{-i}/N scaling (ntt.cpp:214-221) — no butterfly is computed twice and the row never leaves the cluster.
template <int LOGN, int T, int MODE, class IO>
HB_D void inv_cross_pass(u64 *sm, const IO &io, const LimbConst &lc, int row, int B) {
    constexpr NttPlan pl = plan_for(LOGN, false, MODE);
    constexpr int K = pl.k[0], C = 1 << pl.lpre, KL = K - pl.lpre;
    constexpr int LOGG = LOGN - K, GPC = (1 << LOGG) / C;
    static_assert(K >= pl.lpre && LOGG >= 4, "the cross pass covers every cross-CTA stage; strides are multiples of 16 words");
    static_assert(GPC % T == 0 && T % 32 == 0, "whole warps in every step");
    static_assert(inv_s0(pl, pl.npass - 1) == LOGG, "the cross pass is the last pass of the mirrored list");
    constexpr int SJ = sstride(1 << LOGG);
    const ulonglong2 *tw_pass = tw_inv<MODE>(lc) + inv_pass_offset(pl, pl.npass - 1);
#pragma unroll 1
    for (int g = threadIdx.x; g < GPC; g += T) {
        const int t = B * GPC + g;
        const u64 *const smb = sm + sphys(t);
        u64 v[1 << K];
#pragma unroll
        for (int o = 0; o < C; o++) {
            if (o == B) {
#pragma unroll
                for (int jj = 0; jj < (1 << KL); jj++) v[(o << KL) + jj] = smb[jj * SJ];
            } else {
#pragma unroll
                for (int jj = 0; jj < (1 << KL); jj++) v[(o << KL) + jj] = hb_ld_dsmem(smb + jj * SJ, o);
            }
        }
        if constexpr (staged_mode(MODE)) {
            static_assert(!staged_mode(MODE) || GPC == T, "staged cross tables are laid out one group per thread");
            const ulonglong2 *xb = stage_inv<LOGN, MODE>(sm) + inv_stage_local(pl) + threadIdx.x; // [slot][thread], then the scales [j][thread]
            inv_levels<K>(v, TwSmem{xb, T}, lc.nq, lc.q2);
#pragma unroll
            for (int j = 0; j < (1 << K); j++) {
                const ulonglong2 s = xb[((1 << K) - 1 + j) * T]; // psi^{-i}/N, ntt.cpp:219-221
                io.store(row, t + (j << LOGG), harvey_lazy(approx_reduce(v[j], lc), s.x, s.y, lc.nq), lc);
            }
        } else {
            inv_levels<K>(v, TwTable{tw_pass + t, 1 << LOGG}, lc.nq, lc.q2);
#pragma unroll
            for (int j = 0; j < (1 << K); j++) {
                const int i = t + (j << LOGG);
                u64 x = approx_reduce(v[j], lc);              // ntt.cpp:218
                const ulonglong2 s = __ldg(lc.inv_scale + i); // psi^{-i}/N, ntt.cpp:219-221
                io.store(row, i, harvey_lazy(x, s.x, s.y, lc.nq), lc);
            }
        }
    }
}

template <int LOGN, int T, int P, int MODE, class IO>
HB_D void inv_passes(u64 *sm, const IO &io, const LimbConst &lc, int row, int B) {
    constexpr NttPlan pl = plan_for(LOGN, false, MODE);
    if constexpr (pl.xchg && P == pl.npass - 1) {
        hb_cluster_sync(); // the CTA-local stages are done everywhere and their words visible to the cluster
        inv_cross_pass<LOGN, T, MODE>(sm, io, lc, row, B);
        hb_cluster_sync(); // a sibling may still be reading this CTA's shared memory
    } else {
        inv_pass<LOGN, T, P, MODE>(sm, io, lc, row, B);
        if constexpr (P + 1 < pl.npass) {
            if constexpr (pl.xchg && P + 2 == pl.npass) {
                // the cluster barrier in front of the cross pass orders this CTA's stores too
            } else if constexpr (P == 0 && warp_local_pair(inv_k(pl, 0), inv_k(pl, 1))) {
                hb_syncwarp();
            } else {
                __syncthreads();
            }
            inv_passes<LOGN, T, P + 1, MODE>(sm, io, lc, row, B);
        }
    }
}

// ------------------------------------------------------------------------------------------
// Register hand-over between the inverse and the forward transform of one cluster (the fused key-switch kernels of
// ks_pair.cuh).  With exchanging plans whose cross passes are equally wide, the thread that finishes word t + j * G of the
// inverse transform (its cross pass comes last) is the thread that needs that word for the forward transform (its cross
// pass comes first): the row goes from one transform to the next without leaving the registers of the cluster.
// ------------------------------------------------------------------------------------------
template <int LOGN, int MODE>
HB_CX bool plans_hand_over() {
    constexpr NttPlan f = plan_for(LOGN, true, MODE), i = plan_for(LOGN, false, MODE);
    return f.xchg && i.xchg && f.lpre == i.lpre && f.k[0] == i.k[0] && f.threads == i.threads &&
           ((1 << (LOGN - f.k[0])) >> f.lpre) == f.threads; // one group of the cross pass per thread
}
template <int LOGN, int MODE>
constexpr int kHandOverWords = 1 << plan_for(LOGN, false, MODE).k[0];

// the CTA-local passes of the inverse (all but the cross pass)
template <int LOGN, int T, int P, int MODE, class IO>
HB_D void inv_local_passes(u64 *sm, const IO &io, const LimbConst &lc, int row, int B) {
    constexpr NttPlan pl = plan_for(LOGN, false, MODE);
    static_assert(pl.xchg && pl.npass >= 2, "exchanging plans only");
    inv_pass<LOGN, T, P, MODE>(sm, io, lc, row, B);
    HB_PHASE(8 + P);
    if constexpr (P + 2 < pl.npass) {
        if constexpr (P == 0 && warp_local_pair(inv_k(pl, 0), inv_k(pl, 1))) hb_syncwarp();
        else __syncthreads();
        inv_local_passes<LOGN, T, P + 1, MODE>(sm, io, lc, row, B);
    }
}
// The cross pass of the inverse for group t = B * T + threadIdx.x: gathers, runs the stages, finishes the words
// (ntt.cpp:214-221).  On return v[j] is word t + (j << LOGG) of the transform, in [0, 2q).  The caller has run a cluster
// barrier since the local passes; this function ARRIVES at the next one as soon as its remote reads are done (the matching
// wait is the one in front of the next remote write: fwd_cross_from_regs).
template <int LOGN, int T, int MODE>
HB_D void inv_cross_to_regs(const u64 *sm, const LimbConst &lc, int B, u64 (&v)[kHandOverWords<LOGN, MODE>]) {
    constexpr NttPlan pl = plan_for(LOGN, false, MODE);
    constexpr int K = pl.k[0], C = 1 << pl.lpre, KL = K - pl.lpre, LOGG = LOGN - K;
    static_assert(plans_hand_over<LOGN, MODE>() && LOGG >= 4, "see plans_hand_over");
    constexpr int SJ = sstride(1 << LOGG);
    const int t = B * T + (int)threadIdx.x;
    const u64 *const smb = sm + sphys(t);
#pragma unroll
    for (int o = 0; o < C; o++) {
        if (o == B) {
#pragma unroll
            for (int jj = 0; jj < (1 << KL); jj++) v[(o << KL) + jj] = smb[jj * SJ];
        } else {
#pragma unroll
            for (int jj = 0; jj < (1 << KL); jj++) v[(o << KL) + jj] = hb_ld_dsmem(smb + jj * SJ, o);
        }
    }
    hb_cluster_arrive();
    if constexpr (staged_mode(MODE)) {
        const ulonglong2 *xb = stage_inv<LOGN, MODE>(sm) + inv_stage_local(pl) + threadIdx.x; // [slot][thread], then the scales [j][thread]
        inv_levels<K>(v, TwSmem{xb, T}, lc.nq, lc.q2);
#pragma unroll
        for (int j = 0; j < (1 << K); j++) {
            const ulonglong2 s = xb[((1 << K) - 1 + j) * T]; // psi^{-i}/N, ntt.cpp:219-221
            v[j] = harvey_lazy(approx_reduce(v[j], lc), s.x, s.y, lc.nq);
        }
    } else {
        inv_levels<K>(v, TwTable{tw_inv<MODE>(lc) + inv_pass_offset(pl, pl.npass - 1) + t, 1 << LOGG}, lc.nq, lc.q2);
#pragma unroll
        for (int j = 0; j < (1 << K); j++) {
            const ulonglong2 s = __ldg(lc.inv_scale + t + (j << LOGG)); // psi^{-i}/N, ntt.cpp:219-221
            v[j] = harvey_lazy(approx_reduce(v[j], lc), s.x, s.y, lc.nq);
        }
    }
}
// The cross pass of the forward transform from registers: v[j] is input word t + (j << LOGG).  WAITS for the cluster
// barrier every CTA arrived at after its last read of shared memory, then scatters the results to their owners.  The caller
// runs a cluster barrier before the local passes (fwd_passes<.., 1, ..>).
template <int LOGN, int T, int MODE>
HB_D void fwd_cross_from_regs(u64 *sm, const LimbConst &lc, int B, u64 (&v)[kHandOverWords<LOGN, MODE>]) {
    constexpr NttPlan pl = plan_for(LOGN, true, MODE);
    constexpr int K = pl.k[0], C = 1 << pl.lpre, KL = K - pl.lpre, LOGG = LOGN - K;
    constexpr int SJ = sstride(1 << LOGG);
    const int t = B * T + (int)threadIdx.x;
    if constexpr (staged_mode(MODE)) fwd_levels<K>(v, TwSmem{stage_fwd<LOGN, MODE>(sm), 1}, lc.nq, lc.q2);
    else fwd_levels<K>(v, TwTable{tw_fwd<MODE>(lc), 1}, lc.nq, lc.q2);
    hb_cluster_wait();
    u64 *const smb = sm + sphys(t);
#pragma unroll
    for (int o = 0; o < C; o++) {
        if (o == B) {
#pragma unroll
            for (int jj = 0; jj < (1 << KL); jj++) smb[jj * SJ] = v[(o << KL) + jj];
        } else {
#pragma unroll
            for (int jj = 0; jj < (1 << KL); jj++) hb_st_dsmem(smb + jj * SJ, o, v[(o << KL) + jj]);
        }
    }
}

template <int LOGN, class IO, int MODE = 0>
HB_GLOBAL(plan_for(LOGN, false, MODE).threads, plan_for(LOGN, false, MODE).min_blocks)
intt_fast_kernel(const IO io, const LimbConst *__restrict__ limbs) {
    constexpr NttPlan pl = plan_for(LOGN, false, MODE);
    constexpr int NC = 1 << (LOGN - pl.lpre), T = pl.threads;
    HB_SHARED_U64(sm);
    const int row = blockIdx.x >> pl.lpre, B = blockIdx.x & ((1 << pl.lpre) - 1);
    const LimbConst lc = limbs[io.limb(row)];
#if HB_LAT_PREFETCH
    if constexpr (MODE == 1 && pl.xchg) prefetch_inv_twiddles<LOGN, T, MODE>(lc, B);
#endif
    if constexpr (staged_mode(MODE)) stage_inv_tables<LOGN, T, MODE>(sm, lc, B);
    hb_pdl_wait();
    if constexpr (staged_mode(MODE)) {
        hb_cp_async_wait_all();
        __syncthreads();
    }
    inv_passes<LOGN, T, 0, MODE>(sm, io, lc, row, B);
    if constexpr (pl.lpre == 1 && !pl.xchg) {
        // Last stage (gap N/2) pairs word i of CTA 0 with word i of CTA 1 — ntt.cpp:199-206.  Each CTA
        // takes half of the offsets, reads its own words from shared memory and the sibling's through
        // distributed shared memory, and finishes BOTH outputs of every pair (approximate reduction and
        // psi^{-i}/N scaling, ntt.cpp:214-221): no butterfly is computed twice and the row never makes
        // a round trip through L2.
        hb_cluster_sync();
        const ulonglong2 *tw = tw_inv<MODE>(lc) + inv_pass_offset(pl, pl.npass);
        for (int i = B * (NC / 2) + 2 * (int)threadIdx.x; i < (B + 1) * (NC / 2); i += 2 * T) {
            const u64 *p = sm + sphys(i); // i even: words i, i+1 are adjacent
            const ulonglong2 mine = *reinterpret_cast<const ulonglong2 *>(p), other = hb_ld_dsmem2(p, 1 - B);
            const ulonglong2 lo = B ? other : mine, hi = B ? mine : other;
            u64 out_lo[2], out_hi[2];
#pragma unroll
            for (int w = 0; w < 2; w++) {
                const ulonglong2 z = __ldg(tw + i + w);
                const u64 l = w ? lo.y : lo.x, h = w ? hi.y : hi.x;
                const u64 t = harvey_lazy(h, z.x, z.y, lc.nq);
                const ulonglong2 s0 = __ldg(lc.inv_scale + i + w), s1 = __ldg(lc.inv_scale + NC + i + w);
                out_lo[w] = harvey_lazy(approx_reduce(l + t, lc), s0.x, s0.y, lc.nq);
                out_hi[w] = harvey_lazy(approx_reduce(l + lc.q2 - t, lc), s1.x, s1.y, lc.nq);
            }
            if (io.vec) {
                io.store2(row, i, out_lo[0], out_lo[1], lc);
                io.store2(row, NC + i, out_hi[0], out_hi[1], lc);
            } else {
                io.store(row, i, out_lo[0], lc);
                io.store(row, i + 1, out_lo[1], lc);
                io.store(row, NC + i, out_hi[0], lc);
                io.store(row, NC + i + 1, out_hi[1], lc);
            }
        }
        hb_cluster_sync(); // the sibling may still be reading this CTA's shared memory
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
    hb_pdl_wait();
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
    hb_pdl_wait();
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
    LaunchStats *stats;
    int latency_rows;   // launches with at most this many rows take the latency plan (default: half the SM count)
    int latency2_rows;  // ... and with at most this many the mode-2 plans (default: a tenth of the SM count: 8-CTA clusters, one CTA per SM)
    int device;         // kernel attributes (dynamic shared memory opt-in, carveout) are per device
};

// cudaFuncSetAttribute acts on the current device only: remember per kernel instantiation and per device
// what has been configured (a process may hold one context per GPU, cpp/hehub/backend.h set_context).
constexpr int kMaxDevices = 64;
struct PerDeviceConfig {
    std::atomic<int> smem[kMaxDevices];
    // true when the kernel already accepts `bytes` of dynamic shared memory on `device`
    bool covers(int device, int bytes) const { return smem[device & (kMaxDevices - 1)].load(std::memory_order_acquire) >= bytes; }
    void record(int device, int bytes) { smem[device & (kMaxDevices - 1)].store(bytes, std::memory_order_release); }
};

template <int LOGN, bool FWD, class IO, int MODE>
constexpr auto fast_kernel() {
    if constexpr (FWD) return &ntt_fwd_fast_kernel<LOGN, IO, MODE>;
    else return &intt_fast_kernel<LOGN, IO, MODE>;
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

template <int LOGN, bool FWD, int MODE, class IO>
inline cudaError_t launch_fast_mode(const LaunchEnv &env, const IO &io, const LimbConst *limbs, int rows) {
    constexpr NttPlan pl = plan_for(LOGN, FWD, MODE);
#if defined(HB_ABL_TMA) // A/B build: room for the unpadded staging copy of the row behind the working area
    constexpr int smem = smem_words(1 << (LOGN - pl.lpre)) * 8 + ((FWD && pl.lpre == 0) ? (8 << LOGN) : 0);
#else
    constexpr int smem = staged_mode(MODE) ? kStagedSmemBytes<LOGN, MODE> : smem_words(1 << (LOGN - pl.lpre)) * 8;
#endif
    auto kern = fast_kernel<LOGN, FWD, IO, MODE>();
    static PerDeviceConfig configured; // zero-initialised; racing contexts at worst configure twice (idempotent)
    if (!configured.covers(env.device, smem)) {
        cudaError_t e = configure_smem(kern, smem, pl.min_blocks + (FWD ? IoResidency<IO>::extra(LOGN, MODE, true) : 0));
        if (e != cudaSuccess) return e;
        configured.record(env.device, smem);
    }
    env.stats->launches++;
    if constexpr (pl.lpre >= 1) {
        return HB_LAUNCH_CLUSTER(kern, (unsigned)(rows << pl.lpre), pl.threads, smem, env.stream, 1 << pl.lpre, io, limbs);
    } else {
        HB_LAUNCH(kern, rows, pl.threads, smem, env.stream, 1, io, limbs);
        return cudaGetLastError();
    }
}

// few rows (a single ciphertext, a key): the latency plan where the ring size has one, see ntt_plan.h
template <int LOGN, bool FWD, class IO>
inline cudaError_t launch_fast(const LaunchEnv &env, const IO &io, const LimbConst *limbs, int rows) {
    if constexpr (has_latency_plan(LOGN)) {
        // Row counts up to which the latency plan wins (profiles/r3_latency_plans.md), as multiples of the option
        // latency_rows (default half the SM count = 74): N <= 8192 (thin 4-CTA plans) 74 rows; N = 16384 (thin 8-CTA plan)
        // 37 forward rows (64 for the key switch's fan-out, which then runs four CTAs per SM: one wave); N = 32768 (8-CTA plan at 80 registers: finer grain, less of the last wave idles) 148 forward
        // rows; inverse transforms of N >= 16384 (they need their registers: one or two CTAs per SM) 18 rows.
        const long long limit = LOGN <= 13 ? env.latency_rows
                                           : (!FWD ? env.latency_rows / 4
                                                   : (LOGN == 14 ? env.latency_rows * (IoResidency<IO>::extra(LOGN, 1, true) ? 7 : 4) / 8 : 2ll * env.latency_rows));
        // a handful of rows (N = 4096 / 8192): 8-CTA clusters with their tables staged in shared memory, while every CTA gets an
        // SM of its own (mode 2, ntt_plan.h): one N = 8192 row 6.1 -> 4.9 us forward, 6.7 -> 5.2 us inverse
        if constexpr (has_latency2_plan(LOGN)) {
            if (rows <= env.latency2_rows) return launch_fast_mode<LOGN, FWD, 2>(env, io, limbs, rows);
        }
        if (rows <= limit) return launch_fast_mode<LOGN, FWD, 1>(env, io, limbs, rows);
    }
    return launch_fast_mode<LOGN, FWD, 0>(env, io, limbs, rows);
}

template <bool FWD, class IO>
inline cudaError_t launch_generic(const LaunchEnv &env, unsigned logn, const IO &io, const LimbConst *limbs, int rows) {
    const int smem = (1 << logn) * 8;
    int threads = (1 << logn) / 2;
    threads = threads < 32 ? 32 : (threads > 256 ? 256 : threads);
    auto kern = [] {
        if constexpr (FWD) return &ntt_fwd_generic_kernel<IO>;
        else return &intt_generic_kernel<IO>;
    }();
    static PerDeviceConfig configured;
    if (!configured.covers(env.device, smem)) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        configured.record(env.device, smem);
    }
    HB_LAUNCH(kern, rows, threads, smem, env.stream, 1, io, limbs, (int)logn);
    env.stats->launches++;
    return cudaGetLastError();
}

// Dispatch on ring size (the policy's `vec` flag says whether rows move as 128-bit words).  The direction is a
// template parameter so that a policy is only instantiated for the direction it is used in.
template <bool FWD, class IO>
inline cudaError_t launch_ntt(const LaunchEnv &env, unsigned logn, const IO &io, const LimbConst *limbs, int rows) {
    if (rows <= 0) return cudaSuccess;
    if (logn < 1 || logn > kFastLogMax) return cudaErrorInvalidValue;
    if (logn < kFastLogMin || (env.force_generic && logn <= kGenericLogMax)) return launch_generic<FWD>(env, logn, io, limbs, rows);
#define HB_CASE_FAST(LN) \
    case LN: return launch_fast<LN, FWD>(env, io, limbs, rows);
    switch (logn) {
        HB_CASE_FAST(10) HB_CASE_FAST(11) HB_CASE_FAST(12) HB_CASE_FAST(13) HB_CASE_FAST(14) HB_CASE_FAST(15)
    }
#undef HB_CASE_FAST
    return cudaErrorInvalidValue;
}

} // namespace hb
