// ks_pair.cuh — key switch and modulus drop of a FEW ciphertexts as cluster launches (included by ops.cu).
//
// The wave path (ops.cu) runs ckks::mult as six launches — tensor, INTT, fan-out NTT, inner product, INTT of the P limb,
// forward + drop epilogue — each of which streams a batch through HBM at full width.  A few ciphertexts per call have
// nothing to stream: every launch is a handful of rows on an almost idle GPU, and a row costs what its chain of dependent
// butterflies costs on the few SMs it runs on (profiles/r4_pair_path.md: the integer multiplier of those SMs and the
// shared-memory round trips between passes, not the launches).  Here the same arithmetic runs as TWO launches of 8-CTA
// cluster kernels (mode-2 plans, ntt_plan.h: twice the SMs per row, twiddle tables staged in shared memory before the row
// arrives), each an inverse transform handed over IN REGISTERS to forward transforms (ntt_engine.cuh, plans_hand_over),
// with the coefficient-wise steps riding in the loads and stores:
//
//   ks_fan_kernel   cluster (b, p, chunk):  in[b][p]  (the tensor product's c1 * c1' computed in the load: ckks/arith.cpp:55-62;
//                   or read through the Galois permutation: ckks::rotate / conjugate) -> INTT_{q_p}, strict
//                   (rgsw.cpp:103-105) -> registers -> for each target k != p of the chunk: NTT_{q_k} -> dec[b][p][k]
//                   (rgsw.cpp:108-119).  The cluster of chunk 0 also stores in[b][p] as dec[b][p][p], the diagonal term
//                   rgsw.cpp:99-101 keeps.  CTAs beyond those compute the tensor product's d0, d1 (SMs are idle anyway).
//   ks_drop_kernel  cluster (b, h, chunk):  e[h][P] = Mont128(sum_p dec[p][P] * key[p][h][P]) (rgsw.cpp:126-153), all its
//                   operands requested at once -> INTT_P, strict (rescaling.cpp:47-50) -> registers -> for each limb k of
//                   the chunk: centre / Barrett (rescaling.cpp:58-68) -> NTT_{q_k} -> store: (e[h][k] - NTT) / P
//                   (rescaling.cpp:73-74) + d_h[k] (ckks/arith.cpp:70-71), e[h][k] by the same inner product and the
//                   addend both computed BEFORE the transform they meet.
//                   With the ciphertext itself as the source of e (KsSrcPlain) the same kernel is
//                   ckks::rescale_inplace / bgv::mod_switch_inplace of a few ciphertexts in ONE launch.
//
// Every word is produced by the same sequence of operations as on the wave path (the policies below call the very
// functions the wave kernels call), so the results are the same raw words; the inverse transform of in[b][p] and the
// P-limb pipeline are REPEATED by the clusters of different chunks — free on an idle GPU, and the reason this form is
// only taken for small batches (pair_path_wanted, pair_targets_per_cluster in ops.cu).
// Measured, N = 8192, L = 4, one call: ckks::mult 33.2 -> 23.5 us, rotate 31.8 -> 22.4 us, rescale 14.5 -> 8.7 us.
#pragma once

#ifndef HB_PAIR_TRIGGER
#define HB_PAIR_TRIGGER 0 // launching the next grid early measured 4 us slower per pair (profiles/r4_pair_path.md)
#endif

namespace hb {

// ---- where the key switch's input polynomial in[b][p] (NTT form) comes from: aligned pairs of words ----
template <bool GALOIS>
struct KsInPlain {
    const u64 *in;
    size_t batch_stride;
    int logn;
    unsigned ginv;
    HB_D ulonglong2 pair(int b, int p, int i, const LimbConst &) const {
        const u64 *r = in + (size_t)b * batch_stride + ((size_t)p << logn);
        if constexpr (GALOIS) return galois_pair(r, (unsigned)i, ginv, logn, false);
        else return hb_ld_stream2(r + i);
    }
};
struct KsInTensor { // d2 = c1 * c1' (ckks/arith.cpp:55-62, the third polynomial of mult_low_level), never stored
    const u64 *ct1, *ct2; // [batch][2][L][N]
    int L, logn;
    HB_D ulonglong2 pair(int b, int p, int i, const LimbConst &lc) const {
        const size_t off = ((((size_t)b * 2 + 1) * L + p) << logn) + i;
        const ulonglong2 x = hb_ld_stream2(ct1 + off), y = hb_ld_stream2(ct2 + off);
        return make_ulonglong2(mul_hybrid_lazy(x.x, y.x, lc), mul_hybrid_lazy(x.y, y.y, lc));
    }
};

// ---- what is added to polynomial h of the result ----
template <bool GALOIS>
struct KsAddPlain {
    const u64 *addend;
    size_t batch_stride, poly_stride;
    int halves, logn;
    unsigned ginv;
    HB_D bool has(int h) const { return h < halves; }
    HB_D ulonglong2 pair(int b, int h, int k, int i, const LimbConst &) const {
        const u64 *r = addend + (size_t)b * batch_stride + (size_t)h * poly_stride + ((size_t)k << logn);
        if constexpr (GALOIS) return galois_pair(r, (unsigned)i, ginv, logn, true);
        else return hb_ld_ro2(r + i);
    }
};
// d0 = c0 * c0', d1 = c0 * c1' + c1 * c0' of pair b, limb k (ckks/arith.cpp:55-62) -> quad[b][0..1][k], words (i, i + 1).  Run by
// CTAs of the fan-out launch that have no transform to do (the GPU has SMs to spare while one ciphertext is switched), so that
// the drop launch reads its addend instead of multiplying for it.
HB_D void ks_tensor_addend_pair(const u64 *ct1, const u64 *ct2, u64 *quad, const LimbConst *limbs, int L, int logn, size_t unit) {
    const size_t i = (unit & (((size_t)1 << (logn - 1)) - 1)) * 2, row = unit >> (logn - 1); // row = b * L + k
    const size_t b = row / L, k = row - b * L;
    const LimbConst lc = limbs[k];
    const size_t o0 = (((b * 2) * L + k) << logn) + i, o1 = o0 + ((size_t)L << logn);
    const ulonglong2 a0 = hb_ld_stream2(ct1 + o0), b0 = hb_ld_stream2(ct2 + o0), a1 = hb_ld_stream2(ct1 + o1), b1 = hb_ld_stream2(ct2 + o1);
    *reinterpret_cast<ulonglong2 *>(quad + o0) = make_ulonglong2(mul_hybrid_lazy(a0.x, b0.x, lc), mul_hybrid_lazy(a0.y, b0.y, lc));
    *reinterpret_cast<ulonglong2 *>(quad + o1) =
        make_ulonglong2(add_lazy(mul_hybrid_lazy(a0.x, b1.x, lc), mul_hybrid_lazy(a1.x, b0.x, lc), lc.q2),
                        add_lazy(mul_hybrid_lazy(a0.y, b1.y, lc), mul_hybrid_lazy(a1.y, b0.y, lc), lc.q2));
}

// The words a thread moves in the contiguous pass of a transform (first pass of the inverse, last pass of the forward: the
// same words, warp_load / warp_store of ntt_engine.cuh): NP pairs (i0 + 64 k, i0 + 64 k + 1), k < NP.
template <int LOGN, int MODE>
constexpr int kPairsPerThread = (1 << plan_for(LOGN, true, MODE).k[plan_for(LOGN, true, MODE).npass - 1]) / 2;
template <int LOGN, int MODE>
HB_D int ks_first_pair(int B) {
    constexpr NttPlan pl = plan_for(LOGN, true, MODE);
    constexpr int K = pl.k[pl.npass - 1], NC = 1 << (LOGN - pl.lpre);
    static_assert(pl.k[pl.npass - 1] == inv_k(plan_for(LOGN, false, MODE), 0) && (NC >> K) == pl.threads, "one group per thread in the contiguous passes");
    const int lane = (int)threadIdx.x & 31;
    return B * NC + (((int)threadIdx.x - lane) << K) + 2 * lane;
}

// e[k] = words (i0 + 64 k, +1) of Mont128_{q}(sum_p dec[p][.] * key[p][h][.]) — rgsw.cpp:126-153; the sum is exact in 128
// bits and reduced once (ext_mac_unit, ops.cu).  dec_k: &dec[b][0][limb][i0]; key_hk: &key[0][h][limb][i0].  All the
// operands of a digit p (2 NP 128-bit loads) are requested before its products: the thread has nothing else to hide them behind.
template <int NP, int U>
HB_D void ks_mac_pairs(const u64 *dec_k, const u64 *key_hk, int L, int logn, const LimbConst &lc, ulonglong2 *e) {
    const size_t dec_row = (size_t)(L + 1) << logn, key_row = (size_t)(2 * (L + 1)) << logn;
    Acc128 acc[NP][2];
#pragma unroll
    for (int k = 0; k < NP; k++) {
        acc128_clear(acc[k][0]);
        acc128_clear(acc[k][1]);
    }
#pragma unroll U
    for (int p = 0; p < L; p++) {
        ulonglong2 d[NP], w[NP];
#pragma unroll
        for (int k = 0; k < NP; k++) {
            d[k] = hb_ld_ro2(dec_k + p * dec_row + 64 * k);
            w[k] = __ldg(reinterpret_cast<const ulonglong2 *>(key_hk + p * key_row + 64 * k));
        }
#pragma unroll
        for (int k = 0; k < NP; k++) {
            acc128_mac(acc[k][0], d[k].x, w[k].x);
            acc128_mac(acc[k][1], d[k].y, w[k].y);
        }
    }
#pragma unroll
    for (int k = 0; k < NP; k++) {
        u64 lo, hi;
        acc128_fold(acc[k][0], lo, hi);
        e[k].x = montgomery128(lo, hi, lc);
        acc128_fold(acc[k][1], lo, hi);
        e[k].y = montgomery128(lo, hi, lc);
    }
}
// NP pairs, four at a time (the accumulators of four pairs are 56 registers); U digits' operands in flight
template <int NP, int U>
HB_D void ks_mac_thread(const u64 *dec_k, const u64 *key_hk, int L, int logn, const LimbConst &lc, ulonglong2 (&e)[NP]) {
    constexpr int G = NP < 4 ? NP : 4;
#pragma unroll
    for (int c = 0; c < NP; c += G) ks_mac_pairs<G, U>(dec_k + 64 * c, key_hk + 64 * c, L, logn, lc, &e[c]);
}

// ---- where the drop launch takes the polynomial e[h][limb] it divides by the last prime from ----
struct KsSrcMac { // the key switch's inner product over the digits the fan-out launch left in `dec`
    const u64 *dec, *key;
    int L, logn;
    template <int NP, int U>
    HB_D void pairs(int b, int h, int limb, int i0, const LimbConst &lc, ulonglong2 (&e)[NP]) const {
        const size_t L1 = (size_t)L + 1;
        ks_mac_thread<NP, U>(dec + (((size_t)b * L * L1 + limb) << logn) + i0, key + (((size_t)h * L1 + limb) << logn) + i0, L, logn, lc, e);
    }
    template <int NP>
    HB_D void prefetch(int h, int limb, int i0) const { // key words: static data, may be asked for before the predecessor has finished
        if ((threadIdx.x & 7) == 0)
            for (int p = 0; p < L; p++)
#pragma unroll
                for (int kk = 0; kk < NP; kk++) hb_prefetch_l2(key + (((size_t)(p * 2 + h) * (L + 1) + limb) << logn) + i0 + 64 * kk);
    }
};
struct KsSrcPlain { // a ciphertext [batch][2][L + 1][N] as it is: ckks::rescale_inplace / bgv::mod_switch_inplace of a few ciphertexts
    const u64 *ct;
    int L, logn;
    template <int NP, int U>
    HB_D void pairs(int b, int h, int limb, int i0, const LimbConst &, ulonglong2 (&e)[NP]) const {
        const u64 *row = ct + ((((size_t)b * 2 + h) * (L + 1) + limb) << logn) + i0;
#pragma unroll
        for (int k = 0; k < NP; k++) e[k] = hb_ld_stream2(row + 64 * k);
    }
    template <int NP>
    HB_D void prefetch(int, int, int) const {}
};

// ---- transform policies (interface: ntt_engine.cuh).  Built on the device, one per cluster; rows are 16-byte aligned. ----
template <class IN>
struct KsFanLoad {
    const IN &in;
    const LimbConst &lc;
    u64 *diag; // &dec[b][p][p][0] in the cluster that keeps the diagonal term, else null
    int b, p;
    static constexpr bool vec = true;
    HB_D int limb(int) const { return p; }
    HB_D const u64 *src(int) const { return nullptr; }
    HB_D ulonglong2 fetch2(int, int i) const {
        const ulonglong2 v = in.pair(b, p, i, lc);
        if (diag) *reinterpret_cast<ulonglong2 *>(diag + i) = v;
        return v;
    }
    HB_D u64 fetch(int, int i) const { // the transforms take the 128-bit path (vec); kept for the interface
        const ulonglong2 v = fetch2(0, i & ~1);
        return (i & 1) ? v.y : v.x;
    }
    HB_D u64 pre(int, int, u64 raw, const LimbConst &) const { return raw; }
};
// element k of a register array without dynamic indexing (k is a constant once the callers' loops are unrolled; where the
// compiler does not see that, NP - 1 selects instead of a trip through local memory)
template <int NP>
HB_D ulonglong2 ks_pick(const ulonglong2 (&e)[NP], int k) {
    ulonglong2 r = e[0];
#pragma unroll
    for (int j = 1; j < NP; j++)
        if (k == j) r = e[j];
    return r;
}
template <int NP>
struct KsRegLoad { // the inverse transform's input words, computed beforehand (the P limb of the inner product)
    ulonglong2 e[NP];
    int i0;
    static constexpr bool vec = true;
    HB_D int limb(int) const { return 0; }
    HB_D const u64 *src(int) const { return nullptr; }
    HB_D ulonglong2 fetch2(int, int i) const { return ks_pick<NP>(e, (i - i0) >> 6); }
    HB_D u64 fetch(int, int i) const { // the transforms take the 128-bit path (vec); kept for the interface
        const ulonglong2 v = ks_pick<NP>(e, ((i & ~1) - i0) >> 6);
        return (i & 1) ? v.y : v.x;
    }
    HB_D u64 pre(int, int, u64 raw, const LimbConst &) const { return raw; }
};
struct KsRowStore { // forward transform into one row
    u64 *dst;
    static constexpr bool vec = true;
    HB_D void store(int, int i, u64 v, const LimbConst &) const { dst[i] = v; }
    HB_D void store2(int, int i, u64 v0, u64 v1, const LimbConst &) const { *reinterpret_cast<ulonglong2 *>(dst + i) = make_ulonglong2(v0, v1); }
};
template <bool BGV, int NP>
struct KsDropStore { // (e - NTT) / P + addend, e and the addend computed beforehand
    const DropFwdIO<BGV> &drop; // finish(): rescaling.cpp:73-74, mod_switch.cpp:76
    ulonglong2 e[NP], a[NP];
    u64 *dst;           // &out[b][h][k][0]
    const DropConst *d; // limb k
    int i0;
    bool add;
    static constexpr bool vec = true;
    HB_D void store2(int, int i, u64 v0, u64 v1, const LimbConst &lc) const {
        const int k = (i - i0) >> 6;
        const ulonglong2 ek = ks_pick<NP>(e, k);
        ulonglong2 r = make_ulonglong2(drop.finish(ek.x, v0, d, lc), drop.finish(ek.y, v1, d, lc));
        if (add) {
            const ulonglong2 ak = ks_pick<NP>(a, k);
            r.x = add_lazy(r.x, ak.x, lc.q2);
            r.y = add_lazy(r.y, ak.y, lc.q2);
        }
        *reinterpret_cast<ulonglong2 *>(dst + i) = r;
    }
    HB_D void store(int, int i, u64 v, const LimbConst &lc) const { // the transforms take the 128-bit path (vec); kept for the interface
        const int k = ((i & ~1) - i0) >> 6;
        const ulonglong2 ek = ks_pick<NP>(e, k), ak = ks_pick<NP>(a, k);
        u64 r = drop.finish((i & 1) ? ek.y : ek.x, v, d, lc);
        if (add) r = add_lazy(r, (i & 1) ? ak.y : ak.x, lc.q2);
        dst[i] = r;
    }
};

// in[b][p] -> dec[b][p][k], k != p (and the diagonal).  CTAs beyond the `main_ctas` that transform compute the tensor
// product's d0, d1 into `quad` (KsInTensor only).
template <int LOGN, int MODE, class IN>
HB_GLOBAL(plan_for(LOGN, true, MODE).threads, 1)
ks_fan_kernel(const IN in, u64 *__restrict__ dec, const LimbConst *__restrict__ limbs, int L, int tpc, int nchunks, unsigned main_ctas,
              u64 *__restrict__ quad, size_t tensor_units) {
    static_assert(plans_hand_over<LOGN, MODE>(), "plans of this ring size hand over in registers");
    constexpr NttPlan pl = plan_for(LOGN, true, MODE);
    constexpr int T = pl.threads, C = 1 << pl.lpre, W = kHandOverWords<LOGN, MODE>;
    HB_SHARED_U64(sm);
    if (HB_PAIR_TRIGGER) hb_pdl_trigger();
    if (blockIdx.x >= main_ctas) { // whole clusters: none of them reaches a cluster barrier
        if constexpr (std::is_same<IN, KsInTensor>::value) {
            hb_pdl_wait();
            const size_t stride = (size_t)(gridDim.x - main_ctas) * T;
            for (size_t u = (size_t)(blockIdx.x - main_ctas) * T + threadIdx.x; u < tensor_units; u += stride)
                ks_tensor_addend_pair(in.ct1, in.ct2, quad, limbs, L, LOGN, u);
        }
        return;
    }
    const int cid = blockIdx.x >> pl.lpre, B = blockIdx.x & (C - 1);
    const int chunk = cid % nchunks, bp = cid / nchunks, p = bp % L, b = bp / L;
    const LimbConst lcp = limbs[p];
    u64 *const dec_p = dec + (((size_t)(b * L + p) * (L + 1)) << LOGN);
    if constexpr (staged_mode(MODE)) { // both transforms' tables: static data, on their way into shared memory before the row is
        stage_inv_tables<LOGN, T, MODE>(sm, lcp, B);
        stage_fwd_tables<LOGN, T, MODE>(sm, limbs[chunk * tpc < p ? chunk * tpc : chunk * tpc + 1], B);
    }
    HB_PHASE(0);
    hb_pdl_wait();
    if constexpr (staged_mode(MODE)) {
        hb_cp_async_wait_all();
        __syncthreads();
    }
    HB_PHASE(1);
    {
        const KsFanLoad<IN> load{in, lcp, chunk == 0 ? dec_p + ((size_t)p << LOGN) : nullptr, b, p};
        inv_local_passes<LOGN, T, 0, MODE>(sm, load, lcp, 0, B);
    }
    hb_cluster_sync(); // the CTA-local stages are done everywhere and their words visible to the cluster
    HB_PHASE(2);
    u64 w[W];
    inv_cross_to_regs<LOGN, T, MODE>(sm, lcp, B, w);
    HB_PHASE(3);
#pragma unroll
    for (int j = 0; j < W; j++) w[j] = reduce_strict(w[j], lcp.q); // rgsw.cpp:105
    const int t_end = (chunk + 1) * tpc < L ? (chunk + 1) * tpc : L;
#pragma unroll 1
    for (int tt = chunk * tpc; tt < t_end; tt++) {
        const int k = tt < p ? tt : tt + 1;
        const LimbConst lck = limbs[k];
        if constexpr (staged_mode(MODE)) {
            if (tt != chunk * tpc) { // the next target's tables replace the previous one's
                __syncthreads();
                stage_fwd_tables<LOGN, T, MODE>(sm, lck, B);
                hb_cp_async_wait_all();
                __syncthreads();
            }
        }
        u64 v[W];
#pragma unroll
        for (int j = 0; j < W; j++) v[j] = w[j];
        fwd_cross_from_regs<LOGN, T, MODE>(sm, lck, B, v);
        HB_PHASE(4);
        hb_cluster_sync(); // every word has reached its owner
        HB_PHASE(5);
        fwd_passes<LOGN, T, 1, MODE>(sm, KsRowStore{dec_p + ((size_t)k << LOGN)}, lck, 0, B);
        HB_PHASE(6);
        if (tt + 1 < t_end) hb_cluster_arrive(); // done reading shared memory: the next target may scatter into it
    }
}

// e (SRC) -> out[b][h][k]
template <int LOGN, int MODE, bool BGV, class SRC, class ADD>
HB_GLOBAL(plan_for(LOGN, true, MODE).threads, 1)
ks_drop_kernel(const SRC src, const ADD add, u64 *__restrict__ out, const LimbConst *__restrict__ limbs, const DropConst *__restrict__ dc,
               u64 half_qlast, u64 inv_t, u64 inv_t_h, int L, int tpc, int nchunks) {
    static_assert(plans_hand_over<LOGN, MODE>(), "plans of this ring size hand over in registers");
    constexpr NttPlan pl = plan_for(LOGN, true, MODE);
    constexpr int T = pl.threads, C = 1 << pl.lpre, W = kHandOverWords<LOGN, MODE>, LOGG = LOGN - pl.k[0];
    HB_SHARED_U64(sm);
    const int cid = blockIdx.x >> pl.lpre, B = blockIdx.x & (C - 1);
    const int chunk = cid % nchunks, bh = cid / nchunks, h = bh & 1, b = bh >> 1;
    const int L1 = L + 1;
    const LimbConst lcP = limbs[L];
    constexpr int NP = kPairsPerThread<LOGN, MODE>, MAC_U = 4;
    const int i0 = ks_first_pair<LOGN, MODE>(B);
    const int t_end = (chunk + 1) * tpc < L ? (chunk + 1) * tpc : L;
    if (HB_PAIR_TRIGGER) hb_pdl_trigger();
    if constexpr (staged_mode(MODE)) {
        stage_inv_tables<LOGN, T, MODE>(sm, lcP, B);
        stage_fwd_tables<LOGN, T, MODE>(sm, limbs[chunk * tpc], B);
    }
    src.template prefetch<NP>(h, L, i0);
    src.template prefetch<NP>(h, chunk * tpc, i0);
    HB_PHASE(0);
    hb_pdl_wait();
    if constexpr (staged_mode(MODE)) {
        hb_cp_async_wait_all();
        __syncthreads();
    }
    HB_PHASE(1);
    // pre() and finish() of the wave path's policy: the same arithmetic by construction
    const DropFwdIO<BGV> drop{nullptr, nullptr, nullptr, dc, nullptr, 0, 0, half_qlast, L1, LOGN, 0, true, 1u};
    KsDropStore<BGV, NP> st{drop, {}, {}, nullptr, nullptr, i0, add.has(h)};
    const int t0 = B * T + (int)threadIdx.x, t_begin = chunk * tpc;
    u64 z[W];
    // Step 0 is the P limb (inverse transform), steps 1 .. are the limbs of the chunk (forward transforms).  One loop that is
    // not unrolled, so that the code that fetches e[h][limb] — the inner product, all its operands requested before the first
    // product — exists ONCE: the kernel is straight-line code that every warp runs once, and its size is what the
    // instruction cache sees (`no_instruction` was this kernel's first stall reason at 5.4 k instructions:
    // profiles/r4_pair_path.md).  The operands of a forward step's epilogue are thereby in registers before its transform.
#pragma unroll 1
    for (int step = 0; step <= t_end - t_begin; step++) {
        const int limb = step == 0 ? L : t_begin + step - 1;
        const LimbConst lcw = limbs[limb];
        if constexpr (staged_mode(MODE)) {
            if (step >= 2) { // the next limb's forward tables replace the previous one's
                __syncthreads();
                stage_fwd_tables<LOGN, T, MODE>(sm, lcw, B);
            }
        }
        ulonglong2 e[NP];
        src.template pairs<NP, MAC_U>(b, h, limb, i0, lcw, e);
        if (step == 0) {
            KsRegLoad<NP> load;
            load.i0 = i0;
#pragma unroll
            for (int kk = 0; kk < NP; kk++) load.e[kk] = e[kk];
            inv_local_passes<LOGN, T, 0, MODE>(sm, load, lcw, 0, B);
            hb_cluster_sync();
            HB_PHASE(2);
            inv_cross_to_regs<LOGN, T, MODE>(sm, lcw, B, z);
            HB_PHASE(3);
#pragma unroll
            for (int j = 0; j < W; j++) { // DropInttIO::store — rescaling.cpp:47-50, mod_switch.cpp:48-51
                if (BGV) z[j] = harvey_lazy(z[j], inv_t, inv_t_h, lcw.nq);
                z[j] = reduce_strict(z[j], lcw.q);
            }
        } else {
#pragma unroll
            for (int kk = 0; kk < NP; kk++) st.e[kk] = e[kk];
            if (st.add) {
#pragma unroll
                for (int kk = 0; kk < NP; kk++) st.a[kk] = add.pair(b, h, limb, i0 + 64 * kk, lcw);
            }
            if constexpr (staged_mode(MODE)) {
                if (step >= 2) {
                    hb_cp_async_wait_all();
                    __syncthreads();
                }
            }
            st.dst = out + ((((size_t)b * 2 + h) * L + limb) << LOGN);
            st.d = dc + limb;
            u64 v[W];
#pragma unroll
            for (int j = 0; j < W; j++) v[j] = drop.pre(limb, t0 + (j << LOGG), z[j], lcw);
            fwd_cross_from_regs<LOGN, T, MODE>(sm, lcw, B, v);
            HB_PHASE(4);
            hb_cluster_sync();
            HB_PHASE(5);
            fwd_passes<LOGN, T, 1, MODE>(sm, st, lcw, 0, B);
            HB_PHASE(6);
            if (step < t_end - t_begin) hb_cluster_arrive();
        }
    }
}

} // namespace hb
