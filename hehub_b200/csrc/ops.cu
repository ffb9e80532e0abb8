// ops.cu — scheme-level operations composed from the transform engine and fused
// coefficient-wise kernels.  Reference call stacks: SURVEY.md §3.1.
//
//   tensor       ckks::mult_low_level            src/fhe/ckks/arith.cpp:55-62
//   ext_prod     ext_prod_montgomery             src/fhe/primitives/rgsw.cpp:57-156
//   drop_last    rescale_by_one_prime_inplace    src/fhe/ckks/rescaling.cpp:14-78
//                mod_drop_one_prime_inplace      src/fhe/bgv/mod_switch.cpp:13-78
//   relinearize  ckks::relinearize               src/fhe/ckks/arith.cpp:64-73 (bgv/arith.cpp:71-79)
//   galois       cycle / involution              src/fhe/common/permutation.cpp:28-75
//
// Prologue / epilogue arithmetic (Barrett + centring before the forward transforms of a
// rescale, lazy subtract + q_last^{-1} multiply + addend after them, strict reduction after
// the inverse transforms) lives in the IO policies of the transform kernels, so those values
// never make a separate trip through HBM.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "internal.h"

namespace hb {

static inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// ------------------------------------------------------------------------------------------
// tensor product: one pass, 4 loads + 3 stores per coefficient
// ------------------------------------------------------------------------------------------
#ifndef HB_TENSOR_MINB
#define HB_TENSOR_MINB 5
#endif
#ifndef HB_MAC_MINB
#define HB_MAC_MINB 2
#endif
#ifndef HB_MAC_U1 // digits whose operands are in flight together in the one-ciphertext form of the inner product
#define HB_MAC_U1 4
#endif
#ifndef HB_MAC_CPT
#define HB_MAC_CPT 2
#endif
#ifndef HB_MAC_STAGED_MIN_BATCH // from this many ciphertexts per wave on, the key-switch inner product stages its key slice in shared memory
#define HB_MAC_STAGED_MIN_BATCH 8
#endif
constexpr int kMacStagedMinBatch = HB_MAC_STAGED_MIN_BATCH;
#ifndef HB_MAC_STAGED_U
#define HB_MAC_STAGED_U 12
#endif
template <int W>
HB_D void ld_words(u64 (&dst)[W], const u64 *p, bool read_only) {
    if constexpr (W == 2) {
        const ulonglong2 v = read_only ? __ldg(reinterpret_cast<const ulonglong2 *>(p)) : hb_ld_stream2(p);
        dst[0] = v.x;
        dst[1] = v.y;
    } else {
        dst[0] = read_only ? __ldg(p) : hb_ld_stream(p);
    }
}
template <int W>
HB_D void st_words(u64 *p, const u64 (&src)[W]) {
    if constexpr (W == 2) *reinterpret_cast<ulonglong2 *>(p) = make_ulonglong2(src[0], src[1]);
    else p[0] = src[0];
}

// W = 2: one thread per pair of adjacent coefficients (128-bit accesses); W = 1: slabs that are only 8-byte aligned
template <int W>
HB_D void tensor_unit(const u64 *__restrict__ ct1, const u64 *__restrict__ ct2, u64 *__restrict__ quad, const LimbConst *__restrict__ limbs,
                      int L, int logn, size_t gid) {
    constexpr int LW = (W == 2) ? 1 : 0;
    const size_t row = gid >> (logn - LW);              // b * L + l
    const size_t i = (gid & (((size_t)1 << (logn - LW)) - 1)) * W;
    const size_t b = row / L, l = row % L;
    const LimbConst lc = limbs[l];
    const size_t n = (size_t)1 << logn, poly = (size_t)L << logn;
    const size_t in0 = b * 2 * poly + l * n + i, in1 = in0 + poly;
    const size_t o0 = b * 3 * poly + l * n + i;
    u64 a0[W], a1[W], b0[W], b1[W], d0[W], d1[W], d2[W];
    ld_words<W>(a0, ct1 + in0, false);
    ld_words<W>(a1, ct1 + in1, false);
    ld_words<W>(b0, ct2 + in0, false);
    ld_words<W>(b1, ct2 + in1, false);
#pragma unroll
    for (int w = 0; w < W; w++) {
        d0[w] = mul_hybrid_lazy(a0[w], b0[w], lc);
        d1[w] = add_lazy(mul_hybrid_lazy(a0[w], b1[w], lc), mul_hybrid_lazy(a1[w], b0[w], lc), lc.q2);
        d2[w] = mul_hybrid_lazy(a1[w], b1[w], lc);
    }
    st_words<W>(quad + o0, d0);
    st_words<W>(quad + o0 + poly, d1);
    st_words<W>(quad + o0 + 2 * poly, d2);
}
template <int W>
HB_GLOBAL(256, HB_TENSOR_MINB)
tensor_kernel(const u64 *__restrict__ ct1, const u64 *__restrict__ ct2, u64 *__restrict__ quad,
              const LimbConst *__restrict__ limbs, int L, int logn, size_t units_total) {
    hb_pdl_wait();
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= units_total) return;
    tensor_unit<W>(ct1, ct2, quad, limbs, L, logn, gid);
}

int op_ckks_tensor(Context &c, unsigned logn, const u64 *moduli, size_t L, const u64 *ct1, const u64 *ct2, u64 *quad,
                   size_t batch) {
    if (!moduli || !ct1 || !ct2 || !quad) return c.fail(1, "null operand");
    if (batch == 0) return 0;
    for (size_t k = 0; k < L; k++)
        if (!(moduli[k] & 1)) return c.fail(1, "Montgomery multiplication needs odd moduli");
    int err = 0;
    const LimbConst *limbs = c.get_chain(0, moduli, L, &err);
    if (!limbs) return err;
    const bool vec = logn >= 1 && aligned16(ct1) && aligned16(ct2) && aligned16(quad);
    const size_t units = vec ? (batch * L) << (logn - 1) : (batch * L) << logn;
    const size_t blocks = (units + 255) / 256;
    if (blocks > 0x7fffffffull) return c.fail(1, "operand too large for one launch");
    if (vec) {
        HB_LAUNCH(tensor_kernel<2>, (unsigned)blocks, 256, 0, c.stream, 0, ct1, ct2, quad, limbs, (int)L, (int)logn, units);
    } else {
        HB_LAUNCH(tensor_kernel<1>, (unsigned)blocks, 256, 0, c.stream, 0, ct1, ct2, quad, limbs, (int)L, (int)logn, units);
    }
    c.stats.launches++;
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : c.cuda_fail(e, "tensor launch");
}

// ------------------------------------------------------------------------------------------
// Galois automorphisms on NTT-form rows (permutation.cpp:28-75).  Slot j holds the evaluation at psi^(2*brev(j)+1); X -> X^g
// moves root index e to e*g, so output slot j gathers from the slot whose root index is e_j * g^{-1} (mod 2N).  Two
// properties make the gather cheap enough to fuse into whatever reads the row: slots j and j^1 gather from slots f and f^1
// (their root indices differ by N, and g^{-1} is odd), so 128-bit accesses survive as "load the aligned pair, maybe swap";
// and an aligned block of 2^m output slots gathers from ONE aligned block of 2^m input slots, so the gather touches the same
// cache lines a linear read would.  ginv == 1 is the identity.
// ------------------------------------------------------------------------------------------
HB_D unsigned galois_from(unsigned j, unsigned ginv, int logn) {
    const unsigned mask = (2u << logn) - 1;
    const unsigned e = 2 * (__brev(j) >> (32 - logn)) + 1;
    const unsigned src_e = (e * ginv) & mask;
    return __brev((src_e - 1) >> 1) >> (32 - logn);
}
// words (j, j + 1), j even, of the permuted row
HB_D ulonglong2 galois_pair(const u64 *row_words, unsigned j, unsigned ginv, int logn, bool read_only) {
    const unsigned f = galois_from(j, ginv, logn);
    const u64 *p = row_words + (f & ~1u);
    const ulonglong2 v = read_only ? hb_ld_ro2(p) : hb_ld_stream2(p);
    return (f & 1u) ? make_ulonglong2(v.y, v.x) : v;
}

// ------------------------------------------------------------------------------------------
// key switch (ext_prod_montgomery)
// ------------------------------------------------------------------------------------------
// step 1: c[b][p] = strict(INTT_{q_p}(in[b][p]))                     rgsw.cpp:103-105
template <bool GALOIS> // GALOIS: separate instantiation so the plain key switch compiles exactly as without the gather
struct ExtInttIO {
    const u64 *in;
    size_t in_batch_stride;
    u64 *c; // [batch][L][N]
    int L, logn;
    bool vec;
    unsigned ginv; // GALOIS: the input rows are read through the Galois permutation (ckks::rotate / conjugate, ckks/arith.cpp:75-93)
    HB_D int limb(int row) const { return row % L; }
    HB_D const u64 *src(int row) const { return in + (size_t)(row / L) * in_batch_stride + ((size_t)(row % L) << logn); }
    // fetch / fetch2 (ntt_engine.cuh) exist only in the GALOIS instantiation
    template <bool G = GALOIS, class = std::enable_if_t<G>>
    HB_D u64 fetch(int row, int i) const { return hb_ld_stream(src(row) + galois_from((unsigned)i, ginv, logn)); }
    template <bool G = GALOIS, class = std::enable_if_t<G>>
    HB_D ulonglong2 fetch2(int row, int i) const { return galois_pair(src(row), (unsigned)i, ginv, logn, false); }
    HB_D u64 pre(int, int, u64 raw, const LimbConst &) const { return raw; }
    HB_D void store(int row, int i, u64 v, const LimbConst &lc) const { c[((size_t)row << logn) + i] = reduce_strict(v, lc.q); }
    HB_D void store2(int row, int i, u64 v0, u64 v1, const LimbConst &lc) const {
        *reinterpret_cast<ulonglong2 *>(c + ((size_t)row << logn) + i) = make_ulonglong2(reduce_strict(v0, lc.q), reduce_strict(v1, lc.q));
    }
};

// step 2: dec[b][p][k] = NTT_{q_k}(c[b][p]) for k != p, k in [0, L]   rgsw.cpp:108-119
struct ExtFanoutIO {
    const u64 *c; // [batch][L][N]
    u64 *dec;     // [batch][L][L+1][N]; slot k == p is not written (the diagonal keeps in[p])
    int L, logn;
    bool vec;
    HB_D void split(int row, int &b, int &p, int &k) const {
        b = row / (L * L);
        const int rem = row - b * L * L;
        p = rem / L;
        const int kk = rem - p * L;
        k = kk < p ? kk : kk + 1;
    }
    HB_D int limb(int row) const {
        int b, p, k;
        split(row, b, p, k);
        return k;
    }
    HB_D const u64 *src(int row) const {
        int b, p, k;
        split(row, b, p, k);
        return c + ((size_t)(b * L + p) << logn);
    }
    HB_D u64 pre(int, int, u64 raw, const LimbConst &) const { return raw; }
    HB_D void store(int row, int i, u64 v, const LimbConst &) const {
        int b, p, k;
        split(row, b, p, k);
        dec[((size_t)((b * L + p) * (L + 1) + k) << logn) + i] = v;
    }
    HB_D void store2(int row, int i, u64 v0, u64 v1, const LimbConst &) const {
        int b, p, k;
        split(row, b, p, k);
        *reinterpret_cast<ulonglong2 *>(dec + ((size_t)((b * L + p) * (L + 1) + k) << logn) + i) = make_ulonglong2(v0, v1);
    }
};

// One N = 16384 ciphertext has L^2 = 64 fan-out rows at L = 8: 512 CTAs of the 8-CTA thin plan, one wave at FOUR CTAs per SM
// (64 registers, no spills) against 128 long-lived CTAs of the throughput plan: one ckks::mult per call 50.5 -> 47.5 us
// (profiles/r4_pair_path.md).  The drop's forward kernel loses 4 % at four per SM and keeps three.
template <>
struct IoResidency<ExtFanoutIO> {
    static constexpr int extra(int logn, int mode, bool forward) { return (logn == 14 && mode == 1 && forward) ? 1 : 0; }
};

// step 3: out[b][h][k][i] = Mont128_{q_k}( sum_p dec[p][k][i] * key[p][h][k][i] )   rgsw.cpp:126-153
// The sum is exact in 128 bits and reduced once, like the reference (reducing per term would
// change the representative).  One thread owns W adjacent coefficients (W = 2: 128-bit accesses;
// W = 1 for slabs that are only 8-byte aligned) of one limb of CPT consecutive ciphertexts, so a
// key word fetched from L2 feeds CPT products: the key stream (2 L (L+1) N words per ciphertext,
// more than every other operand together) is what bounds this kernel.
template <int CPT, int W, bool GALOIS>
HB_D void ext_mac_unit(const u64 *__restrict__ in, size_t in_batch_stride, const u64 *__restrict__ dec, const u64 *__restrict__ key,
                       u64 *__restrict__ out, const LimbConst *__restrict__ limbs, int L, int logn, size_t batch, size_t gid,
                       unsigned ginv, int nk = 0, bool absolute_gid = false);
template <int CPT, int W, bool GALOIS>
HB_GLOBAL(256, HB_MAC_MINB)
ext_mac_kernel(const u64 *__restrict__ in, size_t in_batch_stride, const u64 *__restrict__ dec, const u64 *__restrict__ key,
               u64 *__restrict__ out, const LimbConst *__restrict__ limbs, int L, int logn, size_t batch, size_t total,
               unsigned groups, unsigned chunks_per_group, unsigned ginv) {
    hb_pdl_wait();
    // gid = (b / CPT, k, i / W).  With `groups` != 0 consecutive CTAs take the SAME 256-thread slice of
    // (k, i) for consecutive ciphertext groups, so CTAs resident together share their key words in L2
    // (the key is read once per group: 8-18 times per wave) instead of streaming it from HBM each time.
    size_t gid;
    if (groups) gid = ((size_t)(blockIdx.x % groups) * chunks_per_group + blockIdx.x / groups) * 256 + threadIdx.x;
    else gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= total) return;
    ext_mac_unit<CPT, W, GALOIS>(in, in_batch_stride, dec, key, out, limbs, L, logn, batch, gid, ginv);
}
template <int CPT, int W, bool GALOIS>
HB_D void ext_mac_unit(const u64 *__restrict__ in, size_t in_batch_stride, const u64 *__restrict__ dec, const u64 *__restrict__ key,
                       u64 *__restrict__ out, const LimbConst *__restrict__ limbs, int L, int logn, size_t batch, size_t gid,
                       unsigned ginv, int nk, bool) {
    const int L1 = L + 1;
    if (nk == 0) nk = L1; // limbs this launch computes (the first nk of the L + 1)
    constexpr int LW = (W == 2) ? 1 : 0;
    const size_t i = (gid & (((size_t)1 << (logn - LW)) - 1)) * W;
    const size_t bk = gid >> (logn - LW);
    const int k = (int)(bk % nk);
    const size_t b0 = (bk / nk) * CPT;
    const LimbConst lc = limbs[k];
    Acc128 acc[CPT][2][W];
#pragma unroll
    for (int c = 0; c < CPT; c++)
#pragma unroll
        for (int h = 0; h < 2; h++)
#pragma unroll
            for (int w = 0; w < W; w++) acc128_clear(acc[c][h][w]);
    const size_t dec_ct = ((size_t)L * L1) << logn, dec_row = (size_t)L1 << logn;
    const size_t key_row = (size_t)(2 * L1) << logn, key_half = (size_t)L1 << logn;
    const u64 *const key_k = key + ((size_t)k << logn) + i;
    // A ragged last group (batch not a multiple of CPT) reads the last valid ciphertext again and drops the result: the
    // loop stays free of per-ciphertext branches.  The diagonal term (p == k) keeps in[p] (rgsw.cpp:99-101), read through
    // the Galois permutation when the key switch carries one: same load, another address, the pair possibly swapped.
    const u64 *dec_c[CPT], *diag_c[CPT];
    bool swap = false;
    {
        size_t diag_off = i;
        if constexpr (GALOIS) {
            const unsigned f = galois_from((unsigned)i, ginv, logn);
            diag_off = W == 2 ? (f & ~1u) : f;
            swap = W == 2 && (f & 1u);
        }
#pragma unroll
        for (int c = 0; c < CPT; c++) {
            const size_t b = (b0 + c < batch) ? b0 + c : batch - 1;
            dec_c[c] = dec + b * dec_ct + ((size_t)k << logn) + i;
            diag_c[c] = in + b * in_batch_stride + ((size_t)k << logn) + diag_off; // only dereferenced when k < L
        }
    }
    auto load = [&](int p, u64 (&d)[CPT][W], u64 (&k0)[W], u64 (&k1)[W]) {
        ld_words<W>(k0, key_k + p * key_row, true);
        ld_words<W>(k1, key_k + p * key_row + key_half, true);
#pragma unroll
        for (int c = 0; c < CPT; c++) {
            ld_words<W>(d[c], (p == k) ? diag_c[c] : dec_c[c] + p * dec_row, false);
            if constexpr (GALOIS && W == 2) {
                if (swap && p == k) {
                    const u64 t = d[c][0];
                    d[c][0] = d[c][1];
                    d[c][1] = t;
                }
            }
        }
    };
    auto mac = [&](const u64 (&d)[CPT][W], const u64 (&k0)[W], const u64 (&k1)[W]) {
#pragma unroll
        for (int c = 0; c < CPT; c++)
#pragma unroll
            for (int w = 0; w < W; w++) {
                acc128_mac(acc[c][0][w], d[c][w], k0[w]);
                acc128_mac(acc[c][1][w], d[c][w], k1[w]);
            }
    };
    int p = 0;
    if constexpr (CPT == 1) {
        // one ciphertext per call: the launch is a single short wave that lives on memory latency — the operands of U digits
        // (3 U 128-bit loads) in flight before the first product
        constexpr int U = HB_MAC_U1;
        for (; p + U <= L; p += U) {
            u64 d[U][CPT][W], k0[U][W], k1[U][W];
#pragma unroll
            for (int u = 0; u < U; u++) load(p + u, d[u], k0[u], k1[u]);
#pragma unroll
            for (int u = 0; u < U; u++) mac(d[u], k0[u], k1[u]);
        }
    }
    for (; p + 1 < L; p += 2) { // two rows' operands in flight before the first product
        u64 dA[CPT][W], kA0[W], kA1[W], dB[CPT][W], kB0[W], kB1[W];
        load(p, dA, kA0, kA1);
        load(p + 1, dB, kB0, kB1);
        mac(dA, kA0, kA1);
        mac(dB, kB0, kB1);
    }
    if (p < L) {
        u64 dA[CPT][W], kA0[W], kA1[W];
        load(p, dA, kA0, kA1);
        mac(dA, kA0, kA1);
    }
#pragma unroll
    for (int c = 0; c < CPT; c++) {
        if (b0 + c < batch) {
            u64 *const o = out + ((((b0 + c) * 2) * L1 + k) << logn) + i;
#pragma unroll
            for (int h = 0; h < 2; h++) {
                u64 r[W];
#pragma unroll
                for (int w = 0; w < W; w++) {
                    u64 lo, hi;
                    acc128_fold(acc[c][h][w], lo, hi);
                    r[w] = montgomery128(lo, hi, lc);
                }
                st_words<W>(o + h * key_half, r);
            }
        }
    }
}

// The same inner product for batches: one CTA owns a slice of 256 * W coefficients of one limb k, brings the 2 L key rows of
// that slice into shared memory ONCE and then walks a group of ciphertexts, so per product the SM pulls 8 bytes of `dec`
// from HBM and the key words come from shared memory — the kernel above fetches every key word from L2 for every pair of
// ciphertexts, and L2 -> SM traffic (as much again as `dec`) was what bounded it (profiles/r3_mac_census.md).
template <int W, bool GALOIS>
HB_GLOBAL(256, 2)
ext_mac_staged_kernel(const u64 *__restrict__ in, size_t in_batch_stride, const u64 *__restrict__ dec, const u64 *__restrict__ key,
                      u64 *__restrict__ out, const LimbConst *__restrict__ limbs, int L, int logn, size_t batch, unsigned ct_chunks,
                      unsigned cts_per_cta, unsigned ginv) {
    HB_SHARED_U64(ks); // [p < L][half][256 * W]
    constexpr int CH = 256 * W;
    const int L1 = L + 1;
    const unsigned slice = blockIdx.x / ct_chunks, chunk = blockIdx.x % ct_chunks; // CTAs of one slice are neighbours: its key rows stay in L2
    const unsigned slices_per_limb = (1u << logn) / CH;
    const int k = (int)(slice / slices_per_limb);
    const size_t i = (size_t)(slice % slices_per_limb) * CH + threadIdx.x * W;
    const LimbConst lc = limbs[k];
    const size_t key_row = (size_t)(2 * L1) << logn, key_half = (size_t)L1 << logn;
    const u64 *const key_k = key + ((size_t)k << logn) + i;
    hb_pdl_wait();
    for (int p = 0; p < L; p++) { // the key is not written by the launches before this one, but keep the prologue simple
        u64 k0[W], k1[W];
        ld_words<W>(k0, key_k + p * key_row, true);
        ld_words<W>(k1, key_k + p * key_row + key_half, true);
        st_words<W>(ks + (size_t)(2 * p) * CH + threadIdx.x * W, k0);
        st_words<W>(ks + (size_t)(2 * p + 1) * CH + threadIdx.x * W, k1);
    }
    // every thread reads back exactly the words it wrote: no barrier needed
    const size_t dec_ct = ((size_t)L * L1) << logn, dec_row = (size_t)L1 << logn;
    size_t diag_off = i;
    bool swap = false;
    if constexpr (GALOIS) {
        const unsigned f = galois_from((unsigned)i, ginv, logn);
        diag_off = W == 2 ? (f & ~1u) : f;
        swap = W == 2 && (f & 1u);
    }
    const size_t b_end = ((size_t)(chunk + 1) * cts_per_cta < batch) ? (size_t)(chunk + 1) * cts_per_cta : batch;
    for (size_t b = (size_t)chunk * cts_per_cta; b < b_end; b++) {
        const u64 *const dec_b = dec + b * dec_ct + ((size_t)k << logn) + i;
        const u64 *const diag_b = in + b * in_batch_stride + ((size_t)k << logn) + diag_off; // only dereferenced when k < L
        Acc128 acc[2][W];
#pragma unroll
        for (int h = 0; h < 2; h++)
#pragma unroll
            for (int w = 0; w < W; w++) acc128_clear(acc[h][w]);
        auto load = [&](int p, u64 (&d)[W]) {
            ld_words<W>(d, (p == k) ? diag_b : dec_b + p * dec_row, false);
            if constexpr (GALOIS && W == 2) {
                if (swap && p == k) {
                    const u64 t = d[0];
                    d[0] = d[1];
                    d[1] = t;
                }
            }
        };
        auto mac = [&](int p, const u64 (&d)[W]) {
            u64 k0[W], k1[W];
            if constexpr (W == 2) {
                const ulonglong2 a = *reinterpret_cast<const ulonglong2 *>(ks + (size_t)(2 * p) * CH + threadIdx.x * 2);
                const ulonglong2 c = *reinterpret_cast<const ulonglong2 *>(ks + (size_t)(2 * p + 1) * CH + threadIdx.x * 2);
                k0[0] = a.x, k0[1] = a.y, k1[0] = c.x, k1[1] = c.y;
            } else {
                k0[0] = ks[(size_t)(2 * p) * CH + threadIdx.x];
                k1[0] = ks[(size_t)(2 * p + 1) * CH + threadIdx.x];
            }
#pragma unroll
            for (int w = 0; w < W; w++) {
                acc128_mac(acc[0][w], d[w], k0[w]);
                acc128_mac(acc[1][w], d[w], k1[w]);
            }
        };
        // The kernel lives on HBM latency (long_scoreboard is its one stall reason): chains of U = 12 rows request all their
        // `dec` words before the first product; shorter chains run the plain loop, which the compiler pipelines well enough
        // (guarding a longer batch with p + u < L measured 2-3 % slower: profiles/r3_mac_census.md)
        constexpr int U = HB_MAC_STAGED_U;
        int p = 0;
        for (; p + U <= L; p += U) {
            u64 d[U][W];
#pragma unroll
            for (int u = 0; u < U; u++) load(p + u, d[u]);
#pragma unroll
            for (int u = 0; u < U; u++) mac(p + u, d[u]);
        }
        for (; p < L; p++) {
            u64 d0[W];
            load(p, d0);
            mac(p, d0);
        }
        u64 *const o = out + (((b * 2) * L1 + k) << logn) + i;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            u64 r[W];
#pragma unroll
            for (int w = 0; w < W; w++) {
                u64 lo, hi;
                acc128_fold(acc[h][w], lo, hi);
                r[w] = montgomery128(lo, hi, lc);
            }
            st_words<W>(o + h * key_half, r);
        }
    }
}

// One ciphertext per call at N >= 16384: the drop's inverse transform of the P limb (2 rows = 16 CTAs, ~13 us alone on the GPU
// between the inner product and the forward transforms) rides in the inner-product launch instead — its clusters compute their
// own P-limb inner product and transform it while the other CTAs of the launch compute the remaining limbs (defined after
// ks_pair.cuh).  `done` tells the drop that z is there.
struct FusedDrop {
    u64 *z; // [2][N]
    u64 inv_t, inv_t_h;
    bool bgv, done;
};
static cudaError_t launch_mac_intt(Context &c, unsigned logn, const LimbConst *limbs, size_t L, const u64 *in, size_t in_batch_stride,
                                   const u64 *dec, const u64 *key, u64 *out, unsigned ginv, const FusedDrop &fd);

// one wave of at most `wave` ciphertexts; scratch slots 0 (c) and 1 (dec)
static int ext_prod_wave(Context &c, unsigned logn, const LimbConst *limbs, size_t L, const u64 *in, size_t in_batch_stride,
                         const u64 *key, u64 *out, size_t batch, u64 *cbuf, u64 *dec, unsigned ginv, FusedDrop *fd = nullptr) {
    const size_t n = (size_t)1 << logn;
    const bool vec1 = aligned16(in) && (in_batch_stride % 2 == 0) && aligned16(cbuf);
    cudaError_t e;
    if (ginv == 1) e = launch_ntt<false>(c.env(), logn, ExtInttIO<false>{in, in_batch_stride, cbuf, (int)L, (int)logn, vec1, 1u}, limbs, (int)(batch * L));
    else e = launch_ntt<false>(c.env(), logn, ExtInttIO<true>{in, in_batch_stride, cbuf, (int)L, (int)logn, vec1, ginv}, limbs, (int)(batch * L));
    if (e != cudaSuccess) return c.cuda_fail(e, "ext_prod: intt launch");
    ExtFanoutIO io2{cbuf, dec, (int)L, (int)logn, aligned16(cbuf) && aligned16(dec)};
    e = launch_ntt<true>(c.env(), logn, io2, limbs, (int)(batch * L * L));
    if (e != cudaSuccess) return c.cuda_fail(e, "ext_prod: ntt launch");
    constexpr int CPT = HB_MAC_CPT;
    const bool vec = aligned16(in) && in_batch_stride % 2 == 0 && aligned16(dec) && aligned16(key) && aligned16(out) && n >= 2;
    // batches: the key slice staged in shared memory (needs whole 256 * W slices per row and room for 2 L rows of one)
    const int SW = vec ? 2 : 1;
    const size_t staged_smem = (size_t)2 * L * 256 * SW * 8;
    if (batch >= (size_t)kMacStagedMinBatch && n % (256 * (size_t)SW) == 0 && staged_smem <= (size_t)100 << 10) {
        // ciphertexts per CTA: enough to amortise the key slice (one slice = the `dec` words of two ciphertexts), few enough
        // that the grid still has several waves
        size_t per_cta = 16;
        while (per_cta > 4 && ((L + 1) * (n / (256 * (size_t)SW))) * ((batch + per_cta - 1) / per_cta) < (size_t)8 * c.sm_count) per_cta /= 2;
        const size_t ct_chunks = (batch + per_cta - 1) / per_cta;
        const size_t blocks = (L + 1) * (n / (256 * (size_t)SW)) * ct_chunks;
        if (blocks <= 0x7fffffffull) {
            auto launch_staged = [&](auto kern) -> cudaError_t {
                // kernels of one signature share the lambda's instantiation: the opt-in is tracked per kernel pointer, in the context
                int &have = c.smem_opt_in[reinterpret_cast<const void *>(kern)];
                if (have < (int)staged_smem) {
                    cudaError_t ce = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)staged_smem);
                    if (ce != cudaSuccess) return ce;
                    have = (int)staged_smem;
                }
                HB_LAUNCH(kern, (unsigned)blocks, 256, staged_smem, c.stream, 0, in, in_batch_stride, dec, key, out, limbs, (int)L, (int)logn, batch,
                          (unsigned)ct_chunks, (unsigned)per_cta, ginv);
                return cudaGetLastError();
            };
            if (vec) e = ginv == 1 ? launch_staged(ext_mac_staged_kernel<2, false>) : launch_staged(ext_mac_staged_kernel<2, true>);
            else e = ginv == 1 ? launch_staged(ext_mac_staged_kernel<1, false>) : launch_staged(ext_mac_staged_kernel<1, true>);
            c.stats.launches++;
            return e == cudaSuccess ? 0 : c.cuda_fail(e, "ext_prod: staged mac launch");
        }
    }
    if (fd && batch == 1 && vec) {
        e = launch_mac_intt(c, logn, limbs, L, in, in_batch_stride, dec, key, out, ginv, *fd);
        if (e != cudaSuccess) return c.cuda_fail(e, "ext_prod: mac + intt launch");
        fd->done = true;
        return 0;
    }
    // one ciphertext per call: no second ciphertext to share key words with, so one per thread (fewer registers, more loads in flight)
    const size_t cpt = batch == 1 ? 1 : CPT;
    const size_t ngroups = (batch + cpt - 1) / cpt;
    const size_t per_group = (L + 1) * (vec ? n / 2 : n), total = ngroups * per_group;
    const bool inter = per_group % 256 == 0 && ngroups > 1;
    auto launch_mac = [&](auto kern) {
        HB_LAUNCH(kern, (unsigned)((total + 255) / 256), 256, 0, c.stream, 0, in, in_batch_stride, dec, key, out, limbs, (int)L, (int)logn,
                  batch, total, inter ? (unsigned)ngroups : 0u, (unsigned)(per_group / 256), ginv);
    };
    if (batch == 1) {
        if (vec) {
            if (ginv == 1) launch_mac(ext_mac_kernel<1, 2, false>);
            else launch_mac(ext_mac_kernel<1, 2, true>);
        } else {
            if (ginv == 1) launch_mac(ext_mac_kernel<1, 1, false>);
            else launch_mac(ext_mac_kernel<1, 1, true>);
        }
    } else if (vec) {
        if (ginv == 1) launch_mac(ext_mac_kernel<CPT, 2, false>);
        else launch_mac(ext_mac_kernel<CPT, 2, true>);
    } else {
        if (ginv == 1) launch_mac(ext_mac_kernel<CPT, 1, false>);
        else launch_mac(ext_mac_kernel<CPT, 1, true>);
    }
    c.stats.launches++;
    e = cudaGetLastError();
    return e == cudaSuccess ? 0 : c.cuda_fail(e, "ext_prod: mac launch");
}

// Ciphertexts per wave: what the scratch cap allows, rounded down to a multiple of a quarter of the SM count
// when the batch has to be cut anyway — the transform launches of a wave run 2, L, L^2 or 2(L-1) rows per
// ciphertext on 1-2 CTAs per row, and with such a wave all of them end on (nearly) full waves of CTAs.
// `rows_per_ct`: the most rows one ciphertext contributes to a single launch — the IO policies index rows (and
// row * limbs) with int, so a wave never exceeds 2^30 of them whatever the scratch cap says.
static size_t wave_size(const Context &c, size_t words_per_ct, size_t batch, size_t rows_per_ct) {
    size_t w = c.scratch_cap_bytes / (words_per_ct * 8);
    const size_t row_cap = ((size_t)1 << 30) / (rows_per_ct ? rows_per_ct : 1);
    if (w > row_cap) w = row_cap;
    if (w < 1) w = 1;
    if (w >= batch) return batch;
    const size_t quantum = (size_t)(c.sm_count > 4 ? c.sm_count / 4 : 1);
    if (w >= quantum) w -= w % quantum;
    return w;
}

static int ext_prod_impl(Context &c, unsigned logn, const u64 *ext_moduli, size_t L, const u64 *in, size_t in_batch_stride, const u64 *key,
                         u64 *out, size_t batch, unsigned ginv, FusedDrop *fd);
int op_ext_prod(Context &c, unsigned logn, const u64 *ext_moduli, size_t L, const u64 *in, size_t in_batch_stride,
                const u64 *key, u64 *out, size_t batch, unsigned ginv) {
    return ext_prod_impl(c, logn, ext_moduli, L, in, in_batch_stride, key, out, batch, ginv, nullptr);
}
static int ext_prod_impl(Context &c, unsigned logn, const u64 *ext_moduli, size_t L, const u64 *in, size_t in_batch_stride, const u64 *key,
                         u64 *out, size_t batch, unsigned ginv, FusedDrop *fd) {
    if (!ext_moduli || !in || !key || !out) return c.fail(1, "null operand");
    if (L == 0) return c.fail(1, "Empty RGSW ciphertext."); // rgsw.cpp:59-61
    if (batch == 0) return 0;
    for (size_t k = 0; k <= L; k++)
        if (!(ext_moduli[k] & 1)) return c.fail(1, "Montgomery reduction needs odd moduli");
    int err = 0;
    const LimbConst *limbs = c.get_chain(logn, ext_moduli, L + 1, &err);
    if (!limbs) return err;
    const size_t n = (size_t)1 << logn;
    const size_t per_ct = L * n + L * (L + 1) * n;
    const size_t wave = wave_size(c, per_ct, batch, L * (L + 1));
    u64 *cbuf = c.get_scratch(0, wave * L * n, &err);
    if (!cbuf) return err;
    u64 *dec = c.get_scratch(1, wave * L * (L + 1) * n, &err);
    if (!dec) return err;
    for (size_t b0 = 0; b0 < batch; b0 += wave) {
        const size_t nb = (batch - b0 < wave) ? batch - b0 : wave;
        if (int rc = ext_prod_wave(c, logn, limbs, L, in + b0 * in_batch_stride, in_batch_stride, key, out + b0 * 2 * (L + 1) * n,
                                   nb, cbuf, dec, ginv, fd))
            return rc;
    }
    return 0;
}

// ------------------------------------------------------------------------------------------
// drop the last prime (CKKS rescale / BGV mod-switch)
// ------------------------------------------------------------------------------------------
// step 1: z[b][h] = strict( [H(., t^{-1})] INTT_{q_last}(ct[b][h][L-1]) )   rescaling.cpp:47-50, mod_switch.cpp:48-51
template <bool BGV>
struct DropInttIO {
    const u64 *ct;
    u64 *z; // [batch][2][N]
    int L, logn;
    u64 inv_t, inv_t_h; // 0 for CKKS
    bool vec;
    HB_D int limb(int) const { return L - 1; }
    HB_D const u64 *src(int row) const { return ct + ((size_t)(row * L + L - 1) << logn); }
    HB_D u64 pre(int, int, u64 raw, const LimbConst &) const { return raw; }
    HB_D void store(int row, int i, u64 v, const LimbConst &lc) const {
        if (BGV) v = harvey_lazy(v, inv_t, inv_t_h, lc.nq);
        z[((size_t)row << logn) + i] = reduce_strict(v, lc.q);
    }
    HB_D void store2(int row, int i, u64 v0, u64 v1, const LimbConst &lc) const {
        if (BGV) {
            v0 = harvey_lazy(v0, inv_t, inv_t_h, lc.nq);
            v1 = harvey_lazy(v1, inv_t, inv_t_h, lc.nq);
        }
        *reinterpret_cast<ulonglong2 *>(z + ((size_t)row << logn) + i) = make_ulonglong2(reduce_strict(v0, lc.q), reduce_strict(v1, lc.q));
    }
};

// step 2: per remaining limb k: r = centre(barrett(z)); NTT; out = H(lazy_sub(ct, r), q_last^{-1}) [...]
// NC: the epilogue's operands come over the non-coherent path (they were written by earlier launches); false inside the
// single-launch kernel, where earlier PHASES of the same launch wrote them
template <bool BGV, bool GALOIS = false, bool NC = true>
struct DropFwdIO {
    const u64 *ct;
    const u64 *z;
    u64 *out; // [batch][2][L-1][N]
    const DropConst *dc;
    const u64 *addend; // optional, lazy-added to the first add_halves polynomials
    size_t add_batch_stride, add_poly_stride;
    u64 half_qlast;
    int L, logn, add_halves;
    bool vec;
    unsigned add_ginv; // GALOIS: the addend is read through the Galois permutation (the c0 of ckks::rotate / conjugate)
    HB_D int limb(int row) const { return row % (L - 1); }
    HB_D const u64 *src(int row) const { return z + ((size_t)(row / (L - 1)) << logn); }
    HB_D u64 pre(int row, int, u64 zz, const LimbConst &lc) const {
        const int k = row % (L - 1);
        // rescaling.cpp:58-59.  z < q_last; when q_last <= q_k the Barrett quotient hi64(z * floor((2^64-1)/q_k)) is 0
        // and the strict reduction is the identity, so the reference's result is z itself.
        u64 r = zz;
        if (!__ldg(&dc[k].z_below_q)) r = reduce_strict(barrett_lazy(zz, lc), lc.q);
        if (zz >= half_qlast) r += lc.q - __ldg(&dc[k].qlast_mod_q); // rescaling.cpp:63-68
        if (BGV) r = harvey_lazy(r, __ldg(&dc[k].t_mod_q), __ldg(&dc[k].t_mod_q_h), lc.nq); // mod_switch.cpp:70
        return r;
    }
    HB_D void store(int row, int i, u64 v, const LimbConst &lc) const {
        const int poly = row / (L - 1), k = row - poly * (L - 1);
        const DropConst *d = dc + k;
        u64 x = finish(NC ? hb_ld_ro(ct + ((size_t)(poly * L + k) << logn) + i) : hb_ld_stream(ct + ((size_t)(poly * L + k) << logn) + i), v, d, lc);
        const int h = poly & 1, b = poly >> 1;
        if (h < add_halves) { // ckks/arith.cpp:70-71, 84, 91
            const unsigned from = GALOIS ? galois_from((unsigned)i, add_ginv, logn) : (unsigned)i;
            const u64 *ap = addend + (size_t)b * add_batch_stride + (size_t)h * add_poly_stride + ((size_t)k << logn) + from;
            x = add_lazy(x, NC ? hb_ld_ro(ap) : hb_ld_stream(ap), lc.q2);
        }
        out[((size_t)row << logn) + i] = x;
    }
    // the epilogue's operands (this row of ct and of the addend) are first touched ~10 us after the CTA
    // starts: ask L2 for them now so the stores at the end do not wait on HBM
    HB_D void prefetch(int row, int first_word, int nwords) const {
        const int poly = row / (L - 1), k = row - poly * (L - 1);
        const u64 *c = ct + ((size_t)(poly * L + k) << logn) + first_word;
        for (int w = (int)threadIdx.x * 16; w < nwords; w += (int)blockDim.x * 16) hb_prefetch_l2(c + w);
        const int h = poly & 1, b = poly >> 1;
        if (h < add_halves) {
            // an aligned block of output slots gathers from one aligned block of the same size (galois_from)
            const unsigned from_block = GALOIS ? (galois_from((unsigned)first_word, add_ginv, logn) & ~(unsigned)(nwords - 1)) : (unsigned)first_word;
            const u64 *a = addend + (size_t)b * add_batch_stride + (size_t)h * add_poly_stride + ((size_t)k << logn) + from_block;
            for (int w = (int)threadIdx.x * 16; w < nwords; w += (int)blockDim.x * 16) hb_prefetch_l2(a + w);
        }
    }
    HB_D u64 finish(u64 x, u64 v, const DropConst *d, const LimbConst &lc) const {
        x = sub_lazy(x, v, lc.q2);                                                         // rescaling.cpp:73
        x = harvey_lazy(x, __ldg(&d->inv_qlast), __ldg(&d->inv_qlast_h), lc.nq);          // rescaling.cpp:74
        if (BGV) x = harvey_lazy(x, __ldg(&d->qlt_mod_q), __ldg(&d->qlt_mod_q_h), lc.nq); // mod_switch.cpp:76
        return x;
    }
    HB_D void store2(int row, int i, u64 v0, u64 v1, const LimbConst &lc) const {
        const int poly = row / (L - 1), k = row - poly * (L - 1);
        const DropConst *d = dc + k;
        const ulonglong2 x = NC ? hb_ld_ro2(ct + ((size_t)(poly * L + k) << logn) + i) : hb_ld_stream2(ct + ((size_t)(poly * L + k) << logn) + i);
        ulonglong2 r = make_ulonglong2(finish(x.x, v0, d, lc), finish(x.y, v1, d, lc));
        const int h = poly & 1, b = poly >> 1;
        if (h < add_halves) { // ckks/arith.cpp:70-71, 84, 91
            const u64 *arow = addend + (size_t)b * add_batch_stride + (size_t)h * add_poly_stride + ((size_t)k << logn);
            const ulonglong2 a = GALOIS ? galois_pair(arow, (unsigned)i, add_ginv, logn, NC) : (NC ? hb_ld_ro2(arow + i) : hb_ld_stream2(arow + i));
            r.x = add_lazy(r.x, a.x, lc.q2);
            r.y = add_lazy(r.y, a.y, lc.q2);
        }
        *reinterpret_cast<ulonglong2 *>(out + ((size_t)row << logn) + i) = r;
    }
};

// few ciphertexts: one cluster launch (defined after ks_pair.cuh); returns -1 when the call does not take that form
static int drop_last_cluster_form(Context &c, unsigned logn, size_t L, u64 t, const u64 *ct, u64 *out, size_t batch, const u64 *addend,
                                  size_t add_batch_stride, size_t add_poly_stride, int add_halves, unsigned add_ginv, const LimbConst *limbs,
                                  const DropSet *ds);

static int drop_last_impl(Context &c, unsigned logn, const u64 *moduli, size_t L, u64 t, const u64 *ct, u64 *out, size_t batch,
                          const u64 *addend, size_t add_batch_stride, size_t add_poly_stride, int add_halves, unsigned add_ginv, const u64 *z_ready);
int op_drop_last(Context &c, unsigned logn, const u64 *moduli, size_t L, u64 t, const u64 *ct, u64 *out, size_t batch,
                 const u64 *addend, size_t add_batch_stride, size_t add_poly_stride, int add_halves, unsigned add_ginv) {
    return drop_last_impl(c, logn, moduli, L, t, ct, out, batch, addend, add_batch_stride, add_poly_stride, add_halves, add_ginv, nullptr);
}
// z_ready: z = strict(INTT(last limb)) of the (single) ciphertext has been computed already (FusedDrop)
static int drop_last_impl(Context &c, unsigned logn, const u64 *moduli, size_t L, u64 t, const u64 *ct, u64 *out, size_t batch,
                          const u64 *addend, size_t add_batch_stride, size_t add_poly_stride, int add_halves, unsigned add_ginv, const u64 *z_ready) {
    if (!moduli || !ct || !out) return c.fail(1, "null operand");
    if (L < 2) return c.fail(1, "Unable to drop the only one prime.");
    if (batch == 0) return 0;
    int err = 0;
    const LimbConst *limbs = c.get_chain(logn, moduli, L, &err);
    if (!limbs) return err;
    const DropSet *ds = c.get_drop(logn, moduli, L, t, &err);
    if (!ds) return err;
    if (!z_ready)
        if (const int rc = drop_last_cluster_form(c, logn, L, t, ct, out, batch, addend, add_batch_stride, add_poly_stride, add_halves, add_ginv, limbs, ds);
            rc >= 0)
            return rc;
    const size_t n = (size_t)1 << logn;
    // waves bound the z scratch ([wave][2][N]) and keep every launch's row count inside int
    const size_t wave = wave_size(c, 2 * n, batch, 2 * L);
    u64 *z = z_ready ? const_cast<u64 *>(z_ready) : c.get_scratch(2, wave * 2 * n, &err);
    if (!z) return err;
    const bool vec = aligned16(ct) && aligned16(z) && aligned16(out) &&
                     (!addend || (aligned16(addend) && add_batch_stride % 2 == 0 && add_poly_stride % 2 == 0));
    const int halves = addend ? add_halves : 0;
    for (size_t b0 = 0; b0 < batch; b0 += wave) {
        const size_t nb = (batch - b0 < wave) ? batch - b0 : wave;
        const u64 *ct_w = ct + b0 * 2 * L * n;
        u64 *out_w = out + b0 * 2 * (L - 1) * n;
        const u64 *add_w = addend ? addend + b0 * add_batch_stride : nullptr;
        cudaError_t e;
        const int rows2 = (int)(nb * 2 * (L - 1));
        if (t) {
            DropInttIO<true> io1{ct_w, z, (int)L, (int)logn, ds->inv_t, ds->inv_t_h, aligned16(ct) && aligned16(z)};
            e = z_ready ? cudaSuccess : launch_ntt<false>(c.env(), logn, io1, limbs, (int)(nb * 2));
            if (e != cudaSuccess) return c.cuda_fail(e, "drop_last: intt launch");
            if (add_ginv != 1) return c.fail(1, "a Galois-permuted addend is a CKKS path (ckks/arith.cpp:75-93)");
            DropFwdIO<true> io2{ct_w, z, out_w, ds->dev, add_w, add_batch_stride, add_poly_stride, ds->half_qlast, (int)L, (int)logn, halves, vec, 1u};
            e = launch_ntt<true>(c.env(), logn, io2, limbs, rows2);
        } else {
            DropInttIO<false> io1{ct_w, z, (int)L, (int)logn, 0, 0, aligned16(ct) && aligned16(z)};
            e = z_ready ? cudaSuccess : launch_ntt<false>(c.env(), logn, io1, limbs, (int)(nb * 2));
            if (e != cudaSuccess) return c.cuda_fail(e, "drop_last: intt launch");
            if (add_ginv == 1) {
                DropFwdIO<false> io2{ct_w, z, out_w, ds->dev, add_w, add_batch_stride, add_poly_stride, ds->half_qlast, (int)L, (int)logn, halves, vec, 1u};
                e = launch_ntt<true>(c.env(), logn, io2, limbs, rows2);
            } else {
                DropFwdIO<false, true> io2{ct_w, z, out_w, ds->dev, add_w, add_batch_stride, add_poly_stride, ds->half_qlast, (int)L, (int)logn, halves, vec, add_ginv};
                e = launch_ntt<true>(c.env(), logn, io2, limbs, rows2);
            }
        }
        if (e != cudaSuccess) return c.cuda_fail(e, "drop_last: ntt launch");
    }
    return 0;
}

} // namespace hb
#include "ks_pair.cuh"
namespace hb {

// ------------------------------------------------------------------------------------------
// one ciphertext per call, N = 16384 / 32768: inner product + inverse transform of the P limb in one launch (FusedDrop)
// ------------------------------------------------------------------------------------------
// Grid: CTAs [0, 2 C) are the two clusters of the inverse transform (emulator: the LAST 2 C, it runs CTAs one after another);
// the others compute the inner product one ciphertext per thread (ext_mac_unit), the P limb's two rows FIRST.  Every CTA
// that has stored P-limb words adds one to sync[0]; the transform clusters wait until the launch's `p_blocks` have arrived,
// then load the rows like the separate launch would (DropInttIO).  The producers never wait, and they are the first CTAs
// after the transform's own in launch order, so the wait cannot starve them.  The last transform CTA past the wait (sync[1])
// puts both words back to zero: every launch — a replay of a captured graph too — starts from the same state.
template <int LOGN, bool BGV, bool GALOIS>
HB_GLOBAL(256, 2)
ext_mac_intt_kernel(const u64 *__restrict__ in, size_t in_batch_stride, const u64 *__restrict__ dec, const u64 *__restrict__ key,
                    u64 *__restrict__ out, const LimbConst *__restrict__ limbs, int L, unsigned ginv, u64 *__restrict__ z, u64 inv_t, u64 inv_t_h,
                    unsigned long long *sync, unsigned p_blocks) {
    constexpr NttPlan pl = plan_for(LOGN, false, 1);
    static_assert(pl.threads == 256 && pl.xchg, "the latency plans of these ring sizes: 8-CTA clusters of 256 threads");
    constexpr unsigned C = 1u << pl.lpre;
    HB_SHARED_U64(sm);
#if defined(HB_KERNEL_SIM)
    const bool transforms = blockIdx.x >= gridDim.x - 2 * C;
    const unsigned tb = blockIdx.x - (gridDim.x - 2 * C), mb = blockIdx.x;
#else
    const bool transforms = blockIdx.x < 2 * C;
    const unsigned tb = blockIdx.x, mb = blockIdx.x - 2 * C;
#endif
    hb_pdl_wait();
    if (transforms) {
        const int h = (int)(tb >> pl.lpre), B = (int)(tb & (C - 1));
#if !defined(HB_KERNEL_SIM)
        if (threadIdx.x == 0) {
            unsigned long long seen;
            do {
                asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(seen) : "l"(sync) : "memory");
            } while (seen < p_blocks);
            if (atomicAdd(sync + 1, 1ull) == 2 * C - 1) { // every producer has arrived, every waiter has seen it
                sync[1] = 0;
                sync[0] = 0;
            }
        }
        __syncthreads();
#endif
        // the P limb's rows of e were written by other CTAs of THIS launch: no non-coherent loads (DropInttIO reads with hb_ld_stream)
        const DropInttIO<BGV> io{out, z, L + 1, LOGN, inv_t, inv_t_h, true};
        inv_passes<LOGN, 256, 0, 1>(sm, io, limbs[L], h, B);
        return;
    }
    // unit order: limb L first, then 0 .. L - 1; a pair of words per thread
    const size_t per_limb = (size_t)1 << (LOGN - 1), unit = (size_t)mb * 256 + threadIdx.x;
    const size_t slot = unit / per_limb; // 0: the P limb
    if (slot <= (size_t)L) {
        const size_t k = slot == 0 ? (size_t)L : slot - 1;
        ext_mac_unit<1, 2, GALOIS>(in, in_batch_stride, dec, key, out, limbs, L, LOGN, 1, k * per_limb + (unit - slot * per_limb), ginv, 0, true);
    }
    if (mb < p_blocks) {
        __syncthreads();
        if (threadIdx.x == 0) {
#if !defined(HB_KERNEL_SIM)
            __threadfence();
            atomicAdd(sync, 1ull);
#endif
        }
    }
}
template <int LOGN>
static cudaError_t launch_mac_intt_logn(Context &c, const LimbConst *limbs, size_t L, const u64 *in, size_t in_batch_stride, const u64 *dec,
                                        const u64 *key, u64 *out, unsigned ginv, const FusedDrop &fd) {
    constexpr NttPlan pl = plan_for(LOGN, false, 1);
    constexpr int C = 1 << pl.lpre, smem = smem_words(1 << (LOGN - pl.lpre)) * 8;
    static_assert(smem <= 48 * 1024, "no opt-in needed");
    const unsigned per_limb_blocks = (unsigned)(((size_t)1 << (LOGN - 1)) / 256), p_blocks = per_limb_blocks; // the P limb: both rows' pairs share a thread
    unsigned blocks = (unsigned)((L + 1) * per_limb_blocks);
    blocks = (blocks + C - 1) / C * C + 2 * C; // whole clusters
    unsigned long long *counters = c.grid_barrier_counter();
    if (!counters) return cudaErrorMemoryAllocation;
    c.stats.launches++;
    auto go = [&](auto kern) {
        return HB_LAUNCH_CLUSTER(kern, blocks, 256, smem, c.stream, C, in, in_batch_stride, dec, key, out, limbs, (int)L, ginv, fd.z, fd.inv_t, fd.inv_t_h,
                                 counters + 1, p_blocks);
    };
    if (fd.bgv) return ginv == 1 ? go(ext_mac_intt_kernel<LOGN, true, false>) : cudaErrorInvalidValue;
    return ginv == 1 ? go(ext_mac_intt_kernel<LOGN, false, false>) : go(ext_mac_intt_kernel<LOGN, false, true>);
}
static cudaError_t launch_mac_intt(Context &c, unsigned logn, const LimbConst *limbs, size_t L, const u64 *in, size_t in_batch_stride,
                                   const u64 *dec, const u64 *key, u64 *out, unsigned ginv, const FusedDrop &fd) {
    return logn == 14 ? launch_mac_intt_logn<14>(c, limbs, L, in, in_batch_stride, dec, key, out, ginv, fd)
                      : launch_mac_intt_logn<15>(c, limbs, L, in, in_batch_stride, dec, key, out, ginv, fd);
}
// Whether a key switch + drop of ONE ciphertext takes that form (option "fused_drop", default on); fills `fd` (z in scratch slot 2).
// Any failure here leaves the decision to the plain path, which reports its own errors in its own order.
static FusedDrop *fused_drop_wanted(Context &c, unsigned logn, const u64 *ext_moduli, size_t L, u64 t, size_t batch, unsigned ginv, FusedDrop *fd) {
    if (!c.fused_drop || batch != 1 || (logn != 14 && logn != 15) || c.force_generic || c.latency_rows == 0 || (t && ginv != 1)) return nullptr;
    int err = 0;
    const DropSet *ds = c.get_drop(logn, ext_moduli, L + 1, t, &err);
    if (!ds) return nullptr;
    u64 *z = c.get_scratch(2, (size_t)2 << logn, &err);
    if (!z) return nullptr;
    *fd = FusedDrop{z, ds->inv_t, ds->inv_t_h, t != 0, false};
    return fd;
}

// ------------------------------------------------------------------------------------------
// few ciphertexts per call: the key switch and the drop of P as two launches (ks_pair.cuh)
// ------------------------------------------------------------------------------------------
static inline bool ranges_overlap(const u64 *a, size_t na, const u64 *b, size_t nb) { return a < b + nb && b < a + na; }

// Whether a call takes the two-launch form.  Option "pair_path": 0 never, 2 whenever the shapes allow, 1 (default) for
// N = 4096 / 8192 while the batch's fan-out rows number at most pair_fill_pct (100) % of the SMs — measured crossovers
// (profiles/r4_pair_path.md): N = 8192, L = 4: 12 ciphertexts per call still gain 6 %, 16 lose; N = 4096, L = 2: 32 gain, 48 lose;
// L = 3: 16 gain, 32 lose; N = 8192, L = 6: a tie from 4 on.  Larger rings are waves of CTAs rather than latency-bound rows even
// for one ciphertext, and the wave path's streaming wins.
static bool pair_path_wanted(const Context &c, unsigned logn, size_t L, size_t batch) {
    if (c.pair_path == 0 || c.force_generic || !has_latency2_plan((int)logn) || L == 0 || L > 64) return false;
    const size_t n = (size_t)1 << logn;
    if (batch * L * (L + 1) * n * 8 > c.scratch_cap_bytes || batch * L * L > ((size_t)1 << 20)) return false;
    if (c.pair_path == 2) return true;
    return batch * L * L <= (size_t)c.sm_count * c.pair_fill_pct / 100;
}
// Targets per cluster.  A cluster runs one inverse transform and then `tpc` forward ones; fewer targets per cluster means more
// clusters (each repeating the inverse transform) and a shorter chain in each.  Pick the split with the shortest estimated
// time: (waves of resident clusters) x (transforms in a chain), ties to the shorter chain.  `cap`: CTAs of this kernel the
// GPU holds at once.
static int pair_targets_per_cluster(const Context &c, size_t cluster, size_t groups, size_t L, size_t cap) {
    if (c.pair_tpc > 0) return c.pair_tpc < (int)L ? c.pair_tpc : (int)L;
    size_t best = 1, best_cost = ~(size_t)0;
    for (size_t tpc = 1; tpc <= L; tpc++) {
        const size_t ctas = groups * ((L + tpc - 1) / tpc) * cluster, waves = (ctas + cap - 1) / cap, cost = waves * (1 + tpc);
        if (cost < best_cost) best = tpc, best_cost = cost;
    }
    return (int)best;
}
// CTAs of a cluster kernel resident at once on this device (asked once per kernel)
template <class K>
static size_t pair_resident_ctas(Context &c, K kern, int threads, int cluster, int smem) {
    int &have = c.cluster_cap[reinterpret_cast<const void *>(kern)];
#if defined(HB_KERNEL_SIM)
    (void)threads, (void)smem;
    if (have == 0) have = c.sm_count >= cluster ? c.sm_count / cluster * cluster : cluster;
#else
    if (have == 0) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned)(cluster * c.sm_count));
        cfg.blockDim = dim3((unsigned)threads);
        cfg.dynamicSmemBytes = (size_t)smem;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = (unsigned)cluster;
        at[0].val.clusterDim.y = at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n < 1) {
            cudaGetLastError();
            n = c.sm_count >= cluster ? c.sm_count / cluster : 1; // one CTA per SM
        }
        have = n * cluster;
        if (std::getenv("HEHUB_B200_DEBUG")) std::fprintf(stderr, "hehub_b200: cluster kernel %p: %d clusters of %d CTAs resident at once (smem %d)\n", (const void *)kern, n, cluster, smem);
    }
#endif
    return (size_t)have;
}

template <int LOGN, int MODE>
constexpr int kPairSmem = staged_mode(MODE) ? kStagedSmemBytes<LOGN, MODE> : smem_words(1 << (LOGN - plan_for(LOGN, true, MODE).lpre)) * 8;
// dynamic shared memory above 48 KB: opt in once per kernel and device (the context is per device)
template <class K>
static cudaError_t pair_opt_in(Context &c, K kern, int smem) {
    if (smem <= 48 * 1024) return cudaSuccess;
    int &have = c.smem_opt_in[reinterpret_cast<const void *>(kern)];
    if (have >= smem) return cudaSuccess;
    const cudaError_t ce = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (ce == cudaSuccess) have = smem;
    return ce;
}
// the drop launch: e (SRC) -> out
template <int LOGN, int MODE, bool BGV, class SRC, class ADD>
static int launch_pair_drop(Context &c, const SRC &src, const ADD &add, u64 *out, const LimbConst *limbs, const DropSet *ds, size_t L, size_t batch) {
    constexpr NttPlan pl = plan_for(LOGN, true, MODE);
    constexpr int C = 1 << pl.lpre, smem = kPairSmem<LOGN, MODE>;
    static_assert(smem <= 200 * 1024, "one CTA per SM");
    auto kern = ks_drop_kernel<LOGN, MODE, BGV, SRC, ADD>;
    if (cudaError_t ce = pair_opt_in(c, kern, smem); ce != cudaSuccess) return c.cuda_fail(ce, "drop launch (cluster form): shared memory");
    const int tpc = pair_targets_per_cluster(c, C, batch * 2, L, pair_resident_ctas(c, kern, pl.threads, C, smem)), chunks = ((int)L + tpc - 1) / tpc;
    const cudaError_t e = HB_LAUNCH_CLUSTER(kern, (unsigned)(batch * 2 * chunks * C), pl.threads, smem, c.stream, C, src, add, out, limbs,
                                            (const DropConst *)ds->dev, ds->half_qlast, ds->inv_t, ds->inv_t_h, (int)L, tpc, chunks);
    c.stats.launches++;
    return e == cudaSuccess ? 0 : c.cuda_fail(e, "drop launch (cluster form)");
}
template <int LOGN, int MODE, class IN, bool BGV, class ADD>
static int launch_pair(Context &c, const IN &in, const ADD &add, const u64 *key, u64 *out, u64 *dec, u64 *quad, const LimbConst *limbs,
                       const DropSet *ds, size_t L, size_t batch) {
    constexpr NttPlan pl = plan_for(LOGN, true, MODE);
    constexpr int C = 1 << pl.lpre, smem = kPairSmem<LOGN, MODE>;
    auto kern = ks_fan_kernel<LOGN, MODE, IN>;
    if (cudaError_t ce = pair_opt_in(c, kern, smem); ce != cudaSuccess) return c.cuda_fail(ce, "key switch (two-launch form): shared memory");
    const int tpc_a = pair_targets_per_cluster(c, C, batch * L, L, pair_resident_ctas(c, kern, pl.threads, C, smem)), chunks_a = ((int)L + tpc_a - 1) / tpc_a;
    const unsigned main_ctas = (unsigned)(batch * L * chunks_a * C);
    // the tensor product's d0, d1 on CTAs of the same launch that have no transform to run: as many as fill the SMs the
    // transforms leave idle (at most one pair of words per thread)
    unsigned extra = 0;
    size_t tensor_units = 0;
    if (quad) {
        tensor_units = (batch * L) << (LOGN - 1);
        const unsigned want = (unsigned)((tensor_units + pl.threads - 1) / pl.threads);
        const unsigned idle = main_ctas < (unsigned)c.sm_count ? (unsigned)c.sm_count - main_ctas : 0;
        extra = want < idle ? want : idle;
        if (extra < (want + 3) / 4) extra = (want + 3) / 4; // no idle SMs (a forced large batch): at most four pairs of words per thread
        extra = (extra + C - 1) / C * C;
    }
    const cudaError_t e = HB_LAUNCH_CLUSTER(kern, main_ctas + extra, pl.threads, smem, c.stream, C, in, dec, limbs, (int)L, tpc_a, chunks_a, main_ctas,
                                            quad, tensor_units);
    c.stats.launches++;
    if (e != cudaSuccess) return c.cuda_fail(e, "key switch (two-launch form): fan-out launch");
    return launch_pair_drop<LOGN, MODE, BGV>(c, KsSrcMac{dec, key, (int)L, LOGN}, add, out, limbs, ds, L, batch);
}
// the mode-2 plans (8-CTA clusters, tables staged in shared memory: ntt_plan.h); the 4-CTA latency plans measured slower at every
// batch size once the targets were split by pair_targets_per_cluster (profiles/r4_pair_path.md)
template <class IN, bool BGV, class ADD>
static int launch_pair_logn(Context &c, unsigned logn, const IN &in, const ADD &add, const u64 *key, u64 *out, u64 *dec, u64 *quad,
                            const LimbConst *limbs, const DropSet *ds, size_t L, size_t batch) {
    switch (logn) {
    case 12: return launch_pair<12, 2, IN, BGV, ADD>(c, in, add, key, out, dec, quad, limbs, ds, L, batch);
    case 13: return launch_pair<13, 2, IN, BGV, ADD>(c, in, add, key, out, dec, quad, limbs, ds, L, batch);
    }
    return c.fail(1, "two-launch form: ring size without a plan that hands over in registers");
}
// ckks::rescale_inplace / bgv::mod_switch_inplace of a few ciphertexts as ONE cluster launch: the inverse transform of the last
// limb handed in registers to the forward transforms of the others (ks_drop_kernel with the ciphertext itself as source).
// Taken while the batch has at most one output row per SM (8-CTA clusters, tables staged in shared memory).
template <bool BGV, class ADD>
static int launch_drop_form(Context &c, unsigned logn, const KsSrcPlain &src, const ADD &add, u64 *out, const LimbConst *limbs, const DropSet *ds,
                            size_t Lk, size_t batch) {
    switch (logn) {
    case 12: return launch_pair_drop<12, 2, BGV>(c, src, add, out, limbs, ds, Lk, batch);
    case 13: return launch_pair_drop<13, 2, BGV>(c, src, add, out, limbs, ds, Lk, batch);
    }
    return -1;
}
static int drop_last_cluster_form(Context &c, unsigned logn, size_t L, u64 t, const u64 *ct, u64 *out, size_t batch, const u64 *addend,
                                  size_t add_batch_stride, size_t add_poly_stride, int add_halves, unsigned add_ginv, const LimbConst *limbs,
                                  const DropSet *ds) {
    if (c.pair_path == 0 || c.force_generic || !has_latency2_plan((int)logn) || L > 64) return -1;
    const size_t n = (size_t)1 << logn, Lk = L - 1, rows = batch * 2 * Lk;
    if (rows > ((size_t)1 << 20)) return -1;
    // measured (profiles/r4_pair_path.md, N = 8192, three remaining limbs): this form wins up to 24 ciphertexts per call
    if (c.pair_path != 2 && rows > (size_t)c.sm_count) return -1;
    if (!aligned16(ct) || !aligned16(out) || ranges_overlap(ct, batch * 2 * L * n, out, batch * 2 * Lk * n)) return -1;
    const int halves = addend ? add_halves : 0;
    if (halves && (!aligned16(addend) || add_batch_stride % 2 || add_poly_stride % 2)) return -1;
    if (t && add_ginv != 1) return -1; // the wave path reports it
    const KsSrcPlain src{ct, (int)Lk, (int)logn};
    if (add_ginv != 1) {
        const KsAddPlain<true> add{addend, add_batch_stride, add_poly_stride, halves, (int)logn, add_ginv};
        return launch_drop_form<false>(c, logn, src, add, out, limbs, ds, Lk, batch);
    }
    const KsAddPlain<false> add{addend, add_batch_stride, add_poly_stride, halves, (int)logn, 1u};
    return t ? launch_drop_form<true>(c, logn, src, add, out, limbs, ds, Lk, batch)
             : launch_drop_form<false>(c, logn, src, add, out, limbs, ds, Lk, batch);
}

// constants and workspace of the two-launch form; the checks come in the order the wave path makes them
static int pair_setup(Context &c, unsigned logn, const u64 *ext_moduli, size_t L, u64 t, size_t batch, const LimbConst **limbs, const DropSet **ds,
                      u64 **dec) {
    for (size_t k = 0; k <= L; k++)
        if (!(ext_moduli[k] & 1)) return c.fail(1, "Montgomery reduction needs odd moduli");
    int err = 0;
    if (!(*limbs = c.get_chain(logn, ext_moduli, L + 1, &err))) return err;
    if (!(*ds = c.get_drop(logn, ext_moduli, L + 1, t, &err))) return err;
    if (!(*dec = c.get_scratch(1, (batch * L * (L + 1)) << logn, &err))) return err;
    return 0;
}

// ------------------------------------------------------------------------------------------
// relinearize / mult
// ------------------------------------------------------------------------------------------
int op_relinearize(Context &c, unsigned logn, const u64 *ext_moduli, size_t L, u64 t, const u64 *quad, const u64 *key,
                   u64 *out, size_t batch) {
    if (!ext_moduli || !quad || !key || !out) return c.fail(1, "null operand");
    if (L == 0) return c.fail(1, "Empty RGSW ciphertext.");
    if (batch == 0) return 0;
    const size_t n = (size_t)1 << logn;
    if (pair_path_wanted(c, logn, L, batch) && aligned16(quad) && aligned16(key) && aligned16(out) &&
        !ranges_overlap(quad, batch * 3 * L * n, out, batch * 2 * L * n)) {
        const LimbConst *limbs;
        const DropSet *ds;
        u64 *dec;
        if (int rc = pair_setup(c, logn, ext_moduli, L, t, batch, &limbs, &ds, &dec)) return rc;
        const KsInPlain<false> in{quad + 2 * L * n, 3 * L * n, (int)logn, 1u};
        const KsAddPlain<false> add{quad, 3 * L * n, L * n, 2, (int)logn, 1u};
        return t ? launch_pair_logn<KsInPlain<false>, true, KsAddPlain<false>>(c, logn, in, add, key, out, dec, nullptr, limbs, ds, L, batch)
                 : launch_pair_logn<KsInPlain<false>, false, KsAddPlain<false>>(c, logn, in, add, key, out, dec, nullptr, limbs, ds, L, batch);
    }
    // waves bound the scratch: per ct  c: L, dec: L(L+1), e: 2(L+1), z: 2  rows of N words
    const size_t per_ct = (L + L * (L + 1) + 2 * (L + 1) + 2) * n;
    const size_t wave = wave_size(c, per_ct, batch, L * (L + 1));
    int err = 0;
    u64 *ebuf = c.get_scratch(3, wave * 2 * (L + 1) * n, &err);
    if (!ebuf) return err;
    FusedDrop fd_store, *fd = fused_drop_wanted(c, logn, ext_moduli, L, t, batch, 1u, &fd_store);
    for (size_t b0 = 0; b0 < batch; b0 += wave) {
        const size_t nb = (batch - b0 < wave) ? batch - b0 : wave;
        const u64 *q = quad + b0 * 3 * L * n;
        if (int rc = ext_prod_impl(c, logn, ext_moduli, L, q + 2 * L * n, 3 * L * n, key, ebuf, nb, 1u, fd)) return rc;
        if (int rc = drop_last_impl(c, logn, ext_moduli, L + 1, t, ebuf, out + b0 * 2 * L * n, nb, q, 3 * L * n, L * n, 2, 1u, fd && fd->done ? fd->z : nullptr))
            return rc;
    }
    return 0;
}

// ------------------------------------------------------------------------------------------
// ONE launch for one ckks::mult (ckks.h:270-274) — N = 4096 / 8192, one ciphertext pair per call; option "single_launch".
// The six steps of the wave path (tensor, INTT of d2, fan-out, inner product, INTT of the P limb, forward + drop epilogue)
// run as phases of one grid of 4-CTA clusters, separated by grid barriers; phases reuse the transform passes and IO
// policies of the multi-launch path unchanged (thin latency plans, clusters exchanging through distributed shared
// memory).  Operands a later phase reads were written by an earlier phase of the SAME launch, so nothing here uses the
// non-coherent load path for them (DropFwdIO<.., NC = false>).
// MEASURED, and therefore OFF by default (profiles/r3_latency_plans.md, C3 shape, one pair per call): six launches chained
// by programmatic dependent launch 33.2 us; this kernel with a counter barrier 39.2 us, with cooperative-groups grid.sync
// 43.1 us; the six launches replayed from a CUDA graph 34.9 us.  A phase costs what its kernel costs (~5 us: one row's
// butterflies on a few SMs, bound by per-warp issue and the multiplier of those SMs); the launch boundaries that this form
// removes were already hidden.
// ------------------------------------------------------------------------------------------
#if !defined(HB_KERNEL_SIM)
} // namespace hb
#include <cooperative_groups.h>
namespace hb {
#ifndef HB_MULT_ONE_COOP
#define HB_MULT_ONE_COOP 0
#endif
// Grid barrier of the single-launch kernel.  Every CTA of the grid is resident (the grid is sized from
// cudaOccupancyMaxActiveClusters and capped at one CTA per SM), so a counter in global memory that only grows is enough:
// barrier number b of a launch that started at count c0 is passed once the count reaches c0 + (b + 1) * gridDim.x.
HB_D void grid_barrier(unsigned long long *counter, unsigned long long &passed) {
    __syncthreads();
    if (threadIdx.x == 0) {
        passed += gridDim.x;
        __threadfence();
        atomicAdd(counter, 1ull);
        unsigned long long seen;
        do {
            asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(seen) : "l"(counter) : "memory");
        } while (seen < passed);
    }
    __syncthreads();
}
struct MultOneArgs {
    unsigned long long *barrier; // grid-barrier counter (device), and its value when this launch starts
    unsigned long long barrier_base;
    const u64 *ct1, *ct2, *key;
    u64 *out;
    u64 *quad, *cbuf, *dec, *ebuf, *z; // workspaces: [3][L][N], [L][N], [L][L+1][N], [2][L+1][N], [2][N]
    const LimbConst *limbs;             // q_0 .. q_{L-1}, P with tables for this ring
    const DropConst *dc;
    u64 half_qlast;
    int L;
};
template <int LOGN>
__global__ void __launch_bounds__(plan_for(LOGN, true, 1).threads, 1) ckks_mult_one_kernel(const MultOneArgs a) {
    constexpr NttPlan plf = plan_for(LOGN, true, 1), pli = plan_for(LOGN, false, 1);
    static_assert(plf.xchg && pli.xchg && plf.threads == pli.threads && plf.lpre == pli.lpre, "one cluster shape for both directions");
    constexpr int T = plf.threads, C = 1 << plf.lpre;
    HB_SHARED_U64(sm);
#if HB_MULT_ONE_COOP
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
#define HB_GRID_SYNC() grid.sync()
#else
    unsigned long long passed = a.barrier_base;
#define HB_GRID_SYNC() grid_barrier(a.barrier, passed)
    hb_pdl_wait();
#endif
    const int cid = blockIdx.x / C, B = blockIdx.x % C, nclusters = gridDim.x / C, L = a.L;
    const size_t n = (size_t)1 << LOGN, Ln = (size_t)L << LOGN;
    const size_t tid = (size_t)blockIdx.x * T + threadIdx.x, nthreads = (size_t)gridDim.x * T;
    // tensor product (ckks/arith.cpp:55-62)
    for (size_t gid = tid; gid < ((size_t)L << (LOGN - 1)); gid += nthreads) tensor_unit<2>(a.ct1, a.ct2, a.quad, a.limbs, L, LOGN, gid);
    HB_GRID_SYNC();
    { // c = strict(INTT(d2))                                           rgsw.cpp:103-105
        const ExtInttIO<false> io{a.quad + 2 * Ln, 3 * Ln, a.cbuf, L, LOGN, true, 1u};
        for (int row = cid; row < L; row += nclusters) inv_passes<LOGN, T, 0, 1>(sm, io, a.limbs[io.limb(row)], row, B);
    }
    HB_GRID_SYNC();
    { // dec[p][k] = NTT_{q_k}(c[p]), k != p                            rgsw.cpp:108-119
        const ExtFanoutIO io{a.cbuf, a.dec, L, LOGN, true};
        for (int row = cid; row < L * L; row += nclusters) {
            hb_cluster_arrive(); // every CTA of the cluster is done with its previous row before words are scattered into it
            fwd_passes<LOGN, T, 0, 1>(sm, io, a.limbs[io.limb(row)], row, B);
        }
    }
    HB_GRID_SYNC();
    // e[h][k] = Mont128(sum_p dec[p][k] * key[p][h][k])                 rgsw.cpp:126-153
    for (size_t gid = tid; gid < (size_t)(L + 1) * (n / 2); gid += nthreads)
        ext_mac_unit<1, 2, false>(a.quad + 2 * Ln, 3 * Ln, a.dec, a.key, a.ebuf, a.limbs, L, LOGN, 1, gid, 1u);
    HB_GRID_SYNC();
    { // z = strict(INTT_P(e[.][L]))                                    rescaling.cpp:47-50
        const DropInttIO<false> io{a.ebuf, a.z, L + 1, LOGN, 0, 0, true};
        for (int row = cid; row < 2; row += nclusters) inv_passes<LOGN, T, 0, 1>(sm, io, a.limbs[io.limb(row)], row, B);
    }
    HB_GRID_SYNC();
    { // out = (e - NTT(centre(z))) / P + (d0, d1)                      rescaling.cpp:52-75, ckks/arith.cpp:70-71
        const DropFwdIO<false, false, false> io{a.ebuf, a.z, a.out, a.dc, a.quad, 3 * Ln, Ln, a.half_qlast, L + 1, LOGN, 2, true, 1u};
        for (int row = cid; row < 2 * L; row += nclusters) {
            hb_cluster_arrive();
            fwd_passes<LOGN, T, 0, 1>(sm, io, a.limbs[io.limb(row)], row, B);
        }
    }
#if !HB_MULT_ONE_COOP
    if (threadIdx.x == 0) atomicAdd(a.barrier, 1ull); // sixth arrival: the count ends at base + 6 * gridDim.x whatever happens next
#endif
}
#undef HB_GRID_SYNC

template <int LOGN>
static int launch_mult_one(Context &c, const MultOneArgs &args) {
    constexpr NttPlan pl = plan_for(LOGN, true, 1);
    constexpr int C = 1 << pl.lpre, smem = smem_words(1 << (LOGN - pl.lpre)) * 8;
    auto kern = ckks_mult_one_kernel<LOGN>;
    cudaLaunchConfig_t cfg{};
    cfg.blockDim = dim3(pl.threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = c.stream;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = C;
    at[0].val.clusterDim.y = at[0].val.clusterDim.z = 1;
#if HB_MULT_ONE_COOP
    at[1].id = cudaLaunchAttributeCooperative;
    at[1].val.cooperative = 1;
#else
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = hb_pdl_enabled() ? 1 : 0;
#endif
    cfg.attrs = at;
    cfg.numAttrs = 2;
    int &clusters = c.mult_one_clusters[LOGN - 12];
    if (clusters == 0) { // how many clusters are resident at once decides the grid: a cooperative grid must fit the GPU whole
        cfg.gridDim = dim3(C * c.sm_count);
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n < 1) {
            cudaGetLastError();
            clusters = -1;
        } else {
            clusters = n < c.sm_count / C ? n : c.sm_count / C; // one CTA per SM is plenty: phases have at most L * L rows
        }
    }
    if (clusters < 0) return -1;
    cfg.gridDim = dim3((unsigned)(clusters * C));
    MultOneArgs a2 = args;
    a2.barrier = c.grid_barrier_counter();
    if (!a2.barrier) return -1;
    a2.barrier_base = c.grid_barrier_count; // launches of one stream run in order: the count after all earlier launches
    c.grid_barrier_count += 6ull * cfg.gridDim.x; // six barriers per launch (the last one keeps the count exact)
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, a2);
    if (e != cudaSuccess) { // cooperative + cluster launch refused on this driver: remember and use the six launches
        cudaGetLastError();
        clusters = -1;
        return -1;
    }
    c.stats.launches++;
    return 0;
}
#endif

int op_mult_relin(Context &c, unsigned logn, const u64 *ext_moduli, size_t L, u64 t, const u64 *ct1, const u64 *ct2,
                  const u64 *key, u64 *out, size_t batch) {
    if (!ext_moduli || !ct1 || !ct2 || !key || !out) return c.fail(1, "null operand");
    if (L == 0) return c.fail(1, "Empty RGSW ciphertext.");
    if (batch == 0) return 0;
    const size_t n = (size_t)1 << logn;
    if (pair_path_wanted(c, logn, L, batch) && !(c.single_launch && batch == 1 && t == 0 && logn <= 13) && aligned16(ct1) && aligned16(ct2) &&
        aligned16(key) && aligned16(out) && !ranges_overlap(ct1, batch * 2 * L * n, out, batch * 2 * L * n) &&
        !ranges_overlap(ct2, batch * 2 * L * n, out, batch * 2 * L * n)) {
        for (size_t k = 0; k < L; k++)
            if (!(ext_moduli[k] & 1)) return c.fail(1, "Montgomery multiplication needs odd moduli");
        const LimbConst *limbs;
        const DropSet *ds;
        u64 *dec;
        if (int rc = pair_setup(c, logn, ext_moduli, L, t, batch, &limbs, &ds, &dec)) return rc;
        int err = 0;
        u64 *d01 = c.get_scratch(4, batch * 2 * L * n, &err); // d0, d1 of the tensor product: [batch][2][L][N]
        if (!d01) return err;
        const KsInTensor in{ct1, ct2, (int)L, (int)logn};
        const KsAddPlain<false> add{d01, 2 * L * n, L * n, 2, (int)logn, 1u};
        return t ? launch_pair_logn<KsInTensor, true, KsAddPlain<false>>(c, logn, in, add, key, out, dec, d01, limbs, ds, L, batch)
                 : launch_pair_logn<KsInTensor, false, KsAddPlain<false>>(c, logn, in, add, key, out, dec, d01, limbs, ds, L, batch);
    }
    const size_t per_ct = (3 * L + L + L * (L + 1) + 2 * (L + 1) + 2) * n;
    const size_t wave = wave_size(c, per_ct, batch, L * (L + 1));
    int err = 0;
    u64 *quad = c.get_scratch(4, wave * 3 * L * n, &err);
    if (!quad) return err;
#if !defined(HB_KERNEL_SIM)
    // one ciphertext pair, CKKS, N = 4096 / 8192, 16-byte aligned operands: the single-launch kernel
    if (batch == 1 && t == 0 && (logn == 12 || logn == 13) && c.single_launch && !c.force_generic && aligned16(ct1) && aligned16(ct2) &&
        aligned16(key) && aligned16(out)) {
        for (size_t k = 0; k <= L; k++)
            if (!(ext_moduli[k] & 1)) return c.fail(1, "Montgomery reduction needs odd moduli");
        const LimbConst *limbs = c.get_chain(logn, ext_moduli, L + 1, &err);
        if (!limbs) return err;
        const DropSet *ds = c.get_drop(logn, ext_moduli, L + 1, 0, &err);
        if (!ds) return err;
        MultOneArgs a{nullptr, 0, ct1, ct2, key, out, quad, nullptr, nullptr, nullptr, nullptr, limbs, ds->dev, ds->half_qlast, (int)L};
        if (!(a.cbuf = c.get_scratch(0, L * n, &err))) return err;
        if (!(a.dec = c.get_scratch(1, L * (L + 1) * n, &err))) return err;
        if (!(a.z = c.get_scratch(2, 2 * n, &err))) return err;
        if (!(a.ebuf = c.get_scratch(3, 2 * (L + 1) * n, &err))) return err;
        const int rc = logn == 12 ? launch_mult_one<12>(c, a) : launch_mult_one<13>(c, a);
        if (rc == 0) return 0; // otherwise: the launch form is not available here, fall through to the six launches
    }
#endif
    for (size_t b0 = 0; b0 < batch; b0 += wave) {
        const size_t nb = (batch - b0 < wave) ? batch - b0 : wave;
        if (int rc = op_ckks_tensor(c, logn, ext_moduli, L, ct1 + b0 * 2 * L * n, ct2 + b0 * 2 * L * n, quad, nb)) return rc;
        if (int rc = op_relinearize(c, logn, ext_moduli, L, t, quad, key, out + b0 * 2 * L * n, nb)) return rc;
    }
    return 0;
}

// ------------------------------------------------------------------------------------------
// RLWE decrypt / encrypt cores — src/fhe/primitives/rlwe.cpp:34-71.
// The products with the secret key and the lazy additions ride in the transforms' load / store.
// ------------------------------------------------------------------------------------------
// pt[b] = strict( INTT( c0 + c1 * sk ) )                                   rlwe.cpp:63-71
struct DecryptIO {
    const u64 *ct; // [batch][2][L][N], NTT form
    const u64 *sk; // [L][N], NTT form, shared by the batch
    u64 *pt;       // [batch][L][N]
    int L, logn;
    bool vec;
    HB_D int limb(int row) const { return row % L; }
    HB_D const u64 *src(int row) const { return ct + (((size_t)(row / L) * 2 * L + row % L) << logn); }
    HB_D u64 pre(int row, int i, u64 c0, const LimbConst &lc) const {
        const int k = row % L;
        const u64 c1 = hb_ld_ro(src(row) + ((size_t)L << logn) + i), s = __ldg(sk + ((size_t)k << logn) + i);
        return add_lazy(c0, mul_hybrid_lazy(c1, s, lc), lc.q2); // rns.cpp:120-140, 58-86
    }
    HB_D void store(int row, int i, u64 v, const LimbConst &lc) const { pt[((size_t)row << logn) + i] = reduce_strict(v, lc.q); }
    HB_D void store2(int row, int i, u64 v0, u64 v1, const LimbConst &lc) const {
        *reinterpret_cast<ulonglong2 *>(pt + ((size_t)row << logn) + i) = make_ulonglong2(reduce_strict(v0, lc.q), reduce_strict(v1, lc.q));
    }
};

// encrypt, step 1: out[b][0] = NTT(e) - c1 * sk, out[b][1] = c1       sampling.cpp:66, rlwe.cpp:50
struct EncryptErrIO {
    const u64 *e, *c1; // [batch][L][N]: error coefficients, NTT-form mask
    const u64 *sk;     // [L][N]
    u64 *out;          // [batch][2][L][N]
    int L, logn;
    bool vec;
    HB_D int limb(int row) const { return row % L; }
    HB_D const u64 *src(int row) const { return e + ((size_t)row << logn); }
    HB_D u64 pre(int, int, u64 raw, const LimbConst &) const { return raw; }
    HB_D void store(int row, int i, u64 v, const LimbConst &lc) const {
        const int b = row / L, k = row - b * L;
        const u64 m = hb_ld_ro(c1 + ((size_t)row << logn) + i), s = __ldg(sk + ((size_t)k << logn) + i);
        u64 *o = out + (((size_t)b * 2 * L + k) << logn) + i;
        o[0] = sub_lazy(v, mul_hybrid_lazy(m, s, lc), lc.q2);
        o[(size_t)L << logn] = m;
    }
    HB_D void store2(int row, int i, u64 v0, u64 v1, const LimbConst &lc) const {
        const int b = row / L, k = row - b * L;
        const ulonglong2 m = hb_ld_ro2(c1 + ((size_t)row << logn) + i);
        const ulonglong2 s = __ldg(reinterpret_cast<const ulonglong2 *>(sk + ((size_t)k << logn) + i));
        u64 *o = out + (((size_t)b * 2 * L + k) << logn) + i;
        *reinterpret_cast<ulonglong2 *>(o) =
            make_ulonglong2(sub_lazy(v0, mul_hybrid_lazy(m.x, s.x, lc), lc.q2), sub_lazy(v1, mul_hybrid_lazy(m.y, s.y, lc), lc.q2));
        *reinterpret_cast<ulonglong2 *>(o + ((size_t)L << logn)) = m;
    }
};

// encrypt, step 2: out[b][0] += NTT(pt)                                  rlwe.cpp:54-58
struct EncryptAddIO {
    const u64 *pt; // [batch][L][N] coefficients
    u64 *out;
    int L, logn;
    bool vec;
    HB_D int limb(int row) const { return row % L; }
    HB_D const u64 *src(int row) const { return pt + ((size_t)row << logn); }
    HB_D u64 pre(int, int, u64 raw, const LimbConst &) const { return raw; }
    HB_D u64 *slot(int row, int i) const {
        const int b = row / L, k = row - b * L;
        return out + (((size_t)b * 2 * L + k) << logn) + i;
    }
    HB_D void store(int row, int i, u64 v, const LimbConst &lc) const {
        u64 *o = slot(row, i);
        *o = add_lazy(*o, v, lc.q2);
    }
    HB_D void store2(int row, int i, u64 v0, u64 v1, const LimbConst &lc) const {
        ulonglong2 *o = reinterpret_cast<ulonglong2 *>(slot(row, i));
        const ulonglong2 c = *o;
        *o = make_ulonglong2(add_lazy(c.x, v0, lc.q2), add_lazy(c.y, v1, lc.q2));
    }
};

int op_rlwe_decrypt_core(Context &c, unsigned logn, const u64 *moduli, size_t L, const u64 *ct, const u64 *sk, u64 *pt,
                         size_t batch) {
    if (!moduli || !ct || !sk || !pt) return c.fail(1, "null operand");
    if (batch == 0) return 0;
    for (size_t k = 0; k < L; k++)
        if (!(moduli[k] & 1)) return c.fail(1, "Montgomery multiplication needs odd moduli");
    int err = 0;
    const LimbConst *limbs = c.get_chain(logn, moduli, L, &err);
    if (!limbs) return err;
    DecryptIO io{ct, sk, pt, (int)L, (int)logn, aligned16(ct) && aligned16(pt)};
    cudaError_t e = launch_ntt<false>(c.env(), logn, io, limbs, (int)(batch * L));
    return e == cudaSuccess ? 0 : c.cuda_fail(e, "decrypt_core: intt launch");
}

int op_rlwe_encrypt_core(Context &c, unsigned logn, const u64 *moduli, size_t L, const u64 *pt, const u64 *sk, const u64 *c1,
                         const u64 *e, u64 *out, size_t batch) {
    if (!moduli || !pt || !sk || !c1 || !e || !out) return c.fail(1, "null operand");
    if (batch == 0) return 0;
    for (size_t k = 0; k < L; k++)
        if (!(moduli[k] & 1)) return c.fail(1, "Montgomery multiplication needs odd moduli");
    int err = 0;
    const LimbConst *limbs = c.get_chain(logn, moduli, L, &err);
    if (!limbs) return err;
    EncryptErrIO io1{e, c1, sk, out, (int)L, (int)logn, aligned16(e) && aligned16(c1) && aligned16(sk) && aligned16(out)};
    cudaError_t rc = launch_ntt<true>(c.env(), logn, io1, limbs, (int)(batch * L));
    if (rc != cudaSuccess) return c.cuda_fail(rc, "encrypt_core: error ntt launch");
    EncryptAddIO io2{pt, out, (int)L, (int)logn, aligned16(pt) && aligned16(out)};
    rc = launch_ntt<true>(c.env(), logn, io2, limbs, (int)(batch * L));
    return rc == cudaSuccess ? 0 : c.cuda_fail(rc, "encrypt_core: plaintext ntt launch");
}

// ------------------------------------------------------------------------------------------
// RNS base transform and key-switch key generation
//   rns_base_transform            src/fhe/common/rns_transform.cpp:11-126
//   RlweKsk::RlweKsk              src/fhe/primitives/keys.cpp:8-36 (+ rgsw.cpp:11-55, rlwe.cpp:34-51)
// ------------------------------------------------------------------------------------------
// one modulus -> many: out[b][k][i] from in[b][i]                       rns_transform.cpp:11-37, :116
// blocks tile one (b, k) row, so the modulus — and the 64-bit division behind `modulus_multiple` — is per block
// The centring constant (q_old / q_k + 1) * q_k - q_old of rns_transform.cpp:20-24 depends only on the modulus pair:
// the host computes it per new modulus and passes up to 32 of them by value (no 64-bit division in the kernel).
struct LiftTable {
    u64 v[32];
};
template <bool VEC>
HB_GLOBAL(256, 1)
base_from_single_kernel(const u64 *__restrict__ in, u64 *__restrict__ out, const LimbConst *__restrict__ limbs, u64 q_old, int Lnew,
                        size_t n, unsigned blocks_per_row, const LiftTable lifts) {
    hb_pdl_wait();
    const size_t bk = blockIdx.x / blocks_per_row;
    const int k = (int)(bk % Lnew);
    const size_t b = bk / Lnew;
    const LimbConst lc = limbs[k];
    const u64 lift = k < 32 ? lifts.v[k] : (q_old / lc.q + 1) * lc.q - q_old; // uniform across the block
    const u64 half = q_old / 2;
    const bool reduce = lc.q < q_old;
    auto f = [&](u64 raw) {
        u64 x = reduce_strict(raw, q_old);
        if (x >= half) x += lift;
        if (reduce) x = barrett_lazy(x, lc);
        return x;
    };
    constexpr int W = VEC ? 2 : 1;
    const size_t i0 = ((size_t)(blockIdx.x % blocks_per_row) * 1024 + threadIdx.x) * W;
#pragma unroll
    for (int u = 0; u < 4; u++) {
        const size_t i = i0 + (size_t)u * 256 * W;
        if (i < n) {
            if constexpr (VEC) {
                const ulonglong2 x = hb_ld_ro2(in + b * n + i);
                *reinterpret_cast<ulonglong2 *>(out + bk * n + i) = make_ulonglong2(f(x.x), f(x.y));
            } else {
                out[bk * n + i] = f(in[b * n + i]);
            }
        }
    }
}

static const char *const kNotSmallMessage =
    "under development: CRT composition of large coefficients (rns_transform.cpp:86-105) is not built";
// many -> one modulus, small-coefficient path; *not_small is set when some coefficient is not the same
// small signed value under every old modulus (the reference then composes big integers)   :39-84, :116
HB_GLOBAL(256, 1)
base_to_single_kernel(const u64 *__restrict__ in, u64 *__restrict__ out, const LimbConst *__restrict__ old_limbs, int L,
                      LimbConst new_lc, size_t n, size_t total, int *not_small) {
    hb_pdl_wait();
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x; // (b, i)
    if (gid >= total) return;
    const size_t i = gid % n, b = gid / n;
    const u64 q0 = old_limbs[0].q, half = q0 / 2;
    const u64 x0 = reduce_strict(in[(b * L) * n + i], q0);
    bool small = true;
    for (int k = 1; k < L; k++) {
        const u64 qk = old_limbs[k].q, xk = reduce_strict(in[(b * L + k) * n + i], qk);
        small &= (x0 < half) ? (xk == x0) : (qk - xk == q0 - x0);
    }
    if (!small) *not_small = 1;
    u64 r = (x0 < half) ? x0 : (q0 / new_lc.q + 1) * new_lc.q - q0 + x0;
    out[gid] = reduce_strict(barrett_lazy(r, new_lc), new_lc.q);
}

int op_base_from_single(Context &c, u64 q_old, const u64 *new_moduli, size_t Lnew, const u64 *in, u64 *out, size_t n, size_t batch) {
    if (!new_moduli || !in || !out) return c.fail(1, "null operand");
    if (q_old < 2 || Lnew == 0) return c.fail(1, "bad moduli");
    const size_t total = batch * Lnew * n;
    if (total == 0) return 0;
    int err = 0;
    const LimbConst *limbs = c.get_chain(0, new_moduli, Lnew, &err);
    if (!limbs) return err;
    LiftTable lifts{};
    for (size_t k = 0; k < Lnew && k < 32; k++) lifts.v[k] = (q_old / new_moduli[k] + 1) * new_moduli[k] - q_old;
    const bool vec = n % 2 == 0 && aligned16(in) && aligned16(out);
    const size_t per_block = vec ? 2048 : 1024;
    const size_t bpr = (n + per_block - 1) / per_block, blocks = batch * Lnew * bpr;
    if (blocks > 0x7fffffffull) return c.fail(1, "operand too large for one launch");
    if (vec) {
        HB_LAUNCH(base_from_single_kernel<true>, (unsigned)blocks, 256, 0, c.stream, 0, in, out, limbs, q_old, (int)Lnew, n, (unsigned)bpr, lifts);
    } else {
        HB_LAUNCH(base_from_single_kernel<false>, (unsigned)blocks, 256, 0, c.stream, 0, in, out, limbs, q_old, (int)Lnew, n, (unsigned)bpr, lifts);
    }
    c.stats.launches++;
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : c.cuda_fail(e, "base transform launch");
}

// `defer_verdict`: the not-small flag is copied to pinned host memory without synchronising; the next
// hehub_b200_ctx_synchronize reports it (stream-async callers: key generation).
int op_base_to_single(Context &c, const u64 *old_moduli, size_t L, u64 new_modulus, const u64 *in, u64 *out, size_t n, size_t batch,
                      bool defer_verdict) {
    if (!old_moduli || !in || !out) return c.fail(1, "null operand");
    if (L == 0 || new_modulus < 2) return c.fail(1, "bad moduli");
    // a single-component input is the one -> many transform with one target (rns_transform.cpp:118-121): lazy Barrett only when
    // the new modulus is the smaller one, no strict reduction
    if (L == 1) return op_base_from_single(c, old_moduli[0], &new_modulus, 1, in, out, n, batch);
    const size_t total = batch * n;
    if (total == 0) return 0;
    int err = 0;
    const LimbConst *limbs = c.get_chain(0, old_moduli, L, &err);
    if (!limbs) return err;
    const ModTables *nt = c.get_tables(new_modulus, 0, &err);
    if (!nt) return err;
    int *flag = reinterpret_cast<int *>(c.get_scratch(6, 2, &err));
    if (!flag) return err;
    cudaError_t e = cudaMemsetAsync(flag, 0, sizeof(int), c.stream);
    if (e != cudaSuccess) return c.cuda_fail(e, "base transform: flag");
    HB_LAUNCH(base_to_single_kernel, (unsigned)((total + 255) / 256), 256, 0, c.stream, 0, in, out, limbs, (int)L, nt->lc, n, total, flag);
    c.stats.launches++;
    if (defer_verdict) {
        int *slot = c.deferred_flag_slot(&err);
        if (!slot) return err;
        e = cudaMemcpyAsync(slot, flag, sizeof(int), cudaMemcpyDeviceToHost, c.stream);
        return e == cudaSuccess ? 0 : c.cuda_fail(e, "base transform to single");
    }
    int host_flag = 0;
    e = cudaMemcpyAsync(&host_flag, flag, sizeof(int), cudaMemcpyDeviceToHost, c.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c.stream);
    if (e != cudaSuccess) return c.cuda_fail(e, "base transform to single");
    if (host_flag) return c.fail(3, kNotSmallMessage);
    return 0;
}

struct PmodTable {
    ulonglong2 v[32];
};
// one row of the key: (c0, c1) = ( (NTT(e) - mask * sk_ext + [k == p] sk_curr * (P mod q_p)) * R, mask * R ), R = 2^64 mod q_k
struct KskRowIO {
    const u64 *errors, *masks; // [L][L+1][N]
    const u64 *sk_ext;         // [L+1][N], NTT form (sk_orig_extended)
    const u64 *sk_curr;        // [L][N], NTT form
    const ulonglong2 *pmod;    // [L]: (P mod q_p, Harvey companion); null: the by-value table below (L <= 32)
    u64 *key;                  // [L][2][L+1][N]
    int L, logn;
    bool vec;
    PmodTable pm; // kernel parameter space: nothing to upload, nothing to wait for
    HB_D int limb(int row) const { return row % (L + 1); }
    HB_D const u64 *src(int row) const { return errors + ((size_t)row << logn); }
    HB_D u64 pre(int, int, u64 raw, const LimbConst &) const { return raw; }
    HB_D void store(int row, int i, u64 v, const LimbConst &lc) const {
        const int p = row / (L + 1), k = row - p * (L + 1);
        const u64 m = hb_ld_ro(masks + ((size_t)row << logn) + i);
        u64 c0 = sub_lazy(v, mul_hybrid_lazy(m, __ldg(sk_ext + ((size_t)k << logn) + i), lc), lc.q2); // rlwe.cpp:50
        // + pt_ntt * basis_p (rgsw.cpp:27): Harvey multiple of sk_curr at k == p, of anything by 0 elsewhere (= 0)
        u64 term = 0;
        if (k == p) {
            const ulonglong2 s = pmod ? __ldg(pmod + p) : pm.v[p];
            term = harvey_lazy(__ldg(sk_curr + ((size_t)k << logn) + i), s.x, s.y, lc.nq);
        }
        c0 = add_lazy(c0, term, lc.q2);
        u64 *o = key + ((size_t)((p * 2) * (L + 1) + k) << logn) + i;
        o[0] = harvey_lazy(c0, lc.r, lc.r_h, lc.nq);                      // rgsw.cpp:47-51
        o[(size_t)(L + 1) << logn] = harvey_lazy(m, lc.r, lc.r_h, lc.nq);
    }
    HB_D void store2(int row, int i, u64 v0, u64 v1, const LimbConst &lc) const {
        store(row, i, v0, lc);
        store(row, i + 1, v1, lc);
    }
};

int op_ksk_generate(Context &c, unsigned logn, const u64 *ext_moduli, size_t L, const u64 *sk_curr, const u64 *sk_orig,
                    const u64 *masks, const u64 *errors, u64 *key) {
    if (!ext_moduli || !sk_curr || !sk_orig || !masks || !errors || !key) return c.fail(1, "null operand");
    if (L == 0) return c.fail(1, "no RNS components");
    for (size_t k = 0; k <= L; k++)
        if (!(ext_moduli[k] & 1)) return c.fail(1, "Montgomery multiplication needs odd moduli");
    const size_t n = (size_t)1 << logn, L1 = L + 1;
    int err = 0;
    const LimbConst *limbs = c.get_chain(logn, ext_moduli, L1, &err);
    if (!limbs) return err;
    u64 *sk_ext = c.get_scratch(7, L1 * n + 2 * L, &err); // sk_orig_extended, then the (P mod q_p) pairs when L > 32
    if (!sk_ext) return err;
    cudaError_t e = cudaMemcpyAsync(sk_ext, sk_orig, L * n * 8, cudaMemcpyDeviceToDevice, c.stream);
    if (e != cudaSuccess) return c.cuda_fail(e, "ksk: copy");
    if (int rc = run_transform(c, false, logn, ext_moduli, L, sk_ext, 1, 0)) return rc;                             // keys.cpp:22
    if (int rc = op_base_to_single(c, ext_moduli, L, ext_moduli[L], sk_ext, sk_ext + L * n, n, 1, true)) return rc; // keys.cpp:23-25
    if (int rc = run_transform(c, true, logn, ext_moduli, L1, sk_ext, 1, 0)) return rc;                             // keys.cpp:26
    KskRowIO io{errors, masks, sk_ext, sk_curr, nullptr, key, (int)L, (int)logn, aligned16(errors), PmodTable{}};
    std::vector<u64> pm(2 * L);
    for (size_t p = 0; p < L; p++) { // keys.cpp:28-33, rns.cpp:162-164
        pm[2 * p] = ext_moduli[L] % ext_moduli[p];
        pm[2 * p + 1] = host_harvey_quotient(pm[2 * p], ext_moduli[p]);
        if (L <= 32) io.pm.v[p] = make_ulonglong2(pm[2 * p], pm[2 * p + 1]);
    }
    if (L > 32) { // more digits than the parameter table holds: upload (pm is a host temporary, so wait for the copy)
        u64 *pmod = sk_ext + L1 * n;
        e = cudaMemcpyAsync(pmod, pm.data(), pm.size() * 8, cudaMemcpyHostToDevice, c.stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c.stream);
        if (e != cudaSuccess) return c.cuda_fail(e, "ksk: constants");
        io.pmod = reinterpret_cast<const ulonglong2 *>(pmod);
    }
    e = launch_ntt<true>(c.env(), logn, io, limbs, (int)(L * L1));
    return e == cudaSuccess ? 0 : c.cuda_fail(e, "ksk: row launch");
}

// ------------------------------------------------------------------------------------------
// Galois permutations on NTT-form polynomials.  Slot j holds the evaluation at psi^(2*brev(j)+1);
// the automorphism X -> X^g moves root index e to e*g, so output slot j gathers from the slot whose
// root index is e_j * g^{-1} (mod 2N).  cycle: g = 3^step (permutation.cpp:42-58); involution:
// g = -1, i.e. slot j <- slot N-1-j (permutation.cpp:70-73).
// ------------------------------------------------------------------------------------------
HB_GLOBAL(256, 1)
galois_kernel(const u64 *__restrict__ in, u64 *__restrict__ out, int logn, unsigned ginv, size_t total) {
    hb_pdl_wait();
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= total) return;
    const unsigned n = 1u << logn, j = (unsigned)(gid & (n - 1));
    const size_t row = gid >> logn;
    out[(row << logn) + j] = in[(row << logn) + galois_from(j, ginv, logn)];
}

static unsigned galois_inverse_factor(unsigned logn, bool conj, size_t step) {
    const unsigned mask = (2u << logn) - 1;
    if (conj) return mask; // -1 is its own inverse
    unsigned f = 1;
    for (size_t i = 0; i < step; i++) f *= 3u; // modulo 2^32, consistent with any smaller 2-power
    f &= mask;
    unsigned inv = f; // Newton: inverse of an odd number modulo 2^32
    for (int i = 0; i < 5; i++) inv *= 2u - f * inv;
    return inv & mask;
}

#if defined(HB_PHASE_CLOCK) && !defined(HB_KERNEL_SIM)
} // namespace hb
// probe builds only: where the kernels of this translation unit log their phase clocks (null: nowhere)
extern "C" int hehub_b200_debug_set_phase_log(unsigned long long *p) { return (int)cudaMemcpyToSymbol(hb::hb_phase_ptr, &p, sizeof(p)); }
namespace hb {
#endif

int op_galois(Context &c, unsigned logn, size_t L, const u64 *in, u64 *out, bool conj, size_t step, size_t batch) {
    if (!in || !out) return c.fail(1, "null operand");
    if (in == out) return c.fail(1, "Galois permutation cannot run in place");
    const size_t total = (batch * L) << logn;
    if (total == 0) return 0;
    HB_LAUNCH(galois_kernel, (unsigned)((total + 255) / 256), 256, 0, c.stream, 0, in, out, (int)logn,
              galois_inverse_factor(logn, conj, step), total);
    c.stats.launches++;
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : c.cuda_fail(e, "galois launch");
}

// ckks::rotate / ckks::conjugate — ckks/arith.cpp:75-93: permute both polynomials, key-switch the permuted c1, drop P, add the
// permuted c0 to the first half only.  Neither permuted polynomial is materialised: the key switch's inverse transforms (and
// the diagonal term of its inner product) read c1 through the permutation, and the drop's epilogue reads c0 through it, so a
// rotation costs what a relinearisation costs.
int op_galois_keyswitch(Context &c, unsigned logn, const u64 *ext_moduli, size_t L, const u64 *ct, const u64 *key, bool conj,
                        size_t step, u64 *out, size_t batch) {
    if (!ext_moduli || !ct || !key || !out) return c.fail(1, "null operand");
    if (L == 0) return c.fail(1, "Empty RGSW ciphertext.");
    if (batch == 0) return 0;
    const size_t n = (size_t)1 << logn;
    const unsigned ginv = galois_inverse_factor(logn, conj, step);
    if (pair_path_wanted(c, logn, L, batch) && aligned16(ct) && aligned16(key) && aligned16(out) &&
        !ranges_overlap(ct, batch * 2 * L * n, out, batch * 2 * L * n)) {
        const LimbConst *limbs;
        const DropSet *ds;
        u64 *dec;
        if (int rc = pair_setup(c, logn, ext_moduli, L, 0, batch, &limbs, &ds, &dec)) return rc;
        const KsInPlain<true> in{ct + L * n, 2 * L * n, (int)logn, ginv};
        const KsAddPlain<true> add{ct, 2 * L * n, L * n, 1, (int)logn, ginv};
        return launch_pair_logn<KsInPlain<true>, false, KsAddPlain<true>>(c, logn, in, add, key, out, dec, nullptr, limbs, ds, L, batch);
    }
    const bool in_place = ct == out; // the epilogue gathers c0 while `out` is written: work from a copy of the wave then
    const size_t per_ct = ((in_place ? 2 * L : 0) + L + L * (L + 1) + 2 * (L + 1) + 2) * n;
    const size_t wave = wave_size(c, per_ct, batch, L * (L + 1));
    int err = 0;
    u64 *ebuf = c.get_scratch(3, wave * 2 * (L + 1) * n, &err);
    if (!ebuf) return err;
    u64 *copy = in_place ? c.get_scratch(5, wave * 2 * L * n, &err) : nullptr;
    if (in_place && !copy) return err;
    for (size_t b0 = 0; b0 < batch; b0 += wave) {
        const size_t nb = (batch - b0 < wave) ? batch - b0 : wave;
        const u64 *ct_w = ct + b0 * 2 * L * n;
        if (in_place) {
            cudaError_t e = cudaMemcpyAsync(copy, ct_w, nb * 2 * L * n * 8, cudaMemcpyDeviceToDevice, c.stream);
            if (e != cudaSuccess) return c.cuda_fail(e, "rotate: copy");
            ct_w = copy;
        }
        FusedDrop fd_store, *fd = fused_drop_wanted(c, logn, ext_moduli, L, 0, nb == batch ? batch : 0, ginv, &fd_store);
        if (int rc = ext_prod_impl(c, logn, ext_moduli, L, ct_w + L * n, 2 * L * n, key, ebuf, nb, ginv, fd)) return rc;
        if (int rc = drop_last_impl(c, logn, ext_moduli, L + 1, 0, ebuf, out + b0 * 2 * L * n, nb, ct_w, 2 * L * n, L * n, 1, ginv, fd && fd->done ? fd->z : nullptr))
            return rc;
    }
    return 0;
}

} // namespace hb
