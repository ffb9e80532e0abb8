// api.cu — the extern "C" boundary (include/hehub_b200.h): context, slabs, transforms and
// coefficient-wise kernels.  Scheme-level ops live in ops.cu.
#include <cstdlib>
#include <cstring>
#include <new>

#include "../../include/hehub_b200.h"
#include "internal.h"

using namespace hb;

namespace hb {

// ------------------------------------------------------------------------------------------
// IO policy of the plain in-place transforms: rows are [batch][L][N], limb = row % L.
// ------------------------------------------------------------------------------------------
template <bool STRICT> // STRICT: reduce_strict in the store (intt_negacyclic_inplace, ntt.h:89-92)
struct RowsIO {
    u64 *x;
    int L;
    int logn;
    bool vec;
    HB_D int limb(int row) const { return row % L; }
    HB_D const u64 *src(int row) const { return x + ((size_t)row << logn); }
    HB_D u64 pre(int, int, u64 raw, const LimbConst &) const { return raw; }
    HB_D void store(int row, int i, u64 v, const LimbConst &lc) const {
#if defined(HB_ABL_NOSTORE) // ablation builds only: results are (almost) never written
        if (v != 0x0123456789abcdefull) return;
#endif
        x[((size_t)row << logn) + i] = STRICT ? reduce_strict(v, lc.q) : v;
    }
    HB_D void store2(int row, int i, u64 v0, u64 v1, const LimbConst &lc) const {
#if defined(HB_ABL_NOSTORE)
        if (v0 != 0x0123456789abcdefull) return;
#endif
        *reinterpret_cast<ulonglong2 *>(x + ((size_t)row << logn) + i) =
            STRICT ? make_ulonglong2(reduce_strict(v0, lc.q), reduce_strict(v1, lc.q)) : make_ulonglong2(v0, v1);
    }
};

// ------------------------------------------------------------------------------------------
// coefficient-wise kernels.  One thread handles two 128-bit vectors (4 words); blocks tile a
// row so the modulus lookup is per block.  HBM-bound: fully coalesced 128-bit accesses.
// ------------------------------------------------------------------------------------------
constexpr int kEwThreads = 256;
// 128-bit vectors per thread are a property of the operation (Op::kVecs), from an A/B sweep on the box
// (profiles/r2m_ew_vectors_ab.log): streaming adds want many small blocks, the multiply-heavy ones fewer and longer.

template <class Op>
HB_GLOBAL(kEwThreads, 1) ew_kernel(const Op op, const LimbConst *__restrict__ limbs, int L, size_t n,
                                                        unsigned blocks_per_row) {
    hb_pdl_wait();
    const size_t row = blockIdx.x / blocks_per_row;
    const unsigned chunk = blockIdx.x % blocks_per_row;
    const LimbConst lc = limbs[row % L];
    const size_t base = row * n;
    constexpr int kEwVecs = Op::kVecs;
    const size_t i0 = (size_t)chunk * (kEwThreads * 2 * kEwVecs) + threadIdx.x * 2;
    if (((n & 1) == 0) && op.aligned) {
#pragma unroll
        for (int u = 0; u < kEwVecs; u++) {
            const size_t i = i0 + (size_t)u * kEwThreads * 2;
            if (i < n) op.apply2(base + i, (int)(row % L), lc);
        }
    } else {
#pragma unroll
        for (int u = 0; u < kEwVecs; u++) {
            const size_t i = i0 + (size_t)u * kEwThreads * 2;
            if (i < n) op.apply1(base + i, (int)(row % L), lc);
            if (i + 1 < n) op.apply1(base + i + 1, (int)(row % L), lc);
        }
    }
}

static inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

struct OpMulHybrid {
    static constexpr int kVecs = 8;
    const u64 *a, *b;
    u64 *c;
    bool aligned;
    HB_D void apply1(size_t i, int, const LimbConst &lc) const { c[i] = mul_hybrid_lazy(a[i], b[i], lc); }
    HB_D void apply2(size_t i, int, const LimbConst &lc) const {
        const ulonglong2 x = *reinterpret_cast<const ulonglong2 *>(a + i), y = *reinterpret_cast<const ulonglong2 *>(b + i);
        *reinterpret_cast<ulonglong2 *>(c + i) = make_ulonglong2(mul_hybrid_lazy(x.x, y.x, lc), mul_hybrid_lazy(x.y, y.y, lc));
    }
};

template <int MODE> // 0 add, 1 sub
struct OpAddSub {
    static constexpr int kVecs = 1;
    u64 *x;
    const u64 *y;
    bool aligned;
    HB_D u64 f(u64 a, u64 b, const LimbConst &lc) const {
        return MODE == 0 ? add_lazy(a, b, lc.q2) : sub_lazy(a, b, lc.q2);
    }
    HB_D void apply1(size_t i, int, const LimbConst &lc) const { x[i] = f(x[i], y[i], lc); }
    HB_D void apply2(size_t i, int, const LimbConst &lc) const {
        const ulonglong2 a = *reinterpret_cast<const ulonglong2 *>(x + i), b = *reinterpret_cast<const ulonglong2 *>(y + i);
        *reinterpret_cast<ulonglong2 *>(x + i) = make_ulonglong2(f(a.x, b.x, lc), f(a.y, b.y, lc));
    }
};

template <int MODE> // 0 strict, 1 barrett lazy, 2 barrett strict
struct OpUnary {
    static constexpr int kVecs = MODE == 0 ? 4 : 2;
    u64 *x;
    bool aligned;
    HB_D u64 f(u64 a, const LimbConst &lc) const {
        if (MODE == 0) return reduce_strict(a, lc.q);
        u64 r = barrett_lazy(a, lc);
        return MODE == 1 ? r : reduce_strict(r, lc.q);
    }
    HB_D void apply1(size_t i, int, const LimbConst &lc) const { x[i] = f(x[i], lc); }
    HB_D void apply2(size_t i, int, const LimbConst &lc) const {
        const ulonglong2 a = *reinterpret_cast<const ulonglong2 *>(x + i);
        *reinterpret_cast<ulonglong2 *>(x + i) = make_ulonglong2(f(a.x, lc), f(a.y, lc));
    }
};

struct OpMulScalar {
    static constexpr int kVecs = 1;
    u64 *x;
    const ulonglong2 *scalars; // [L] (s mod q, harvey companion)
    bool aligned;
    HB_D void apply1(size_t i, int limb, const LimbConst &lc) const {
        const ulonglong2 s = __ldg(scalars + limb);
        x[i] = harvey_lazy(x[i], s.x, s.y, lc.nq);
    }
    HB_D void apply2(size_t i, int limb, const LimbConst &lc) const {
        const ulonglong2 s = __ldg(scalars + limb);
        const ulonglong2 a = *reinterpret_cast<const ulonglong2 *>(x + i);
        *reinterpret_cast<ulonglong2 *>(x + i) = make_ulonglong2(harvey_lazy(a.x, s.x, s.y, lc.nq), harvey_lazy(a.y, s.x, s.y, lc.nq));
    }
};

struct OpMontgomery {
    static constexpr int kVecs = 2;
    const u64 *in; // (lo, hi) pairs
    u64 *out;
    bool aligned;
    HB_D void apply1(size_t i, int, const LimbConst &lc) const {
        const ulonglong2 a = *reinterpret_cast<const ulonglong2 *>(in + 2 * i);
        out[i] = montgomery128(a.x, a.y, lc);
    }
    HB_D void apply2(size_t i, int limb, const LimbConst &lc) const {
        apply1(i, limb, lc);
        apply1(i + 1, limb, lc);
    }
};

template <class Op>
static int launch_ew(Context &c, const Op &op, const LimbConst *limbs, size_t L, size_t n, size_t rows) {
    if (rows == 0 || n == 0) return 0;
    constexpr size_t kEwWordsPerBlock = (size_t)kEwThreads * 2 * Op::kVecs;
    const size_t bpr = (n + kEwWordsPerBlock - 1) / kEwWordsPerBlock;
    const size_t blocks = rows * bpr;
    if (blocks > 0x7fffffffull) return c.fail(HEHUB_B200_ERR_INVALID, "operand too large for one launch");
    auto kern = ew_kernel<Op>;
    HB_LAUNCH(kern, (unsigned)blocks, kEwThreads, 0, c.stream, 0, op, limbs, (int)L, n, (unsigned)bpr);
    c.stats.launches++;
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : c.cuda_fail(e, "coefficient-wise kernel launch");
}

// SURVEY Appendix B generator, parallelised by jumping the LCG ahead per 64-word chunk.
HB_GLOBAL(128, 1) lcg_fill_kernel(u64 *x, const LimbConst *__restrict__ limbs, int L, size_t n, size_t rows, u64 seed0,
                                u64 seed_stride) {
    hb_pdl_wait();
    constexpr u64 A = 6364136223846793005ull, C = 1442695040888963407ull;
    constexpr int CH = 64;
    const size_t chunks_per_row = (n + CH - 1) / CH;
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= rows * chunks_per_row) return;
    const size_t row = gid / chunks_per_row, start = (gid % chunks_per_row) * CH;
    const u64 q = limbs[row % L].q;
    // affine map x -> a x + c composed `start` times
    u64 a = 1, cc = 0, pa = A, pc = C;
    for (size_t k = start; k; k >>= 1) {
        if (k & 1) {
            cc = cc * pa + pc;
            a = a * pa;
        }
        pc = pc * pa + pc;
        pa = pa * pa;
    }
    u64 s = a * (seed0 + row * seed_stride) + cc;
    const size_t end = (start + CH < n) ? start + CH : n;
    for (size_t i = start; i < end; i++) {
        s = s * A + C;
        x[row * n + i] = s % q;
    }
}

int check_ring(Context &c, unsigned logn, size_t L, size_t batch) {
    if (logn < 1 || logn > 16) return c.fail(HEHUB_B200_ERR_INVALID, "dimension should be a 2-power (2..65536)");
    if (logn > (unsigned)kFastLogMax) return c.fail(HEHUB_B200_ERR_UNSUPPORTED, "ring dimension above 2^15 is not supported");
    if (L == 0) return c.fail(HEHUB_B200_ERR_INVALID, "no RNS components");
    if ((batch * L) >> 30) return c.fail(HEHUB_B200_ERR_INVALID, "batch too large for one call");
    return 0;
}

} // namespace hb

namespace hb {
int run_transform(Context &c, bool forward, unsigned logn, const u64 *moduli, size_t L, u64 *x, size_t batch, int strict) {
    if (int rc = check_ring(c, logn, L, batch)) return rc;
    if (batch == 0) return HEHUB_B200_OK;
    if (!moduli || !x) return c.fail(HEHUB_B200_ERR_INVALID, "null operand");
    int err = 0;
    const LimbConst *limbs = c.get_chain(logn, moduli, L, &err);
    if (!limbs) return err;
    cudaError_t e;
    if (strict && !forward) {
        e = launch_ntt<false>(c.env(), logn, RowsIO<true>{x, (int)L, (int)logn, aligned16(x)}, limbs, (int)(batch * L));
    } else { // the forward transform has no strict variant in the reference
        const RowsIO<false> io{x, (int)L, (int)logn, aligned16(x)};
        e = forward ? launch_ntt<true>(c.env(), logn, io, limbs, (int)(batch * L)) : launch_ntt<false>(c.env(), logn, io, limbs, (int)(batch * L));
    }
    return e == cudaSuccess ? 0 : c.cuda_fail(e, forward ? "ntt launch" : "intt launch");
}
} // namespace hb

extern "C" {

const char *hehub_b200_version(void) { return "hehub_b200 0.1 (sm_100a)"; }

int hehub_b200_ctx_create(hehub_b200_ctx **out, int device, void *stream) {
    if (!out) return HEHUB_B200_ERR_INVALID;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0 || device < 0 || device >= count) return HEHUB_B200_ERR_CUDA;
    e = cudaSetDevice(device);
    if (e != cudaSuccess) return HEHUB_B200_ERR_CUDA;
    hehub_b200_ctx *ctx = new (std::nothrow) hehub_b200_ctx();
    if (!ctx) return HEHUB_B200_ERR_NOMEM;
    ctx->c.device = device;
    if (stream) {
        ctx->c.stream = static_cast<cudaStream_t>(stream);
    } else {
        e = cudaStreamCreateWithFlags(&ctx->c.stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) {
            delete ctx;
            return HEHUB_B200_ERR_CUDA;
        }
        ctx->c.owns_stream = true;
    }
    const char *fg = std::getenv("HEHUB_B200_FORCE_GENERIC");
    ctx->c.force_generic = fg && fg[0] == '1';
#if defined(HB_KERNEL_SIM)
    ctx->c.sm_count = 2; // small persistent grids so the emulator exercises the row loop
#else
    {
        int sms = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && sms > 0) ctx->c.sm_count = sms;
    }
#endif
    *out = ctx;
    return HEHUB_B200_OK;
}

int hehub_b200_ctx_destroy(hehub_b200_ctx *ctx) {
    if (!ctx) return HEHUB_B200_OK;
    cudaSetDevice(ctx->c.device);
    cudaStreamSynchronize(ctx->c.stream);
    delete ctx;
    return HEHUB_B200_OK;
}

int hehub_b200_ctx_set_stream(hehub_b200_ctx *ctx, void *stream) {
    CTX_GUARD(ctx);
    cudaStreamSynchronize(c.stream);
    if (c.owns_stream && c.stream) cudaStreamDestroy(c.stream);
    c.owns_stream = false;
    c.stream = static_cast<cudaStream_t>(stream);
    return HEHUB_B200_OK;
}

int hehub_b200_ctx_synchronize(hehub_b200_ctx *ctx) {
    CTX_GUARD(ctx);
    cudaError_t e = cudaStreamSynchronize(c.stream);
    if (e != cudaSuccess) return c.cuda_fail(e, "cudaStreamSynchronize");
    if (c.take_deferred()) // verdict of a stream-asynchronous rns_base_transform (key generation), see context.h
        return c.fail(HEHUB_B200_ERR_UNSUPPORTED,
                      "under development: CRT composition of large coefficients (rns_transform.cpp:86-105) is not built "
                      "(raised by an earlier asynchronous key generation)");
    return HEHUB_B200_OK;
}

const char *hehub_b200_last_error(const hehub_b200_ctx *ctx) { return ctx ? ctx->c.last_error.c_str() : "null context"; }

int hehub_b200_ctx_set_option(hehub_b200_ctx *ctx, const char *name, int64_t value) {
    CTX_GUARD(ctx);
    if (!name) return c.fail(HEHUB_B200_ERR_INVALID, "null option name");
    if (!std::strcmp(name, "force_generic")) {
        c.force_generic = value != 0;
    } else if (!std::strcmp(name, "latency_rows")) { // -1: default (half the SM count); 0: never use the latency plans
        c.latency_rows = (int)value;
    } else if (!std::strcmp(name, "latency2_rows")) { // -1: default (a tenth of the SM count); 0: never use the mode-2 plans
        c.latency2_rows = (int)value;
    } else if (!std::strcmp(name, "pair_path")) { // 0 / 1 / 2: never / automatic / always take the two-launch key switch (context.h)
        c.pair_path = (int)value;
    } else if (!std::strcmp(name, "fused_drop")) {
        c.fused_drop = value != 0;
    } else if (!std::strcmp(name, "pair_tpc")) {
        c.pair_tpc = (int)value;
    } else if (!std::strcmp(name, "pair_fill_pct")) {
        c.pair_fill_pct = (int)value;
    } else if (!std::strcmp(name, "single_launch")) { // 0: one ckks::mult pair per call runs as six launches (A/B)
        c.single_launch = value != 0;
    } else if (!std::strcmp(name, "host_chunk_kib")) {
        if (value < 1) return c.fail(HEHUB_B200_ERR_INVALID, "host_chunk_kib must be positive");
        c.host_chunk_bytes = (size_t)value << 10;
    } else if (!std::strcmp(name, "scratch_cap_mib")) {
        if (value < 1) return c.fail(HEHUB_B200_ERR_INVALID, "scratch_cap_mib must be positive");
        c.scratch_cap_bytes = (size_t)value << 20;
    } else {
        return c.fail(HEHUB_B200_ERR_INVALID, std::string("unknown option ") + name);
    }
    return HEHUB_B200_OK;
}

uint64_t hehub_b200_launch_count(const hehub_b200_ctx *ctx) { return ctx ? ctx->c.stats.launches : 0; }

int hehub_b200_tables_prepare(hehub_b200_ctx *ctx, unsigned logn, const uint64_t *moduli, size_t nmod) {
    CTX_GUARD(ctx);
    if (int rc = check_ring(c, logn, nmod ? nmod : 1, 1)) return rc;
    for (size_t k = 0; k < nmod; k++) {
        int err = 0;
        if (!c.get_tables(moduli[k], logn, &err)) return err;
    }
    return HEHUB_B200_OK;
}

// ---- slabs ---------------------------------------------------------------------------------
int hehub_b200_slab_alloc(hehub_b200_ctx *ctx, size_t n_words, uint64_t **out_dev) {
    CTX_GUARD(ctx);
    if (!out_dev || n_words == 0) return c.fail(HEHUB_B200_ERR_INVALID, "bad slab request");
    auto &fl = c.slab_free[n_words];
    u64 *p = nullptr;
    if (!fl.empty()) {
        p = fl.back();
        fl.pop_back();
    } else {
        cudaError_t e = cudaMalloc(&p, n_words * sizeof(u64));
        if (e != cudaSuccess) {
            cudaGetLastError();
            return c.fail(HEHUB_B200_ERR_NOMEM, std::string("cudaMalloc(slab): ") + cudaGetErrorString(e));
        }
    }
    c.slab_live[p] = n_words;
    *out_dev = reinterpret_cast<uint64_t *>(p);
    return HEHUB_B200_OK;
}

int hehub_b200_slab_free(hehub_b200_ctx *ctx, uint64_t *dev) {
    CTX_GUARD(ctx);
    if (!dev) return HEHUB_B200_OK;
    auto it = c.slab_live.find(reinterpret_cast<u64 *>(dev));
    if (it == c.slab_live.end()) return c.fail(HEHUB_B200_ERR_INVALID, "slab not owned by this context");
    // stream order makes immediate reuse safe: every user of the slab was enqueued before
    c.slab_free[it->second].push_back(it->first);
    c.slab_live.erase(it);
    return HEHUB_B200_OK;
}

int hehub_b200_slab_h2d(hehub_b200_ctx *ctx, uint64_t *dev, const uint64_t *host, size_t n_words) {
    CTX_GUARD(ctx);
    cudaError_t e = cudaMemcpyAsync(dev, host, n_words * 8, cudaMemcpyHostToDevice, c.stream);
    return e == cudaSuccess ? 0 : c.cuda_fail(e, "h2d");
}

int hehub_b200_slab_d2h(hehub_b200_ctx *ctx, uint64_t *host, const uint64_t *dev, size_t n_words) {
    CTX_GUARD(ctx);
    cudaError_t e = cudaMemcpyAsync(host, dev, n_words * 8, cudaMemcpyDeviceToHost, c.stream);
    return e == cudaSuccess ? 0 : c.cuda_fail(e, "d2h");
}

int hehub_b200_slab_d2d(hehub_b200_ctx *ctx, uint64_t *dst, const uint64_t *src, size_t n_words) {
    CTX_GUARD(ctx);
    cudaError_t e = cudaMemcpyAsync(dst, src, n_words * 8, cudaMemcpyDeviceToDevice, c.stream);
    return e == cudaSuccess ? 0 : c.cuda_fail(e, "d2d");
}

int hehub_b200_host_alloc(hehub_b200_ctx *ctx, size_t n_words, uint64_t **out_host) {
    CTX_GUARD(ctx);
    if (!out_host || n_words == 0) return c.fail(HEHUB_B200_ERR_INVALID, "bad host buffer request");
    void *p = nullptr;
    cudaError_t e = cudaMallocHost(&p, n_words * 8);
    if (e != cudaSuccess) return c.cuda_fail(e, "cudaMallocHost");
    *out_host = static_cast<uint64_t *>(p);
    return HEHUB_B200_OK;
}

int hehub_b200_host_free(hehub_b200_ctx *ctx, uint64_t *host) {
    CTX_GUARD(ctx);
    cudaError_t e = cudaFreeHost(host);
    return e == cudaSuccess ? 0 : c.cuda_fail(e, "cudaFreeHost");
}

// ---- transforms ----------------------------------------------------------------------------
static int run_plain_ntt(hehub_b200_ctx *ctx, bool forward, unsigned logn, const uint64_t *moduli, size_t L, uint64_t *x,
                         size_t batch, int strict) {
    CTX_GUARD(ctx);
    return run_transform(c, forward, logn, reinterpret_cast<const u64 *>(moduli), L, reinterpret_cast<u64 *>(x), batch, strict);
}

int hehub_b200_ntt_fwd_lazy(hehub_b200_ctx *ctx, unsigned logn, const uint64_t *moduli, size_t L, uint64_t *x, size_t batch) {
    return run_plain_ntt(ctx, true, logn, moduli, L, x, batch, 0);
}

int hehub_b200_intt_lazy(hehub_b200_ctx *ctx, unsigned logn, const uint64_t *moduli, size_t L, uint64_t *x, size_t batch,
                         int strict) {
    return run_plain_ntt(ctx, false, logn, moduli, L, x, batch, strict);
}

// ---- coefficient-wise ----------------------------------------------------------------------
#define EW_PROLOGUE()                                                                              \
    CTX_GUARD(ctx);                                                                                \
    if (!moduli || L == 0) return c.fail(HEHUB_B200_ERR_INVALID, "no moduli");                     \
    int err = 0;                                                                                   \
    const LimbConst *limbs = c.get_chain(0, reinterpret_cast<const u64 *>(moduli), L, &err);       \
    if (!limbs) return err;

int hehub_b200_mulmod_hybrid_lazy(hehub_b200_ctx *ctx, size_t n, const uint64_t *moduli, size_t L, const uint64_t *a,
                                  const uint64_t *b, uint64_t *cc, size_t batch) {
    EW_PROLOGUE();
    for (size_t k = 0; k < L; k++)
        if (!(moduli[k] & 1)) return c.fail(HEHUB_B200_ERR_INVALID, "Montgomery multiplication needs odd moduli");
    OpMulHybrid op{(const u64 *)a, (const u64 *)b, (u64 *)cc, aligned16(a) && aligned16(b) && aligned16(cc)};
    return launch_ew(c, op, limbs, L, n, batch * L);
}

int hehub_b200_add_lazy(hehub_b200_ctx *ctx, size_t n, const uint64_t *moduli, size_t L, uint64_t *x, const uint64_t *y,
                        size_t batch) {
    EW_PROLOGUE();
    OpAddSub<0> op{(u64 *)x, (const u64 *)y, aligned16(x) && aligned16(y)};
    return launch_ew(c, op, limbs, L, n, batch * L);
}

int hehub_b200_sub_lazy(hehub_b200_ctx *ctx, size_t n, const uint64_t *moduli, size_t L, uint64_t *x, const uint64_t *y,
                        size_t batch) {
    EW_PROLOGUE();
    OpAddSub<1> op{(u64 *)x, (const u64 *)y, aligned16(x) && aligned16(y)};
    return launch_ew(c, op, limbs, L, n, batch * L);
}

int hehub_b200_mul_scalar_lazy(hehub_b200_ctx *ctx, size_t n, const uint64_t *moduli, size_t L, uint64_t *x,
                               const uint64_t *scalars, size_t batch) {
    EW_PROLOGUE();
    if (!scalars) return c.fail(HEHUB_B200_ERR_INVALID, "null scalars");
    std::vector<u64> key;
    key.insert(key.end(), moduli, moduli + L);
    key.insert(key.end(), scalars, scalars + L);
    auto it = c.scalar_sets.find(key);
    if (it == c.scalar_sets.end()) {
        if (c.scalar_sets.size() > 4096) { // bounded cache; entries are tiny
            cudaStreamSynchronize(c.stream);
            for (auto &kv : c.scalar_sets) cudaFree(kv.second);
            c.scalar_sets.clear();
        }
        std::vector<u64> host(2 * L);
        for (size_t k = 0; k < L; k++) { // rns.cpp:162-164
            host[2 * k] = scalars[k] % moduli[k];
            host[2 * k + 1] = host_harvey_quotient(host[2 * k], moduli[k]);
        }
        u64 *dev = nullptr;
        cudaError_t e = cudaMalloc(&dev, host.size() * 8);
        if (e != cudaSuccess) return c.cuda_fail(e, "cudaMalloc(scalars)");
        e = cudaMemcpyAsync(dev, host.data(), host.size() * 8, cudaMemcpyHostToDevice, c.stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c.stream);
        if (e != cudaSuccess) return c.cuda_fail(e, "upload scalars");
        it = c.scalar_sets.emplace(std::move(key), dev).first;
    }
    OpMulScalar op{(u64 *)x, reinterpret_cast<const ulonglong2 *>(it->second), aligned16(x)};
    return launch_ew(c, op, limbs, L, n, batch * L);
}

int hehub_b200_reduce_strict(hehub_b200_ctx *ctx, size_t n, const uint64_t *moduli, size_t L, uint64_t *x, size_t batch) {
    EW_PROLOGUE();
    OpUnary<0> op{(u64 *)x, aligned16(x)};
    return launch_ew(c, op, limbs, L, n, batch * L);
}

int hehub_b200_barrett_lazy(hehub_b200_ctx *ctx, size_t n, const uint64_t *moduli, size_t L, uint64_t *x, size_t batch) {
    EW_PROLOGUE();
    OpUnary<1> op{(u64 *)x, aligned16(x)};
    return launch_ew(c, op, limbs, L, n, batch * L);
}

int hehub_b200_barrett(hehub_b200_ctx *ctx, size_t n, const uint64_t *moduli, size_t L, uint64_t *x, size_t batch) {
    EW_PROLOGUE();
    OpUnary<2> op{(u64 *)x, aligned16(x)};
    return launch_ew(c, op, limbs, L, n, batch * L);
}

int hehub_b200_montgomery128_lazy(hehub_b200_ctx *ctx, uint64_t q, size_t n, const uint64_t *in_lohi, uint64_t *out) {
    const uint64_t *moduli = &q;
    const size_t L = 1;
    EW_PROLOGUE();
    if (!(q & 1)) return c.fail(HEHUB_B200_ERR_INVALID, "Montgomery reduction needs an odd modulus");
    if (!aligned16(in_lohi)) return c.fail(HEHUB_B200_ERR_INVALID, "128-bit input must be 16-byte aligned");
    OpMontgomery op{(const u64 *)in_lohi, (u64 *)out, true};
    return launch_ew(c, op, limbs, 1, n, 1);
}

int hehub_b200_lcg_fill(hehub_b200_ctx *ctx, size_t n, const uint64_t *moduli, size_t L, uint64_t *x, size_t rows,
                        uint64_t seed0, uint64_t seed_stride) {
    EW_PROLOGUE();
    if (rows == 0 || n == 0) return 0;
    const size_t threads = rows * ((n + 63) / 64);
    HB_LAUNCH(lcg_fill_kernel, (unsigned)((threads + 127) / 128), 128, 0, c.stream, 0, (u64 *)x, limbs, (int)L, n, rows, seed0,
              seed_stride);
    c.stats.launches++;
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : c.cuda_fail(e, "lcg_fill launch");
}

// ---- scheme-level ops: argument checks here, work in ops.cu --------------------------------
int hehub_b200_galois_cycle(hehub_b200_ctx *ctx, unsigned logn, size_t L, const uint64_t *in, uint64_t *out, size_t step,
                            size_t batch) {
    CTX_GUARD(ctx);
    if (int rc = check_ring(c, logn, L, batch)) return rc;
    if (step >= ((size_t)1 << 17)) return c.fail(HEHUB_B200_ERR_INVALID, "rotation step out of table range");
    return op_galois(c, logn, L, (const u64 *)in, (u64 *)out, false, step, batch);
}

int hehub_b200_galois_involution(hehub_b200_ctx *ctx, unsigned logn, size_t L, const uint64_t *in, uint64_t *out,
                                 size_t batch) {
    CTX_GUARD(ctx);
    if (int rc = check_ring(c, logn, L, batch)) return rc;
    return op_galois(c, logn, L, (const u64 *)in, (u64 *)out, true, 0, batch);
}

int hehub_b200_ckks_tensor(hehub_b200_ctx *ctx, unsigned logn, const uint64_t *moduli, size_t L, const uint64_t *ct1,
                           const uint64_t *ct2, uint64_t *quad, size_t batch) {
    CTX_GUARD(ctx);
    if (int rc = check_ring(c, logn, L, batch)) return rc;
    return op_ckks_tensor(c, logn, (const u64 *)moduli, L, (const u64 *)ct1, (const u64 *)ct2, (u64 *)quad, batch);
}

int hehub_b200_ext_prod_montgomery(hehub_b200_ctx *ctx, unsigned logn, const uint64_t *ext_moduli, size_t L,
                                   const uint64_t *in, const uint64_t *key, uint64_t *out, size_t batch) {
    CTX_GUARD(ctx);
    if (int rc = check_ring(c, logn, L + 1, batch)) return rc;
    return op_ext_prod(c, logn, (const u64 *)ext_moduli, L, (const u64 *)in, L << logn, (const u64 *)key, (u64 *)out, batch);
}

int hehub_b200_ckks_rescale(hehub_b200_ctx *ctx, unsigned logn, const uint64_t *moduli, size_t L, const uint64_t *ct,
                            uint64_t *out, size_t batch) {
    CTX_GUARD(ctx);
    if (int rc = check_ring(c, logn, L, batch)) return rc;
    if (L < 2) return c.fail(HEHUB_B200_ERR_INVALID, "Unable to drop the only one prime."); // rescaling.cpp:27-29
    return op_drop_last(c, logn, (const u64 *)moduli, L, 0, (const u64 *)ct, (u64 *)out, batch, nullptr, 0, 0, 0);
}

int hehub_b200_bgv_mod_switch(hehub_b200_ctx *ctx, unsigned logn, const uint64_t *moduli, size_t L, uint64_t t,
                              const uint64_t *ct, uint64_t *out, size_t batch) {
    CTX_GUARD(ctx);
    if (int rc = check_ring(c, logn, L, batch)) return rc;
    if (L < 2) return c.fail(HEHUB_B200_ERR_INVALID, "Unable to drop the only one prime."); // mod_switch.cpp:26-28
    if (t == 0) return c.fail(HEHUB_B200_ERR_INVALID, "plain modulus must be positive");
    return op_drop_last(c, logn, (const u64 *)moduli, L, t, (const u64 *)ct, (u64 *)out, batch, nullptr, 0, 0, 0);
}

int hehub_b200_ckks_relinearize(hehub_b200_ctx *ctx, unsigned logn, const uint64_t *ext_moduli, size_t L,
                                const uint64_t *quad, const uint64_t *key, uint64_t *out, size_t batch) {
    CTX_GUARD(ctx);
    if (int rc = check_ring(c, logn, L + 1, batch)) return rc;
    return op_relinearize(c, logn, (const u64 *)ext_moduli, L, 0, (const u64 *)quad, (const u64 *)key, (u64 *)out, batch);
}

int hehub_b200_bgv_relinearize(hehub_b200_ctx *ctx, unsigned logn, const uint64_t *ext_moduli, size_t L, uint64_t t,
                               const uint64_t *quad, const uint64_t *key, uint64_t *out, size_t batch) {
    CTX_GUARD(ctx);
    if (int rc = check_ring(c, logn, L + 1, batch)) return rc;
    if (t == 0) return c.fail(HEHUB_B200_ERR_INVALID, "plain modulus must be positive");
    return op_relinearize(c, logn, (const u64 *)ext_moduli, L, t, (const u64 *)quad, (const u64 *)key, (u64 *)out, batch);
}

int hehub_b200_ckks_mult_relin(hehub_b200_ctx *ctx, unsigned logn, const uint64_t *ext_moduli, size_t L,
                               const uint64_t *ct1, const uint64_t *ct2, const uint64_t *key, uint64_t *out, size_t batch) {
    CTX_GUARD(ctx);
    if (int rc = check_ring(c, logn, L + 1, batch)) return rc;
    return op_mult_relin(c, logn, (const u64 *)ext_moduli, L, 0, (const u64 *)ct1, (const u64 *)ct2, (const u64 *)key,
                         (u64 *)out, batch);
}

int hehub_b200_bgv_mult_relin(hehub_b200_ctx *ctx, unsigned logn, const uint64_t *ext_moduli, size_t L, uint64_t t,
                              const uint64_t *ct1, const uint64_t *ct2, const uint64_t *key, uint64_t *out, size_t batch) {
    CTX_GUARD(ctx);
    if (int rc = check_ring(c, logn, L + 1, batch)) return rc;
    if (t == 0) return c.fail(HEHUB_B200_ERR_INVALID, "plain modulus must be positive");
    return op_mult_relin(c, logn, (const u64 *)ext_moduli, L, t, (const u64 *)ct1, (const u64 *)ct2, (const u64 *)key,
                         (u64 *)out, batch);
}

int hehub_b200_rlwe_decrypt_core(hehub_b200_ctx *ctx, unsigned logn, const uint64_t *moduli, size_t L, const uint64_t *ct,
                                 const uint64_t *sk, uint64_t *pt, size_t batch) {
    CTX_GUARD(ctx);
    if (int rc = check_ring(c, logn, L, batch)) return rc;
    return op_rlwe_decrypt_core(c, logn, (const u64 *)moduli, L, (const u64 *)ct, (const u64 *)sk, (u64 *)pt, batch);
}

int hehub_b200_rlwe_encrypt_core(hehub_b200_ctx *ctx, unsigned logn, const uint64_t *moduli, size_t L, const uint64_t *pt,
                                 const uint64_t *sk, const uint64_t *c1, const uint64_t *e, uint64_t *out, size_t batch) {
    CTX_GUARD(ctx);
    if (int rc = check_ring(c, logn, L, batch)) return rc;
    return op_rlwe_encrypt_core(c, logn, (const u64 *)moduli, L, (const u64 *)pt, (const u64 *)sk, (const u64 *)c1,
                                (const u64 *)e, (u64 *)out, batch);
}

int hehub_b200_rns_base_transform_from_single(hehub_b200_ctx *ctx, uint64_t q_old, const uint64_t *new_moduli, size_t Lnew,
                                              const uint64_t *in, uint64_t *out, size_t n, size_t batch) {
    CTX_GUARD(ctx);
    return op_base_from_single(c, q_old, (const u64 *)new_moduli, Lnew, (const u64 *)in, (u64 *)out, n, batch);
}

int hehub_b200_rns_base_transform_to_single(hehub_b200_ctx *ctx, const uint64_t *old_moduli, size_t L, uint64_t new_modulus,
                                            const uint64_t *in, uint64_t *out, size_t n, size_t batch) {
    CTX_GUARD(ctx);
    return op_base_to_single(c, (const u64 *)old_moduli, L, new_modulus, (const u64 *)in, (u64 *)out, n, batch);
}

int hehub_b200_ksk_generate(hehub_b200_ctx *ctx, unsigned logn, const uint64_t *ext_moduli, size_t L, const uint64_t *sk_curr,
                            const uint64_t *sk_orig, const uint64_t *masks, const uint64_t *errors, uint64_t *key) {
    CTX_GUARD(ctx);
    if (int rc = check_ring(c, logn, L + 1, 1)) return rc;
    return op_ksk_generate(c, logn, (const u64 *)ext_moduli, L, (const u64 *)sk_curr, (const u64 *)sk_orig, (const u64 *)masks,
                           (const u64 *)errors, (u64 *)key);
}

int hehub_b200_ckks_rotate(hehub_b200_ctx *ctx, unsigned logn, const uint64_t *ext_moduli, size_t L, const uint64_t *ct,
                           const uint64_t *key, size_t step, uint64_t *out, size_t batch) {
    CTX_GUARD(ctx);
    if (int rc = check_ring(c, logn, L + 1, batch)) return rc;
    if (step >= ((size_t)1 << 17)) return c.fail(HEHUB_B200_ERR_INVALID, "rotation step out of table range");
    return op_galois_keyswitch(c, logn, (const u64 *)ext_moduli, L, (const u64 *)ct, (const u64 *)key, false, step,
                               (u64 *)out, batch);
}

int hehub_b200_ckks_conjugate(hehub_b200_ctx *ctx, unsigned logn, const uint64_t *ext_moduli, size_t L, const uint64_t *ct,
                              const uint64_t *key, uint64_t *out, size_t batch) {
    CTX_GUARD(ctx);
    if (int rc = check_ring(c, logn, L + 1, batch)) return rc;
    return op_galois_keyswitch(c, logn, (const u64 *)ext_moduli, L, (const u64 *)ct, (const u64 *)key, true, 0, (u64 *)out,
                               batch);
}

} // extern "C"
