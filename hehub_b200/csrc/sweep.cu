// sweep.cu — the batched-ciphertext sweep driver (BASELINE config 5): ckks::mult + relinearize over a
// large set of independent synthetic ciphertext pairs, cut into contiguous per-rank ranges.
//
// The reference has no counterpart (it is a single-process library whose callers loop over ciphertexts,
// src/fhe/ckks/ckks.h:270-274); the driver is the multi-GPU harness north_star asks for.  Every ciphertext is an
// independent unit, so the data path has NO collective: each rank generates its inputs on its own device from
// (seed, global row index), runs hehub_b200_ckks_mult_relin in waves and reduces every result to one 64-bit
// checksum with a kernel of this library.  Collectives appear only at the two ends:
//   * broadcast of the key-switch key from rank 0 (78 MiB at C5), and
//   * all-gather of the per-ciphertext checksums so rank 0 can verify the whole sweep.
// They go through a small provider interface (hehub_b200_collectives): the product provider is NCCL over NVLink
// (csrc/nccl_provider.cpp -> libhehub_b200_nccl.so); the CPU test-suite plugs in gloo, where the "device" is the CTA
// emulator.  No PyTorch anywhere in this file.
#include <chrono>
#include <vector>

#include "../../include/hehub_b200.h"
#include "internal.h"

using namespace hb;

namespace hb {

constexpr u64 kCheckMult = 0x9E3779B97F4A7C15ull; // odd: position weights w_j = (2j + 1) * kCheckMult mod 2^64

// sums[ct] += sum_j words[ct][j] * (2j + 1) * kCheckMult (mod 2^64).  Wrapping addition is associative and
// commutative, so the order in which the blocks of one ciphertext arrive does not matter.
HB_GLOBAL(256, 1)
ct_checksum_kernel(const u64 *__restrict__ words, size_t words_per_ct, unsigned blocks_per_ct, unsigned long long *__restrict__ sums) {
    hb_pdl_wait();
    const size_t ct = blockIdx.x / blocks_per_ct;
    const unsigned chunk = blockIdx.x % blocks_per_ct;
    const u64 *w = words + ct * words_per_ct;
    u64 acc = 0;
    for (size_t j = (size_t)chunk * 256 + threadIdx.x; j < words_per_ct; j += (size_t)blocks_per_ct * 256)
        acc += w[j] * ((2 * (u64)j + 1) * kCheckMult);
#if defined(HB_KERNEL_SIM)
    __atomic_fetch_add(sums + ct, acc, __ATOMIC_RELAXED);
#else
#pragma unroll
    for (int off = 16; off; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if ((threadIdx.x & 31) == 0) atomicAdd(sums + ct, (unsigned long long)acc);
#endif
}

int op_ct_checksums(Context &c, const u64 *words, size_t words_per_ct, size_t cts, u64 *sums) {
    if (!words || !sums) return c.fail(1, "null operand");
    if (cts == 0 || words_per_ct == 0) return 0;
    cudaError_t e = cudaMemsetAsync(sums, 0, cts * sizeof(u64), c.stream);
    if (e != cudaSuccess) return c.cuda_fail(e, "checksums: clear");
    size_t bpc = (words_per_ct + 256 * 16 - 1) / (256 * 16); // ~16 words per thread
    if (bpc < 1) bpc = 1;
    if (bpc > 1024) bpc = 1024;
    if (cts * bpc > 0x7fffffffull) return c.fail(1, "operand too large for one launch");
    HB_LAUNCH(ct_checksum_kernel, (unsigned)(cts * bpc), 256, 0, c.stream, 0, words, words_per_ct, (unsigned)bpc,
              reinterpret_cast<unsigned long long *>(sums));
    c.stats.launches++;
    e = cudaGetLastError();
    return e == cudaSuccess ? 0 : c.cuda_fail(e, "checksum launch");
}

} // namespace hb

struct hehub_b200_sweep {
    hehub_b200_ctx *ctx = nullptr;
    unsigned logn = 0;
    size_t L = 0, n = 0, ct_words = 0;
    std::vector<u64> ext; // q_0 .. q_{L-1}, P
    u64 seed = 0;
    int rank = 0, world = 1;
    hehub_b200_collectives coll{};
    bool has_coll = false;
    u64 *key = nullptr; // device, [L][2][L+1][N]
    bool key_ready = false;
};

extern "C" {

void hehub_b200_shard_range(size_t total, int world, int rank, size_t *first, size_t *count) {
    // contiguous balanced partition: the first (total % world) ranks take one extra unit
    const size_t w = (size_t)(world > 0 ? world : 1), r = (size_t)(rank > 0 ? rank : 0);
    const size_t base = total / w, extra = total % w;
    if (first) *first = r * base + (r < extra ? r : extra);
    if (count) *count = base + (r < extra ? 1 : 0);
}

int hehub_b200_ct_checksums(hehub_b200_ctx *ctx, const uint64_t *words, size_t words_per_ct, size_t cts, uint64_t *sums_dev) {
    CTX_GUARD(ctx);
    return op_ct_checksums(c, (const u64 *)words, words_per_ct, cts, (u64 *)sums_dev);
}

int hehub_b200_sweep_create(hehub_b200_sweep **out, hehub_b200_ctx *ctx, unsigned logn, const uint64_t *ext_moduli, size_t L,
                            uint64_t seed, int rank, int world, const hehub_b200_collectives *coll) {
    if (!out) return HEHUB_B200_ERR_INVALID;
    *out = nullptr;
    CTX_GUARD(ctx);
    if (int rc = check_ring(c, logn, L + 1, 1)) return rc;
    if (!ext_moduli) return c.fail(HEHUB_B200_ERR_INVALID, "null moduli");
    if (world < 1 || rank < 0 || rank >= world) return c.fail(HEHUB_B200_ERR_INVALID, "bad rank / world size");
    if (world > 1 && (!coll || !coll->broadcast || !coll->allgather))
        return c.fail(HEHUB_B200_ERR_INVALID, "a sweep over several ranks needs a collectives provider");
    hehub_b200_sweep *s = new (std::nothrow) hehub_b200_sweep();
    if (!s) return HEHUB_B200_ERR_NOMEM;
    s->ctx = ctx;
    s->logn = logn;
    s->L = L;
    s->n = (size_t)1 << logn;
    s->ct_words = 2 * L * s->n;
    s->ext.assign(ext_moduli, ext_moduli + L + 1);
    s->seed = seed;
    s->rank = rank;
    s->world = world;
    if (coll) {
        s->coll = *coll;
        s->has_coll = true;
    }
    int rc = hehub_b200_slab_alloc(ctx, L * 2 * (L + 1) * s->n, reinterpret_cast<uint64_t **>(&s->key));
    if (rc) {
        delete s;
        return rc;
    }
    *out = s;
    return HEHUB_B200_OK;
}

int hehub_b200_sweep_destroy(hehub_b200_sweep *s) {
    if (!s) return HEHUB_B200_OK;
    if (s->key) hehub_b200_slab_free(s->ctx, reinterpret_cast<uint64_t *>(s->key));
    delete s;
    return HEHUB_B200_OK;
}

const uint64_t *hehub_b200_sweep_key(const hehub_b200_sweep *s) { return s ? reinterpret_cast<const uint64_t *>(s->key) : nullptr; }

// rank 0 fills the key ([L][2][L+1][N], limb L = special modulus) from the seed; everyone else receives it
int hehub_b200_sweep_make_key(hehub_b200_sweep *s) {
    if (!s) return HEHUB_B200_ERR_INVALID;
    CTX_GUARD(s->ctx);
    const size_t words = s->L * 2 * (s->L + 1) * s->n;
    if (s->rank == 0 || s->world == 1) {
        if (int rc = hehub_b200_lcg_fill(s->ctx, s->n, reinterpret_cast<const uint64_t *>(s->ext.data()), s->L + 1, reinterpret_cast<uint64_t *>(s->key), s->L * 2 * (s->L + 1),
                                         s->seed + 1000, 1))
            return rc;
    } else {
        cudaError_t e = cudaMemsetAsync(s->key, 0, words * 8, c.stream);
        if (e != cudaSuccess) return c.cuda_fail(e, "sweep: clear key");
    }
    if (s->world > 1) {
        if (s->coll.broadcast(s->coll.self, s->key, words * 8, 0, c.stream)) return c.fail(HEHUB_B200_ERR_CUDA, "sweep: key broadcast failed");
    }
    cudaError_t e = cudaStreamSynchronize(c.stream);
    if (e != cudaSuccess) return c.cuda_fail(e, "sweep: key");
    s->key_ready = true;
    return HEHUB_B200_OK;
}

// row seeds: ciphertext i of operand a (0 / 1) has rows [i * 2L, (i + 1) * 2L) of stream a, so the words of a pair do not
// depend on how the batch is sharded
int hehub_b200_sweep_fill_inputs(hehub_b200_sweep *s, uint64_t *ct1, uint64_t *ct2, size_t first_ct, size_t count) {
    if (!s) return HEHUB_B200_ERR_INVALID;
    const size_t rows = count * 2 * s->L;
    for (int operand = 0; operand < 2; operand++) {
        const u64 seed0 = s->seed + (u64)operand * 0x5851F42D4C957F2Dull + (u64)first_ct * 2 * s->L;
        if (int rc = hehub_b200_lcg_fill(s->ctx, s->n, reinterpret_cast<const uint64_t *>(s->ext.data()), s->L, operand ? ct2 : ct1, rows, seed0, 1)) return rc;
    }
    return HEHUB_B200_OK;
}

int hehub_b200_sweep_run(hehub_b200_sweep *s, size_t total, size_t wave, uint64_t *my_checksums_host, uint64_t *all_checksums_host,
                         size_t *first_out, size_t *count_out, double *op_seconds) {
    if (!s) return HEHUB_B200_ERR_INVALID;
    CTX_GUARD(s->ctx);
    if (!s->key_ready)
        if (int rc = hehub_b200_sweep_make_key(s)) return rc;
    size_t first = 0, count = 0;
    hehub_b200_shard_range(total, s->world, s->rank, &first, &count);
    if (first_out) *first_out = first;
    if (count_out) *count_out = count;
    if (wave < 1) wave = 1;
    if (wave > count && count) wave = count;
    size_t cap = 0; // the largest share: every rank contributes that many slots to the gather
    hehub_b200_shard_range(total, s->world, 0, nullptr, &cap);
    int err = 0;
    // workspaces: two input waves, one result wave, the checksums of this rank (padded) and of everyone
    u64 *ct1 = c.get_scratch(8, wave * s->ct_words, &err);
    if (!ct1) return err;
    u64 *ct2 = c.get_scratch(9, wave * s->ct_words, &err);
    if (!ct2) return err;
    u64 *res = c.get_scratch(10, wave * s->ct_words, &err);
    if (!res) return err;
    u64 *sums = c.get_scratch(11, (cap ? cap : 1) * (size_t)(s->world + 1), &err);
    if (!sums) return err;
    cudaError_t e = cudaMemsetAsync(sums, 0, (cap ? cap : 1) * 8, c.stream);
    if (e != cudaSuccess) return c.cuda_fail(e, "sweep: clear checksums");
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    double seconds = 0;
#if !defined(HB_KERNEL_SIM)
    if (op_seconds) {
        if (cudaEventCreate(&ev0) != cudaSuccess || cudaEventCreate(&ev1) != cudaSuccess) return c.fail(HEHUB_B200_ERR_CUDA, "sweep: events");
    }
#endif
    for (size_t done = 0; done < count; done += wave) {
        const size_t nb = (count - done < wave) ? count - done : wave;
        if (int rc = hehub_b200_sweep_fill_inputs(s, (uint64_t *)ct1, (uint64_t *)ct2, first + done, nb)) return rc;
#if defined(HB_KERNEL_SIM)
        const auto t0 = std::chrono::steady_clock::now();
#else
        if (op_seconds) cudaEventRecord(ev0, c.stream);
#endif
        if (int rc = op_mult_relin(c, s->logn, s->ext.data(), s->L, 0, ct1, ct2, s->key, res, nb)) return rc;
#if defined(HB_KERNEL_SIM)
        seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
#else
        if (op_seconds) {
            cudaEventRecord(ev1, c.stream);
            if ((e = cudaEventSynchronize(ev1)) != cudaSuccess) return c.cuda_fail(e, "sweep: wave");
            float ms = 0;
            cudaEventElapsedTime(&ms, ev0, ev1);
            seconds += ms * 1e-3;
        }
#endif
        if (int rc = op_ct_checksums(c, res, s->ct_words, nb, sums + done)) return rc;
    }
#if !defined(HB_KERNEL_SIM)
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
#endif
    if (op_seconds) *op_seconds = seconds;
    if (my_checksums_host && count) {
        e = cudaMemcpyAsync(my_checksums_host, sums, count * 8, cudaMemcpyDeviceToHost, c.stream);
        if (e != cudaSuccess) return c.cuda_fail(e, "sweep: checksums d2h");
    }
    if (s->world > 1) { // ragged gather: every rank's vector is padded to the largest share
        u64 *all = sums + (cap ? cap : 1);
        if (s->coll.allgather(s->coll.self, sums, all, (cap ? cap : 1) * 8, c.stream)) return c.fail(HEHUB_B200_ERR_CUDA, "sweep: all-gather failed");
        if (s->rank == 0 && all_checksums_host) {
            for (int r = 0; r < s->world; r++) {
                size_t f = 0, n_r = 0;
                hehub_b200_shard_range(total, s->world, r, &f, &n_r);
                if (!n_r) continue;
                e = cudaMemcpyAsync(all_checksums_host + f, all + (size_t)r * (cap ? cap : 1), n_r * 8, cudaMemcpyDeviceToHost, c.stream);
                if (e != cudaSuccess) return c.cuda_fail(e, "sweep: gathered checksums d2h");
            }
        }
    } else if (all_checksums_host && count) {
        e = cudaMemcpyAsync(all_checksums_host, sums, count * 8, cudaMemcpyDeviceToHost, c.stream);
        if (e != cudaSuccess) return c.cuda_fail(e, "sweep: checksums d2h");
    }
    e = cudaStreamSynchronize(c.stream);
    return e == cudaSuccess ? HEHUB_B200_OK : c.cuda_fail(e, "sweep: synchronize");
}

} // extern "C"
