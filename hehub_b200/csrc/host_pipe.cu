// host_pipe.cu — host-buffer entry points: the call a reference-style caller makes when its
// RnsPolynomial / CkksCt words live in host memory (as they always do in the reference,
// src/fhe/common/rns.cpp:25-27).
//
// A batch is cut into chunks; chunk c is copied in on the copy-in stream, transformed on the
// context's compute stream and copied out on the copy-out stream, up to four chunks in flight, so
// PCIe traffic in both directions overlaps the kernels.  The calls return once the results are
// in host memory.  Host buffers should come from hehub_b200_host_alloc (pinned); pageable
// memory works but serialises the copies.
#include "../../include/hehub_b200.h"
#include <vector>

#include "internal.h"

using namespace hb;

namespace hb {

int Context::ensure_pipe() {
    if (pipe_ready) return 0;
    cudaError_t e = cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking);
    for (int i = 0; i < kPipeSlots && e == cudaSuccess; i++) {
        e = cudaEventCreateWithFlags(&ev_in[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev_k[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev_out[i], cudaEventDisableTiming);
    }
    if (e != cudaSuccess) return cuda_fail(e, "create copy streams");
    pipe_ready = true;
    return 0;
}

// Runs `work(slot_buffers..., first, count)` over [0, units) in chunks of `chunk` units.
// in_words / out_words: words per unit copied in (per input operand) and out.
template <class Work>
static int run_pipeline(Context &c, size_t units, size_t chunk, int n_in, const u64 *const *host_in, size_t in_words,
                        u64 *host_out, size_t out_words, int scratch_base, bool inplace, Work work) {
    if (int rc = c.ensure_pipe()) return rc;
    constexpr int S = Context::kPipeSlots;
    int err = 0;
    // one allocation per role, S slots each
    u64 *din[2] = {nullptr, nullptr};
    for (int k = 0; k < n_in; k++) {
        din[k] = c.get_scratch(scratch_base + k, (size_t)S * chunk * in_words, &err);
        if (!din[k]) return err;
    }
    // in place: the result is read back from the first input's staging slab
    u64 *dout = inplace ? din[0] : c.get_scratch(scratch_base + 2, (size_t)S * chunk * out_words, &err);
    if (!dout) return err;
    // everything queued earlier on the compute stream precedes the first kernel anyway; the copy
    // streams only touch the staging slabs
    // Chunk schedule: the first copy-in and the last copy-out overlap nothing, so the pipeline starts
    // and ends with small chunks (chunk/8, chunk/4, chunk/2) and runs full-size chunks in between.
    std::vector<size_t> sched;
    {
        std::vector<size_t> ramp;
        size_t used = 0;
        for (size_t cc = chunk / 8 ? chunk / 8 : 1; cc < chunk; cc *= 2) {
            ramp.push_back(cc);
            used += 2 * cc;
        }
        if (used + chunk <= units) {
            sched = ramp;
            for (size_t left = units - used; left > 0;) {
                const size_t cc = left < chunk ? left : chunk;
                sched.push_back(cc);
                left -= cc;
            }
            sched.insert(sched.end(), ramp.rbegin(), ramp.rend());
        } else {
            for (size_t left = units; left > 0;) {
                const size_t cc = left < chunk ? left : chunk;
                sched.push_back(cc);
                left -= cc;
            }
        }
    }
    size_t idx = 0, first = 0;
    for (; idx < sched.size(); first += sched[idx], idx++) {
        const size_t cnt = sched[idx];
        const int slot = (int)(idx % S);
        cudaError_t e = cudaSuccess;
        if (idx >= (size_t)S) e = cudaStreamWaitEvent(c.s_in, c.ev_out[slot], 0); // slot drained
        for (int k = 0; k < n_in && e == cudaSuccess; k++)
            e = cudaMemcpyAsync(din[k] + (size_t)slot * chunk * in_words, host_in[k] + first * in_words, cnt * in_words * 8,
                                cudaMemcpyHostToDevice, c.s_in);
        if (e == cudaSuccess) e = cudaEventRecord(c.ev_in[slot], c.s_in);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(c.stream, c.ev_in[slot], 0);
        if (e != cudaSuccess) return c.cuda_fail(e, "host pipeline: copy in");
        if (int rc = work(din[0] + (size_t)slot * chunk * in_words, n_in > 1 ? din[1] + (size_t)slot * chunk * in_words : nullptr,
                          dout + (size_t)slot * chunk * out_words, cnt))
            return rc;
        e = cudaEventRecord(c.ev_k[slot], c.stream);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(c.s_out, c.ev_k[slot], 0);
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(host_out + first * out_words, dout + (size_t)slot * chunk * out_words, cnt * out_words * 8,
                                cudaMemcpyDeviceToHost, c.s_out);
        if (e == cudaSuccess) e = cudaEventRecord(c.ev_out[slot], c.s_out);
        if (e != cudaSuccess) return c.cuda_fail(e, "host pipeline: copy out");
    }
    cudaError_t e = cudaStreamSynchronize(c.s_out);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c.stream);
    return e == cudaSuccess ? 0 : c.cuda_fail(e, "host pipeline: drain");
}

static size_t pick_chunk(size_t units, size_t words_per_unit, size_t target_bytes) {
    size_t chunk = target_bytes / (words_per_unit * 8);
    if (chunk < 1) chunk = 1;
    // at least kPipeSlots chunks when the batch allows, so the three stages overlap
    const size_t third = (units + Context::kPipeSlots - 1) / Context::kPipeSlots;
    if (chunk > third && third > 0) chunk = third;
    return chunk;
}

} // namespace hb

extern "C" {

int hehub_b200_ntt_host(hehub_b200_ctx *ctx, int forward, unsigned logn, const uint64_t *moduli, size_t L,
                        const uint64_t *host_in, uint64_t *host_out, size_t batch, int strict) {
    CTX_GUARD(ctx);
    if (int rc = check_ring(c, logn, L, batch)) return rc;
    if (batch == 0) return HEHUB_B200_OK;
    if (!moduli || !host_in || !host_out) return c.fail(HEHUB_B200_ERR_INVALID, "null operand");
    { // validate the moduli (and build tables) before any copy is queued
        int err = 0;
        if (!c.get_chain(logn, reinterpret_cast<const u64 *>(moduli), L, &err)) return err;
    }
    const size_t words = L << logn; // one unit = one polynomial of L limbs
    const size_t chunk = pick_chunk(batch, words, c.host_chunk_bytes);
    const u64 *ins[1] = {reinterpret_cast<const u64 *>(host_in)};
    return run_pipeline(c, batch, chunk, 1, ins, words, reinterpret_cast<u64 *>(host_out), words, 8, true,
                        [&](u64 *a, u64 *, u64 *, size_t cnt) -> int {
                            return run_transform(c, forward != 0, logn, reinterpret_cast<const u64 *>(moduli), L, a, cnt, strict);
                        });
}

int hehub_b200_ckks_mult_relin_host(hehub_b200_ctx *ctx, unsigned logn, const uint64_t *ext_moduli, size_t L,
                                    const uint64_t *host_ct1, const uint64_t *host_ct2, const uint64_t *dev_key,
                                    uint64_t *host_out, size_t batch) {
    CTX_GUARD(ctx);
    if (int rc = check_ring(c, logn, L + 1, batch)) return rc;
    if (batch == 0) return HEHUB_B200_OK;
    if (!ext_moduli || !host_ct1 || !host_ct2 || !dev_key || !host_out) return c.fail(HEHUB_B200_ERR_INVALID, "null operand");
    {
        int err = 0;
        if (!c.get_chain(logn, reinterpret_cast<const u64 *>(ext_moduli), L + 1, &err)) return err;
    }
    const size_t words = (2 * L) << logn; // one ciphertext
    const size_t chunk = pick_chunk(batch, words, c.host_chunk_bytes);
    const u64 *ins[2] = {reinterpret_cast<const u64 *>(host_ct1), reinterpret_cast<const u64 *>(host_ct2)};
    return run_pipeline(c, batch, chunk, 2, ins, words, reinterpret_cast<u64 *>(host_out), words, 8, false,
                        [&](u64 *a, u64 *b, u64 *out, size_t cnt) -> int {
                            return op_mult_relin(c, logn, reinterpret_cast<const u64 *>(ext_moduli), L, 0, a, b,
                                                 reinterpret_cast<const u64 *>(dev_key), out, cnt);
                        });
}

} // extern "C"
