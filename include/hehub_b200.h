/*
 * hehub_b200.h — C ABI of the B200-native RNS polynomial-arithmetic backend for HEhub.
 *
 * This is the drop-in boundary for the ciphertext-op hot path of primihub/hehub
 * (negacyclic NTT/INTT, lazy coefficient-wise modular arithmetic, RNS rescale /
 * mod-switch, key-switch inner product).  The reference has no FFI layer: its seam
 * is the set of hehub:: free functions cited per entry point below (paths relative
 * to the reference root).  A maintainer binds these symbols from the reference's
 * C++; hehub_b200/cpp/hehub/ holds that host-side mirror.
 *
 * Conventions
 *   - Every polynomial operand is a DEVICE pointer to a contiguous little-endian
 *     u64 slab indexed [batch][poly][limb][N]; `moduli` are HOST arrays.  A
 *     key-switch key is [row p < L][half < 2][limb k <= L][N], limb L being the
 *     special modulus P (reference: RlweKsk = vector<RlweCt>, keys.h:19-32).
 *   - All arithmetic is integer mod q_i and lazy exactly like the reference:
 *     results are the same raw u64 representatives the reference CPU path
 *     produces on the same inputs (bit-exact), not merely congruent values.
 *   - Calls are asynchronous on the context's CUDA stream and return a status
 *     code; no exception crosses the ABI.  A context is thread-compatible (one
 *     thread at a time), like the reference (which is single-threaded).
 *   - There is no CPU fallback: without a CUDA device every compute entry point
 *     fails with HEHUB_B200_ERR_CUDA.
 */
#ifndef HEHUB_B200_H
#define HEHUB_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HEHUB_B200_OK 0
#define HEHUB_B200_ERR_INVALID 1     /* std::invalid_argument in the reference      */
#define HEHUB_B200_ERR_CUDA 2        /* CUDA runtime failure / no device            */
#define HEHUB_B200_ERR_UNSUPPORTED 3 /* the reference throws const char* here       */
#define HEHUB_B200_ERR_NOMEM 4

typedef struct hehub_b200_ctx hehub_b200_ctx;

/* ---- context ------------------------------------------------------------- */
const char *hehub_b200_version(void);
/* stream: a cudaStream_t (as void*) to run on, or NULL for a context-owned stream. */
int hehub_b200_ctx_create(hehub_b200_ctx **out, int device, void *stream);
int hehub_b200_ctx_destroy(hehub_b200_ctx *ctx);
int hehub_b200_ctx_set_stream(hehub_b200_ctx *ctx, void *stream);
int hehub_b200_ctx_synchronize(hehub_b200_ctx *ctx);
const char *hehub_b200_last_error(const hehub_b200_ctx *ctx);
/* options: "force_generic" (0/1) routes transforms through the one-level-per-sweep kernels
 * (an on-device cross-check of the fast path); "scratch_cap_mib" bounds the workspace a
 * batched call may use (large batches are processed in waves);
 * "host_chunk_kib" sets the chunk size of the host-buffer pipeline (default 16384; measured best on PCIe Gen5, profiles/r1d_e2e_chunk_sweep.log);
 * "latency_rows": transform launches with at most this many rows (one limb of one polynomial each) split every
 *   row over a 4- / 8-CTA cluster (default -1 = half the SM count; 0 = never).
 * "latency2_rows": ... and with at most this many rows (N = 4096 / 8192) over an 8-CTA cluster with its twiddle tables staged
 *   in shared memory (default -1 = a tenth of the SM count; 0 = never).
 * "pair_path" (default 1): key switch + drop of the last prime (ckks / bgv relinearize, mult_relin, rotate, conjugate) of a few
 *   ciphertexts per call at N = 4096 / 8192 as TWO cluster launches, rescale / mod_switch as ONE (csrc/ks_pair.cuh);
 *   0 = always the wave path, 2 = whenever the shapes allow.  "pair_fill_pct" (default 100): the form is taken while
 *   batch * L * L is at most this percentage of the SM count.  "pair_tpc": forward transforms per cluster (0 = automatic).
 * "fused_drop" (default 1): one ciphertext per call at N = 16384 / 32768 — the inverse transform of the special prime's limb runs
 *   inside the key switch's inner-product launch (its clusters wait on a counter for that limb's words) instead of alone on the
 *   GPU between two launches; 0 = separate launch (A/B).
 * "single_launch" (default 0): 1 makes hehub_b200_ckks_mult_relin with batch 1 at N = 4096 / 8192 run as ONE launch (grid
 *   barriers between its six phases) instead of six programmatically chained launches; measured slower (39 vs 33 us at
 *   N = 8192, L = 4), kept for A/B.
 * Environment: HEHUB_B200_PDL=0 turns programmatic dependent launch off (A/B only); HEHUB_B200_DEBUG=1 prints, once per cluster
 *   kernel, how many of its clusters the device holds at once (stderr). */
int hehub_b200_ctx_set_option(hehub_b200_ctx *ctx, const char *name, int64_t value);
/* number of kernels this context has launched since creation (bench bookkeeping) */
uint64_t hehub_b200_launch_count(const hehub_b200_ctx *ctx);

/* Build and upload the twiddle tables and per-modulus constants for (logN, q) for
 * every q in moduli[].  Optional (tables are built on first use), like
 * cache_ntt_factors_strict — src/fhe/common/ntt.cpp:225-231.  Fails with
 * ERR_INVALID when round(log2 q) > 59 or 2N does not divide q-1 (ntt.cpp:26-47). */
int hehub_b200_tables_prepare(hehub_b200_ctx *ctx, unsigned logn, const uint64_t *moduli, size_t nmod);

/* ---- parameter selection (host only; usable without a device) --------------------------------
 * prime_row: the reference's prime table, src/fhe/common/primelists.cpp:5-192 — row `bits` holds the primes
 *   create_params draws from, in order (rows 27..59; the table's own irregularities are reproduced, see
 *   hehub_b200/csrc/params.cu).  Writes min(row length, capacity) entries to out and returns the row length.
 * pick_moduli: ckks::create_params' selection, src/fhe/ckks/basics.cpp:14-38 — the additional modulus is drawn
 *   first, then the chain in order, one cursor per bit-size row.  additional_out == NULL: no additional
 *   modulus is drawn (hehub::create_params, src/fhe/primitives/rlwe.cpp:9-29).  ERR_UNSUPPORTED when a row
 *   runs out ("No suitable primes in the library."). */
int hehub_b200_prime_row(unsigned bits, size_t capacity, uint64_t *out);
int hehub_b200_pick_moduli(const unsigned *moduli_bits, size_t L, unsigned additional_bits, uint64_t *moduli_out,
                           uint64_t *additional_out);

/* ---- device slabs (the device analogue of SmartArray / FixedBlockAllocator,
 *      src/fhe/common/allocator.h:12-223: pooled per block size, reused on free) -- */
int hehub_b200_slab_alloc(hehub_b200_ctx *ctx, size_t n_words, uint64_t **out_dev);
int hehub_b200_slab_free(hehub_b200_ctx *ctx, uint64_t *dev);
int hehub_b200_slab_h2d(hehub_b200_ctx *ctx, uint64_t *dev, const uint64_t *host, size_t n_words);
int hehub_b200_slab_d2h(hehub_b200_ctx *ctx, uint64_t *host, const uint64_t *dev, size_t n_words);
int hehub_b200_slab_d2d(hehub_b200_ctx *ctx, uint64_t *dst, const uint64_t *src, size_t n_words);
/* pinned host staging buffers for the host-buffer entry points */
int hehub_b200_host_alloc(hehub_b200_ctx *ctx, size_t n_words, uint64_t **out_host);
int hehub_b200_host_free(hehub_b200_ctx *ctx, uint64_t *host);

/* ---- transforms ------------------------------------------------------------
 * x: [batch][L][N] in place.  Forward: natural-order coefficients in, bit-reversed
 * evaluation order out, values < 2q under the reference's approximate reduction.
 *   ntt_negacyclic_inplace_lazy  — src/fhe/common/ntt.cpp:145-176, ntt.h:41-51
 *   intt_negacyclic_inplace_lazy — src/fhe/common/ntt.cpp:178-223, ntt.h:72-82
 *   strict != 0 adds reduce_strict (intt_negacyclic_inplace, ntt.h:89-92). */
int hehub_b200_ntt_fwd_lazy(hehub_b200_ctx *ctx, unsigned logn, const uint64_t *moduli, size_t L,
                            uint64_t *x, size_t batch);
int hehub_b200_intt_lazy(hehub_b200_ctx *ctx, unsigned logn, const uint64_t *moduli, size_t L,
                         uint64_t *x, size_t batch, int strict);

/* ---- coefficient-wise kernels (n = words per limb; slabs are [batch][L][n]) --
 *   mulmod_hybrid_lazy : batched_mul_mod_hybrid_lazy, mod_arith.cpp:64-92 via
 *                        RnsIntVec operator*, rns.cpp:120-140  (c may alias a or b)
 *   add_lazy / sub_lazy: operator+= / operator-=, rns.cpp:58-118   (x op= y)
 *   mul_scalar_lazy    : operator*=(vector<u64>), rns.cpp:155-171; scalars[L] host
 *   reduce_strict      : batched_reduce_strict, mod_arith.h:58-72
 *   barrett_lazy/barrett: mod_arith.cpp:9-17, mod_arith.h:18-25
 *   montgomery128_lazy : mod_arith.cpp:113-134; in = n (lo,hi) pairs, one modulus */
int hehub_b200_mulmod_hybrid_lazy(hehub_b200_ctx *ctx, size_t n, const uint64_t *moduli, size_t L,
                                  const uint64_t *a, const uint64_t *b, uint64_t *c, size_t batch);
int hehub_b200_add_lazy(hehub_b200_ctx *ctx, size_t n, const uint64_t *moduli, size_t L,
                        uint64_t *x, const uint64_t *y, size_t batch);
int hehub_b200_sub_lazy(hehub_b200_ctx *ctx, size_t n, const uint64_t *moduli, size_t L,
                        uint64_t *x, const uint64_t *y, size_t batch);
int hehub_b200_mul_scalar_lazy(hehub_b200_ctx *ctx, size_t n, const uint64_t *moduli, size_t L,
                               uint64_t *x, const uint64_t *scalars, size_t batch);
int hehub_b200_reduce_strict(hehub_b200_ctx *ctx, size_t n, const uint64_t *moduli, size_t L,
                             uint64_t *x, size_t batch);
int hehub_b200_barrett_lazy(hehub_b200_ctx *ctx, size_t n, const uint64_t *moduli, size_t L,
                            uint64_t *x, size_t batch);
int hehub_b200_barrett(hehub_b200_ctx *ctx, size_t n, const uint64_t *moduli, size_t L,
                       uint64_t *x, size_t batch);
int hehub_b200_montgomery128_lazy(hehub_b200_ctx *ctx, uint64_t q, size_t n, const uint64_t *in_lohi,
                                  uint64_t *out);

/* ---- Galois permutations on NTT-form polynomials (next row §8(f).1) ----------
 *   cycle / involution — src/fhe/common/permutation.cpp:28-75; in/out [batch][L][N] */
int hehub_b200_galois_cycle(hehub_b200_ctx *ctx, unsigned logn, size_t L, const uint64_t *in,
                            uint64_t *out, size_t step, size_t batch);
int hehub_b200_galois_involution(hehub_b200_ctx *ctx, unsigned logn, size_t L, const uint64_t *in,
                                 uint64_t *out, size_t batch);

/* ---- scheme-level ops --------------------------------------------------------
 * ckks_tensor: ckks::mult_low_level, src/fhe/ckks/arith.cpp:55-62 (== bgv::mult_low_level,
 *   bgv/arith.cpp:59-69).  ct1, ct2: [batch][2][L][N] -> quad: [batch][3][L][N].
 * ext_prod_montgomery: src/fhe/primitives/rgsw.cpp:57-156.  ext_moduli has L+1
 *   entries (q_0..q_{L-1}, P); in: [batch][L][N]; key: [L][2][L+1][N] shared by
 *   the batch; out: [batch][2][L+1][N].
 * ckks_rescale: ckks::rescale_inplace(ct, 1), src/fhe/ckks/rescaling.cpp:14-90.
 *   ct: [batch][2][L][N] -> out: [batch][2][L-1][N]  (L >= 2).
 * bgv_mod_switch: bgv::mod_switch_inplace(ct, 1), src/fhe/bgv/mod_switch.cpp:13-90.
 * ckks_relinearize: src/fhe/ckks/arith.cpp:64-73; quad [batch][3][L][N] -> out [batch][2][L][N].
 * bgv_relinearize: src/fhe/bgv/arith.cpp:71-79 with plain modulus t for the internal
 *   mod-switch (the reference always runs it with the default t = 1).
 * ckks_mult_relin: ckks::mult, src/fhe/ckks/ckks.h:270-274 = tensor + relinearize,
 *   fused so the degree-2 ciphertext never round-trips through host code.
 * bgv_mult_relin: bgv::mult, src/fhe/bgv/bgv.h (mult_low_level, bgv/arith.cpp:59-69, then
 *   relinearize, :71-79) with plain modulus t for the internal mod-switch (the reference: t = 1).
 * ckks_rotate / ckks_conjugate: src/fhe/ckks/arith.cpp:75-93 (next row §8(f).1). */
int hehub_b200_ckks_tensor(hehub_b200_ctx *ctx, unsigned logn, const uint64_t *moduli, size_t L,
                           const uint64_t *ct1, const uint64_t *ct2, uint64_t *quad, size_t batch);
int hehub_b200_ext_prod_montgomery(hehub_b200_ctx *ctx, unsigned logn, const uint64_t *ext_moduli,
                                   size_t L, const uint64_t *in, const uint64_t *key, uint64_t *out,
                                   size_t batch);
int hehub_b200_ckks_rescale(hehub_b200_ctx *ctx, unsigned logn, const uint64_t *moduli, size_t L,
                            const uint64_t *ct, uint64_t *out, size_t batch);
int hehub_b200_bgv_mod_switch(hehub_b200_ctx *ctx, unsigned logn, const uint64_t *moduli, size_t L,
                              uint64_t plain_modulus, const uint64_t *ct, uint64_t *out, size_t batch);
int hehub_b200_ckks_relinearize(hehub_b200_ctx *ctx, unsigned logn, const uint64_t *ext_moduli, size_t L,
                                const uint64_t *quad, const uint64_t *key, uint64_t *out, size_t batch);
int hehub_b200_bgv_relinearize(hehub_b200_ctx *ctx, unsigned logn, const uint64_t *ext_moduli, size_t L,
                               uint64_t plain_modulus, const uint64_t *quad, const uint64_t *key,
                               uint64_t *out, size_t batch);
int hehub_b200_ckks_mult_relin(hehub_b200_ctx *ctx, unsigned logn, const uint64_t *ext_moduli, size_t L,
                               const uint64_t *ct1, const uint64_t *ct2, const uint64_t *key,
                               uint64_t *out, size_t batch);
int hehub_b200_bgv_mult_relin(hehub_b200_ctx *ctx, unsigned logn, const uint64_t *ext_moduli, size_t L,
                              uint64_t plain_modulus, const uint64_t *ct1, const uint64_t *ct2,
                              const uint64_t *key, uint64_t *out, size_t batch);
int hehub_b200_ckks_rotate(hehub_b200_ctx *ctx, unsigned logn, const uint64_t *ext_moduli, size_t L,
                           const uint64_t *ct, const uint64_t *key, size_t step, uint64_t *out,
                           size_t batch);
int hehub_b200_ckks_conjugate(hehub_b200_ctx *ctx, unsigned logn, const uint64_t *ext_moduli, size_t L,
                              const uint64_t *ct, const uint64_t *key, uint64_t *out, size_t batch);

/* ---- RLWE cores (next row §8(f).3) — src/fhe/primitives/rlwe.cpp:34-71 ---------------
 * decrypt_core: pt = reduce_strict(INTT(c0 + c1 * sk)); ct [batch][2][L][N] and sk [L][N] in NTT
 *   form, pt [batch][L][N] coefficients < q.
 * encrypt_core: ct = (NTT(e) - c1 * sk + NTT(pt), c1).  The reference draws the mask c1 (uniform,
 *   NTT form) and the error e (small coefficients, reduced mod each q) from a process-global RNG
 *   (sampling.cpp:12-69); here the caller supplies them, so the op is deterministic.
 *   pt, e, c1: [batch][L][N]; out: [batch][2][L][N]. */
int hehub_b200_rlwe_decrypt_core(hehub_b200_ctx *ctx, unsigned logn, const uint64_t *moduli, size_t L,
                                 const uint64_t *ct, const uint64_t *sk, uint64_t *pt, size_t batch);
int hehub_b200_rlwe_encrypt_core(hehub_b200_ctx *ctx, unsigned logn, const uint64_t *moduli, size_t L,
                                 const uint64_t *pt, const uint64_t *sk, const uint64_t *c1, const uint64_t *e,
                                 uint64_t *out, size_t batch);

/* ---- RNS base transform and key generation (next row §8(f).2) ----------------------------
 * rns_base_transform — src/fhe/common/rns_transform.cpp:11-126 on coefficient-form words (the
 *   strict reduction of :116 included).  from_single: in [batch][n] under q_old -> out
 *   [batch][Lnew][n].  to_single: in [batch][L][n] -> out [batch][n]; only the reference's
 *   small-coefficient path (:47-84) is built: ERR_UNSUPPORTED when some coefficient is not the same
 *   small signed value under every old modulus (the reference then composes big integers).
 *   to_single synchronises the stream (it has to read the verdict).
 * ksk_generate — RlweKsk::RlweKsk, src/fhe/primitives/keys.cpp:8-36 with rgsw_encrypt_montgomery
 *   (rgsw.cpp:11-55).  The reference draws one RLWE sample per row from a process-global RNG; here
 *   the caller supplies them: masks [L][L+1][N] (uniform, NTT form), errors [L][L+1][N] (small
 *   coefficients reduced mod each modulus).  sk_curr, sk_orig: [L][N] NTT form; key: [L][2][L+1][N]
 *   (the layout ext_prod_montgomery takes). */
int hehub_b200_rns_base_transform_from_single(hehub_b200_ctx *ctx, uint64_t q_old, const uint64_t *new_moduli,
                                              size_t Lnew, const uint64_t *in, uint64_t *out, size_t n, size_t batch);
int hehub_b200_rns_base_transform_to_single(hehub_b200_ctx *ctx, const uint64_t *old_moduli, size_t L,
                                            uint64_t new_modulus, const uint64_t *in, uint64_t *out, size_t n,
                                            size_t batch);
int hehub_b200_ksk_generate(hehub_b200_ctx *ctx, unsigned logn, const uint64_t *ext_moduli, size_t L,
                            const uint64_t *sk_curr, const uint64_t *sk_orig, const uint64_t *masks,
                            const uint64_t *errors, uint64_t *key);

/* ---- host-buffer variants --------------------------------------------------------
 * The reference keeps every RnsPolynomial in host memory (rns.cpp:25-27), so a caller that has
 * not moved its data to device slabs calls these: operands are HOST pointers (pinned memory from
 * hehub_b200_host_alloc gives full PCIe speed), the batch is streamed through device staging
 * slabs in chunks with copy-in, kernels and copy-out overlapped on three streams, and the call
 * returns when host_out holds the result.  ntt_host: forward != 0 -> ntt_fwd_lazy, else intt_lazy
 * (+strict); host_in may equal host_out.  ckks_mult_relin_host: ciphertexts on the host, the
 * key-switch key resident on the device (it is reused by every call). */
int hehub_b200_ntt_host(hehub_b200_ctx *ctx, int forward, unsigned logn, const uint64_t *moduli, size_t L,
                        const uint64_t *host_in, uint64_t *host_out, size_t batch, int strict);
int hehub_b200_ckks_mult_relin_host(hehub_b200_ctx *ctx, unsigned logn, const uint64_t *ext_moduli, size_t L,
                                    const uint64_t *host_ct1, const uint64_t *host_ct2, const uint64_t *dev_key,
                                    uint64_t *host_out, size_t batch);

/* ---- harness helpers (synthetic inputs generated on the device) --------------
 * lcg_fill: row r of x ([rows][n]) = SURVEY Appendix B LCG with seed seed0 + r*seed_stride,
 *   reduced mod moduli[r % L].  fnv1a: FNV-1a over the words of x (device reduction
 *   is order-dependent, so this one runs single-threaded per row then chains on host). */
int hehub_b200_lcg_fill(hehub_b200_ctx *ctx, size_t n, const uint64_t *moduli, size_t L, uint64_t *x,
                        size_t rows, uint64_t seed0, uint64_t seed_stride);

/* ---- batched-ciphertext sweep across GPUs (BASELINE config 5) -----------------------------
 * ckks::mult + relinearize (src/fhe/ckks/ckks.h:270-274) over `total` independent synthetic ciphertext pairs, cut
 * into contiguous per-rank ranges: one process per GPU, no data-path collective.  Inputs are generated on the
 * device from (seed, global pair index), so pair i has the same words however the batch is sharded; every result
 * is reduced on the device to one checksum, sum_j word_j * (2j + 1) * 0x9E3779B97F4A7C15 mod 2^64.  Collectives
 * only broadcast the key-switch key from rank 0 and all-gather the checksums; they go through a provider:
 * NCCL over NVLink in the product (hehub_b200_nccl_*, libhehub_b200_nccl.so), gloo in the CPU test-suite.
 *   shard_range: contiguous balanced partition (the first total % world ranks take one extra unit).
 *   sweep_run: this rank's share in waves of at most `wave` pairs.  my_checksums_host[count] (may be NULL);
 *     all_checksums_host[total] is filled on rank 0 (NULL elsewhere); op_seconds (may be NULL) receives the device time of
 *     the mult+relin calls alone (CUDA events on the context's stream).  Synchronises the stream before returning. */
typedef struct hehub_b200_collectives {
    void *self;
    int (*broadcast)(void *self, void *dev_buf, size_t bytes, int root, void *stream); /* in place; 0 = ok */
    int (*allgather)(void *self, const void *dev_send, void *dev_recv, size_t bytes_per_rank, void *stream);
} hehub_b200_collectives;
typedef struct hehub_b200_sweep hehub_b200_sweep;
void hehub_b200_shard_range(size_t total, int world, int rank, size_t *first, size_t *count);
int hehub_b200_ct_checksums(hehub_b200_ctx *ctx, const uint64_t *words, size_t words_per_ct, size_t cts, uint64_t *sums_dev);
int hehub_b200_sweep_create(hehub_b200_sweep **out, hehub_b200_ctx *ctx, unsigned logn, const uint64_t *ext_moduli, size_t L,
                            uint64_t seed, int rank, int world, const hehub_b200_collectives *coll /* NULL: world == 1 */);
int hehub_b200_sweep_destroy(hehub_b200_sweep *s);
int hehub_b200_sweep_make_key(hehub_b200_sweep *s);
const uint64_t *hehub_b200_sweep_key(const hehub_b200_sweep *s); /* device pointer, [L][2][L+1][N] */
int hehub_b200_sweep_fill_inputs(hehub_b200_sweep *s, uint64_t *ct1, uint64_t *ct2, size_t first_ct, size_t count);
int hehub_b200_sweep_run(hehub_b200_sweep *s, size_t total, size_t wave, uint64_t *my_checksums_host,
                         uint64_t *all_checksums_host, size_t *first, size_t *count, double *op_seconds);

/* The NCCL provider — implemented by hehub_b200/libhehub_b200_nccl.so (csrc/nccl_provider.cpp), not by the core library.
 * Rank 0 calls unique_id and passes the bytes to the other ranks (MPI_Bcast, a file, the launcher's process group); every
 * rank then calls create, which runs ncclCommInitRank.  version: NCCL's version code (e.g. 22703). */
#define HEHUB_B200_NCCL_ID_BYTES 128
int hehub_b200_nccl_version(void);
int hehub_b200_nccl_unique_id(uint8_t id[HEHUB_B200_NCCL_ID_BYTES]);
int hehub_b200_nccl_create(hehub_b200_collectives *out, int rank, int world, const uint8_t id[HEHUB_B200_NCCL_ID_BYTES], int device);
int hehub_b200_nccl_destroy(hehub_b200_collectives *c);

#ifdef __cplusplus
}
#endif
#endif
