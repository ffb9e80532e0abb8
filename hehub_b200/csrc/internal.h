// internal.h — declarations shared between the translation units of libhehub_b200.so.
#pragma once
#include "context.h"

struct hehub_b200_ctx {
    hb::Context c;
};

#define CTX_GUARD(ctx)                                        \
    if (!(ctx)) return 1;                \
    Context &c = (ctx)->c;                                    \
    {                                                         \
        cudaError_t e__ = cudaSetDevice(c.device);            \
        if (e__ != cudaSuccess) return c.cuda_fail(e__, "cudaSetDevice"); \
    }

namespace hb {

// Common argument checks; returns 0 or an error code with ctx.last_error set.
int check_ring(Context &c, unsigned logn, size_t L, size_t batch);

// ops.cu — composite operations (device pointers, already validated)
int op_ckks_tensor(Context &c, unsigned logn, const u64 *moduli, size_t L, const u64 *ct1, const u64 *ct2,
                   u64 *quad, size_t batch);
// ginv != 1: `in` is read through the Galois permutation with inverse factor ginv (rotate / conjugate)
int op_ext_prod(Context &c, unsigned logn, const u64 *ext_moduli, size_t L, const u64 *in, size_t in_batch_stride,
                const u64 *key, u64 *out, size_t batch, unsigned ginv = 1);
// drop the last prime of ct [batch][2][L][N] -> out [batch][2][L-1][N]; t == 0: CKKS rescale,
// else BGV mod-switch.  Optional addend (lazy-added to the result): [batch][..][L-1][N] with
// the given strides, applied to the first add_halves polynomials.
int op_drop_last(Context &c, unsigned logn, const u64 *moduli, size_t L, u64 t, const u64 *ct, u64 *out,
                 size_t batch, const u64 *addend, size_t add_batch_stride, size_t add_poly_stride, int add_halves,
                 unsigned add_ginv = 1);
int op_relinearize(Context &c, unsigned logn, const u64 *ext_moduli, size_t L, u64 t, const u64 *quad,
                   const u64 *key, u64 *out, size_t batch);
int op_mult_relin(Context &c, unsigned logn, const u64 *ext_moduli, size_t L, u64 t, const u64 *ct1, const u64 *ct2,
                  const u64 *key, u64 *out, size_t batch);
int op_rlwe_decrypt_core(Context &c, unsigned logn, const u64 *moduli, size_t L, const u64 *ct, const u64 *sk, u64 *pt,
                         size_t batch);
int op_rlwe_encrypt_core(Context &c, unsigned logn, const u64 *moduli, size_t L, const u64 *pt, const u64 *sk, const u64 *c1,
                         const u64 *e, u64 *out, size_t batch);
int op_base_from_single(Context &c, u64 q_old, const u64 *new_moduli, size_t Lnew, const u64 *in, u64 *out, size_t n, size_t batch);
int op_base_to_single(Context &c, const u64 *old_moduli, size_t L, u64 new_modulus, const u64 *in, u64 *out, size_t n, size_t batch,
                      bool defer_verdict = false);
int op_ksk_generate(Context &c, unsigned logn, const u64 *ext_moduli, size_t L, const u64 *sk_curr, const u64 *sk_orig,
                    const u64 *masks, const u64 *errors, u64 *key);
int op_galois(Context &c, unsigned logn, size_t L, const u64 *in, u64 *out, bool conj, size_t step, size_t batch);
int op_galois_keyswitch(Context &c, unsigned logn, const u64 *ext_moduli, size_t L, const u64 *ct, const u64 *key,
                        bool conj, size_t step, u64 *out, size_t batch);


// api.cu — plain in-place transform on device rows [batch][L][N]
int run_transform(Context &c, bool forward, unsigned logn, const u64 *moduli, size_t L, u64 *x, size_t batch, int strict);

} // namespace hb
