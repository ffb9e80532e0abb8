/*
 * hehub_oracle.h — CPU oracle for the HEhub RNS polynomial-arithmetic hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C restatement of the reference's
 * algorithms (primihub/hehub, src/fhe/...).  It exists so that the CUDA path in
 * hehub_b200/ can be checked word-for-word on identical inputs.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load it; the product library never links or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks every function here against
 * (a) the known-answer values recorded from the unmodified reference
 * (tests/golden/reference_kat.json, produced by oracle/make_golden.py running the
 * real reference built into oracle/_ref/), (b) the reference's own unit-test
 * properties (tests/ntt_t.cpp, tests/mod_arith_t.cpp, tests/ckks_t.cpp:136-175),
 * and (c) when oracle/_ref/libhehub_ref.so is present, a randomized differential
 * test against the real reference.
 *
 * Layout conventions (shared with include/hehub_b200.h): every polynomial slab is
 * a contiguous little-endian u64 array indexed [poly][limb][N]; a key-switch key
 * is [row p < L][half h < 2][limb k <= L][N], the last limb being the special
 * modulus P.
 */
#ifndef HEHUB_ORACLE_H
#define HEHUB_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef uint64_t orc_u64;

/* ---- word-level primitives (mod_arith.h / mod_arith.cpp) ---- */
orc_u64 orc_pow_mod(orc_u64 q, orc_u64 base, orc_u64 e);
orc_u64 orc_root_2n(orc_u64 q, orc_u64 n); /* 0 if 2n does not divide q-1 */
orc_u64 orc_inverse_mod_prime(orc_u64 elem, orc_u64 prime);
orc_u64 orc_harvey_quotient(orc_u64 w, orc_u64 q); /* floor(w*2^64/q) */
orc_u64 orc_harvey_lazy(orc_u64 q, orc_u64 x, orc_u64 w, orc_u64 w_harvey);
void orc_mont_consts(orc_u64 q, orc_u64 *minus_qinv, orc_u64 *r, orc_u64 *r_harvey);
void orc_barrett_lazy(orc_u64 q, size_t n, orc_u64 *x);
void orc_barrett(orc_u64 q, size_t n, orc_u64 *x);
void orc_reduce_strict(orc_u64 q, size_t n, orc_u64 *x);
void orc_mul_hybrid_lazy(orc_u64 q, size_t n, const orc_u64 *a, const orc_u64 *b, orc_u64 *c);
void orc_mul_barrett_lazy(orc_u64 q, size_t n, const orc_u64 *a, const orc_u64 *b, orc_u64 *c);
/* in = n 128-bit words stored as (lo, hi) u64 pairs */
void orc_montgomery128_lazy(orc_u64 q, size_t n, const orc_u64 *in_lohi, orc_u64 *out);
void orc_add_lazy(orc_u64 q, size_t n, orc_u64 *x, const orc_u64 *y);
void orc_sub_lazy(orc_u64 q, size_t n, orc_u64 *x, const orc_u64 *y);
void orc_mul_scalar_lazy(orc_u64 q, size_t n, orc_u64 *x, orc_u64 scalar);

/* ---- transforms (ntt.cpp) ---- returns 0 ok, nonzero on parameter error */
int orc_ntt_fwd_lazy(unsigned logn, orc_u64 q, orc_u64 *x);
int orc_intt_lazy(unsigned logn, orc_u64 q, orc_u64 *x);
/* same dataflow as orc_intt_lazy without the two bit-reversal permutations
 * (the form the CUDA kernels use); must be word-identical to orc_intt_lazy */
int orc_intt_lazy_folded(unsigned logn, orc_u64 q, orc_u64 *x);
/* raw tables, reference index order: fwd[i], fwd_h[i] i<N ; inv[i], inv_h[i] i<2N */
int orc_ntt_tables(unsigned logn, orc_u64 q, orc_u64 *fwd, orc_u64 *fwd_h,
                   orc_u64 *inv, orc_u64 *inv_h);
void orc_clear_caches(void);

int orc_bench_ntt(unsigned logn, orc_u64 q, orc_u64 *x, size_t rows, int forward);

/* ---- composite ops on [poly][limb][N] slabs ---- */
int orc_poly_ntt_fwd(unsigned logn, size_t L, const orc_u64 *moduli, orc_u64 *x);
int orc_poly_intt(unsigned logn, size_t L, const orc_u64 *moduli, orc_u64 *x, int strict);
int orc_ckks_tensor(unsigned logn, size_t L, const orc_u64 *moduli,
                    const orc_u64 *ct1, const orc_u64 *ct2, orc_u64 *quad);
int orc_ext_prod(unsigned logn, size_t L, const orc_u64 *ext_moduli,
                 const orc_u64 *in, const orc_u64 *key, orc_u64 *out);
int orc_ckks_rescale(unsigned logn, size_t L, const orc_u64 *moduli,
                     const orc_u64 *ct, orc_u64 *out);
int orc_bgv_mod_switch(unsigned logn, size_t L, const orc_u64 *moduli, orc_u64 t,
                       const orc_u64 *ct, orc_u64 *out);
int orc_ckks_relinearize(unsigned logn, size_t L, const orc_u64 *ext_moduli,
                         const orc_u64 *quad, const orc_u64 *key, orc_u64 *out);
int orc_bgv_relinearize(unsigned logn, size_t L, const orc_u64 *ext_moduli, orc_u64 t,
                        const orc_u64 *quad, const orc_u64 *key, orc_u64 *out);
int orc_ckks_mult_relin(unsigned logn, size_t L, const orc_u64 *ext_moduli,
                        const orc_u64 *ct1, const orc_u64 *ct2, const orc_u64 *key,
                        orc_u64 *out);
/* RLWE cores (rlwe.cpp:34-71); encrypt takes the samples (mask c1 in NTT form, error e in
 * coefficient form) from the caller so it is deterministic */
int orc_rlwe_decrypt_core(unsigned logn, size_t L, const orc_u64 *moduli, const orc_u64 *ct, const orc_u64 *sk, orc_u64 *pt);
int orc_rlwe_encrypt_core(unsigned logn, size_t L, const orc_u64 *moduli, const orc_u64 *pt, const orc_u64 *sk,
                          const orc_u64 *c1, const orc_u64 *e, orc_u64 *out);
/* rns_base_transform (rns_transform.cpp:11-84; many->one: small-coefficient path, rc 3 otherwise) and
 * RlweKsk construction (keys.cpp:8-36) on caller-supplied samples */
int orc_base_transform_from_single(orc_u64 q_old, size_t n, const orc_u64 *in, const orc_u64 *new_moduli, size_t Lnew, orc_u64 *out);
int orc_base_transform_to_single(size_t n, size_t L, const orc_u64 *old_moduli, const orc_u64 *in, orc_u64 new_modulus, orc_u64 *out);
int orc_ksk_generate(unsigned logn, size_t L, const orc_u64 *ext_moduli, const orc_u64 *sk_curr, const orc_u64 *sk_orig,
                     const orc_u64 *masks, const orc_u64 *errors, orc_u64 *key);
int orc_galois_cycle(unsigned logn, size_t L, const orc_u64 *in, orc_u64 *out, size_t step);
int orc_galois_involution(unsigned logn, size_t L, const orc_u64 *in, orc_u64 *out);
/* rotate / conjugate = permutation + ext_prod + rescale + add (ckks/arith.cpp:75-93) */
int orc_ckks_rotate(unsigned logn, size_t L, const orc_u64 *ext_moduli, const orc_u64 *ct,
                    const orc_u64 *key, size_t step, orc_u64 *out);
int orc_ckks_conjugate(unsigned logn, size_t L, const orc_u64 *ext_moduli, const orc_u64 *ct,
                       const orc_u64 *key, orc_u64 *out);

/* ---- harness helpers ---- */
void orc_lcg_fill(orc_u64 seed, orc_u64 q, size_t n, orc_u64 *x);
orc_u64 orc_fnv1a(const orc_u64 *x, size_t n, orc_u64 h);
/* the reference's prime table restated as a rule: row `bits` holds, in descending
 * order, the primes p < 2^bits with p = 1 (mod 2^16); writes up to `count`. */
int orc_prime_row(unsigned bits, size_t count, orc_u64 *out);
/* ckks::create_params(dimension, moduli_bits, additional_mod_bits): P first. */
int orc_ckks_pick_moduli(const unsigned *moduli_bits, size_t L, unsigned additional_bits,
                         orc_u64 *moduli_out, orc_u64 *additional_out);

#ifdef __cplusplus
}
#endif
#endif
