/*
 * hehub_oracle.c — CPU oracle (TEST INFRASTRUCTURE ONLY, see hehub_oracle.h).
 *
 * A from-scratch C restatement of the reference's hot-path arithmetic.  Every
 * function cites the reference file:line whose word-level behaviour it follows
 * (paths relative to the reference root).  Raw lazy representatives — not just
 * residues — must match the reference, so the op sequences below are kept
 * arithmetic-for-arithmetic identical even where a cheaper form exists.
 */
#include "hehub_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef orc_u64 u64;
typedef unsigned __int128 u128;

/* ------------------------------------------------------------------ */
/* word-level primitives                                               */
/* ------------------------------------------------------------------ */

static inline u64 mulmod(u64 a, u64 b, u64 q) { return (u64)((u128)a * b % q); }

/* src/fhe/common/ntt.cpp:9-24 (square-and-multiply; any correct pow gives the
 * same canonical residue) */
u64 orc_pow_mod(u64 q, u64 base, u64 e) {
    u64 r = 1 % q;
    base %= q;
    while (e) {
        if (e & 1) r = mulmod(r, base, q);
        base = mulmod(base, base, q);
        e >>= 1;
    }
    return r;
}

/* src/fhe/common/ntt.cpp:26-39: least g >= 2 with g^((q-1)/2) == q-1, then
 * psi = g^((q-1)/(2n)). */
u64 orc_root_2n(u64 q, u64 n) {
    if (n == 0 || (q - 1) % (2 * n) != 0) return 0;
    u64 g = 2;
    while (orc_pow_mod(q, g, (q - 1) / 2) != q - 1) g++;
    return orc_pow_mod(q, g, (q - 1) / (2 * n));
}

/* src/fhe/common/mod_arith.cpp:138-149 (xgcd in the reference; the canonical
 * inverse in [0, prime) is unique, here via Fermat). */
u64 orc_inverse_mod_prime(u64 elem, u64 prime) {
    if (prime == 1) return 0;
    return orc_pow_mod(prime, elem % prime, prime - 2);
}

/* w' = floor(w * 2^64 / q): ntt.cpp:57,72,84 ; rns.cpp:146,164 */
u64 orc_harvey_quotient(u64 w, u64 q) { return (u64)(((u128)w << 64) / q); }

/* src/fhe/common/mod_arith.h:74-78 */
u64 orc_harvey_lazy(u64 q, u64 x, u64 w, u64 wh) {
    u64 qhat = (u64)(((u128)x * wh) >> 64);
    return (u64)((u128)x * w) - (u64)((u128)qhat * q);
}

/* src/fhe/common/mod_arith.cpp:49-62: -q^{-1} mod 2^64, 2^64 mod q and its
 * Harvey companion. */
void orc_mont_consts(u64 q, u64 *minus_qinv, u64 *r, u64 *r_harvey) {
    /* Newton iteration for q^{-1} mod 2^64 (q odd) */
    u64 inv = q;
    for (int i = 0; i < 6; i++) inv *= 2 - q * inv;
    if (minus_qinv) *minus_qinv = (u64)0 - inv;
    u64 rr = ((u64)(-1LL) % q) + 1;
    if (r) *r = rr;
    if (r_harvey) *r_harvey = (u64)(((u128)rr << 64) / q);
}

/* src/fhe/common/mod_arith.cpp:9-17 */
void orc_barrett_lazy(u64 q, size_t n, u64 *x) {
    u64 c = (u64)(-1) / q;
    for (size_t i = 0; i < n; i++) {
        u64 qhat = (u64)(((u128)x[i] * c) >> 64);
        x[i] -= q * qhat;
    }
}

/* src/fhe/common/mod_arith.h:58-63 */
void orc_reduce_strict(u64 q, size_t n, u64 *x) {
    for (size_t i = 0; i < n; i++) x[i] -= (x[i] >= q) ? q : 0;
}

/* src/fhe/common/mod_arith.h:18-25 */
void orc_barrett(u64 q, size_t n, u64 *x) {
    orc_barrett_lazy(q, n, x);
    orc_reduce_strict(q, n, x);
}

/* src/fhe/common/mod_arith.cpp:64-92: Montgomery reduce a*b (-> *2^-64) then
 * Harvey-multiply by 2^64 mod q. */
void orc_mul_hybrid_lazy(u64 q, size_t n, const u64 *a, const u64 *b, u64 *c) {
    u64 mqi, r, rh;
    orc_mont_consts(q, &mqi, &r, &rh);
    for (size_t i = 0; i < n; i++) {
        u128 p = (u128)a[i] * b[i];
        u64 u = (u64)p * mqi;
        u64 t = (u64)((p + (u128)u * q) >> 64);
        c[i] = orc_harvey_lazy(q, t, r, rh);
    }
}

/* src/fhe/common/mod_arith.cpp:94-111 (128-bit Barrett; off the hot path, kept
 * for the mod_arith_t.cpp KATs). */
void orc_mul_barrett_lazy(u64 q, size_t n, const u64 *a, const u64 *b, u64 *c) {
    u128 cc = (u128)(-1) / q;
    u64 ch = (u64)(cc >> 64), cl = (u64)cc;
    for (size_t i = 0; i < n; i++) {
        u128 p = (u128)a[i] * b[i];
        u64 ah = (u64)(p >> 64), al = (u64)p;
        u64 qhat = ah * ch + (u64)((((u128)ah * cl) + ((u128)al * ch)) >> 64);
        c[i] = (u64)(p - (u128)q * qhat);
    }
}

/* src/fhe/common/mod_arith.cpp:113-134 */
void orc_montgomery128_lazy(u64 q, size_t n, const u64 *in, u64 *out) {
    u64 mqi;
    orc_mont_consts(q, &mqi, NULL, NULL);
    for (size_t i = 0; i < n; i++) {
        u128 a = ((u128)in[2 * i + 1] << 64) | in[2 * i];
        u64 u = (u64)a * mqi;
        out[i] = (u64)((a + (u128)u * q) >> 64);
    }
}

/* src/fhe/common/rns.cpp:78-84 */
void orc_add_lazy(u64 q, size_t n, u64 *x, const u64 *y) {
    u64 q2 = 2 * q;
    for (size_t i = 0; i < n; i++) {
        x[i] += y[i];
        x[i] -= (x[i] >= q2) ? q2 : 0;
    }
}

/* src/fhe/common/rns.cpp:109-115 */
void orc_sub_lazy(u64 q, size_t n, u64 *x, const u64 *y) {
    u64 q2 = 2 * q;
    for (size_t i = 0; i < n; i++) {
        x[i] += q2 - y[i];
        x[i] -= (x[i] >= q2) ? q2 : 0;
    }
}

/* src/fhe/common/rns.cpp:142-171 (scalar reduced, then Harvey multiply) */
void orc_mul_scalar_lazy(u64 q, size_t n, u64 *x, u64 scalar) {
    u64 s = scalar % q;
    u64 sh = orc_harvey_quotient(s, q);
    for (size_t i = 0; i < n; i++) x[i] = orc_harvey_lazy(q, x[i], s, sh);
}

/* ------------------------------------------------------------------ */
/* NTT tables and transforms                                           */
/* ------------------------------------------------------------------ */

/* src/fhe/common/permutation.h:41-55 */
static inline u64 bitrev(u64 x, unsigned bits) {
    u64 r = 0;
    for (unsigned i = 0; i < bits; i++) r |= ((x >> i) & 1) << (bits - 1 - i);
    return r;
}

typedef struct {
    u64 q;
    unsigned logn;
    u64 *fwd, *fwd_h; /* N entries, index i: psi^{bitrev(i)} */
    u64 *inv, *inv_h; /* 2N entries: per-level inverse twiddles then psi^{-i}/N */
} tables_t;

static tables_t *g_tables = NULL;
static size_t g_ntables = 0;

void orc_clear_caches(void) {
    for (size_t i = 0; i < g_ntables; i++) {
        free(g_tables[i].fwd);
        free(g_tables[i].fwd_h);
        free(g_tables[i].inv);
        free(g_tables[i].inv_h);
    }
    free(g_tables);
    g_tables = NULL;
    g_ntables = 0;
}

/* src/fhe/common/ntt.cpp:41-105 (NTTFactors).  Powers are produced by repeated
 * multiplication instead of one pow_mod per entry; the residues are canonical
 * either way. */
static const tables_t *get_tables(unsigned logn, u64 q) {
    for (size_t i = 0; i < g_ntables; i++)
        if (g_tables[i].q == q && g_tables[i].logn == logn) return &g_tables[i];

    /* ntt.cpp:43-47 */
    const u64 log_modulus = (u64)(log2((double)q) + 0.5);
    if (log_modulus > 59 || logn == 0 || logn > 16) return NULL;
    const size_t n = (size_t)1 << logn;
    const u64 psi = orc_root_2n(q, n);
    if (psi == 0) return NULL;

    tables_t t;
    t.q = q;
    t.logn = logn;
    t.fwd = malloc(n * sizeof(u64));
    t.fwd_h = malloc(n * sizeof(u64));
    t.inv = malloc(2 * n * sizeof(u64));
    t.inv_h = malloc(2 * n * sizeof(u64));
    u64 *pw = malloc(n * sizeof(u64));

    pw[0] = 1;
    for (size_t k = 1; k < n; k++) pw[k] = mulmod(pw[k - 1], psi, q);
    for (size_t i = 0; i < n; i++) { /* ntt.cpp:54-58 */
        t.fwd[i] = pw[bitrev(i, logn)];
        t.fwd_h[i] = orc_harvey_quotient(t.fwd[i], q);
    }

    const u64 psi_inv = orc_pow_mod(q, psi, 2 * n - 1); /* ntt.cpp:62-63 */
    pw[0] = 1;
    for (size_t k = 1; k < n; k++) pw[k] = mulmod(pw[k - 1], psi_inv, q);
    for (unsigned l = 0; l < logn; l++) { /* ntt.cpp:64-74 */
        size_t start = ((size_t)1 << l) - 1;
        size_t factor = (size_t)1 << (logn - l);
        for (size_t i = 0; i < ((size_t)1 << l); i++) {
            t.inv[start + i] = pw[(bitrev(i, l) * factor) & (n - 1)];
            /* exponent bitrev(i,l)*factor < n except l == 0 (exponent 0) */
            t.inv_h[start + i] = orc_harvey_quotient(t.inv[start + i], q);
        }
    }
    t.inv[n - 1] = 0; /* never read (ntt.cpp:216 skips this slot) */
    t.inv_h[n - 1] = 0;
    const u64 n_inv = q - ((q - 1) >> logn); /* ntt.cpp:75 */
    const u64 n_inv_h = orc_harvey_quotient(n_inv, q);
    for (size_t i = 0; i < n; i++) { /* ntt.cpp:78-85 */
        u64 s = orc_harvey_lazy(q, pw[i], n_inv, n_inv_h);
        s -= (s >= q) ? q : 0;
        t.inv[n + i] = s;
        t.inv_h[n + i] = orc_harvey_quotient(s, q);
    }
    free(pw);

    g_tables = realloc(g_tables, (g_ntables + 1) * sizeof(tables_t));
    g_tables[g_ntables] = t;
    return &g_tables[g_ntables++];
}

int orc_ntt_tables(unsigned logn, u64 q, u64 *fwd, u64 *fwd_h, u64 *inv, u64 *inv_h) {
    const tables_t *t = get_tables(logn, q);
    if (!t) return 1;
    size_t n = (size_t)1 << logn;
    if (fwd) memcpy(fwd, t->fwd, n * sizeof(u64));
    if (fwd_h) memcpy(fwd_h, t->fwd_h, n * sizeof(u64));
    if (inv) memcpy(inv, t->inv, 2 * n * sizeof(u64));
    if (inv_h) memcpy(inv_h, t->inv_h, 2 * n * sizeof(u64));
    return 0;
}

/* the sweep at ntt.cpp:171-175 / :214-218 */
static inline void approx_reduce(u64 q, size_t n, u64 *x) {
    const u64 k = (u64)(log2((double)q) + 0.5);
    const u64 fix = (q >= ((u64)1 << k)) ? 1 : 0;
    for (size_t i = 0; i < n; i++) x[i] -= ((x[i] >> k) - fix) * q;
}

/* the butterfly network shared by both directions: ntt.cpp:155-169 / :194-208.
 * `idx` is the running table index (1 for forward, 0 for inverse). */
static void ct_network(unsigned logn, u64 q, u64 *x, const u64 *w, const u64 *wh, size_t idx) {
    const size_t n = (size_t)1 << logn;
    size_t step = n;
    for (unsigned level = 1; level <= logn; level++, step >>= 1) {
        size_t gap = step / 2;
        for (size_t start = 0; start < n; start += step, idx++) {
            u64 z = w[idx], zh = wh[idx];
            for (size_t l = start; l < start + gap; l++) {
                size_t h = l + gap;
                u64 t = orc_harvey_lazy(q, x[h], z, zh);
                x[h] = x[l] + 2 * q - t;
                x[l] = x[l] + t;
            }
        }
    }
}

/* src/fhe/common/ntt.cpp:145-176 */
int orc_ntt_fwd_lazy(unsigned logn, u64 q, u64 *x) {
    const tables_t *t = get_tables(logn, q);
    if (!t) return 1;
    ct_network(logn, q, x, t->fwd, t->fwd_h, 1);
    approx_reduce(q, (size_t)1 << logn, x);
    return 0;
}

/* src/fhe/common/ntt.cpp:178-223 */
int orc_intt_lazy(unsigned logn, u64 q, u64 *x) {
    const tables_t *t = get_tables(logn, q);
    if (!t) return 1;
    const size_t n = (size_t)1 << logn;
    u64 *y = malloc(n * sizeof(u64));
    for (size_t i = 0; i < n; i++) y[i] = x[bitrev(i, logn)]; /* :185-189 */
    ct_network(logn, q, y, t->inv, t->inv_h, 0);              /* :191-208 */
    for (size_t i = 0; i < n; i++) x[i] = y[bitrev(i, logn)]; /* :210-212 */
    free(y);
    approx_reduce(q, n, x); /* :214-218 */
    for (size_t i = 0; i < n; i++) /* :219-221 */
        x[i] = orc_harvey_lazy(q, x[i], t->inv[n + i], t->inv_h[n + i]);
    return 0;
}

/* The same dataflow graph as orc_intt_lazy with the permutations folded into
 * the indexing (SURVEY Appendix A "Folded INTT"): stage s pairs (p, p + 2^{s-1}),
 * twiddle = inv[2^{s-1} - 1 + bitrev(p mod 2^{s-1}, s-1)]. */
int orc_intt_lazy_folded(unsigned logn, u64 q, u64 *x) {
    const tables_t *t = get_tables(logn, q);
    if (!t) return 1;
    const size_t n = (size_t)1 << logn;
    for (unsigned s = 1; s <= logn; s++) {
        size_t gap = (size_t)1 << (s - 1);
        for (size_t blk = 0; blk < n; blk += 2 * gap) {
            for (size_t j = 0; j < gap; j++) {
                size_t ti = gap - 1 + bitrev(j, s - 1);
                size_t lo = blk + j, hi = lo + gap;
                u64 v = orc_harvey_lazy(q, x[hi], t->inv[ti], t->inv_h[ti]);
                x[hi] = x[lo] + 2 * q - v;
                x[lo] = x[lo] + v;
            }
        }
    }
    approx_reduce(q, n, x);
    for (size_t i = 0; i < n; i++)
        x[i] = orc_harvey_lazy(q, x[i], t->inv[n + i], t->inv_h[n + i]);
    return 0;
}

/* ------------------------------------------------------------------ */
/* composite ops                                                       */
/* ------------------------------------------------------------------ */

/* src/fhe/common/ntt.h:41-51 */
/* timing helper for bench.py's cpu_baseline leg: `rows` single-limb transforms back to back */
int orc_bench_ntt(unsigned logn, u64 q, u64 *x, size_t rows, int forward) {
    const size_t n = (size_t)1 << logn;
    for (size_t r = 0; r < rows; r++) {
        int rc = forward ? orc_ntt_fwd_lazy(logn, q, x + r * n) : orc_intt_lazy(logn, q, x + r * n);
        if (rc) return rc;
    }
    return 0;
}

int orc_poly_ntt_fwd(unsigned logn, size_t L, const u64 *moduli, u64 *x) {
    const size_t n = (size_t)1 << logn;
    for (size_t k = 0; k < L; k++)
        if (orc_ntt_fwd_lazy(logn, moduli[k], x + k * n)) return 1;
    return 0;
}

/* src/fhe/common/ntt.h:72-92 */
int orc_poly_intt(unsigned logn, size_t L, const u64 *moduli, u64 *x, int strict) {
    const size_t n = (size_t)1 << logn;
    for (size_t k = 0; k < L; k++) {
        if (orc_intt_lazy(logn, moduli[k], x + k * n)) return 1;
        if (strict) orc_reduce_strict(moduli[k], n, x + k * n);
    }
    return 0;
}

/* src/fhe/ckks/arith.cpp:55-62 == src/fhe/bgv/arith.cpp:59-69; the products via
 * rns.cpp:120-140, the sum via rns.cpp:58-87. */
int orc_ckks_tensor(unsigned logn, size_t L, const u64 *moduli, const u64 *ct1,
                    const u64 *ct2, u64 *quad) {
    const size_t n = (size_t)1 << logn, pn = L * n;
    u64 *tmp = malloc(n * sizeof(u64));
    for (size_t k = 0; k < L; k++) {
        u64 q = moduli[k];
        const u64 *a0 = ct1 + k * n, *a1 = ct1 + pn + k * n;
        const u64 *b0 = ct2 + k * n, *b1 = ct2 + pn + k * n;
        orc_mul_hybrid_lazy(q, n, a0, b0, quad + k * n);
        orc_mul_hybrid_lazy(q, n, a0, b1, quad + pn + k * n);
        orc_mul_hybrid_lazy(q, n, a1, b0, tmp);
        orc_add_lazy(q, n, quad + pn + k * n, tmp);
        orc_mul_hybrid_lazy(q, n, a1, b1, quad + 2 * pn + k * n);
    }
    free(tmp);
    return 0;
}

/* src/fhe/primitives/rgsw.cpp:57-156.  in: [L][N] (NTT form), key:
 * [L][2][L+1][N], out: [2][L+1][N]. */
int orc_ext_prod(unsigned logn, size_t L, const u64 *ext_moduli, const u64 *in,
                 const u64 *key, u64 *out) {
    const size_t n = (size_t)1 << logn, E = L + 1;
    u64 *coef = malloc(L * n * sizeof(u64));
    u64 *dec = malloc(L * E * n * sizeof(u64)); /* [p][k][N] */
    u128 *acc = malloc(n * sizeof(u128));
    int rc = 0;

    memcpy(coef, in, L * n * sizeof(u64)); /* rgsw.cpp:103-105 */
    if (orc_poly_intt(logn, L, ext_moduli, coef, 1)) { rc = 1; goto done; }

    for (size_t p = 0; p < L; p++) { /* rgsw.cpp:99-101,108-119 */
        for (size_t k = 0; k < E; k++) {
            u64 *d = dec + (p * E + k) * n;
            if (k == p) {
                memcpy(d, in + p * n, n * sizeof(u64));
            } else {
                memcpy(d, coef + p * n, n * sizeof(u64));
                if (orc_ntt_fwd_lazy(logn, ext_moduli[k], d)) { rc = 1; goto done; }
            }
        }
    }
    for (size_t h = 0; h < 2; h++) { /* rgsw.cpp:126-153 */
        for (size_t k = 0; k < E; k++) {
            memset(acc, 0, n * sizeof(u128));
            for (size_t p = 0; p < L; p++) {
                const u64 *d = dec + (p * E + k) * n;
                const u64 *kk = key + ((p * 2 + h) * E + k) * n;
                for (size_t i = 0; i < n; i++) acc[i] += (u128)d[i] * kk[i];
            }
            orc_montgomery128_lazy(ext_moduli[k], n, (const u64 *)acc, out + (h * E + k) * n);
        }
    }
done:
    free(coef);
    free(dec);
    free(acc);
    return rc;
}

/* shared body of ckks/rescaling.cpp:31-77 and bgv/mod_switch.cpp:30-77.
 * t == 0 selects the CKKS variant. */
static int drop_last_prime(unsigned logn, size_t L, const u64 *moduli, u64 t,
                           const u64 *ct, u64 *out) {
    if (L < 2) return 2;
    const size_t n = (size_t)1 << logn;
    const u64 q_last = moduli[L - 1], half = q_last / 2;
    u64 *z = malloc(n * sizeof(u64));
    u64 *r = malloc(n * sizeof(u64));
    int rc = 0;
    for (size_t h = 0; h < 2; h++) {
        const u64 *poly = ct + h * L * n;
        u64 *dst = out + h * (L - 1) * n;
        memcpy(z, poly + (L - 1) * n, n * sizeof(u64));
        if (orc_intt_lazy(logn, q_last, z)) { rc = 1; break; }
        if (t) /* mod_switch.cpp:49: *= t^{-1} mod q_last */
            orc_mul_scalar_lazy(q_last, n, z, orc_inverse_mod_prime(t, q_last));
        orc_reduce_strict(q_last, n, z);
        for (size_t k = 0; k + 1 < L; k++) {
            u64 q = moduli[k], q_last_red = q_last % q;
            memcpy(r, z, n * sizeof(u64));
            orc_barrett(q, n, r);
            for (size_t i = 0; i < n; i++) /* centre: rescaling.cpp:63-68 */
                if (z[i] >= half) r[i] += q - q_last_red;
            if (t) orc_mul_scalar_lazy(q, n, r, t); /* mod_switch.cpp:70 */
            if (orc_ntt_fwd_lazy(logn, q, r)) { rc = 1; break; }
            memcpy(dst + k * n, poly + k * n, n * sizeof(u64));
            orc_sub_lazy(q, n, dst + k * n, r); /* rescaling.cpp:73 */
            /* rescaling.cpp:74: *= q_last^{-1} mod q_i (vector form reduces
             * the scalar mod q_i first; it already is) */
            orc_mul_scalar_lazy(q, n, dst + k * n, orc_inverse_mod_prime(q_last, q));
            if (t) orc_mul_scalar_lazy(q, n, dst + k * n, q_last % t); /* mod_switch.cpp:76 */
        }
        if (rc) break;
    }
    free(z);
    free(r);
    return rc;
}

int orc_ckks_rescale(unsigned logn, size_t L, const u64 *moduli, const u64 *ct, u64 *out) {
    return drop_last_prime(logn, L, moduli, 0, ct, out);
}

int orc_bgv_mod_switch(unsigned logn, size_t L, const u64 *moduli, u64 t, const u64 *ct,
                       u64 *out) {
    if (t == 0) return 2;
    return drop_last_prime(logn, L, moduli, t, ct, out);
}

/* src/fhe/ckks/arith.cpp:64-73 (t == 0) and src/fhe/bgv/arith.cpp:71-79 (t != 0;
 * note the reference calls mod_switch_inplace on a BgvCt whose plain_modulus
 * is still the default 1 — callers pass the t they want reproduced). */
static int relinearize(unsigned logn, size_t L, const u64 *ext_moduli, u64 t, const u64 *quad,
                       const u64 *key, u64 *out) {
    const size_t n = (size_t)1 << logn, E = L + 1;
    u64 *e = malloc(2 * E * n * sizeof(u64));
    int rc = orc_ext_prod(logn, L, ext_moduli, quad + 2 * L * n, key, e);
    if (!rc) rc = drop_last_prime(logn, E, ext_moduli, t, e, out);
    if (!rc)
        for (size_t h = 0; h < 2; h++)
            for (size_t k = 0; k < L; k++)
                orc_add_lazy(ext_moduli[k], n, out + (h * L + k) * n, quad + (h * L + k) * n);
    free(e);
    return rc;
}

int orc_ckks_relinearize(unsigned logn, size_t L, const u64 *ext_moduli, const u64 *quad,
                         const u64 *key, u64 *out) {
    return relinearize(logn, L, ext_moduli, 0, quad, key, out);
}

int orc_bgv_relinearize(unsigned logn, size_t L, const u64 *ext_moduli, u64 t,
                        const u64 *quad, const u64 *key, u64 *out) {
    if (t == 0) return 2;
    return relinearize(logn, L, ext_moduli, t, quad, key, out);
}

/* src/fhe/ckks/ckks.h:270-274 */
int orc_ckks_mult_relin(unsigned logn, size_t L, const u64 *ext_moduli, const u64 *ct1,
                        const u64 *ct2, const u64 *key, u64 *out) {
    const size_t n = (size_t)1 << logn;
    u64 *quad = malloc(3 * L * n * sizeof(u64));
    int rc = orc_ckks_tensor(logn, L, ext_moduli, ct1, ct2, quad);
    if (!rc) rc = orc_ckks_relinearize(logn, L, ext_moduli, quad, key, out);
    free(quad);
    return rc;
}

/* src/fhe/common/permutation.cpp:28-60.  3^i (mod 2^32) masks to 3^i mod 2N. */
int orc_galois_cycle(unsigned logn, size_t L, const u64 *in, u64 *out, size_t step) {
    if (logn < 1 || logn > 16 || step >= ((size_t)1 << 17)) return 1;
    const size_t n = (size_t)1 << logn;
    const uint32_t mask = ((uint32_t)1 << (logn + 1)) - 1;
    uint32_t factor = 1;
    for (size_t i = 0; i < step; i++) factor *= 3;
    factor &= mask;
    uint32_t g = 1;
    for (size_t i = 0; i < n / 2; i++, g *= 3) {
        uint32_t old_idx = g & mask;
        size_t from = bitrev((old_idx - 1) / 2, logn);
        uint32_t new_idx = (old_idx * factor) & mask;
        size_t to = bitrev((new_idx - 1) / 2, logn);
        for (size_t k = 0; k < L; k++) {
            out[k * n + to] = in[k * n + from];
            out[k * n + n - 1 - to] = in[k * n + n - 1 - from];
        }
    }
    return 0;
}

/* src/fhe/common/permutation.cpp:62-75 */
int orc_galois_involution(unsigned logn, size_t L, const u64 *in, u64 *out) {
    const size_t n = (size_t)1 << logn;
    for (size_t k = 0; k < L; k++)
        for (size_t i = 0; i < n; i++) out[k * n + i] = in[k * n + n - 1 - i];
    return 0;
}

/* shared body of ckks::rotate / ckks::conjugate (ckks/arith.cpp:75-93):
 * permute both polys, key-switch the permuted c1, drop P, add permuted c0. */
static int galois_keyswitch(unsigned logn, size_t L, const u64 *ext_moduli, const u64 *ct,
                            const u64 *key, int conj, size_t step, u64 *out) {
    const size_t n = (size_t)1 << logn, E = L + 1;
    u64 *perm = malloc(2 * L * n * sizeof(u64));
    u64 *e = malloc(2 * E * n * sizeof(u64));
    int rc = 0;
    for (size_t h = 0; h < 2 && !rc; h++)
        rc = conj ? orc_galois_involution(logn, L, ct + h * L * n, perm + h * L * n)
                  : orc_galois_cycle(logn, L, ct + h * L * n, perm + h * L * n, step);
    if (!rc) rc = orc_ext_prod(logn, L, ext_moduli, perm + L * n, key, e);
    if (!rc) rc = drop_last_prime(logn, E, ext_moduli, 0, e, out);
    if (!rc)
        for (size_t k = 0; k < L; k++)
            orc_add_lazy(ext_moduli[k], n, out + k * n, perm + k * n);
    free(perm);
    free(e);
    return rc;
}

int orc_ckks_rotate(unsigned logn, size_t L, const u64 *ext_moduli, const u64 *ct,
                    const u64 *key, size_t step, u64 *out) {
    return galois_keyswitch(logn, L, ext_moduli, ct, key, 0, step, out);
}

int orc_ckks_conjugate(unsigned logn, size_t L, const u64 *ext_moduli, const u64 *ct,
                       const u64 *key, u64 *out) {
    return galois_keyswitch(logn, L, ext_moduli, ct, key, 1, 0, out);
}

/* ------------------------------------------------------------------ */
/* RLWE encrypt / decrypt cores                   */
/* ------------------------------------------------------------------ */

/* decrypt_core — src/fhe/primitives/rlwe.cpp:63-71: pt = c0 + c1 * sk (hybrid mulmod, lazy add),
 * INTT, reduce_strict.  ct: [2][L][N] NTT form; sk: [L][N] NTT form; pt: [L][N] coefficients < q. */
int orc_rlwe_decrypt_core(unsigned logn, size_t L, const u64 *moduli, const u64 *ct, const u64 *sk, u64 *pt) {
    const size_t n = (size_t)1 << logn;
    u64 *prod = malloc(n * sizeof(u64));
    for (size_t k = 0; k < L; k++) {
        const u64 q = moduli[k];
        orc_mul_hybrid_lazy(q, n, ct + (L + k) * n, sk + k * n, prod); /* c1 * sk, rns.cpp:120-140 */
        memcpy(pt + k * n, ct + k * n, n * sizeof(u64));               /* c0 + ..., rns.cpp:58-86   */
        orc_add_lazy(q, n, pt + k * n, prod);
    }
    free(prod);
    return orc_poly_intt(logn, L, moduli, pt, 1);
}

/* encrypt_core with the samples supplied by the caller — rlwe.cpp:34-61 and sampling.cpp:47-69:
 * ex = NTT(e) for the error coefficients e (already reduced mod each q), c0 = ex - c1 * sk,
 * c0 += NTT(pt).  pt, e: [L][N] coefficient form; c1: [L][N] uniform NTT-form mask; out: [2][L][N]. */
int orc_rlwe_encrypt_core(unsigned logn, size_t L, const u64 *moduli, const u64 *pt, const u64 *sk, const u64 *c1,
                          const u64 *e, u64 *out) {
    const size_t n = (size_t)1 << logn;
    u64 *prod = malloc(n * sizeof(u64));
    u64 *ptn = malloc(L * n * sizeof(u64));
    int rc = 0;
    memcpy(out, e, L * n * sizeof(u64));
    memcpy(ptn, pt, L * n * sizeof(u64));
    if (orc_poly_ntt_fwd(logn, L, moduli, out)) { rc = 1; goto done; }  /* sampling.cpp:66 */
    if (orc_poly_ntt_fwd(logn, L, moduli, ptn)) { rc = 1; goto done; }  /* rlwe.cpp:54-55   */
    for (size_t k = 0; k < L; k++) {
        const u64 q = moduli[k];
        orc_mul_hybrid_lazy(q, n, c1 + k * n, sk + k * n, prod);
        orc_sub_lazy(q, n, out + k * n, prod);                          /* rlwe.cpp:50      */
        orc_add_lazy(q, n, out + k * n, ptn + k * n);                   /* rlwe.cpp:58      */
    }
    memcpy(out + L * n, c1, L * n * sizeof(u64));
done:
    free(prod);
    free(ptn);
    return rc;
}

/* ------------------------------------------------------------------ */
/* RNS base transform and key-switch key generation */
/* ------------------------------------------------------------------ */

/* rns_base_transform, one modulus -> many — src/fhe/common/rns_transform.cpp:11-37 (after the strict
 * reduction at :116).  in: [N] coefficients (any lazy value); out: [Lnew][N]. */
int orc_base_transform_from_single(u64 q_old, size_t n, const u64 *in, const u64 *new_moduli, size_t Lnew, u64 *out) {
    const u64 half = q_old / 2;
    for (size_t k = 0; k < Lnew; k++) {
        const u64 q = new_moduli[k], multiple = (q_old / q + 1) * q;
        for (size_t i = 0; i < n; i++) {
            u64 x = in[i];
            x -= (x >= q_old) ? q_old : 0; /* reduce_strict of the by-value argument, :116 */
            out[k * n + i] = (x < half) ? x : multiple - q_old + x;
        }
        if (q < q_old) orc_barrett_lazy(q, n, out + k * n);
    }
    return 0;
}

/* rns_base_transform, many -> one modulus, small-coefficient path — rns_transform.cpp:39-84.
 * Returns 3 when the coefficients are not all small (the reference then composes big integers). */
int orc_base_transform_to_single(size_t n, size_t L, const u64 *old_moduli, const u64 *in, u64 new_modulus, u64 *out) {
    u64 *x = malloc(L * n * sizeof(u64));
    memcpy(x, in, L * n * sizeof(u64));
    for (size_t k = 0; k < L; k++) orc_reduce_strict(old_moduli[k], n, x + k * n); /* :116 */
    const u64 q0 = old_moduli[0], half = q0 / 2;
    int small = 1;
    for (size_t i = 0; i < n && small; i++)
        for (size_t k = 1; k < L; k++) {
            if (x[i] < half && x[k * n + i] != x[i]) small = 0;
            if (x[i] >= half && old_moduli[k] - x[k * n + i] != q0 - x[i]) small = 0;
        }
    if (!small) {
        free(x);
        return 3;
    }
    const u64 multiple = (q0 / new_modulus + 1) * new_modulus;
    for (size_t i = 0; i < n; i++) out[i] = (x[i] < half) ? x[i] : multiple - q0 + x[i];
    orc_barrett(new_modulus, n, out);
    free(x);
    return 0;
}

/* RlweKsk::RlweKsk — src/fhe/primitives/keys.cpp:8-36, with rgsw_encrypt_montgomery (rgsw.cpp:11-55)
 * and get_rlwe_sample (rlwe.cpp:34-51) on caller-supplied samples.
 *   sk_curr, sk_orig: [L][N] NTT form; masks: [L][L+1][N] uniform NTT-form c1 per row;
 *   errors: [L][L+1][N] error coefficients per row; key: [L][2][L+1][N]. */
int orc_ksk_generate(unsigned logn, size_t L, const u64 *ext_moduli, const u64 *sk_curr, const u64 *sk_orig,
                     const u64 *masks, const u64 *errors, u64 *key) {
    const size_t n = (size_t)1 << logn, L1 = L + 1;
    const u64 P = ext_moduli[L];
    int rc = 0;
    u64 *sk_ext = malloc(L1 * n * sizeof(u64)); /* sk_orig_extended */
    u64 *prod = malloc(n * sizeof(u64));
    u64 *term = malloc(n * sizeof(u64));
    memcpy(sk_ext, sk_orig, L * n * sizeof(u64));
    if (orc_poly_intt(logn, L, ext_moduli, sk_ext, 0)) { rc = 1; goto done; }                       /* keys.cpp:22 */
    /* keys.cpp:23-25 calls rns_base_transform, which hands a single-component input to the one -> many transform
     * (rns_transform.cpp:118-121): lazy Barrett only when P < q_0, no strict reduction */
    if (L == 1) {
        orc_reduce_strict(ext_moduli[0], n, sk_ext);                                                 /* :116 */
        orc_base_transform_from_single(ext_moduli[0], n, sk_ext, &P, 1, sk_ext + n);
    } else {
        rc = orc_base_transform_to_single(n, L, ext_moduli, sk_ext, P, sk_ext + L * n);
    }
    if (rc) goto done;
    if (orc_poly_ntt_fwd(logn, L1, ext_moduli, sk_ext)) { rc = 1; goto done; }                     /* :26        */
    for (size_t p = 0; p < L; p++) {
        u64 *c0 = key + (p * 2 + 0) * L1 * n, *c1 = key + (p * 2 + 1) * L1 * n;
        memcpy(c0, errors + p * L1 * n, L1 * n * sizeof(u64));
        memcpy(c1, masks + p * L1 * n, L1 * n * sizeof(u64));
        if (orc_poly_ntt_fwd(logn, L1, ext_moduli, c0)) { rc = 1; goto done; }                     /* sampling.cpp:66 */
        for (size_t k = 0; k < L1; k++) {
            const u64 q = ext_moduli[k];
            orc_mul_hybrid_lazy(q, n, c1 + k * n, sk_ext + k * n, prod);
            orc_sub_lazy(q, n, c0 + k * n, prod);                                                    /* rlwe.cpp:50 */
            /* pt_ntt * basis_p (rgsw.cpp:27, rns.cpp:155-171): basis_p[k] = P mod q_p at k == p, else 0.  The
             * extra limb of sk_curr_extended is uninitialised in the reference but multiplied by 0. */
            if (k < L) memcpy(term, sk_curr + k * n, n * sizeof(u64));
            else memset(term, 0, n * sizeof(u64));
            orc_mul_scalar_lazy(q, n, term, (k == p) ? P % q : 0);
            orc_add_lazy(q, n, c0 + k * n, term);
            const u64 r = ((u64)(-1LL) % q) + 1;                                                     /* rgsw.cpp:36-39 */
            orc_mul_scalar_lazy(q, n, c0 + k * n, r);                                               /* rgsw.cpp:47-51 */
            orc_mul_scalar_lazy(q, n, c1 + k * n, r);
        }
    }
done:
    free(sk_ext);
    free(prod);
    free(term);
    return rc;
}

/* ------------------------------------------------------------------ */
/* harness helpers                                                     */
/* ------------------------------------------------------------------ */

/* SURVEY Appendix B input generator */
void orc_lcg_fill(u64 seed, u64 q, size_t n, u64 *x) {
    u64 s = seed;
    for (size_t i = 0; i < n; i++) {
        s = s * 6364136223846793005ULL + 1442695040888963407ULL;
        x[i] = s % q;
    }
}

u64 orc_fnv1a(const u64 *x, size_t n, u64 h) {
    if (h == 0) h = 1469598103934665603ULL;
    for (size_t i = 0; i < n; i++) {
        h ^= x[i];
        h *= 1099511628211ULL;
    }
    return h;
}

static int is_prime_u64(u64 n) {
    static const u64 bases[] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37};
    if (n < 2) return 0;
    for (size_t i = 0; i < 12; i++) {
        if (n == bases[i]) return 1;
        if (n % bases[i] == 0) return 0;
    }
    u64 d = n - 1;
    unsigned r = 0;
    while (!(d & 1)) { d >>= 1; r++; }
    for (size_t i = 0; i < 12; i++) {
        u64 x = orc_pow_mod(n, bases[i], d);
        if (x == 1 || x == n - 1) continue;
        int comp = 1;
        for (unsigned j = 1; j < r; j++) {
            x = mulmod(x, x, n);
            if (x == n - 1) { comp = 0; break; }
        }
        if (comp) return 0;
    }
    return 1;
}

/* src/fhe/common/primelists.cpp:5-192 restated as its generating rule.  (The
 * table's two typo entries and the short row 45 are data errors of the
 * reference, not part of the rule; see DESIGN.md.) */
int orc_prime_row(unsigned bits, size_t count, u64 *out) {
    if (bits < 17 || bits > 62) return 0;
    u64 top = (u64)1 << bits;
    u64 c = top - 65536 + 1; /* largest value = 1 (mod 2^16) below 2^bits */
    size_t found = 0;
    while (found < count && c > ((u64)1 << (bits - 1))) {
        if (is_prime_u64(c)) out[found++] = c;
        c -= 65536;
    }
    return (int)found;
}

/* src/fhe/ckks/basics.cpp:14-38: the additional modulus is drawn first, then the
 * chain in order, each from its bit-size row with a per-row cursor. */
int orc_ckks_pick_moduli(const unsigned *moduli_bits, size_t L, unsigned additional_bits,
                         u64 *moduli_out, u64 *additional_out) {
    size_t cursor[64] = {0};
    u64 row[20];
    if (additional_bits >= 64) return 1;
    if (orc_prime_row(additional_bits, 20, row) < 1) return 1;
    *additional_out = row[cursor[additional_bits]++];
    for (size_t k = 0; k < L; k++) {
        unsigned b = moduli_bits[k];
        if (b >= 64) return 1;
        int have = orc_prime_row(b, 20, row);
        if ((size_t)have <= cursor[b]) return 1;
        moduli_out[k] = row[cursor[b]++];
    }
    return 0;
}
