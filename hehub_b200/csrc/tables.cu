// tables.cu — twiddle-table construction and the host-side context caches.
//
// Table *values* follow the reference's NTTFactors (src/fhe/common/ntt.cpp:41-105) exactly:
// psi from the least generator candidate (ntt.cpp:26-39), forward T[i] = psi^{bitrev(i)},
// inverse per-level U[2^l - 1 + i] = psi^{-bitrev(i,l) * N/2^l}, scale S[i] = strict(psi^{-i}/N),
// each with its Harvey companion floor(w * 2^64 / q).  Table *layout* is ours: besides the
// reference order (used by the generic kernels) each table is stored pass-major /
// slot-major (ntt_plan.h) so that a warp's lanes read consecutive 16-byte entries.
#include <cmath>
#include <cstring>

#include "context.h"

namespace hb {

typedef unsigned __int128 u128;

static inline u64 mulmod(u64 a, u64 b, u64 q) { return (u64)((u128)a * b % q); }

u64 host_pow_mod(u64 q, u64 base, u64 e) {
    u64 r = 1 % q;
    base %= q;
    while (e) {
        if (e & 1) r = mulmod(r, base, q);
        base = mulmod(base, base, q);
        e >>= 1;
    }
    return r;
}

// ntt.cpp:26-39
u64 host_root_2n(u64 q, u64 n) {
    if (n == 0 || q < 3 || (q - 1) % (2 * n) != 0) return 0;
    u64 g = 2;
    for (;; g++) {
        if (g >= q) return 0; // q is not prime / has no such element
        if (host_pow_mod(q, g, (q - 1) / 2) == q - 1) break;
    }
    return host_pow_mod(q, g, (q - 1) / (2 * n));
}

// mod_arith.cpp:138-149 (canonical inverse; Fermat instead of xgcd)
u64 host_inverse_mod_prime(u64 elem, u64 prime) {
    if (prime <= 1) return 0;
    return host_pow_mod(prime, elem % prime, prime - 2);
}

static inline unsigned bitrev(unsigned x, unsigned bits) {
    unsigned r = 0;
    for (unsigned i = 0; i < bits; i++) r |= ((x >> i) & 1u) << (bits - 1 - i);
    return r;
}

static void fill_consts(LimbConst &lc, u64 q) {
    std::memset(&lc, 0, sizeof(lc));
    lc.q = q;
    lc.q2 = 2 * q;
    lc.nq = (u64)0 - q;
    if (q & 1) { // mod_arith.cpp:49-62
        u64 inv = q;
        for (int i = 0; i < 6; i++) inv *= 2 - q * inv;
        lc.minus_qinv = (u64)0 - inv;
    }
    lc.r = ((u64)(-1LL) % q) + 1;
    lc.r_h = (u64)(((u128)lc.r << 64) / q);
    lc.barrett_c = (u64)(-1) / q;
    lc.logq = (u32)(u64)(std::log2((double)q) + 0.5); // ntt.cpp:171
    lc.fix = (lc.logq < 64 && q >= ((u64)1 << lc.logq)) ? 1u : 0u;
}

Context::~Context() {
    for (auto &kv : tables)
        if (kv.second.dev_block) cudaFree(kv.second.dev_block);
    for (auto &kv : chains) cudaFree(kv.second);
    for (auto &kv : drops) cudaFree(kv.second.dev);
    for (auto &kv : scalar_sets) cudaFree(kv.second);
    for (auto &kv : slab_free)
        for (u64 *p : kv.second) cudaFree(p);
    for (auto &kv : slab_live) cudaFree(kv.first);
    for (auto &s : scratch)
        if (s.first) cudaFree(s.first);
    if (deferred_flags) cudaFreeHost(deferred_flags);
    if (grid_barrier_dev) cudaFree(grid_barrier_dev);
    if (pipe_ready) {
        for (int i = 0; i < kPipeSlots; i++) {
            cudaEventDestroy(ev_in[i]);
            cudaEventDestroy(ev_k[i]);
            cudaEventDestroy(ev_out[i]);
        }
        cudaStreamDestroy(s_in);
        cudaStreamDestroy(s_out);
    }
    if (owns_stream && stream) cudaStreamDestroy(stream);
}

unsigned long long *Context::grid_barrier_counter() {
    if (!grid_barrier_dev) {
        // [0]: the single-launch kernel's grid barrier (only grows); [1], [2]: arrivals / released clusters of ext_mac_intt_kernel
        // (reset by the kernel itself: nothing on the host to keep in step, replays of a captured graph included)
        if (cudaMalloc(&grid_barrier_dev, 4 * sizeof(unsigned long long)) != cudaSuccess) {
            cudaGetLastError();
            return nullptr;
        }
        cudaMemsetAsync(grid_barrier_dev, 0, 4 * sizeof(unsigned long long), stream);
        grid_barrier_count = 0;
    }
    return grid_barrier_dev;
}

int *Context::deferred_flag_slot(int *err) {
    *err = 0;
    if (!deferred_flags) {
        void *p = nullptr;
        cudaError_t e = cudaMallocHost(&p, kDeferredSlots * sizeof(int));
        if (e != cudaSuccess) {
            *err = cuda_fail(e, "cudaMallocHost(deferred verdicts)");
            return nullptr;
        }
        deferred_flags = static_cast<int *>(p);
        for (int i = 0; i < kDeferredSlots; i++) deferred_flags[i] = 0;
    }
    if (deferred_used == kDeferredSlots) { // all slots pending: drain them (rare: 64 key generations without a synchronize)
        cudaStreamSynchronize(stream);
        int raised = 0;
        for (int i = 0; i < kDeferredSlots; i++) raised |= deferred_flags[i];
        for (int i = 0; i < kDeferredSlots; i++) deferred_flags[i] = 0;
        deferred_used = 0;
        if (raised) deferred_flags[deferred_used++] = 1; // keep the verdict for the caller's next synchronize
    }
    return deferred_flags + deferred_used++;
}

int Context::take_deferred() {
    int raised = 0;
    for (int i = 0; i < deferred_used; i++) {
        raised |= deferred_flags[i];
        deferred_flags[i] = 0;
    }
    deferred_used = 0;
    return raised;
}

int Context::cuda_fail(cudaError_t e, const char *where) {
    last_error = std::string(where) + ": " + cudaGetErrorString(e);
    return 2;
}

const ModTables *Context::get_tables(u64 q, unsigned logn, int *err) {
    *err = 0;
    auto key = std::make_pair(q, logn);
    auto it = tables.find(key);
    if (it != tables.end()) return &it->second;
    if (q < 2) {
        *err = fail(1, "modulus must be >= 2");
        return nullptr;
    }

    ModTables mt;
    mt.dev_block = nullptr;
    fill_consts(mt.lc, q);
    if (logn == 0) { // constants only (coefficient-wise kernels)
        return &tables.emplace(key, mt).first->second;
    }
    if (mt.lc.logq > 59) { // ntt.cpp:43-47
        *err = fail(1, "NTT not supporting primes with bit size > 59 currently.");
        return nullptr;
    }
    if (logn > (unsigned)kFastLogMax) {
        *err = fail(3, "ring dimension above 2^15 is not supported by the device kernels");
        return nullptr;
    }
    const size_t n = (size_t)1 << logn;
    const u64 psi = host_root_2n(q, n);
    if (psi == 0) { // ntt.cpp:27-29
        *err = fail(1, "2N doesn't divide (modulus - 1)");
        return nullptr;
    }

    const bool fast = logn >= (unsigned)kFastLogMin;
    const bool lat = fast && has_latency_plan((int)logn);
    const bool lat2 = lat && has_latency2_plan((int)logn);
    // fwd_nat, inv_nat, inv_scale [, fwd fast, inv fast [, latency-plan pair [, mode-2 pair (staged layout, 2 n entries each)]]]
    const size_t ntab = fast ? (lat ? (lat2 ? 11 : 7) : 5) : 3;
    std::vector<ulonglong2> host(ntab * n, make_ulonglong2(0, 0));
    ulonglong2 *fwd_nat = host.data(), *inv_nat = fwd_nat + n, *inv_scale = inv_nat + n;
    auto pair_of = [&](u64 w) { return make_ulonglong2(w, host_harvey_quotient(w, q)); };

    std::vector<u64> pw(n);
    pw[0] = 1;
    for (size_t k = 1; k < n; k++) pw[k] = mulmod(pw[k - 1], psi, q);
    for (size_t i = 0; i < n; i++) fwd_nat[i] = pair_of(pw[bitrev((unsigned)i, logn)]); // ntt.cpp:54-58

    const u64 psi_inv = host_pow_mod(q, psi, 2 * n - 1); // ntt.cpp:62-63
    pw[0] = 1;
    for (size_t k = 1; k < n; k++) pw[k] = mulmod(pw[k - 1], psi_inv, q);
    for (unsigned l = 0; l < logn; l++) // ntt.cpp:64-74
        for (size_t i = 0; i < ((size_t)1 << l); i++)
            inv_nat[((size_t)1 << l) - 1 + i] = pair_of(pw[(size_t)bitrev((unsigned)i, l) << (logn - l)]);
    const u64 n_inv = q - ((q - 1) >> logn); // ntt.cpp:75
    for (size_t i = 0; i < n; i++)           // ntt.cpp:78-85 (canonical residue of psi^{-i} / N)
        inv_scale[i] = pair_of(mulmod(pw[i], n_inv, q));

    // pass-major / slot-major layouts of one plan pair (forward, inverse) — see ntt_plan.h
    auto fill_plan_tables = [&](ulonglong2 *ff, ulonglong2 *fi, int mode) {
        NttPlan pl = plan_for((int)logn, true, mode);
        if (pl.lpre && !pl.xchg) ff[0] = fwd_nat[1];
        for (int p = 0; p < pl.npass; p++) {
            const int K = pl.k[p], l0g = fwd_glevel0(pl, p), off = fwd_pass_offset(pl, p);
            for (int m = 1; m <= K; m++)
                for (int blk = 0; blk < (1 << (m - 1)); blk++) {
                    const int slot = (1 << (m - 1)) - 1 + blk;
                    for (int hb = 0; hb < (1 << l0g); hb++)
                        ff[off + (slot << l0g) + hb] = fwd_nat[((size_t)1 << (l0g + m - 1)) + ((size_t)hb << (m - 1)) + blk];
                }
        }
        pl = plan_for((int)logn, false, mode);
        for (int p = 0; p <= pl.npass; p++) {
            if (p == pl.npass && (!pl.lpre || pl.xchg)) break;
            const int K = (p == pl.npass) ? 1 : inv_k(pl, p);
            const int S0 = (p == pl.npass) ? (int)logn - 1 : inv_s0(pl, p), off = inv_pass_offset(pl, p);
            for (int m = 1; m <= K; m++)
                for (int jj = 0; jj < (1 << (m - 1)); jj++) {
                    const int slot = (1 << (m - 1)) - 1 + jj, s = S0 + m;
                    for (int lo = 0; lo < (1 << S0); lo++) {
                        const unsigned pos = ((unsigned)jj << S0) + (unsigned)lo;
                        fi[off + (slot << S0) + lo] = inv_nat[((size_t)1 << (s - 1)) - 1 + bitrev(pos, s - 1)];
                    }
                }
        }
    };
    if (fast) fill_plan_tables(inv_scale + n, inv_scale + 2 * n, 0);
    if (lat) fill_plan_tables(inv_scale + 3 * n, inv_scale + 4 * n, 1);
    if (lat2) { // per-CTA staged layout, see ntt_plan.h
        ulonglong2 *ff = inv_scale + 5 * n, *fi = inv_scale + 7 * n;
        NttPlan pl = plan_for((int)logn, true, 2);
        const int C = 1 << pl.lpre;
        if ((size_t)fwd_stage_total(pl) > 2 * n || (size_t)inv_stage_total(plan_for((int)logn, false, 2)) > 2 * n) {
            *err = fail(3, "internal: staged tables do not fit their allocation");
            return nullptr;
        }
        for (int B = 0; B < C; B++) {
            ulonglong2 *blk = ff + (size_t)B * fwd_stage_block(pl);
            for (int p = 0; p < pl.npass; p++) {
                const int K = pl.k[p], l0g = fwd_glevel0(pl, p), off = fwd_stage_offset(pl, p);
                const int lam = p == 0 ? 0 : fwd_lambda0(pl, p); // CTA-local levels done: 2^lam blocks per CTA
                for (int m = 1; m <= K; m++)
                    for (int bk = 0; bk < (1 << (m - 1)); bk++) {
                        const int slot = (1 << (m - 1)) - 1 + bk;
                        if (p == 0) { // the cross pass: group index below the smallest gap, one twiddle per slot
                            blk[off + slot] = fwd_nat[((size_t)1 << (m - 1)) + bk];
                        } else {
                            for (int hb = 0; hb < (1 << lam); hb++) {
                                const size_t ghb = ((size_t)B << lam) + hb; // block index in the full row
                                blk[off + (slot << lam) + hb] = fwd_nat[((size_t)1 << (l0g + m - 1)) + (ghb << (m - 1)) + bk];
                            }
                        }
                    }
            }
        }
        pl = plan_for((int)logn, false, 2);
        for (int p = 0; p + 1 < pl.npass; p++) { // the local passes: same entries as the mode-independent layout
            const int K = inv_k(pl, p), S0 = inv_s0(pl, p), off = inv_pass_offset(pl, p);
            for (int m = 1; m <= K; m++)
                for (int jj = 0; jj < (1 << (m - 1)); jj++) {
                    const int slot = (1 << (m - 1)) - 1 + jj, s = S0 + m;
                    for (int lo = 0; lo < (1 << S0); lo++) {
                        const unsigned pos = ((unsigned)jj << S0) + (unsigned)lo;
                        fi[off + (slot << S0) + lo] = inv_nat[((size_t)1 << (s - 1)) - 1 + bitrev(pos, s - 1)];
                    }
                }
        }
        {
            const int K = pl.k[0], LOGG = (int)logn - K, T = pl.threads, S0 = LOGG;
            for (int B = 0; B < C; B++) {
                ulonglong2 *blk = fi + inv_stage_local(pl) + (size_t)B * inv_stage_cross_block(pl);
                for (int tid = 0; tid < T; tid++) {
                    const int t = B * T + tid;
                    for (int m = 1; m <= K; m++)
                        for (int jj = 0; jj < (1 << (m - 1)); jj++) {
                            const int slot = (1 << (m - 1)) - 1 + jj, s = S0 + m;
                            const unsigned pos = ((unsigned)jj << S0) + (unsigned)t;
                            blk[slot * T + tid] = inv_nat[((size_t)1 << (s - 1)) - 1 + bitrev(pos, s - 1)];
                        }
                    for (int j = 0; j < (1 << K); j++) blk[((1 << K) - 1 + j) * T + tid] = inv_scale[t + ((size_t)j << LOGG)];
                }
            }
        }
    }

    void *dev = nullptr;
    cudaError_t e = cudaMalloc(&dev, host.size() * sizeof(ulonglong2));
    if (e != cudaSuccess) {
        *err = cuda_fail(e, "cudaMalloc(tables)");
        return nullptr;
    }
    e = cudaMemcpyAsync(dev, host.data(), host.size() * sizeof(ulonglong2), cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream); // host vector dies at scope exit
    if (e != cudaSuccess) {
        cudaFree(dev);
        *err = cuda_fail(e, "upload tables");
        return nullptr;
    }
    mt.dev_block = dev;
    const ulonglong2 *d = static_cast<const ulonglong2 *>(dev);
    mt.lc.fwd_nat = d;
    mt.lc.inv_nat = d + n;
    mt.lc.inv_scale = d + 2 * n;
    mt.lc.fwd = fast ? d + 3 * n : nullptr;
    mt.lc.inv = fast ? d + 4 * n : nullptr;
    mt.lc.fwd_lat = lat ? d + 5 * n : nullptr;
    mt.lc.inv_lat = lat ? d + 6 * n : nullptr;
    mt.lc.fwd_lat2 = lat2 ? d + 7 * n : nullptr;
    mt.lc.inv_lat2 = lat2 ? d + 9 * n : nullptr;
    return &tables.emplace(key, mt).first->second;
}

const LimbConst *Context::get_chain(unsigned logn, const u64 *moduli, size_t L, int *err) {
    *err = 0;
    std::vector<u64> key;
    key.reserve(L + 1);
    key.push_back(logn);
    key.insert(key.end(), moduli, moduli + L);
    auto it = chains.find(key);
    if (it != chains.end()) return it->second;
    std::vector<LimbConst> host(L);
    for (size_t k = 0; k < L; k++) {
        const ModTables *mt = get_tables(moduli[k], logn, err);
        if (!mt) return nullptr;
        host[k] = mt->lc;
    }
    LimbConst *dev = nullptr;
    cudaError_t e = cudaMalloc(&dev, L * sizeof(LimbConst));
    if (e != cudaSuccess) {
        *err = cuda_fail(e, "cudaMalloc(chain)");
        return nullptr;
    }
    e = cudaMemcpyAsync(dev, host.data(), L * sizeof(LimbConst), cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) {
        cudaFree(dev);
        *err = cuda_fail(e, "upload chain");
        return nullptr;
    }
    chains.emplace(std::move(key), dev);
    return dev;
}

// rescaling.cpp:31-44 / mod_switch.cpp:30-44.  t == 0 selects the CKKS variant.
const DropSet *Context::get_drop(unsigned logn, const u64 *moduli, size_t L, u64 t, int *err) {
    *err = 0;
    std::vector<u64> key;
    key.push_back(logn);
    key.push_back(t);
    key.insert(key.end(), moduli, moduli + L);
    auto it = drops.find(key);
    if (it != drops.end()) return &it->second;
    const u64 q_last = moduli[L - 1];
    std::vector<DropConst> host(L - 1);
    for (size_t k = 0; k + 1 < L; k++) {
        const u64 q = moduli[k];
        DropConst &d = host[k];
        d.qlast_mod_q = q_last % q;
        d.inv_qlast = host_inverse_mod_prime(q_last, q) % q; // operator*=(vector) reduces mod q_i, rns.cpp:163
        d.inv_qlast_h = host_harvey_quotient(d.inv_qlast, q);
        d.t_mod_q = t ? t % q : 0; // operator*=(u64), rns.cpp:145-146
        d.t_mod_q_h = host_harvey_quotient(d.t_mod_q, q);
        d.qlt_mod_q = t ? (q_last % t) % q : 0;
        d.qlt_mod_q_h = host_harvey_quotient(d.qlt_mod_q, q);
        d.z_below_q = q_last <= q ? 1 : 0;
    }
    DropSet ds;
    ds.half_qlast = q_last / 2;
    ds.inv_t = t ? host_inverse_mod_prime(t, q_last) % q_last : 0;
    ds.inv_t_h = host_harvey_quotient(ds.inv_t, q_last);
    ds.dev = nullptr;
    cudaError_t e = cudaMalloc(&ds.dev, (L - 1) * sizeof(DropConst));
    if (e != cudaSuccess) {
        *err = cuda_fail(e, "cudaMalloc(drop)");
        return nullptr;
    }
    e = cudaMemcpyAsync(ds.dev, host.data(), (L - 1) * sizeof(DropConst), cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) {
        cudaFree(ds.dev);
        *err = cuda_fail(e, "upload drop constants");
        return nullptr;
    }
    return &drops.emplace(std::move(key), ds).first->second;
}

u64 *Context::get_scratch(size_t slot, size_t words, int *err) {
    *err = 0;
    if (scratch.size() <= slot) scratch.resize(slot + 1, {nullptr, 0});
    auto &s = scratch[slot];
    if (s.second >= words) return s.first;
    if (s.first) {
        // earlier launches on the stream may still read the old buffer
        cudaStreamSynchronize(stream);
        cudaFree(s.first);
        s = {nullptr, 0};
    }
    u64 *p = nullptr;
    cudaError_t e = cudaMalloc(&p, words * sizeof(u64));
    if (e != cudaSuccess) {
        *err = cuda_fail(e, "cudaMalloc(scratch)");
        return nullptr;
    }
    s = {p, words};
    return p;
}

} // namespace hb
