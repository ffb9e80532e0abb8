// context.h — host-side state behind the opaque hehub_b200_ctx handle.
//
// Holds the CUDA stream, the per-(modulus, ring size) twiddle tables (the device analogue of
// the reference's global NTTFactors caches, src/fhe/common/ntt.cpp:107-143), per-chain
// LimbConst arrays, a pooled slab allocator (device analogue of FixedBlockAllocator,
// src/fhe/common/allocator.h:12-103) and grow-only scratch workspaces.
#pragma once
#include "compat.h"

#include <cstdint>
#include <map>
#include <string>
#include <utility>
#include <vector>

#include "modarith.cuh"
#include "ntt_engine.cuh"

namespace hb {

struct ModTables {
    LimbConst lc;     // host copy; pointer members are device addresses
    void *dev_block;  // one allocation holding every table of this (q, logn)
};

// extra per-limb constants of rescale / mod-switch (rescaling.cpp:36-44, mod_switch.cpp:36-44)
struct DropConst {
    u64 qlast_mod_q;        // q_last mod q_i
    u64 inv_qlast, inv_qlast_h; // q_last^{-1} mod q_i and Harvey companion
    u64 t_mod_q, t_mod_q_h;     // BGV: t mod q_i
    u64 qlt_mod_q, qlt_mod_q_h; // BGV: (q_last mod t) mod q_i
    u64 z_below_q;              // q_last <= q_i: every strictly reduced z is already < q_i
};

struct DropSet {
    DropConst *dev;   // [L-1]
    u64 half_qlast;   // q_last / 2
    u64 inv_t, inv_t_h; // BGV: t^{-1} mod q_last
};

struct Context {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool owns_stream = false;
    bool force_generic = false;
    int sm_count = 148;
    int latency_rows = -1; // option "latency_rows"; -1: half the SM count
    // option "latency2_rows"; -1: a tenth of the SM count (0 when latency_rows is 0).  15 rows of 8-CTA clusters run one CTA per SM
    // on a B200; the 16th cluster shares its SMs and the launch takes 7.9 instead of 4.4 us (profiles/r4_pair_path.md)
    int latency2_rows = -1;
    // option "single_launch": one ckks::mult pair per call as ONE launch with grid barriers (N = 4096 / 8192).  Off by default:
    // measured slower than the six programmatically chained launches (39 vs 33 us at C3, profiles/r3_latency_plans.md)
    bool single_launch = false;
    // option "pair_path": few ciphertexts per call run the key switch + drop as two cluster launches (ks_pair.cuh); 0 never,
    // 1 (default) for batches small enough to gain (ops.cu, pair_path_wanted), 2 whenever the shapes allow.  "pair_tpc": targets
    // per cluster (0: automatic).  "pair_fill_pct": fan-out rows per call, in % of the SM count, up to which the form is taken.
    int pair_path = 1, pair_tpc = 0, pair_fill_pct = 100;
    // option "fused_drop" (default 1): one ciphertext per call at N = 16384 / 32768 — the drop's inverse transform of the P limb
    // rides in the key switch's inner-product launch (ext_mac_intt_kernel, ops.cu)
    bool fused_drop = true;
    unsigned long long *grid_barrier_dev = nullptr; // counter of the single-launch kernel's grid barriers (only grows)
    unsigned long long grid_barrier_count = 0;       // its value once every launch enqueued so far has finished
    unsigned long long *grid_barrier_counter();
    int mult_one_clusters[2] = {0, 0}; // resident clusters of that kernel on this device (0: not asked yet, -1: launch form unavailable)
    LaunchEnv env() {
        const int lat = latency_rows < 0 ? sm_count / 2 : latency_rows;
        return LaunchEnv{stream, sm_count, force_generic, &stats, lat, latency2_rows < 0 ? (lat ? sm_count / 10 : 0) : latency2_rows, device};
    }
    // bound on the per-call workspace (option "scratch_cap_mib"); batches run in waves.  32 GiB of the 180 GB: a C5 wave of 296
    // ciphertexts (23 GiB with its key-switch digits) runs 1.7 % faster per ciphertext than four waves of 74 (profiles/r3_cluster_plans.md)
    size_t scratch_cap_bytes = (size_t)32 << 30;
    std::string last_error;
    LaunchStats stats;

    std::map<std::pair<u64, unsigned>, ModTables> tables;
    std::map<std::vector<u64>, LimbConst *> chains;      // key: {logn, q0, q1, ...}
    std::map<std::vector<u64>, DropSet> drops;           // key: {logn, t, q0, ..., q_last}
    std::map<std::vector<u64>, u64 *> scalar_sets;       // uploaded (s, s') pairs for mul_scalar
    std::map<const void *, int> cluster_cap;             // cluster kernel -> CTAs of it resident at once on this device
    std::map<const void *, int> smem_opt_in;             // kernel -> dynamic shared memory it has been opted in to (this device)
    std::map<size_t, std::vector<u64 *>> slab_free;      // pooled slabs by size
    std::map<u64 *, size_t> slab_live;
    std::vector<std::pair<u64 *, size_t>> scratch;       // grow-only workspaces, by slot

    // host-buffer entry points (host_pipe.cu): copy-in / copy-out streams and per-slot events
#ifndef HB_PIPE_SLOTS
#define HB_PIPE_SLOTS 4
#endif
    static constexpr int kPipeSlots = HB_PIPE_SLOTS;
    size_t host_chunk_bytes = (size_t)16 << 20; // option "host_chunk_kib": bytes per operand and chunk in the host-buffer pipeline
    cudaStream_t s_in = nullptr, s_out = nullptr;
    cudaEvent_t ev_in[kPipeSlots] = {}, ev_k[kPipeSlots] = {}, ev_out[kPipeSlots] = {};
    bool pipe_ready = false;
    int ensure_pipe();

    // Verdicts of stream-asynchronous operations (rns_base_transform many -> one inside key generation): each lands in
    // a pinned host word when the stream reaches it and is reported by the next hehub_b200_ctx_synchronize.
    int *deferred_flags = nullptr; // pinned, kDeferredSlots words
    static constexpr int kDeferredSlots = 64;
    int deferred_used = 0;
    int *deferred_flag_slot(int *err);
    // after a stream synchronisation: non-zero when some deferred verdict was raised (and clears them)
    int take_deferred();

    ~Context();
    int fail(int code, const std::string &msg) {
        last_error = msg;
        return code;
    }
    int cuda_fail(cudaError_t e, const char *where);

    // returns nullptr and sets last_error on invalid modulus
    const ModTables *get_tables(u64 q, unsigned logn, int *err);
    // device array of LimbConst for the chain (tables built on demand)
    const LimbConst *get_chain(unsigned logn, const u64 *moduli, size_t L, int *err);
    const DropSet *get_drop(unsigned logn, const u64 *moduli, size_t L, u64 t, int *err);
    u64 *get_scratch(size_t slot, size_t words, int *err);
};

// host helpers (exact restatements of the reference's parameter arithmetic)
u64 host_pow_mod(u64 q, u64 base, u64 e);
u64 host_root_2n(u64 q, u64 n);
u64 host_inverse_mod_prime(u64 elem, u64 prime);
inline u64 host_harvey_quotient(u64 w, u64 q) { return (u64)(((unsigned __int128)w << 64) / q); }

} // namespace hb
