// ntt_plan.h — compile-time decomposition of an N-point negacyclic NTT into register-resident
// passes, shared by the device kernels (ntt_engine.cuh) and the host table builder (tables.cu).
//
// A row of N = 2^logn coefficients is processed by 2^lpre CTAs, each owning NC = N >> lpre
// coefficients in shared memory.  The logn - lpre CTA-local butterfly levels are split into
// `npass` passes; in pass p every thread keeps 2^k[p] coefficients in registers and runs k[p]
// levels on them before the CTA exchanges through shared memory.  The forward transform runs
// the passes in the listed order (gaps shrink N/2 -> 1, last pass touches contiguous
// coefficients); the inverse transform runs the mirrored list (gaps grow 1 -> N/2).  The two directions
// have separate tables and may use different plans.
//
// The twiddle tables are laid out per pass as [slot][block] so that the lanes of a warp read
// consecutive 16-byte (w, w') pairs in every pass (see tables.cu).
#pragma once
#include "compat.h"

namespace hb {

constexpr int kMaxPasses = 5;

struct NttPlan {
    int logn;
    int lpre;    // levels handled outside the CTA-local network (0, or 1 for N = 32768)
    int npass;   // CTA-local passes
    int k[kMaxPasses];
    int threads; // CTA size
    int min_blocks;
};

constexpr int kFastLogMin = 10;
constexpr int kFastLogMax = 15;
constexpr int kGenericLogMax = 14; // one row must fit one CTA's shared memory

// Latency plans: when a launch has fewer rows than half the SMs, a row of N = 4096 / 8192 is split over a 2-CTA
// cluster (half a row per CTA, each on its own SM) — the level the two CTAs share is computed twice, which costs
// nothing on an otherwise idle GPU, and the row's critical path halves.  N >= 16384 already runs as clusters.
HB_CX bool has_latency_plan(int logn) { return logn == 12 || logn == 13; }

HB_CX NttPlan plan_for(int logn, bool fwd, int mode = 0) {
    if (mode == 1 && logn == 12) return NttPlan{12, 1, 3, {3, 4, 4, 0, 0}, 128, 1};
    if (mode == 1 && logn == 13) return NttPlan{13, 1, 3, {4, 4, 4, 0, 0}, 256, 1};
    // Measured on B200 (profiles/r1_plan_sweep.md).  Passes of 4-5 levels keep 16-32 words per thread
    // in registers; ending with two passes of equal width keeps their exchange inside a warp.
    switch (logn) {
    case 10: return NttPlan{10, 0, 3, {3, 3, 4, 0, 0}, 64, 8};
#if defined(HB_PLAN11) && HB_PLAN11 == 1
    case 11: return NttPlan{11, 0, 3, {4, 3, 4, 0, 0}, 128, 6};
#else
    case 11: return NttPlan{11, 0, 3, {3, 4, 4, 0, 0}, 128, 6};
#endif
    case 12: return NttPlan{12, 0, 3, {4, 4, 4, 0, 0}, 256, 3};
#if defined(HB_PLAN13) && HB_PLAN13 == 0
    case 13: return NttPlan{13, 0, 3, {5, 4, 4, 0, 0}, 256, 2};
#else // forward: four narrower passes and three CTAs per SM (+2 %, more for the fused kernels); the inverse loses 3 % with it
    // (inverse {4,5,4}: +0.7 % over {5,4,4}, profiles/r2n_plan13_inverse_ab.log)
    case 13: return fwd ? NttPlan{13, 0, 4, {3, 3, 3, 4, 0}, 256, 3} : NttPlan{13, 0, 3, {4, 5, 4, 0, 0}, 256, 2};
#endif
#if defined(HB_PLAN14) && HB_PLAN14 == 0 // one CTA per row: nothing else shares the SM, load/compute/store phases do not overlap
    case 14: return NttPlan{14, 0, 3, {5, 5, 4, 0, 0}, 512, 1};
#elif defined(HB_PLAN14) && HB_PLAN14 == 1 // 2-CTA cluster per row, half a row per CTA, two CTAs per SM (+5 % over one CTA per row)
    case 14: return NttPlan{14, 1, 3, {5, 4, 4, 0, 0}, 256, 2};
#else // cluster form with four narrower passes: three CTAs (of up to three different rows) per SM, another +2-4 %
    case 14: return NttPlan{14, 1, 4, {3, 3, 3, 4, 0}, 256, 3}; // inverse {4,5,4} / {5,4,4} x2: -4 % / -2 % (profiles/r2n_plan14_15_inverse_ab.log)
#endif
#if defined(HB_PLAN15) && HB_PLAN15 == 0 // 16 warps per CTA (one CTA per SM): too few to hide the fused epilogues' loads
    default: return NttPlan{15, 1, 4, {4, 3, 3, 4, 0}, 512, 1};
#else // 32 warps per CTA at 64 registers: C5 rescale -11 %, mult+relin -1.6 % (profiles/r1_plan_sweep.md)
    default: return NttPlan{15, 1, 4, {3, 3, 4, 4, 0}, 1024, 1};
#endif
    }
}

// ---- forward layout --------------------------------------------------------------------
// local levels completed before CTA pass p
HB_CX int fwd_lambda0(const NttPlan &pl, int p) {
    int s = 0;
    for (int i = 0; i < p; i++) s += pl.k[i];
    return s;
}
// entry offset of CTA pass p in the forward table: the lpre pre-level (one entry, T[1]) first
HB_CX int fwd_pass_offset(const NttPlan &pl, int p) {
    int off = pl.lpre ? 1 : 0;
    for (int i = 0; i < p; i++) off += ((1 << pl.k[i]) - 1) << (pl.lpre + fwd_lambda0(pl, i));
    return off;
}

// ---- inverse layout (mirrored pass list) -------------------------------------------------
HB_CX int inv_k(const NttPlan &pl, int p) { return pl.k[pl.npass - 1 - p]; }
HB_CX int inv_s0(const NttPlan &pl, int p) {
    int s = 0;
    for (int i = 0; i < p; i++) s += inv_k(pl, i);
    return s;
}
// entry offset of inverse pass p; p == npass addresses the lpre post-stage (gap N/2)
HB_CX int inv_pass_offset(const NttPlan &pl, int p) {
    int off = 0;
    for (int i = 0; i < p; i++) off += ((1 << inv_k(pl, i)) - 1) << inv_s0(pl, i);
    return off;
}

// shared-memory padding: 2 words after every 16 keeps 128-bit accesses of 16-word-strided
// owners and 64-bit accesses of consecutive lanes conflict-free
HB_CX int smem_phys(int i) { return i + ((i >> 4) << 1); }
HB_CX int smem_words(int nc) { return nc + (nc >> 3); }

} // namespace hb
