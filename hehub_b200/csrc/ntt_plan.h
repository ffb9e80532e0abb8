// ntt_plan.h — compile-time decomposition of an N-point negacyclic NTT into register-resident
// passes, shared by the device kernels (ntt_engine.cuh) and the host table builder (tables.cu).
//
// A row of N = 2^logn coefficients is processed by 2^lpre CTAs (a thread-block cluster when lpre > 0), each
// owning NC = N >> lpre coefficients in shared memory.  The butterfly levels are split into `npass` passes; in
// pass p every thread keeps 2^k[p] coefficients in registers and runs k[p] levels on them before the CTA
// exchanges through shared memory.  The forward transform runs the passes in the listed order (gaps shrink
// N/2 -> 1, last pass touches contiguous coefficients); the inverse transform runs the mirrored list (gaps grow
// 1 -> N/2).  The two directions have separate tables and may use different plans.
//
// How the lpre levels that pair words of different CTAs are done (`xchg`):
//   xchg = 0  (lpre <= 1) the level is not part of the pass list.  Forward: both CTAs read the whole row and each
//             computes its half of level 1 (one extra multiply per word); inverse: a finishing stage pulls the
//             sibling's words through distributed shared memory.
//   xchg = 1  the cross-CTA levels are the leading levels of k[0], a pass of the FULL row that the CTAs of the
//             cluster share (each takes 1/2^lpre of its butterfly groups).  Forward: the pass reads global memory
//             and scatters each result into the shared memory of the CTA that owns it (st.shared::cluster);
//             inverse (mirrored: the last pass) gathers its operands from the owners' shared memory
//             (ld.shared::cluster), finishes the words and stores them.  No level is computed twice.
//
// The twiddle tables are laid out per pass as [slot][block] so that the lanes of a warp read
// consecutive 16-byte (w, w') pairs in every pass (see tables.cu).
#pragma once
#include "compat.h"

namespace hb {

constexpr int kMaxPasses = 7;

struct NttPlan {
    int logn;
    int lpre;    // levels handled outside the CTA-local network (0, or 1 for N = 32768)
    int npass;   // CTA-local passes
    int k[kMaxPasses];
    int threads; // CTA size
    int min_blocks;
    int xchg;    // see above
};

constexpr int kFastLogMin = 10;
constexpr int kFastLogMax = 15;
constexpr int kGenericLogMax = 14; // one row must fit one CTA's shared memory

// Latency plans: when a launch has few rows (a single ciphertext, a key), a row is split over more CTAs than the
// throughput plan uses, each on its own SM, so the row's critical path shrinks while the GPU would otherwise idle.
// N = 4096 / 8192: a 2-CTA cluster (half a row per CTA; the level the two share is computed twice, which costs nothing
// on an idle GPU).  N = 16384 / 32768: an 8-CTA cluster that exchanges through distributed shared memory
// (4096 / 2048 words per CTA: a single N = 32768 row takes ~1/4 of the time of the 2-CTA form).
HB_CX bool has_latency_plan(int logn) { return logn >= 12 && logn <= 15; }
// Mode 2: N = 4096 / 8192 rows on an 8-CTA cluster of four warps — one warp per scheduler on twice the SMs of the mode-1 plans.
// The transforms of one row are bound by the integer multiplier of the SMs they run on (profiles/r4_pair_path.md), so this
// halves them; it is the form of the two-launch key switch of ONE ciphertext (ks_pair.cuh), whose 16 + 8 rows then cover the
// GPU once.  More rows per launch than that and the 4-CTA plans win (fewer, fuller CTAs).
HB_CX bool has_latency2_plan(int logn) { return logn == 12 || logn == 13; }

#ifndef HB_PLAN14
#define HB_PLAN14 3
#endif
#ifndef HB_PLAN15
#define HB_PLAN15 2
#endif

#ifndef HB_PLAN14I
#define HB_PLAN14I 0
#endif
#ifndef HB_PLAN15I
#define HB_PLAN15I 0
#endif

HB_CX NttPlan plan_for(int logn, bool fwd, int mode = 0) {
#if HB_PLAN14I == 1
    if (mode == 0 && logn == 14 && !fwd) return NttPlan{14, 1, 4, {3, 3, 4, 4, 0}, 256, 2, 1};
#elif HB_PLAN14I == 2
    if (mode == 0 && logn == 14 && !fwd) return NttPlan{14, 1, 3, {5, 5, 4, 0, 0}, 256, 2, 1};
#elif HB_PLAN14I == 3
    if (mode == 0 && logn == 14 && !fwd) return NttPlan{14, 1, 4, {4, 3, 3, 4, 0}, 256, 3, 1};
#endif
#if HB_PLAN15I == 1
    if (mode == 0 && logn == 15 && !fwd) return NttPlan{15, 2, 4, {4, 3, 4, 4, 0}, 256, 2, 1};
#elif HB_PLAN15I == 2
    if (mode == 0 && logn == 15 && !fwd) return NttPlan{15, 2, 4, {5, 3, 3, 4, 0}, 256, 2, 1};
#elif HB_PLAN15I == 3
    if (mode == 0 && logn == 15 && !fwd) return NttPlan{15, 2, 4, {4, 4, 3, 4, 0}, 256, 3, 1};
#endif
    if (mode == 2 && logn == 12) return NttPlan{12, 3, 4, {3, 3, 3, 3, 0}, 64, 1, 1};
    if (mode == 2 && logn == 13) return NttPlan{13, 3, 5, {3, 3, 3, 1, 3}, 128, 1, 1};
    if (mode == 2) mode = 1;
#ifndef HB_LAT_THIN
#define HB_LAT_THIN 1
#endif
#if HB_LAT_THIN
    // Thin latency plans: 8 words per thread (passes of 3 levels, one of 1-2 where the count needs it) on N/8 threads per row,
    // spread over a 4- or 8-CTA cluster.  A row alone on its SMs is bound by how fast ONE warp issues dependent instructions
    // (~6 cycles each at 1-2 warps per scheduler: profiles/r3_latency_plans.md), so its latency is the instruction count per
    // thread: 52-60 butterflies here against 112 with 16 words per thread on N/16 threads.
#if defined(HB_LAT_C8) // A/B: eight CTAs of four warps per row
    if (mode == 1 && logn == 12) return NttPlan{12, 3, 4, {3, 3, 3, 3, 0}, 64, 1, 1};
    if (mode == 1 && logn == 13) return NttPlan{13, 3, 5, {3, 3, 3, 1, 3}, 128, 1, 1};
#endif
#if HB_LAT_THIN == 2 // four words per thread on N/4 threads: six / seven passes of two levels
    if (mode == 1 && logn == 12) return NttPlan{12, 2, 6, {2, 2, 2, 2, 2, 2, 0}, 256, 1, 1};
    if (mode == 1 && logn == 13) return NttPlan{13, 2, 7, {2, 2, 2, 2, 2, 1, 2}, 512, 1, 1};
#endif
    if (mode == 1 && logn == 12) return NttPlan{12, 2, 4, {3, 3, 3, 3, 0}, 128, 1, 1};
    if (mode == 1 && logn == 13) return NttPlan{13, 2, 5, {3, 3, 3, 1, 3}, 256, 1, 1};
    if (mode == 1 && logn == 14) return NttPlan{14, 3, 5, {3, 3, 3, 2, 3}, 256, fwd ? 3 : 2, 1};
    // N = 32768: five thin passes on 512 threads lose to four passes of 16 words (one pair per call 143 -> 173 us): the
    // launches of one C5 ciphertext are not latency-bound rows but one to two waves of CTAs
#ifndef HB_LAT15_MINB
#define HB_LAT15_MINB 3
#endif
    if (mode == 1 && logn == 15) return NttPlan{15, 3, 4, {4, 4, 3, 4, 0}, 256, fwd ? HB_LAT15_MINB : 1, 1};
#else
    if (mode == 1 && logn == 12) return NttPlan{12, 1, 3, {3, 4, 4, 0, 0}, 128, 1, 0};
    if (mode == 1 && logn == 13) return NttPlan{13, 1, 3, {4, 4, 4, 0, 0}, 256, 1, 0};
    // 80 registers (three 256-thread CTAs per SM): launches of up to ~2 waves of such CTAs still gain from the finer grain
    // (C5: one pair per call 164 -> 148 us, two per call 115 -> 108 us/ct; profiles/r3_latency_plans.md)
    if (mode == 1 && logn == 14) return NttPlan{14, 3, 4, {4, 3, 3, 4, 0}, 128, fwd ? 6 : 2, 1};
    if (mode == 1 && logn == 15) return NttPlan{15, 3, 4, {4, 4, 3, 4, 0}, 256, fwd ? 3 : 1, 1};
#endif
    // Measured on B200 (profiles/r1_plan_sweep.md, profiles/r3_cluster_plans.md).  Passes of 3-5 levels keep 8-32 words
    // per thread in registers; ending with two passes of equal width keeps their exchange inside a warp.
    switch (logn) {
    case 10: return NttPlan{10, 0, 3, {3, 3, 4, 0, 0}, 64, 8, 0};
#if defined(HB_PLAN11) && HB_PLAN11 == 1
    case 11: return NttPlan{11, 0, 3, {4, 3, 4, 0, 0}, 128, 6, 0};
#else
    case 11: return NttPlan{11, 0, 3, {3, 4, 4, 0, 0}, 128, 6, 0};
#endif
    case 12: return NttPlan{12, 0, 3, {4, 4, 4, 0, 0}, 256, 3, 0};
#if defined(HB_PLAN13) && HB_PLAN13 == 0
    case 13: return NttPlan{13, 0, 3, {5, 4, 4, 0, 0}, 256, 2, 0};
#else // forward: four narrower passes and three CTAs per SM (+2 %, more for the fused kernels); the inverse loses 3 % with it
    // (inverse {4,5,4}: +0.7 % over {5,4,4}, profiles/r2n_plan13_inverse_ab.log)
    case 13: return fwd ? NttPlan{13, 0, 4, {3, 3, 3, 4, 0}, 256, 3, 0} : NttPlan{13, 0, 3, {4, 5, 4, 0, 0}, 256, 2, 0};
#endif
#if HB_PLAN14 == 0 // one CTA per row: nothing else shares the SM, load/compute/store phases do not overlap
    case 14: return NttPlan{14, 0, 3, {5, 5, 4, 0, 0}, 512, 1, 0};
#elif HB_PLAN14 == 1 // 2-CTA cluster per row, half a row per CTA, two CTAs per SM (+5 % over one CTA per row)
    case 14: return NttPlan{14, 1, 3, {5, 4, 4, 0, 0}, 256, 2, 0};
#elif HB_PLAN14 == 2 // the same with four narrower passes: three CTAs (of up to three different rows) per SM, another +2-4 %; level 1 twice
    case 14: return NttPlan{14, 1, 4, {3, 3, 3, 4, 0}, 256, 3, 0}; // inverse {4,5,4} / {5,4,4} x2: -4 % / -2 % (profiles/r2n_plan14_15_inverse_ab.log)
#elif HB_PLAN14 == 3 // 2-CTA cluster exchanging through distributed shared memory: no level twice
    case 14: return NttPlan{14, 1, 4, {3, 3, 4, 4, 0}, 256, 3, 1};
#else // 4-CTA cluster, a quarter row per CTA
    case 14: return NttPlan{14, 2, 4, {3, 3, 4, 4, 0}, 256, 3, 1};
#endif
#if HB_PLAN15 == 0 // 2-CTA cluster, 16 warps per CTA (one CTA per SM): too few to hide the fused epilogues' loads
    default: return NttPlan{15, 1, 4, {4, 3, 3, 4, 0}, 512, 1, 0};
#elif HB_PLAN15 == 1 // 32 warps per CTA at 64 registers: C5 rescale -11 %, mult+relin -1.6 % (profiles/r1_plan_sweep.md); level 1 twice
    default: return NttPlan{15, 1, 4, {3, 3, 4, 4, 0}, 1024, 1, 0};
#elif HB_PLAN15 == 2 // 4-CTA cluster exchanging through distributed shared memory: a quarter row (72 KB) per CTA, three CTAs of
      // different rows per SM so their load / compute / store phases overlap, and no level is computed twice
    // forward {3,4,4,4}: +1 % over {4,3,4,4}; the inverse prefers the wider cross pass (more loads in flight per thread): +5 %
    default: return fwd ? NttPlan{15, 2, 4, {3, 4, 4, 4, 0}, 256, 3, 1} : NttPlan{15, 2, 4, {4, 3, 4, 4, 0}, 256, 3, 1};
#elif HB_PLAN15 == 3
    default: return NttPlan{15, 2, 4, {3, 4, 4, 4, 0}, 256, 3, 1};
#elif HB_PLAN15 == 4 // 2-CTA cluster, exchanging instead of recomputing
    default: return NttPlan{15, 1, 4, {3, 4, 4, 4, 0}, 1024, 1, 1};
#else // 8-CTA cluster, 4096 words per CTA
    default: return NttPlan{15, 3, 4, {4, 4, 3, 4, 0}, 256, 3, 1};
#endif
    }
}

// CTAs per row
HB_CX int plan_cluster(const NttPlan &pl) { return 1 << pl.lpre; }

// ---- forward layout --------------------------------------------------------------------
// levels of the full row completed before pass p
HB_CX int fwd_glevel0(const NttPlan &pl, int p) {
    int s = pl.xchg ? 0 : pl.lpre;
    for (int i = 0; i < p; i++) s += pl.k[i];
    return s;
}
// CTA-local levels completed before pass p (p >= 1 when the plan exchanges)
HB_CX int fwd_lambda0(const NttPlan &pl, int p) { return fwd_glevel0(pl, p) - pl.lpre; }
// entry offset of pass p in the forward table (xchg = 0: the pre-level's one entry, T[1], comes first)
HB_CX int fwd_pass_offset(const NttPlan &pl, int p) {
    int off = (pl.lpre && !pl.xchg) ? 1 : 0;
    for (int i = 0; i < p; i++) off += ((1 << pl.k[i]) - 1) << fwd_glevel0(pl, i);
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

// ---- staged tables (mode 2) ----------------------------------------------------------------
// The mode-2 plans serve kernels that run ONE row per cluster on an otherwise idle SM (ks_pair.cuh): nothing hides a table
// load there, and every pass waited ~300 cycles for twiddles no other CTA of the SM had touched (profiles/r4_pair_path.md).
// Their tables are therefore laid out per CTA, in the order the CTA uses them, so that a CTA copies its slice into shared
// memory with a handful of contiguous asynchronous copies before its row arrives:
//   forward  [B][ cross pass: 2^k0 - 1 entries, 8 reserved | pass 1 [slot][hb] | pass 2 ... ], each block padded to kStagePad
//   inverse  [ local passes: the mode-independent [pass][slot][lo] block, padded ] then
//            [B][ cross twiddles [slot][thread] | psi^-i / N scales [j][thread] ]
// (entries are (w, w') pairs of 16 bytes).
constexpr int kStagePad = 256;
HB_CX int stage_pad(int entries) { return (entries + kStagePad - 1) / kStagePad * kStagePad; }
HB_CX int fwd_stage_offset(const NttPlan &pl, int p) { // within a CTA's block
    int off = 0;
    for (int i = 0; i < p; i++) off += i == 0 ? 8 : ((1 << pl.k[i]) - 1) << fwd_lambda0(pl, i);
    return off;
}
HB_CX int fwd_stage_block(const NttPlan &pl) { return stage_pad(fwd_stage_offset(pl, pl.npass)); }
HB_CX int inv_stage_local(const NttPlan &pl) { return stage_pad(inv_pass_offset(pl, pl.npass - 1)); }
HB_CX int inv_stage_cross_block(const NttPlan &pl) { return ((2 << pl.k[0]) - 1) * pl.threads; }
HB_CX int fwd_stage_total(const NttPlan &pl) { return fwd_stage_block(pl) << pl.lpre; }
HB_CX int inv_stage_total(const NttPlan &pl) { return inv_stage_local(pl) + (inv_stage_cross_block(pl) << pl.lpre); }

// shared-memory padding: 2 words after every 16 keeps 128-bit accesses of 16-word-strided
// owners and 64-bit accesses of consecutive lanes conflict-free
HB_CX int smem_phys(int i) { return i + ((i >> 4) << 1); }
HB_CX int smem_words(int nc) { return nc + (nc >> 3); }

} // namespace hb
