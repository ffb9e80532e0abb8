// int_peak.cu — measures the integer-pipe ceilings that bound the NTT butterflies on this GPU:
// 32-bit IMAD, IMAD.WIDE and IADD3 issue rates and the rate of register-resident Harvey
// butterflies (no memory traffic).  Prints one JSON line.  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I hehub_b200/csrc tools/int_peak.cu -o tools/int_peak
#include <cstdio>
#include <cuda_runtime.h>
#include "modarith.cuh"
using namespace hb;

constexpr int ITERS = 4096;

__global__ void k_imad(unsigned *out, unsigned a, unsigned b) {
    unsigned x[8];
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b));
    }
    unsigned s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += x[i];
    if (s == 0x12345678) out[0] = s;
}

__global__ void k_imad_wide(unsigned long long *out, unsigned a) {
    unsigned long long x[8];
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            unsigned lo = (unsigned)x[i];
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x[i]) : "r"(lo), "r"(a));
        }
    }
    unsigned long long s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += x[i];
    if (s == 0x12345678) out[0] = s;
}

__global__ void k_iadd3(unsigned *out, unsigned a, unsigned b) {
    unsigned x[8];
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) asm volatile("add.u32 %0, %0, %1; xor.b32 %0, %0, %2;" : "+r"(x[i]) : "r"(a), "r"(b));
    }
    unsigned s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += x[i];
    if (s == 0x12345678) out[0] = s;
}

__global__ void k_imad_hi(unsigned *out, unsigned a, unsigned b) {
    unsigned x[8];
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b));
    }
    unsigned s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += x[i];
    if (s == 0x12345678) out[0] = s;
}


// mul.wide.u32 with no addend (independent products, results xor-folded on the ALU pipe)
__global__ void k_mul_wide_noadd(unsigned long long *out, unsigned a) {
    unsigned x[8];
    unsigned long long acc[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { x[i] = threadIdx.x + i; acc[i] = 0; }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            unsigned long long p;
            asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(x[i]), "r"(a));
            x[i] = (unsigned)(p >> 32) ^ (unsigned)p;
        }
    }
    unsigned long long s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += x[i];
    if (s == 0x12345678) out[0] = s;
}

// fma-pipe IMAD and alu-pipe IADD3/LOP3 streams interleaved: do the two pipes overlap fully?
__global__ void k_imad_plus_alu(unsigned *out, unsigned a, unsigned b) {
    unsigned x[8], y[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { x[i] = threadIdx.x + i; y[i] = threadIdx.x * 3 + i; }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b));
            asm volatile("add.u32 %0, %0, %1;" : "+r"(y[i]) : "r"(a));
            asm volatile("xor.b32 %0, %0, %1;" : "+r"(y[i]) : "r"(b));
        }
    }
    unsigned s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += x[i] + y[i];
    if (s == 0x12345678) out[0] = s;
}

// IMAD.WIDE stream plus three ALU ops per wide multiply
__global__ void k_wide_plus_alu(unsigned long long *out, unsigned a, unsigned b) {
    unsigned long long x[8];
    unsigned y[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { x[i] = threadIdx.x + i; y[i] = threadIdx.x * 3 + i; }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            unsigned lo = (unsigned)x[i];
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x[i]) : "r"(lo), "r"(a));
            asm volatile("add.u32 %0, %0, %1;" : "+r"(y[i]) : "r"(a));
            asm volatile("xor.b32 %0, %0, %1;" : "+r"(y[i]) : "r"(b));
            asm volatile("add.u32 %0, %0, %1;" : "+r"(y[i]) : "r"(b));
        }
    }
    unsigned long long s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += x[i] + y[i];
    if (s == 0x12345678) out[0] = s;
}

// exact hi64 from four independent 32x32 products, sums on the ALU pipe
__device__ __forceinline__ u64 umul64hi_indep(u64 x, u64 w) {
    unsigned x0 = (unsigned)x, x1 = (unsigned)(x >> 32), w0 = (unsigned)w, w1 = (unsigned)(w >> 32);
    u64 p00, p01, p10, p11;
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(p00) : "r"(x0), "r"(w0));
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(p01) : "r"(x0), "r"(w1));
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(p10) : "r"(x1), "r"(w0));
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(p11) : "r"(x1), "r"(w1));
    u64 mid = p01 + (p00 >> 32);
    u64 mid2 = p10 + (mid & 0xffffffffull);
    return p11 + (mid >> 32) + (mid2 >> 32);
}
__device__ __forceinline__ u64 harvey_lazy_indep(u64 x, u64 w, u64 wh, u64 nq) { return mul2_lo64(x, w, umul64hi_indep(x, wh), nq); }
// plain C++ form: whatever the compiler makes of it
__device__ __forceinline__ u64 harvey_lazy_plain(u64 x, u64 w, u64 wh, u64 q) { return x * w - __umul64hi(x, wh) * q; }

// 8 values per thread, 3 levels (12 butterflies) per iteration, all in registers
template <int VARIANT>
__global__ void k_bfly(u64 *out, ulonglong2 tw, u64 nq, u64 q2) {
    u64 v[8];
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = threadIdx.x * 977 + i;
    for (int it = 0; it < ITERS / 4; it++) {
#pragma unroll
        for (int lvl = 4; lvl >= 1; lvl >>= 1) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (i & lvl) continue;
                u64 t;
                if (VARIANT == 0) t = harvey_lazy(v[i + lvl], tw.x, tw.y, nq);
                else if (VARIANT == 1) t = harvey_lazy_split(v[i + lvl], tw.x, tw.y, nq);
                else if (VARIANT == 2) t = harvey_lazy_indep(v[i + lvl], tw.x, tw.y, nq);
                else t = harvey_lazy_plain(v[i + lvl], tw.x, tw.y, 0 - nq);
                v[i + lvl] = v[i] + q2 - t;
                v[i] = v[i] + t;
            }
        }
    }
    u64 s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s ^= v[i];
    if (s == 0x12345678) out[0] = s;
}


// butterflies plus EXTRA independent ALU-pipe instructions per butterfly: how much non-multiply
// overhead (address arithmetic, moves) the issue port tolerates before the FMA pipe starves
template <int EXTRA>
__global__ void k_bfly_overhead(u64 *out, ulonglong2 tw, u64 nq, u64 q2, unsigned a) {
    u64 v[8];
    unsigned y[4];
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = threadIdx.x * 977 + i;
#pragma unroll
    for (int i = 0; i < 4; i++) y[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS / 4; it++) {
#pragma unroll
        for (int lvl = 4; lvl >= 1; lvl >>= 1) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (i & lvl) continue;
                u64 t = harvey_lazy(v[i + lvl], tw.x, tw.y, nq);
                v[i + lvl] = v[i] + q2 - t;
                v[i] = v[i] + t;
#pragma unroll
                for (int e = 0; e < EXTRA; e++) {
                    asm volatile("prmt.b32 %0, %0, %1, 0x1230;" : "+r"(y[e & 3]) : "r"(a));
                }
            }
        }
    }
    u64 s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s ^= v[i];
#pragma unroll
    for (int i = 0; i < 4; i++) s ^= y[i];
    if (s == 0x12345678) out[0] = s;
}

// 64-bit adds (IADD3 + IADD3.X pairs)
__global__ void k_add64(u64 *out, u64 a) {
    u64 x[8];
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) asm volatile("add.u64 %0, %0, %1;" : "+l"(x[i]) : "l"(a));
    }
    u64 s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += x[i];
    if (s == 0x12345678) out[0] = s;
}

template <class F>
static double time_ms(F f) {
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    f();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(a);
        f();
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount, blocks = sms * 8, threads = 256;
    void *out;
    cudaMalloc(&out, 64);
    const double lanes = (double)blocks * threads;
    double t;
    t = time_ms([&] { k_imad<<<blocks, threads>>>((unsigned *)out, 3, 5); });
    const double imad = lanes * ITERS * 8 / (t * 1e-3);
    t = time_ms([&] { k_imad_wide<<<blocks, threads>>>((unsigned long long *)out, 3); });
    const double wide = lanes * ITERS * 8 / (t * 1e-3);
    t = time_ms([&] { k_iadd3<<<blocks, threads>>>((unsigned *)out, 3, 5); });
    const double alu = lanes * ITERS * 8 * 2 / (t * 1e-3);
    const u64 q = 576460752272228353ull;
    ulonglong2 tw = make_ulonglong2(123456789123456789ull % q, 0);
    tw.y = (u64)(((unsigned __int128)tw.x << 64) / q);
    t = time_ms([&] { k_bfly<0><<<blocks, threads>>>((u64 *)out, tw, 0 - q, 2 * q); });
    const double bfly = lanes * (ITERS / 4) * 12 / (t * 1e-3);
    t = time_ms([&] { k_bfly<1><<<blocks, threads>>>((u64 *)out, tw, 0 - q, 2 * q); });
    const double bfly_split = lanes * (ITERS / 4) * 12 / (t * 1e-3);
    t = time_ms([&] { k_imad_hi<<<blocks, threads>>>((unsigned *)out, 3, 5); });
    const double imad_hi = lanes * ITERS * 8 / (t * 1e-3);
    t = time_ms([&] { k_bfly<2><<<blocks, threads>>>((u64 *)out, tw, 0 - q, 2 * q); });
    const double bfly_indep = lanes * (ITERS / 4) * 12 / (t * 1e-3);
    t = time_ms([&] { k_bfly<3><<<blocks, threads>>>((u64 *)out, tw, 0 - q, 2 * q); });
    const double bfly_plain = lanes * (ITERS / 4) * 12 / (t * 1e-3);
    t = time_ms([&] { k_mul_wide_noadd<<<blocks, threads>>>((unsigned long long *)out, 3); });
    const double wide_noadd = lanes * ITERS * 8 / (t * 1e-3);
    t = time_ms([&] { k_imad_plus_alu<<<blocks, threads>>>((unsigned *)out, 3, 5); });
    const double imad_alu = lanes * ITERS * 8 / (t * 1e-3); // groups of (1 IMAD + 2 ALU)
    t = time_ms([&] { k_wide_plus_alu<<<blocks, threads>>>((unsigned long long *)out, 3, 5); });
    const double wide_alu = lanes * ITERS * 8 / (t * 1e-3); // groups of (1 IMAD.WIDE + 3 ALU)
    double ovh[5];
    t = time_ms([&] { k_bfly_overhead<0><<<blocks, threads>>>((u64 *)out, tw, 0 - q, 2 * q, 3); });
    ovh[0] = lanes * (ITERS / 4) * 12 / (t * 1e-3);
    t = time_ms([&] { k_bfly_overhead<2><<<blocks, threads>>>((u64 *)out, tw, 0 - q, 2 * q, 3); });
    ovh[1] = lanes * (ITERS / 4) * 12 / (t * 1e-3);
    t = time_ms([&] { k_bfly_overhead<4><<<blocks, threads>>>((u64 *)out, tw, 0 - q, 2 * q, 3); });
    ovh[2] = lanes * (ITERS / 4) * 12 / (t * 1e-3);
    t = time_ms([&] { k_bfly_overhead<8><<<blocks, threads>>>((u64 *)out, tw, 0 - q, 2 * q, 3); });
    ovh[3] = lanes * (ITERS / 4) * 12 / (t * 1e-3);
    t = time_ms([&] { k_bfly_overhead<12><<<blocks, threads>>>((u64 *)out, tw, 0 - q, 2 * q, 3); });
    ovh[4] = lanes * (ITERS / 4) * 12 / (t * 1e-3);
    t = time_ms([&] { k_add64<<<blocks, threads>>>((u64 *)out, 12345); });
    const double add64 = lanes * ITERS * 8 / (t * 1e-3);
    // occupancy sweep: how many resident warps per scheduler the butterfly stream needs to fill the pipe
    char occ[512];
    int off = 0;
    for (int bps : {1, 2, 3, 4, 6, 8}) {
        const int nb = sms * bps;
        t = time_ms([&] { k_bfly<0><<<nb, threads>>>((u64 *)out, tw, 0 - q, 2 * q); });
        const double r = (double)nb * threads * (ITERS / 4) * 12 / (t * 1e-3);
        off += snprintf(occ + off, sizeof(occ) - off, "%s\"%d\": %.4g", off ? ", " : "", bps * threads / 32 / 4, r);
    }
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz_max\": %d, \"imad_per_s\": %.4g, \"imad_wide_per_s\": %.4g, "
           "\"alu_ops_per_s\": %.4g, \"harvey_butterflies_per_s\": %.4g, \"harvey_butterflies_split_hi_per_s\": %.4g, \"imad_hi_per_s\": %.4g, "
           "\"imad_per_clk_per_sm_at_max_clock\": %.1f, \"imad_wide_per_clk_per_sm_at_max_clock\": %.1f, "
           "\"ntt4096_per_s_alu_ceiling\": %.4g, \"butterflies_per_s_by_warps_per_scheduler\": {%s}, "
           "\"harvey_butterflies_indep_products_per_s\": %.4g, \"harvey_butterflies_plain_cxx_per_s\": %.4g, "
           "\"mul_wide_noaddend_per_s\": %.4g, \"groups_imad_plus_2alu_per_s\": %.4g, \"groups_wide_plus_3alu_per_s\": %.4g, "
           "\"add64_per_s\": %.4g, \"butterflies_per_s_with_extra_alu_instr_per_butterfly\": {\"0\": %.4g, \"2\": %.4g, \"4\": %.4g, \"8\": %.4g, \"12\": %.4g}}\n",
           p.name, sms, clk, imad, wide, alu, bfly, bfly_split, imad_hi, imad / sms / (clk * 1e3), wide / sms / (clk * 1e3), bfly / 24576.0, occ, bfly_indep, bfly_plain, wide_noadd, imad_alu, wide_alu, add64, ovh[0], ovh[1], ovh[2], ovh[3], ovh[4]);
    return 0;
}
