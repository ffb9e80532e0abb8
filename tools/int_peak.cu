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
                u64 t = VARIANT ? harvey_lazy_split(v[i + lvl], tw.x, tw.y, nq) : harvey_lazy(v[i + lvl], tw.x, tw.y, nq);
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
           "\"ntt4096_per_s_alu_ceiling\": %.4g, \"butterflies_per_s_by_warps_per_scheduler\": {%s}}\n",
           p.name, sms, clk, imad, wide, alu, bfly, bfly_split, imad_hi, imad / sms / (clk * 1e3), wide / sms / (clk * 1e3), bfly / 24576.0, occ);
    return 0;
}
