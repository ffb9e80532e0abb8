// fp64_assist.cu — can the FP64 pipe of sm_100 take one of the six wide multiplies of an exact Harvey
// butterfly off the FMA-heavy pipe?  hi32(x0*w0) is exactly the low mantissa word of
// fma.rz.f64(double(x0), double(w0), 2^84): the sum 2^84 + p (p < 2^64) has ulp 2^32 and round-toward-zero
// truncates p to a multiple of 2^32.  This tool (1) checks that identity against __umul64hi on random
// words, (2) times DFMA / DADD alone and interleaved with IMAD.WIDE, (3) times register-resident
// butterflies with and without the FP64-assisted high product.  Prints one JSON line.  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I hehub_b200/csrc tools/fp64_assist.cu -o tools/fp64_assist
#include <cstdio>
#include <cuda_runtime.h>
#include "modarith.cuh"
using namespace hb;

constexpr int ITERS = 4096;

__device__ __forceinline__ double u32_to_f64(u32 v) { return __hiloint2double(0x43300000, (int)v) - 4503599627370496.0; }

// exact hi64(x*w) given c = hi32(x0*w0): m = x1*w0 + x0*w1 + c (66 bits), result x1*w1 + (m >> 32).
// Written as carry chains so that ptxas keeps three IMAD.WIDE (one with carry-out) and no fourth.
__device__ __forceinline__ u64 hi64_from_c(u32 x0, u32 x1, u32 w0, u32 w1, u32 c) {
    u32 r0, r1;
    asm("{\n\t.reg .u32 m0, m1, m2, t;\n\t.reg .u64 mm;\n\t"
        "mul.wide.u32 mm, %2, %5;\n\t"
        "mov.b64 {m0, m1}, mm;\n\t"
        "mad.lo.cc.u32 m0, %3, %4, m0;\n\t"
        "madc.hi.cc.u32 m1, %3, %4, m1;\n\t"
        "addc.u32 m2, 0, 0;\n\t"
        "add.cc.u32 t, m0, %6;\n\t"
        "addc.cc.u32 m1, m1, 0;\n\t"
        "addc.u32 m2, m2, 0;\n\t"
        "mad.lo.cc.u32 %0, %3, %5, m1;\n\t"
        "madc.hi.u32 %1, %3, %5, m2;\n\t}"
        : "=r"(r0), "=r"(r1) : "r"(x0), "r"(x1), "r"(w0), "r"(w1), "r"(c));
    return ((u64)r1 << 32) | r0;
}
// c from the FP64 pipe; CVT = 0: I2F.F64.U32 conversion, 1: magic-number DADD conversion
template <int CVT>
__device__ __forceinline__ double u32_as_f64(u32 v) {
    if (CVT == 0) return (double)v;
    return u32_to_f64(v);
}
template <int CVT>
__device__ __forceinline__ u64 umul64hi_dfma(u64 x, u64 w, double w0d) {
    const u32 x0 = (u32)x, x1 = (u32)(x >> 32), w0 = (u32)w, w1 = (u32)(w >> 32);
    const u32 c = (u32)__double2loint(__fma_rz(u32_as_f64<CVT>(x0), w0d, 0x1p84));
    return hi64_from_c(x0, x1, w0, w1, c);
}
// same chain with c from IMAD.HI (integer only): isolates what the restructured chain itself costs
__device__ __forceinline__ u64 umul64hi_chain_int(u64 x, u64 w) {
    const u32 x0 = (u32)x, x1 = (u32)(x >> 32), w0 = (u32)w, w1 = (u32)(w >> 32);
    return hi64_from_c(x0, x1, w0, w1, __umulhi(x0, w0));
}

__global__ void k_check(unsigned long long *bad, u64 seed, int per_thread) {
    u64 s = seed + (blockIdx.x * (u64)blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull;
    unsigned long long nbad = 0;
    for (int i = 0; i < per_thread; i++) {
        s = s * 6364136223846793005ull + 1442695040888963407ull;
        u64 x = s ^ (s >> 29);
        s = s * 6364136223846793005ull + 1442695040888963407ull;
        u64 w = s ^ (s >> 31);
        if ((i & 15) == 3) x |= 0xffffffffull;
        if ((i & 15) == 5) w |= 0xffffffffull;
        if ((i & 15) == 7) { x |= 0xffffffffull; w |= 0xffffffffull; }
        if ((i & 15) == 9) x &= ~0xffffffffull;
        if ((i & 31) == 11) { x = ~0ull; w = ~0ull; }
        const u64 want = __umul64hi(x, w);
        const u64 got = umul64hi_dfma<0>(x, w, (double)(u32)w);
        const u64 got2 = umul64hi_dfma<1>(x, w, u32_to_f64((u32)w));
        const u64 got3 = umul64hi_chain_int(x, w);
        nbad += (got != want) + (got2 != want) + (got3 != want);
    }
    if (nbad) atomicAdd(bad, nbad);
}

__global__ void k_dfma(double *out, double a, double b) {
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) asm volatile("fma.rz.f64 %0, %0, %1, %2;" : "+d"(x[i]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += x[i];
    if (s == 0.12345) out[0] = s;
}

__global__ void k_dadd(double *out, double a) {
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(x[i]) : "d"(a));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += x[i];
    if (s == 0.12345) out[0] = s;
}

// one IMAD.WIDE + NF DFMAs per group: do the FP64 and FMA-heavy pipes overlap?
template <int NF>
__global__ void k_wide_plus_dfma(double *out, unsigned a, double b, double c) {
    unsigned long long x[8];
    double y[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { x[i] = threadIdx.x + i; y[i] = threadIdx.x * 3 + i; }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            unsigned lo = (unsigned)x[i];
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x[i]) : "r"(lo), "r"(a));
#pragma unroll
            for (int f = 0; f < NF; f++) asm volatile("fma.rz.f64 %0, %0, %1, %2;" : "+d"(y[i]) : "d"(b), "d"(c));
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += y[i] + (double)x[i];
    if (s == 0.12345) out[0] = s;
}

// 8 values per thread, 3 levels (12 butterflies) per iteration, all in registers
template <int VARIANT>
__global__ void k_bfly(u64 *out, ulonglong2 tw, u64 nq, u64 q2) {
    u64 v[8];
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = threadIdx.x * 977 + i;
    // variant 2 converts the twiddle's low word once per level like a transform pass would
    for (int it = 0; it < ITERS / 4; it++) {
#pragma unroll
        for (int lvl = 4; lvl >= 1; lvl >>= 1) {
            double w0d = 0;
            if (VARIANT == 1 || VARIANT == 2) {
                u32 lo = (u32)tw.y;
                asm volatile("" : "+r"(lo)); // keep the twiddle's conversion inside the loop, once per level
                w0d = (VARIANT == 1) ? (double)lo : u32_to_f64(lo);
            }
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (i & lvl) continue;
                u64 t;
                if (VARIANT == 0) t = harvey_lazy(v[i + lvl], tw.x, tw.y, nq);
                else if (VARIANT == 1) t = mul2_lo64(v[i + lvl], tw.x, umul64hi_dfma<0>(v[i + lvl], tw.y, w0d), nq);
                else if (VARIANT == 2) t = mul2_lo64(v[i + lvl], tw.x, umul64hi_dfma<1>(v[i + lvl], tw.y, w0d), nq);
                else t = mul2_lo64(v[i + lvl], tw.x, umul64hi_chain_int(v[i + lvl], tw.y), nq);
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
    cudaMemset(out, 0, 64);
    k_check<<<sms * 4, 256>>>((unsigned long long *)out, 12345, 8192);
    unsigned long long bad = 0;
    cudaMemcpy(&bad, out, 8, cudaMemcpyDeviceToHost);
    const double checked = (double)sms * 4 * 256 * 8192;
    const double lanes = (double)blocks * threads;
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double per_clk_sm = 1.0 / sms / (clk * 1e3);
    double t;
    t = time_ms([&] { k_dfma<<<blocks, threads>>>((double *)out, 1.0000001, 0.5); });
    const double dfma = lanes * ITERS * 8 / (t * 1e-3);
    t = time_ms([&] { k_dadd<<<blocks, threads>>>((double *)out, 0.5); });
    const double dadd = lanes * ITERS * 8 / (t * 1e-3);
    t = time_ms([&] { k_wide_plus_dfma<0><<<blocks, threads>>>((double *)out, 3, 1.0000001, 0.5); });
    const double w0 = lanes * ITERS * 8 / (t * 1e-3);
    t = time_ms([&] { k_wide_plus_dfma<1><<<blocks, threads>>>((double *)out, 3, 1.0000001, 0.5); });
    const double w1 = lanes * ITERS * 8 / (t * 1e-3);
    t = time_ms([&] { k_wide_plus_dfma<2><<<blocks, threads>>>((double *)out, 3, 1.0000001, 0.5); });
    const double w2 = lanes * ITERS * 8 / (t * 1e-3);
    const u64 q = 576460752272228353ull;
    ulonglong2 tw = make_ulonglong2(123456789123456789ull % q, 0);
    tw.y = (u64)(((unsigned __int128)tw.x << 64) / q);
    double bf[4];
    t = time_ms([&] { k_bfly<0><<<blocks, threads>>>((u64 *)out, tw, 0 - q, 2 * q); });
    bf[0] = lanes * (ITERS / 4) * 12 / (t * 1e-3);
    t = time_ms([&] { k_bfly<1><<<blocks, threads>>>((u64 *)out, tw, 0 - q, 2 * q); });
    bf[1] = lanes * (ITERS / 4) * 12 / (t * 1e-3);
    t = time_ms([&] { k_bfly<2><<<blocks, threads>>>((u64 *)out, tw, 0 - q, 2 * q); });
    bf[2] = lanes * (ITERS / 4) * 12 / (t * 1e-3);
    t = time_ms([&] { k_bfly<3><<<blocks, threads>>>((u64 *)out, tw, 0 - q, 2 * q); });
    bf[3] = lanes * (ITERS / 4) * 12 / (t * 1e-3);
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz_max\": %d, \"hi64_identity_checked\": %.4g, \"hi64_identity_mismatches\": %llu, "
           "\"dfma_per_s\": %.4g, \"dfma_per_clk_per_sm\": %.1f, \"dadd_per_s\": %.4g, \"dadd_per_clk_per_sm\": %.1f, "
           "\"groups_wide_only_per_s\": %.4g, \"groups_wide_plus_1dfma_per_s\": %.4g, \"groups_wide_plus_2dfma_per_s\": %.4g, "
           "\"butterflies_per_s\": {\"int_only\": %.4g, \"dfma_low_product_i2f\": %.4g, \"dfma_low_product_magic_dadd\": %.4g, \"int_only_same_chain_imad_hi\": %.4g}}\n",
           p.name, sms, clk, checked, bad, dfma, dfma * per_clk_sm, dadd, dadd * per_clk_sm, w0, w1, w2, bf[0], bf[1], bf[2], bf[3]);
    return 0;
}
