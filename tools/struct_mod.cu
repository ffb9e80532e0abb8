// struct_mod.cu — what a modulus of the form q = 2^k - delta (delta < 2^32; every prime of the reference's
// tables) would save in the Harvey butterfly: -q has the high word -(2^(k-32)) mod 2^32, so the narrow
// multiply qhat0 * nq1 becomes a shift and a subtract on the ALU pipe.  Register-resident butterflies,
// same harness as tools/int_peak.cu.  Prints one JSON line.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I hehub_b200/csrc tools/struct_mod.cu -o tools/struct_mod
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
#include "modarith.cuh"
using namespace hb;
constexpr int ITERS = 4096;

// lo64(x*w + h*n) with n = (n0, -(1 << s)): three narrow IMADs instead of four
__device__ __forceinline__ u64 mul2_lo64_struct(u64 x, u64 w, u64 h, u32 n0, u32 s) {
    u64 t;
    asm("{\n\t"
        ".reg .u64 acc;\n\t"
        ".reg .u32 x0, x1, w0, w1, h0, h1, lo, hi, sh;\n\t"
        "mov.b64 {x0, x1}, %1;\n\t"
        "mov.b64 {w0, w1}, %2;\n\t"
        "mov.b64 {h0, h1}, %3;\n\t"
        "mul.wide.u32 acc, x0, w0;\n\t"
        "mad.wide.u32 acc, h0, %4, acc;\n\t"
        "mov.b64 {lo, hi}, acc;\n\t"
        "mad.lo.u32 hi, x0, w1, hi;\n\t"
        "mad.lo.u32 hi, x1, w0, hi;\n\t"
        "mad.lo.u32 hi, h1, %4, hi;\n\t"
        "shl.b32 sh, h0, %5;\n\t"
        "sub.u32 hi, hi, sh;\n\t"
        "mov.b64 %0, {lo, hi};\n\t"
        "}"
        : "=l"(t)
        : "l"(x), "l"(w), "l"(h), "r"(n0), "r"(s));
    return t;
}

template <int VARIANT>
__global__ void k_bfly(u64 *out, ulonglong2 tw, u64 nq, u64 q2, u32 s) {
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
                else t = mul2_lo64_struct(v[i + lvl], tw.x, __umul64hi(v[i + lvl], tw.y), (u32)nq, s);
                v[i + lvl] = v[i] + q2 - t;
                v[i] = v[i] + t;
            }
        }
    }
    u64 r = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) r ^= v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
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
    u64 *out;
    cudaMalloc(&out, (size_t)blocks * threads * 8);
    const u64 q = 576460752272228353ull, nq = 0 - q;
    const u32 s = 27; // -q = 0xF8000000_01DBFFFF: high word -(1 << 27)
    ulonglong2 tw = make_ulonglong2(123456789123456789ull % q, 0);
    tw.y = (u64)(((unsigned __int128)tw.x << 64) / q);
    const double lanes = (double)blocks * threads;
    std::vector<u64> h0((size_t)blocks * threads), h1(h0.size());
    double t = time_ms([&] { k_bfly<0><<<blocks, threads>>>(out, tw, nq, 2 * q, s); });
    cudaMemcpy(h0.data(), out, h0.size() * 8, cudaMemcpyDeviceToHost);
    const double b0 = lanes * (ITERS / 4) * 12 / (t * 1e-3);
    t = time_ms([&] { k_bfly<1><<<blocks, threads>>>(out, tw, nq, 2 * q, s); });
    cudaMemcpy(h1.data(), out, h1.size() * 8, cudaMemcpyDeviceToHost);
    const double b1 = lanes * (ITERS / 4) * 12 / (t * 1e-3);
    size_t bad = 0;
    for (size_t i = 0; i < h0.size(); i++) bad += h0[i] != h1[i];
    printf("{\"gpu\": \"%s\", \"butterflies_per_s\": {\"generic\": %.4g, \"structured_modulus\": %.4g}, \"ratio\": %.4f, \"result_mismatches\": %zu}\n",
           p.name, b0, b1, b1 / b0, bad);
    return 0;
}
