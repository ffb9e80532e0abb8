// uniform_tw.cu — does it matter whether the twiddle operands of the butterfly stream sit in uniform registers
// (kernel parameters, as in tools/int_peak.cu) or in per-thread registers (as in the transform kernels, where
// they are loaded from the tables)?  Same register-resident butterfly loop, three operand placements.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I hehub_b200/csrc tools/uniform_tw.cu -o tools/uniform_tw
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
#include "modarith.cuh"
using namespace hb;
constexpr int ITERS = 4096;

// MODE 0: twiddle and modulus are kernel parameters (uniform registers)
// MODE 1: twiddle per thread (loaded from global memory), modulus uniform
// MODE 2: twiddle and modulus constants per thread
template <int MODE>
__global__ void k_bfly(u64 *out, const ulonglong2 *__restrict__ tws, ulonglong2 twu, u64 nqu, u64 q2u, const u64 *__restrict__ qs) {
    u64 v[8];
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = threadIdx.x * 977 + i;
    ulonglong2 tw = twu;
    u64 nq = nqu, q2 = q2u;
    if (MODE >= 1) tw = tws[threadIdx.x & 31];
    if (MODE >= 2) { nq = qs[threadIdx.x & 1]; q2 = qs[2 + (threadIdx.x & 1)]; }
    for (int it = 0; it < ITERS / 4; it++) {
#pragma unroll
        for (int lvl = 4; lvl >= 1; lvl >>= 1) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (i & lvl) continue;
                u64 t = harvey_lazy(v[i + lvl], tw.x, tw.y, nq);
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
    u64 *out, *qs;
    ulonglong2 *tws;
    cudaMalloc(&out, (size_t)blocks * threads * 8);
    cudaMalloc(&tws, 32 * sizeof(ulonglong2));
    cudaMalloc(&qs, 4 * 8);
    const u64 q = 576460752272228353ull, nq = 0 - q;
    ulonglong2 tw = make_ulonglong2(123456789123456789ull % q, 0);
    tw.y = (u64)(((unsigned __int128)tw.x << 64) / q);
    std::vector<ulonglong2> h(32, tw);
    cudaMemcpy(tws, h.data(), 32 * sizeof(ulonglong2), cudaMemcpyHostToDevice);
    const u64 hq[4] = {nq, nq, 2 * q, 2 * q};
    cudaMemcpy(qs, hq, sizeof(hq), cudaMemcpyHostToDevice);
    const double lanes = (double)blocks * threads;
    double r[3];
    double t = time_ms([&] { k_bfly<0><<<blocks, threads>>>(out, tws, tw, nq, 2 * q, qs); });
    r[0] = lanes * (ITERS / 4) * 12 / (t * 1e-3);
    t = time_ms([&] { k_bfly<1><<<blocks, threads>>>(out, tws, tw, nq, 2 * q, qs); });
    r[1] = lanes * (ITERS / 4) * 12 / (t * 1e-3);
    t = time_ms([&] { k_bfly<2><<<blocks, threads>>>(out, tws, tw, nq, 2 * q, qs); });
    r[2] = lanes * (ITERS / 4) * 12 / (t * 1e-3);
    printf("{\"gpu\": \"%s\", \"butterflies_per_s\": {\"twiddle_and_modulus_uniform\": %.4g, \"twiddle_per_thread\": %.4g, "
           "\"twiddle_and_modulus_per_thread\": %.4g}, \"ntt4096_per_s_ceiling\": {\"uniform\": %.4g, \"twiddle_per_thread\": %.4g, \"all_per_thread\": %.4g}}\n",
           p.name, r[0], r[1], r[2], r[0] / 24576, r[1] / 24576, r[2] / 24576);
    return 0;
}
