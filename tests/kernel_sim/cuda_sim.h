// cuda_sim.h — a minimal CTA emulator for the CPU test suite.  TEST INFRASTRUCTURE ONLY.
//
// Lets hehub_b200/csrc/*.cu be compiled with g++ (-DHB_KERNEL_SIM) so that `pytest -m "not gpu"`
// can drive the very same kernel bodies, table builders and host logic through the same C ABI,
// one CTA at a time, and compare against the oracle.  It checks index arithmetic and layouts;
// it says nothing about performance and is never part of the product (see compat.h).
//
// Model: blocks run one after another; a kernel launched with sync != 0 runs its `block`
// threads as real std::threads meeting at a barrier for __syncthreads(); other kernels run
// their threads sequentially.  "Device" memory is host memory; streams are synchronous.
#pragma once
#include <condition_variable>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

typedef unsigned long long hbsim_u64;
struct alignas(16) ulonglong2 {
    hbsim_u64 x, y;
};
static inline ulonglong2 make_ulonglong2(hbsim_u64 x, hbsim_u64 y) { return ulonglong2{x, y}; }
struct hbsim_dim3 {
    unsigned x, y, z;
};

namespace hbsim {
extern thread_local hbsim_dim3 t_threadIdx, t_blockIdx, t_blockDim, t_gridDim;
hbsim_u64 *shared_u64();
hbsim_u64 *shared_u64_of(size_t cluster_rank);
void sync_threads();
void cluster_sync();
void launch(size_t grid, size_t block, size_t smem_bytes, int sync, const std::function<void()> &body, size_t cluster = 1);
} // namespace hbsim

#define threadIdx (::hbsim::t_threadIdx)
#define blockIdx (::hbsim::t_blockIdx)
#define blockDim (::hbsim::t_blockDim)
#define gridDim (::hbsim::t_gridDim)
#define __syncthreads() ::hbsim::sync_threads()
#define __restrict__
#define __forceinline__ inline
#define __device__
#define __host__

template <class T>
static inline T __ldg(const T *p) { return *p; }
static inline hbsim_u64 __umul64hi(hbsim_u64 a, hbsim_u64 b) { return (hbsim_u64)(((unsigned __int128)a * b) >> 64); }
static inline unsigned __brev(unsigned x) {
    unsigned r = 0;
    for (int i = 0; i < 32; i++) r |= ((x >> i) & 1u) << (31 - i);
    return r;
}

// ---- the sliver of the CUDA runtime the host code uses ----
typedef int cudaError_t;
typedef void *cudaStream_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2 };
enum { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
enum { cudaStreamNonBlocking = 1 };
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaFuncAttributePreferredSharedMemoryCarveout = 9 };
static inline const char *cudaGetErrorString(cudaError_t e) { return e ? "simulated CUDA error" : "no error"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
template <class T>
static inline cudaError_t cudaMalloc(T **p, size_t bytes) {
    void *q = nullptr;
    if (posix_memalign(&q, 256, bytes ? bytes : 1)) return cudaErrorMemoryAllocation;
    *p = static_cast<T *>(q);
    return cudaSuccess;
}
static inline cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMallocHost(void **p, size_t bytes) { return cudaMalloc(p, bytes); }
static inline cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, int, cudaStream_t) {
    memmove(d, s, n);
    return cudaSuccess;
}
static inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t) {
    memset(d, v, n);
    return cudaSuccess;
}
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, int) { *s = (void *)1; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
typedef void *cudaEvent_t;
enum { cudaEventDisableTiming = 2 };
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, int) { *e = (void *)1; return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, int) { return cudaSuccess; }
template <class F>
static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }
