// compat.h — the few macros that let the kernel sources be compiled twice:
//
//   * by nvcc for sm_100a — the product (libhehub_b200.so);
//   * by g++ against tests/kernel_sim/cuda_sim.h — a CTA emulator used ONLY by the CPU test
//     suite to check the kernels' index arithmetic, table layouts and host logic without a
//     GPU (tests/kernel_sim/README.md).  The emulator is never built into, linked with or
//     loaded by the product library; the product has no CPU path.
#pragma once

#if defined(HB_KERNEL_SIM)
#include <cstring>

#include "cuda_sim.h" // tests/kernel_sim
#define HB_D inline
#define HB_CX constexpr
#define HB_GLOBAL(threads, minblocks) static void
#define HB_SHARED_U64(name) u64 *name = ::hbsim::shared_u64()
// sync != 0: the kernel calls __syncthreads() (the emulator then runs real threads)
#define HB_LAUNCH(kern, grid, block, smem, stream, sync, ...) \
    ::hbsim::launch((grid), (block), (smem), (sync), [&]() { kern(__VA_ARGS__); })
// thread-block cluster of `cluster` consecutive CTAs (the emulator runs them concurrently)
#define HB_LAUNCH_CLUSTER(kern, grid, block, smem, stream, cluster, ...) \
    (::hbsim::launch((grid), (block), (smem), 1, [&]() { kern(__VA_ARGS__); }, (cluster)), cudaSuccess)
inline void hb_cluster_sync() { ::hbsim::cluster_sync(); }
// split form: arrive early, wait where the guarantee is needed (the emulator's barrier is one-shot: it meets at the wait)
inline void hb_cluster_arrive() {}
inline void hb_cluster_wait() { ::hbsim::cluster_sync(); }
// one word of CTA `rank`'s shared memory at the offset `p` has in this CTA's
inline unsigned long long hb_ld_dsmem(const unsigned long long *p, unsigned rank) {
    return ::hbsim::shared_u64_of(rank)[p - ::hbsim::shared_u64()];
}
inline void hb_st_dsmem(unsigned long long *p, unsigned rank, unsigned long long v) {
    ::hbsim::shared_u64_of(rank)[p - ::hbsim::shared_u64()] = v;
}
// two adjacent words of CTA `rank`'s shared memory at the offset `p` has in this CTA's (DSMEM)
inline ulonglong2 hb_ld_dsmem2(const unsigned long long *p, unsigned rank) {
    return *reinterpret_cast<const ulonglong2 *>(::hbsim::shared_u64_of(rank) + (p - ::hbsim::shared_u64()));
}
inline void hb_prefetch_l2(const void *) {}
inline void hb_prefetch_l1(const void *) {}
inline void hb_pdl_wait() {}
inline void hb_pdl_trigger() {}
inline void hb_cp_async16(void *dst_shared, const void *src_global) { std::memcpy(dst_shared, src_global, 16); }
inline void hb_cp_async_wait_all() {}
inline void hb_syncwarp() { __syncthreads(); } // the emulator has no warps: a CTA barrier is a superset
inline unsigned long long hb_ld_stream(const unsigned long long *p) { return *p; }
inline ulonglong2 hb_ld_stream2(const unsigned long long *p) { return *reinterpret_cast<const ulonglong2 *>(p); }
inline unsigned long long hb_ld_ro(const unsigned long long *p) { return *p; }
inline ulonglong2 hb_ld_ro2(const unsigned long long *p) { return *reinterpret_cast<const ulonglong2 *>(p); }
template <class T>
inline T hb_ldcg(const T *p) { return *p; }
#else
#include <cuda_runtime.h>
#define HB_D __device__ __forceinline__
#define HB_CX __host__ __device__ constexpr
#define HB_GLOBAL(threads, minblocks) __global__ void __launch_bounds__(threads, minblocks)
#define HB_SHARED_U64(name) extern __shared__ __align__(16) unsigned long long name[]
// Every launch goes through cudaLaunchKernelEx with programmatic stream serialization: the next kernel of the
// stream may be scheduled while this one drains, and runs its prologue (index arithmetic, per-limb constants) up
// to hb_pdl_wait(), which returns once every earlier grid has completed and its writes are visible.  The
// transforms of one ciphertext are six launches on nearly empty grids: launch latency is most of that case.
#define HB_LAUNCH(kern, grid, block, smem, stream, sync, ...) \
    ((void)::hb_launch_ex(kern, (unsigned)(grid), (unsigned)(block), (smem), (stream), 1u, __VA_ARGS__))
#define HB_LAUNCH_CLUSTER(kern, grid, block, smem, stream, cluster, ...) \
    ::hb_launch_ex(kern, (unsigned)(grid), (unsigned)(block), (smem), (stream), (unsigned)(cluster), __VA_ARGS__)
#include <cstdlib>
#include <utility>
inline bool hb_pdl_enabled() {
    static const bool on = [] {
        const char *e = std::getenv("HEHUB_B200_PDL");
        return !(e && e[0] == '0');
    }();
    return on;
}
template <class... KArgs, class... Args>
inline cudaError_t hb_launch_ex(void (*kern)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t st,
                                unsigned cluster, Args &&...args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[2];
    unsigned na = 0;
    if (cluster > 1) {
        at[na].id = cudaLaunchAttributeClusterDimension;
        at[na].val.clusterDim.x = cluster;
        at[na].val.clusterDim.y = 1;
        at[na].val.clusterDim.z = 1;
        na++;
    }
    if (hb_pdl_enabled()) {
        at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[na].val.programmaticStreamSerializationAllowed = 1;
        na++;
    }
    cfg.attrs = at;
    cfg.numAttrs = na;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(std::forward<Args>(args))...);
}
// returns once all grids launched earlier on the stream have completed and flushed (no-op without the attribute)
// (an explicit early griddepcontrol.launch_dependents measured no better: profiles/r2c_pdl_ab.log)
__device__ __forceinline__ void hb_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// lets the next grid of the stream become resident now and run up to ITS hb_pdl_wait() (which still waits for this grid to
// complete): used by kernels whose successor has a prologue worth overlapping (ks_pair.cuh)
__device__ __forceinline__ void hb_pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// read-only operands that never alias the kernel's outputs (non-coherent path: the compiler may
// hoist these above stores), read once: no L1 allocation
__device__ __forceinline__ unsigned long long hb_ld_ro(const unsigned long long *p) {
    unsigned long long v;
    asm("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ ulonglong2 hb_ld_ro2(const unsigned long long *p) {
    ulonglong2 v;
    asm("ld.global.nc.L1::no_allocate.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p));
    return v;
}
// barrier over the thread-block cluster with release/acquire ordering of global and shared writes
__device__ __forceinline__ void hb_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void hb_cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void hb_cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
// the address of `p` (this CTA's shared memory) in CTA `rank`'s shared memory, as a shared::cluster address
__device__ __forceinline__ unsigned hb_dsmem_addr(const unsigned long long *p, unsigned rank) {
    unsigned local = (unsigned)__cvta_generic_to_shared(p), remote;
    asm("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(rank));
    return remote;
}
// one word of CTA `rank`'s shared memory at the offset `p` has in this CTA's (distributed shared memory)
__device__ __forceinline__ unsigned long long hb_ld_dsmem(const unsigned long long *p, unsigned rank) {
    unsigned long long v;
    asm volatile("ld.shared::cluster.u64 %0, [%1];" : "=l"(v) : "r"(hb_dsmem_addr(p, rank)) : "memory");
    return v;
}
__device__ __forceinline__ void hb_st_dsmem(unsigned long long *p, unsigned rank, unsigned long long v) {
    asm volatile("st.shared::cluster.u64 [%0], %1;" ::"r"(hb_dsmem_addr(p, rank)), "l"(v) : "memory");
}
// two adjacent words of CTA `rank`'s shared memory at the offset `p` has in this CTA's (distributed shared memory)
__device__ __forceinline__ ulonglong2 hb_ld_dsmem2(const unsigned long long *p, unsigned rank) {
    unsigned local = (unsigned)__cvta_generic_to_shared(p), remote;
    asm("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(rank));
    ulonglong2 v;
    asm volatile("ld.shared::cluster.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "r"(remote) : "memory");
    return v;
}
// ask L2 for a line the CTA will read much later (epilogue operands of the fused transforms)
__device__ __forceinline__ void hb_prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// pull a line the thread will read soon into L1 (twiddles of the latency plans: tables never change, so this may run before
// the programmatic-dependent-launch wait)
__device__ __forceinline__ void hb_prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
// 16 bytes global -> this CTA's shared memory without passing through registers (LDGSTS); completion: hb_cp_async_wait_all by the
// issuing thread, then a barrier before other threads read
__device__ __forceinline__ void hb_cp_async16(void *dst_shared, const void *src_global) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst_shared)), "l"(src_global) : "memory");
}
__device__ __forceinline__ void hb_cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void hb_syncwarp() { __syncwarp(); }
// row words are read once: keep them out of L1, which holds the twiddle tables
#if defined(HB_NO_STREAM_LD) // A/B builds only
__device__ __forceinline__ unsigned long long hb_ld_stream(const unsigned long long *p) { return *p; }
__device__ __forceinline__ ulonglong2 hb_ld_stream2(const unsigned long long *p) { return *reinterpret_cast<const ulonglong2 *>(p); }
#else
__device__ __forceinline__ unsigned long long hb_ld_stream(const unsigned long long *p) {
    unsigned long long v;
    asm("ld.global.L1::no_allocate.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ ulonglong2 hb_ld_stream2(const unsigned long long *p) {
    ulonglong2 v;
    asm("ld.global.L1::no_allocate.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p));
    return v;
}
#endif
template <class T>
__device__ __forceinline__ T hb_ldcg(const T *p) { return __ldcg(p); }
#endif
