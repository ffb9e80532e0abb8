// nccl_provider.cpp — the NCCL implementation of hehub_b200_collectives (include/hehub_b200.h), built into its own small
// library (hehub_b200/libhehub_b200_nccl.so) so that libhehub_b200.so itself does not depend on NCCL.
//
// One communicator per process (one process per GPU).  Rank 0 obtains the 128-byte unique id (hehub_b200_nccl_unique_id)
// and hands it to the other ranks by whatever means the launcher offers (bench.py: the process group torchrun set up; a
// pure C++ job: MPI_Bcast or a file); every rank then calls hehub_b200_nccl_create.  The sweep driver (csrc/sweep.cu) uses
// the two collectives the path needs: ncclBroadcast of the key-switch key and ncclAllGather of per-ciphertext checksums —
// both over NVLink / NVSwitch on a B200 node.  There is no data-path collective.
#include <cuda_runtime.h>
#include <nccl.h>

#include <cstdio>
#include <cstring>
#include <new>

#include "../../include/hehub_b200.h"

namespace {

struct NcclState {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1, device = 0;
};

int nccl_broadcast(void *self, void *dev_buf, size_t bytes, int root, void *stream) {
    auto *st = static_cast<NcclState *>(self);
    return ncclBroadcast(dev_buf, dev_buf, bytes, ncclUint8, root, st->comm, static_cast<cudaStream_t>(stream)) == ncclSuccess ? 0 : 1;
}

int nccl_allgather(void *self, const void *dev_send, void *dev_recv, size_t bytes_per_rank, void *stream) {
    auto *st = static_cast<NcclState *>(self);
    return ncclAllGather(dev_send, dev_recv, bytes_per_rank, ncclUint8, st->comm, static_cast<cudaStream_t>(stream)) == ncclSuccess ? 0 : 1;
}

} // namespace

extern "C" {

int hehub_b200_nccl_version(void) {
    int v = 0;
    ncclGetVersion(&v);
    return v;
}

int hehub_b200_nccl_unique_id(uint8_t id[HEHUB_B200_NCCL_ID_BYTES]) {
    static_assert(sizeof(ncclUniqueId) == HEHUB_B200_NCCL_ID_BYTES, "ncclUniqueId is 128 bytes");
    ncclUniqueId uid;
    if (ncclGetUniqueId(&uid) != ncclSuccess) return HEHUB_B200_ERR_CUDA;
    std::memcpy(id, &uid, sizeof(uid));
    return HEHUB_B200_OK;
}

int hehub_b200_nccl_create(hehub_b200_collectives *out, int rank, int world, const uint8_t id[HEHUB_B200_NCCL_ID_BYTES], int device) {
    if (!out || !id || world < 1 || rank < 0 || rank >= world) return HEHUB_B200_ERR_INVALID;
    if (cudaSetDevice(device) != cudaSuccess) return HEHUB_B200_ERR_CUDA;
    auto *st = new (std::nothrow) NcclState();
    if (!st) return HEHUB_B200_ERR_NOMEM;
    st->rank = rank;
    st->world = world;
    st->device = device;
    ncclUniqueId uid;
    std::memcpy(&uid, id, sizeof(uid));
    const ncclResult_t rc = ncclCommInitRank(&st->comm, world, uid, rank);
    if (rc != ncclSuccess) {
        std::fprintf(stderr, "hehub_b200 sweep: ncclCommInitRank failed on rank %d: %s\n", rank, ncclGetErrorString(rc));
        delete st;
        return HEHUB_B200_ERR_CUDA;
    }
    int v = 0, n = 0;
    ncclGetVersion(&v);
    ncclCommCount(st->comm, &n);
    std::fprintf(stderr, "hehub_b200 sweep: NCCL %d.%d.%d communicator, rank %d of %d, device %d (C++ driver)\n", v / 10000, (v / 100) % 100,
                 v % 100, rank, n, device);
    out->self = st;
    out->broadcast = nccl_broadcast;
    out->allgather = nccl_allgather;
    return HEHUB_B200_OK;
}

int hehub_b200_nccl_destroy(hehub_b200_collectives *c) {
    if (!c || !c->self) return HEHUB_B200_OK;
    auto *st = static_cast<NcclState *>(c->self);
    if (st->comm) ncclCommDestroy(st->comm);
    delete st;
    c->self = nullptr;
    return HEHUB_B200_OK;
}

} // extern "C"
