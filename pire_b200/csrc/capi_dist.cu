// capi_dist.cu -- the multi-GPU form of the hot path behind the C ABI.
//
// The path shards by string: every Run() depends only on its own bytes and the replicated scanner tables, so there
// is no data-path collective.  One process (or thread) per GPU scans its contiguous shard of the batch -- shard
// boundaries are multiples of 32 strings, so bitmap words never straddle ranks -- straight into its slot of the
// full-length match bitmap, and ONE in-place ncclAllGather of the equal-sized slots makes the bitmap complete on
// every rank: 1/N of the bytes of an all-reduce of the zero-initialised bitmap, no zeroing, same result (the
// shards are disjoint and word aligned, so SUM == OR == concatenation).  The exchange runs on the communicator's
// own stream; with PIRE_GPU_RUN_ASYNC_EXCHANGE it overlaps whatever the caller enqueues next (the next batch's
// scan) until pire_gpu_comm_wait joins it.
//
// NCCL is bound at run time (dlopen of libnccl.so.2 -- in a process that already loaded NCCL, e.g. through
// PyTorch, that very copy is used), so libpire_b200.so itself loads on boxes without NCCL and single-GPU users
// carry no dependency.
#include "capi_internal.hpp"

#include <dlfcn.h>
#include <nccl.h>

#include <cstring>

using namespace pire_b200;

namespace {

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommCount)(const ncclComm_t, int*) = nullptr;
    ncclResult_t (*CommUserRank)(const ncclComm_t, int*) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string error;
    bool ok = false;
};

const NcclApi& Nccl()
{
    static const NcclApi api = [] {
        NcclApi a;
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);      // the copy this process already uses
        if (!h)
            h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h)
            h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) {
            a.error = std::string("NCCL is not available: ") + dlerror();
            return a;
        }
        a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
        a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
        a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
        a.AllGather = reinterpret_cast<decltype(a.AllGather)>(dlsym(h, "ncclAllGather"));
        a.CommCount = reinterpret_cast<decltype(a.CommCount)>(dlsym(h, "ncclCommCount"));
        a.CommUserRank = reinterpret_cast<decltype(a.CommUserRank)>(dlsym(h, "ncclCommUserRank"));
        a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
        a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllGather && a.CommCount && a.CommUserRank && a.GetErrorString;
        if (!a.ok)
            a.error = "NCCL is not available: libnccl.so.2 lacks an expected symbol";
        return a;
    }();
    return api;
}

int FailNccl(ncclResult_t r, const char* where)
{
    return Fail(PIRE_GPU_ECUDA, std::string(where) + ": NCCL: " + Nccl().GetErrorString(r));
}

#define NCCL_TRY(expr)                          \
    do {                                        \
        ncclResult_t r__ = (expr);              \
        if (r__ != ncclSuccess)                 \
            return FailNccl(r__, #expr);        \
    } while (0)

} // namespace

struct pire_gpu_comm {
    ncclComm_t comm = nullptr;
    bool owned = false;
    int world = 1, rank = 0, device = 0;
    cudaStream_t stream = nullptr;      // the exchange runs here
    cudaEvent_t scanned = nullptr;      // the caller's stream has written this rank's slot
    cudaEvent_t exchanged = nullptr;    // the gathered bitmap is complete
    bool pending = false;
};

extern "C" {

void pire_gpu_shard_bounds(uint64_t n_global, int world, int rank, uint64_t* lo, uint64_t* hi)
{
    if (world < 1)
        world = 1;
    uint64_t per = (n_global + (uint64_t) world - 1) / (uint64_t) world;
    per = (per + 31) / 32 * 32;
    const uint64_t l = (uint64_t) rank * per;         // on the 32-string grid even when the shard is empty
    uint64_t h = l + per < n_global ? l + per : n_global;
    if (h < l)
        h = l;
    if (lo)
        *lo = l;
    if (hi)
        *hi = h;
}

uint64_t pire_gpu_sharded_words(uint64_t n_global, int world)
{
    if (world < 1)
        world = 1;
    uint64_t per = (n_global + (uint64_t) world - 1) / (uint64_t) world;
    per = (per + 31) / 32 * 32;
    return per / 32 * (uint64_t) world;
}

int pire_gpu_comm_get_id(void* id_out)
{
    if (!id_out)
        return Fail(PIRE_GPU_EINVAL, "null id buffer");
    if (!Nccl().ok)
        return Fail(PIRE_GPU_EUNSUPPORTED, Nccl().error);
    ncclUniqueId id;
    NCCL_TRY(Nccl().GetUniqueId(&id));
    static_assert(sizeof(id) == PIRE_GPU_COMM_ID_BYTES, "ncclUniqueId size");
    std::memcpy(id_out, &id, sizeof(id));
    return PIRE_GPU_OK;
}

static int FinishComm(pire_gpu_comm* c, pire_gpu_comm** out)
{
    cudaError_t ce = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (ce == cudaSuccess)
        ce = cudaEventCreateWithFlags(&c->scanned, cudaEventDisableTiming);
    if (ce == cudaSuccess)
        ce = cudaEventCreateWithFlags(&c->exchanged, cudaEventDisableTiming);
    if (ce != cudaSuccess) {
        pire_gpu_comm_destroy(c);
        return FailCuda(ce, "pire_gpu_comm: stream/event creation");
    }
    *out = c;
    return PIRE_GPU_OK;
}

int pire_gpu_comm_create(const void* id_bytes, int world, int rank, int device, pire_gpu_comm** out)
{
    if (!out || !id_bytes || world < 1 || rank < 0 || rank >= world || device < 0)
        return Fail(PIRE_GPU_EINVAL, "bad communicator arguments");
    *out = nullptr;
    if (!Nccl().ok)
        return Fail(PIRE_GPU_EUNSUPPORTED, Nccl().error);
    CUDA_TRY(cudaSetDevice(device));
    pire_gpu_comm* c = new (std::nothrow) pire_gpu_comm;
    if (!c)
        return Fail(PIRE_GPU_EINVAL, "out of memory");
    c->world = world;
    c->rank = rank;
    c->device = device;
    c->owned = true;
    ncclUniqueId id;
    std::memcpy(&id, id_bytes, sizeof(id));
    ncclResult_t r = Nccl().CommInitRank(&c->comm, world, id, rank);
    if (r != ncclSuccess) {
        delete c;
        return FailNccl(r, "ncclCommInitRank");
    }
    return FinishComm(c, out);
}

int pire_gpu_comm_adopt(void* nccl_comm, int device, pire_gpu_comm** out)
{
    if (!out || !nccl_comm || device < 0)
        return Fail(PIRE_GPU_EINVAL, "bad communicator arguments");
    *out = nullptr;
    if (!Nccl().ok)
        return Fail(PIRE_GPU_EUNSUPPORTED, Nccl().error);
    CUDA_TRY(cudaSetDevice(device));
    pire_gpu_comm* c = new (std::nothrow) pire_gpu_comm;
    if (!c)
        return Fail(PIRE_GPU_EINVAL, "out of memory");
    c->comm = static_cast<ncclComm_t>(nccl_comm);
    c->owned = false;
    c->device = device;
    ncclResult_t r = Nccl().CommCount(c->comm, &c->world);
    if (r == ncclSuccess)
        r = Nccl().CommUserRank(c->comm, &c->rank);
    if (r != ncclSuccess) {
        delete c;
        return FailNccl(r, "ncclCommCount/ncclCommUserRank");
    }
    return FinishComm(c, out);
}

void pire_gpu_comm_destroy(pire_gpu_comm* c)
{
    if (!c)
        return;
    cudaSetDevice(c->device);
    if (c->stream)
        cudaStreamSynchronize(c->stream);
    if (c->owned && c->comm && Nccl().ok)
        Nccl().CommDestroy(c->comm);
    if (c->scanned)
        cudaEventDestroy(c->scanned);
    if (c->exchanged)
        cudaEventDestroy(c->exchanged);
    if (c->stream)
        cudaStreamDestroy(c->stream);
    delete c;
}

int pire_gpu_comm_info(const pire_gpu_comm* c, int* world, int* rank)
{
    if (!c)
        return Fail(PIRE_GPU_EINVAL, "null communicator");
    if (world)
        *world = c->world;
    if (rank)
        *rank = c->rank;
    return PIRE_GPU_OK;
}

int pire_gpu_comm_wait(pire_gpu_comm* c, void* stream)
{
    if (!c)
        return Fail(PIRE_GPU_EINVAL, "null communicator");
    if (!c->pending)
        return PIRE_GPU_OK;
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaStreamWaitEvent(static_cast<cudaStream_t>(stream), c->exchanged, 0));
    return PIRE_GPU_OK;
}

// The exchange alone: this rank's slot of d_match_bits_all has been written on `stream` (by whatever means, e.g.
// uploaded after pire_gpu_run_batch_host); gather every rank's slot.
int pire_gpu_comm_gather_bits(pire_gpu_comm* c, uint64_t n_global, uint32_t* d_match_bits_all, uint32_t flags, void* stream)
{
    if (!c || !d_match_bits_all)
        return Fail(PIRE_GPU_EINVAL, "null communicator or bitmap");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CUDA_TRY(cudaSetDevice(c->device));
    if (c->world <= 1)
        return PIRE_GPU_OK;
    const uint64_t words_per = pire_gpu_sharded_words(n_global, c->world) / (uint64_t) c->world;
    uint32_t* slot = d_match_bits_all + (size_t) c->rank * words_per;
    CUDA_TRY(cudaEventRecord(c->scanned, st));
    CUDA_TRY(cudaStreamWaitEvent(c->stream, c->scanned, 0));
    NCCL_TRY(Nccl().AllGather(slot, d_match_bits_all, (size_t) words_per, ncclUint32, c->comm, c->stream));
    CUDA_TRY(cudaEventRecord(c->exchanged, c->stream));
    c->pending = true;
    if (!(flags & PIRE_GPU_RUN_ASYNC_EXCHANGE))
        CUDA_TRY(cudaStreamWaitEvent(st, c->exchanged, 0));
    return PIRE_GPU_OK;
}

int pire_gpu_run_sharded(const pire_gpu_scanner* sc, pire_gpu_comm* c, const uint8_t* d_corpus, const uint64_t* d_offsets,
                         uint64_t fixed_len, uint64_t n_global, uint32_t flags, uint32_t* d_match_bits_all,
                         uint32_t* d_accept_masks, uint32_t* d_state_idx, void* stream)
{
    int rc = CheckRunnable(sc);
    if (rc != PIRE_GPU_OK)
        return rc;
    if (!c)
        return Fail(PIRE_GPU_EINVAL, "null communicator");
    if (c->device != sc->device)
        return Fail(PIRE_GPU_EINVAL, "scanner and communicator live on different devices");
    if (!d_match_bits_all)
        return Fail(PIRE_GPU_EINVAL, "the sharded run gathers the match bitmap: d_match_bits_all must not be null");
    const bool async = (flags & PIRE_GPU_RUN_ASYNC_EXCHANGE) != 0;
    flags &= ~PIRE_GPU_RUN_ASYNC_EXCHANGE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CUDA_TRY(cudaSetDevice(sc->device));
    uint64_t lo, hi;
    pire_gpu_shard_bounds(n_global, c->world, c->rank, &lo, &hi);
    const uint64_t n_local = hi - lo;
    const uint64_t words_per = pire_gpu_sharded_words(n_global, c->world) / (uint64_t) c->world;
    uint32_t* slot = d_match_bits_all + (size_t) c->rank * words_per;
    const uint64_t valid_words = (n_local + 31) / 32;
    // an earlier exchange may still be reading this buffer: the scan that overwrites the slot comes after it
    if (c->pending)
        CUDA_TRY(cudaStreamWaitEvent(st, c->exchanged, 0));
    if (valid_words < words_per)            // the slack of a short (or empty) last shard
        CUDA_TRY(cudaMemsetAsync(slot + valid_words, 0, (size_t) (words_per - valid_words) * 4, st));
    rc = pire_gpu_run_batch(sc, d_corpus, d_offsets, fixed_len, n_local, flags, slot, d_accept_masks, d_state_idx, st);
    if (rc != PIRE_GPU_OK)
        return rc;
    if (c->world > 1) {
        CUDA_TRY(cudaEventRecord(c->scanned, st));
        CUDA_TRY(cudaStreamWaitEvent(c->stream, c->scanned, 0));
        NCCL_TRY(Nccl().AllGather(slot, d_match_bits_all, (size_t) words_per, ncclUint32, c->comm, c->stream));
        CUDA_TRY(cudaEventRecord(c->exchanged, c->stream));
        c->pending = true;
        if (!async)
            CUDA_TRY(cudaStreamWaitEvent(st, c->exchanged, 0));
    }
    return PIRE_GPU_OK;
}

} // extern "C"
