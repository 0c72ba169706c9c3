// psdr-b200: the one exchange step of the path on several GPUs — a sum (or gather) of the per-rank film and ONE sum of the flat
// gradient vector at the end of renderD's reverse pass (BASELINE.json north_star; the reference is single-GPU) —
// enqueued with NCCL on the context's stream, i.e. directly behind the last adjoint kernel, no host synchronisation in between.
//
// NCCL is resolved at run time (dlopen of the libnccl.so.2 already mapped into the process by torch, else the system one), so the
// library builds and loads on a box without NCCL and single-GPU users never touch it. The handful of prototypes below are NCCL 2.x's
// stable C ABI (nccl.h: ncclGetUniqueId / ncclCommInitRank / ncclAllReduce / ncclAllGather / ncclCommDestroy / ncclGetErrorString).
#include <dlfcn.h>

#include <cstring>
#include <mutex>

#include "../../include/psdr_b200.h"
#include "pb_host.h"

namespace {

typedef struct ncclComm *ncclComm_t;
struct ncclUniqueId { char internal[128]; };
enum { ncclSuccess = 0 };
enum { ncclFloat32 = 7 };
enum { ncclSum = 0 };

struct NcclApi {
    void *handle = nullptr;
    int (*GetUniqueId)(ncclUniqueId *) = nullptr;
    int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*CommCount)(const ncclComm_t, int *) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};

NcclApi &nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);   // the copy torch.distributed already uses, if any
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return;
        api.handle = h;
        api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
        api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
        api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
        api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(dlsym(h, "ncclAllReduce"));
        api.AllGather = reinterpret_cast<decltype(api.AllGather)>(dlsym(h, "ncclAllGather"));
        api.CommCount = reinterpret_cast<decltype(api.CommCount)>(dlsym(h, "ncclCommCount"));
        api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
    });
    return api;
}

NcclApi &require_nccl() {
    NcclApi &a = nccl();
    if (!a.handle || !a.GetUniqueId || !a.CommInitRank || !a.AllReduce || !a.AllGather || !a.CommDestroy)
        throw pb::Error("NCCL is not available: libnccl.so.2 could not be loaded (multi-GPU calls need it; single-GPU rendering does not)");
    return a;
}

void check(NcclApi &a, int rc, const char *what) {
    if (rc != ncclSuccess) throw pb::Error(std::string("NCCL error in ") + what + ": " + (a.GetErrorString ? a.GetErrorString(rc) : "?"));
}

template <class F> int guard(pb_ctx *c, F &&f) {
    if (!c) return 1;   // every call here needs a context
    try { f(); return 0; }
    catch (const std::exception &e) { if (c) c->error = e.what(); return 1; }
}

}  // namespace

extern "C" {

int pb_dist_available(void) {
    NcclApi &a = nccl();
    return (a.handle && a.AllReduce) ? 1 : 0;
}

int pb_dist_unique_id(pb_ctx *c, char *id128) {
    return guard(c, [&] {
        PB_ASSERT_MSG(id128, "Null argument");
        NcclApi &a = require_nccl();
        ncclUniqueId id;
        check(a, a.GetUniqueId(&id), "ncclGetUniqueId");
        std::memcpy(id128, id.internal, sizeof(id.internal));
    });
}

int pb_dist_init(pb_ctx *c, const char *id128, int rank, int world) {
    return guard(c, [&] {
        PB_ASSERT_MSG(id128 && world >= 1 && rank >= 0 && rank < world, "Invalid arguments");
        NcclApi &a = require_nccl();
        PB_CUDA(cudaSetDevice(c->device));
        if (c->nccl_comm && c->nccl_owned) a.CommDestroy(static_cast<ncclComm_t>(c->nccl_comm));
        c->nccl_comm = nullptr; c->nccl_owned = false;
        ncclUniqueId id;
        std::memcpy(id.internal, id128, sizeof(id.internal));
        ncclComm_t comm = nullptr;
        check(a, a.CommInitRank(&comm, world, id, rank), "ncclCommInitRank");
        c->nccl_comm = comm; c->nccl_owned = true;
        c->rank = rank; c->world = world; c->retained_valid = false;
    });
}

int pb_dist_adopt_comm(pb_ctx *c, void *nccl_comm, int rank, int world) {
    return guard(c, [&] {
        PB_ASSERT_MSG(world >= 1 && rank >= 0 && rank < world, "Invalid arguments");
        NcclApi &a = require_nccl();
        if (c->nccl_comm && c->nccl_owned) a.CommDestroy(static_cast<ncclComm_t>(c->nccl_comm));
        c->nccl_comm = nccl_comm; c->nccl_owned = false;
        c->rank = rank; c->world = world; c->retained_valid = false;
    });
}

int pb_dist_finalize(pb_ctx *c) {
    return guard(c, [&] {
        if (c->nccl_comm && c->nccl_owned) {
            PB_CUDA(cudaSetDevice(c->device));
            PB_CUDA(cudaStreamSynchronize(c->stream));
            require_nccl().CommDestroy(static_cast<ncclComm_t>(c->nccl_comm));
        }
        c->nccl_comm = nullptr; c->nccl_owned = false;
    });
}

int pb_allreduce_grads(pb_ctx *c, float *d_grad, int64_t count) {
    return guard(c, [&] {
        PB_ASSERT_MSG(d_grad || count == 0, "Null argument");
        if (c->world <= 1 || count <= 0) return;   // one GPU: the vector is already complete
        PB_ASSERT_MSG(c->nccl_comm, "pb_allreduce_grads: no communicator (pb_dist_init / pb_dist_adopt_comm)");
        NcclApi &a = require_nccl();
        PB_CUDA(cudaSetDevice(c->device));
        check(a, a.AllReduce(d_grad, d_grad, (size_t)count, ncclFloat32, ncclSum, static_cast<ncclComm_t>(c->nccl_comm), c->stream), "ncclAllReduce");
        c->collectives++;
    });
}

int pb_allreduce_image(pb_ctx *c, float *d_image) {
    return guard(c, [&] {
        PB_ASSERT_MSG(d_image, "Null argument");
        if (c->world <= 1) return;
        PB_ASSERT_MSG(c->nccl_comm, "pb_allreduce_image: no communicator (pb_dist_init / pb_dist_adopt_comm)");
        NcclApi &a = require_nccl();
        PB_CUDA(cudaSetDevice(c->device));
        ncclComm_t comm = static_cast<ncclComm_t>(c->nccl_comm);
        const int64_t row = (int64_t)c->width * 3;
        const int tile = c->tile_rows > 0 ? c->tile_rows : (c->height + c->world - 1) / c->world;
        if (c->shard_mode == PB_SHARD_PIXELS && (int64_t)tile * c->world == c->height) {
            // contiguous row blocks of equal size: every rank's block is already in place, one in-place all-gather moves each byte once
            check(a, a.AllGather(d_image + (int64_t)c->rank * tile * row, d_image, (size_t)(tile * row), ncclFloat32, comm, c->stream), "ncclAllGather");
        } else {
            // sample shards (partial sums of every pixel) or interleaved tiles (zeros outside the own rows: x + 0 is exact)
            check(a, a.AllReduce(d_image, d_image, (size_t)(c->height * row), ncclFloat32, ncclSum, comm, c->stream), "ncclAllReduce");
        }
        c->collectives++;
    });
}

int64_t pb_stats_collectives(pb_ctx *c) { return c->collectives; }

}  // extern "C"
