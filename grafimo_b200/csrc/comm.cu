// comm.cu -- the one exchange step of a multi-GPU scan, inside the C ABI.
//
// The reference funnels every process's rows through a Manager dict and merges them in the parent before the
// Benjamini-Hochberg step (src/grafimo/score_sequences.py:115-118,171-188,194-198).  Here every GPU scores its own
// shard; the global multiset of p-values is the element-wise sum of the per-GPU score histograms, so the exchange is ONE
// ncclAllReduce (sum, uint64) of span+1 counters per motif on the context's stream, and -- for the merged report table --
// an ncclAllGather of the fixed-width hit columns.  The context owns the communicator.
//
// NCCL is bound at run time (dlopen of libnccl.so.2: the copy the process already loaded, e.g. torch's, else the system
// one; GB2_NCCL_LIB overrides), so the library has no link-time dependency on it and single-GPU callers never touch it.
#include <dlfcn.h>
#include <stdlib.h>

#include "internal.cuh"

namespace {
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[GB2_NCCL_ID_BYTES]; } ncclUniqueId;
typedef int ncclResult_t;
enum { NCCL_UINT8 = 1, NCCL_UINT64 = 5, NCCL_FLOAT64 = 8 };  // ncclDataType_t (nccl.h)
enum { NCCL_SUM = 0, NCCL_MAX = 2 };                         // ncclRedOp_t

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    char why[256] = {0};
};

NcclApi *nccl_api()
{
    static NcclApi api;
    static bool tried = false;
    if (tried) return api.handle ? &api : nullptr;
    tried = true;
    const char *names[3] = {getenv("GB2_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
        if (!n || !*n) continue;
        api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (api.handle) break;
    }
    if (!api.handle) {
        snprintf(api.why, sizeof(api.why), "libnccl.so.2 not found (%s)", dlerror());
        return nullptr;
    }
#define GB2_SYM(field, name)                                                   \
    *(void **)(&api.field) = dlsym(api.handle, name);                          \
    if (!api.field) {                                                          \
        snprintf(api.why, sizeof(api.why), "NCCL symbol %s is missing", name); \
        api.handle = nullptr;                                                  \
        return nullptr;                                                        \
    }
    GB2_SYM(GetUniqueId, "ncclGetUniqueId")
    GB2_SYM(CommInitRank, "ncclCommInitRank")
    GB2_SYM(CommDestroy, "ncclCommDestroy")
    GB2_SYM(AllReduce, "ncclAllReduce")
    GB2_SYM(AllGather, "ncclAllGather")
    GB2_SYM(GetErrorString, "ncclGetErrorString")
#undef GB2_SYM
    return &api;
}
}  // namespace

#define GB2_NCCL(ctx, api, call)                                                                      \
    do {                                                                                              \
        ncclResult_t r__ = (call);                                                                    \
        if (r__ != 0) {                                                                               \
            GB2_SET_ERR(ctx, "%s:%d: %s failed: %s", __FILE__, __LINE__, #call, (api)->GetErrorString(r__)); \
            return GB2_ERR_CUDA;                                                                      \
        }                                                                                             \
    } while (0)

extern "C" int gb2_comm_unique_id(uint8_t *id)
{
    if (!id) return GB2_ERR_ARG;
    NcclApi *api = nccl_api();
    if (!api) return GB2_ERR_STATE;
    ncclUniqueId u;
    if (api->GetUniqueId(&u) != 0) return GB2_ERR_CUDA;
    memcpy(id, u.internal, GB2_NCCL_ID_BYTES);
    return GB2_OK;
}

extern "C" int gb2_comm_init(gb2_ctx *ctx, const uint8_t *id, int rank, int world)
{
    if (!ctx) return GB2_ERR_ARG;
    GB2_REQUIRE(ctx, world >= 1 && rank >= 0 && rank < world, "gb2_comm_init: rank %d outside a world of %d", rank, world);
    gb2_comm_release(ctx);
    ctx->comm_rank = rank;
    ctx->comm_world = world;
    if (world == 1) return GB2_OK;  // nothing to exchange: every collective below is a no-op
    GB2_REQUIRE(ctx, id != nullptr, "gb2_comm_init: null unique id");
    NcclApi *api = nccl_api();
    if (!api) {
        static NcclApi *probe = nullptr;
        (void)probe;
        GB2_SET_ERR(ctx, "gb2_comm_init: NCCL is not available in this process");
        return GB2_ERR_STATE;
    }
    GB2_CUDA(ctx, cudaSetDevice(ctx->device));
    ncclUniqueId u;
    memcpy(u.internal, id, GB2_NCCL_ID_BYTES);
    ncclComm_t comm = nullptr;
    GB2_NCCL(ctx, api, api->CommInitRank(&comm, world, u, rank));
    ctx->nccl_comm = comm;
    // NCCL connects its peers lazily, inside the first collective of each kind (0.5-1.5 s at 8 ranks: measured as most of the
    // first merged report table of a run).  Pay that here, where a caller expects set-up cost, not inside the first scan.
    uint64_t *d_warm = nullptr;
    GB2_CUDA(ctx, cudaMalloc(&d_warm, sizeof(uint64_t) * (size_t)(world + 1)));
    GB2_CUDA(ctx, cudaMemsetAsync(d_warm, 0, sizeof(uint64_t) * (size_t)(world + 1), ctx->stream));
    ncclResult_t r1 = api->AllReduce(d_warm, d_warm, 1, NCCL_UINT64, NCCL_SUM, comm, ctx->stream);
    ncclResult_t r2 = api->AllGather(d_warm, d_warm + 1, sizeof(uint64_t), NCCL_UINT8, comm, ctx->stream);
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_warm);
    if (r1 != 0 || r2 != 0 || e != cudaSuccess) {
        GB2_SET_ERR(ctx, "gb2_comm_init: the first collectives failed (%s / %s / %s)", api->GetErrorString(r1), api->GetErrorString(r2),
                    cudaGetErrorString(e));
        return GB2_ERR_CUDA;
    }
    return GB2_OK;
}

void gb2_comm_release(gb2_ctx *ctx)
{
    if (ctx && ctx->nccl_comm) {
        NcclApi *api = nccl_api();
        if (api) {
            cudaSetDevice(ctx->device);
            cudaStreamSynchronize(ctx->stream);
            api->CommDestroy((ncclComm_t)ctx->nccl_comm);
        }
        ctx->nccl_comm = nullptr;
    }
    if (ctx) { ctx->comm_rank = 0; ctx->comm_world = 1; }
}

extern "C" int gb2_comm_destroy(gb2_ctx *ctx)
{
    if (!ctx) return GB2_ERR_ARG;
    gb2_comm_release(ctx);
    return GB2_OK;
}

extern "C" int gb2_comm_info(const gb2_ctx *ctx, int *rank, int *world)
{
    if (!ctx) return GB2_ERR_ARG;
    if (rank) *rank = ctx->comm_rank;
    if (world) *world = ctx->comm_world;
    return GB2_OK;
}

static int allreduce(gb2_ctx *ctx, void *d_buf, int64_t n, int dtype, int op, const char *who)
{
    if (!ctx) return GB2_ERR_ARG;
    GB2_REQUIRE(ctx, n >= 0, "%s: negative count", who);
    if (ctx->comm_world == 1 || n == 0) return GB2_OK;
    GB2_REQUIRE(ctx, d_buf != nullptr, "%s: null buffer", who);
    if (!ctx->nccl_comm) {
        GB2_SET_ERR(ctx, "%s: gb2_comm_init has not been called on this context", who);
        return GB2_ERR_STATE;
    }
    NcclApi *api = nccl_api();
    GB2_CUDA(ctx, cudaSetDevice(ctx->device));
    GB2_NCCL(ctx, api, api->AllReduce(d_buf, d_buf, (size_t)n, dtype, op, (ncclComm_t)ctx->nccl_comm, ctx->stream));
    ctx->launches++;  // one NCCL kernel
    return GB2_OK;
}

extern "C" int gb2_allreduce_hist(gb2_ctx *ctx, uint64_t *d_hist, int64_t n)
{
    return allreduce(ctx, d_hist, n, NCCL_UINT64, NCCL_SUM, "gb2_allreduce_hist");
}

extern "C" int gb2_allreduce_max_f64(gb2_ctx *ctx, double *d_values, int64_t n)
{
    return allreduce(ctx, d_values, n, NCCL_FLOAT64, NCCL_MAX, "gb2_allreduce_max_f64");
}

extern "C" int gb2_allgather_bytes(gb2_ctx *ctx, const void *d_send, void *d_recv, int64_t bytes_per_rank)
{
    if (!ctx) return GB2_ERR_ARG;
    GB2_REQUIRE(ctx, bytes_per_rank >= 0, "gb2_allgather_bytes: negative size");
    if (bytes_per_rank == 0) return GB2_OK;
    GB2_REQUIRE(ctx, d_send && d_recv, "gb2_allgather_bytes: null buffer");
    GB2_CUDA(ctx, cudaSetDevice(ctx->device));
    if (ctx->comm_world == 1) {
        if (d_send != d_recv)
            GB2_CUDA(ctx, cudaMemcpyAsync(d_recv, d_send, (size_t)bytes_per_rank, cudaMemcpyDeviceToDevice, ctx->stream));
        return GB2_OK;
    }
    if (!ctx->nccl_comm) {
        GB2_SET_ERR(ctx, "gb2_allgather_bytes: gb2_comm_init has not been called on this context");
        return GB2_ERR_STATE;
    }
    NcclApi *api = nccl_api();
    GB2_NCCL(ctx, api, api->AllGather(d_send, d_recv, (size_t)bytes_per_rank, NCCL_UINT8, (ncclComm_t)ctx->nccl_comm, ctx->stream));
    ctx->launches++;
    return GB2_OK;
}
