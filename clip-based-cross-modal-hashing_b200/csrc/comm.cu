// Multi-GPU transport of the sharded paths (include/cmh_b200.h, "multi-GPU transport"): the built-in NCCL
// implementation of the cmh_comm function table.  The reference has no distributed code; the north star
// asks for the database sharded over the GPUs of one box with an NCCL exchange of the per-shard results.
//
// libnccl.so.2 is bound at run time (dlopen): the copy the process has already loaded is preferred (a PyTorch process
// has its bundled NCCL mapped; two NCCL builds in one process would each bring their own proxy threads), then the
// system library.  Nothing here links against NCCL, so the library loads - and everything single-GPU works - on a box
// without it.
#include <dlfcn.h>
#include <nccl.h>

#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace cmh {

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*);
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*GroupStart)();
    ncclResult_t (*GroupEnd)();
    const char* (*GetErrorString)(ncclResult_t);
    bool ok = false;
};

static NcclApi g_nccl;
static std::once_flag g_nccl_once;

static void load_nccl() {
    void* h = nullptr;
    const char* env = getenv("CMH_NCCL_LIB");
    if (env && env[0]) h = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);      // whatever the process already uses
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return;
    bool all = true;
#define CMH_SYM(field, name)                                                   \
    do {                                                                       \
        g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(h, name)); \
        all = all && g_nccl.field != nullptr;                                  \
    } while (0)
    CMH_SYM(GetUniqueId, "ncclGetUniqueId");
    CMH_SYM(CommInitRank, "ncclCommInitRank");
    CMH_SYM(CommInitAll, "ncclCommInitAll");
    CMH_SYM(CommDestroy, "ncclCommDestroy");
    CMH_SYM(AllReduce, "ncclAllReduce");
    CMH_SYM(AllGather, "ncclAllGather");
    CMH_SYM(Send, "ncclSend");
    CMH_SYM(Recv, "ncclRecv");
    CMH_SYM(GroupStart, "ncclGroupStart");
    CMH_SYM(GroupEnd, "ncclGroupEnd");
    CMH_SYM(GetErrorString, "ncclGetErrorString");
#undef CMH_SYM
    g_nccl.ok = all;
}

static bool nccl_ready() {
    std::call_once(g_nccl_once, load_nccl);
    return g_nccl.ok;
}

static int nccl_fail(ncclResult_t r, const char* what) {
    set_error("%s: NCCL error %d (%s)", what, (int)r, g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
    return 1000 + (int)r;     // positive: passthrough of the library's error code (1000 + ncclResult_t)
}

#define CMH_NCCL(expr)                                      \
    do {                                                    \
        ncclResult_t _r = (expr);                           \
        if (_r != ncclSuccess) return nccl_fail(_r, #expr); \
    } while (0)

struct NcclCtx {
    ncclComm_t comm;
    int rank, world;
};

static int nccl_all_reduce_u32(void* ctx, uint32_t* buf, int64_t count, int op, void* stream) {
    NcclCtx* c = static_cast<NcclCtx*>(ctx);
    if (count <= 0) return CMH_OK;
    CMH_NCCL(g_nccl.AllReduce(buf, buf, (size_t)count, ncclUint32, op == 1 ? ncclMax : ncclSum, c->comm, (cudaStream_t)stream));
    return CMH_OK;
}

static int nccl_all_gather(void* ctx, const void* send, void* recv, int64_t bytes, void* stream) {
    NcclCtx* c = static_cast<NcclCtx*>(ctx);
    if (bytes <= 0) return CMH_OK;
    CMH_NCCL(g_nccl.AllGather(send, recv, (size_t)bytes, ncclUint8, c->comm, (cudaStream_t)stream));
    return CMH_OK;
}

static int nccl_all_to_all(void* ctx, const void* send, void* recv, int64_t bytes, void* stream) {
    NcclCtx* c = static_cast<NcclCtx*>(ctx);
    if (bytes <= 0) return CMH_OK;
    const char* s = static_cast<const char*>(send);
    char* r = static_cast<char*>(recv);
    CMH_NCCL(g_nccl.GroupStart());
    for (int p = 0; p < c->world; ++p) {
        CMH_NCCL(g_nccl.Send(s + (size_t)p * bytes, (size_t)bytes, ncclUint8, p, c->comm, (cudaStream_t)stream));
        CMH_NCCL(g_nccl.Recv(r + (size_t)p * bytes, (size_t)bytes, ncclUint8, p, c->comm, (cudaStream_t)stream));
    }
    CMH_NCCL(g_nccl.GroupEnd());
    return CMH_OK;
}

// ---- loopback transport (measurement aid) ---------------------------------------------------------------------------
// One GPU plays rank `rank` of `world` statistically identical shards: a sum over the ranks is world x the local value, a
// max is the local value, gathered / exchanged blocks are copies of the local block.  Lets the per-shard GPU work of an
// N-GPU search (every launch, every small kernel, the merge of N lists) be timed on ONE GPU; results are NOT a ranking
// of any database.  scripts/shard_emul.py.
struct LoopCtx { int rank, world; };

__global__ void __launch_bounds__(256) loop_scale_kernel(uint32_t* buf, int64_t n, uint32_t factor) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) buf[i] *= factor;
}
static int loop_all_reduce_u32(void* ctx, uint32_t* buf, int64_t count, int op, void* stream) {
    LoopCtx* c = static_cast<LoopCtx*>(ctx);
    if (count <= 0 || op == 1) return CMH_OK;
    loop_scale_kernel<<<(unsigned)ceil_div(count, 256), 256, 0, (cudaStream_t)stream>>>(buf, count, (uint32_t)c->world);
    CMH_LAUNCH_CHECK("loop_scale_kernel");
    return CMH_OK;
}
static int loop_all_gather(void* ctx, const void* send, void* recv, int64_t bytes, void* stream) {
    LoopCtx* c = static_cast<LoopCtx*>(ctx);
    for (int r = 0; r < c->world; ++r) {
        char* dst = static_cast<char*>(recv) + (size_t)r * bytes;
        if (dst != send) CMH_CUDA(cudaMemcpyAsync(dst, send, (size_t)bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    }
    return CMH_OK;
}
static int loop_all_to_all(void* ctx, const void* send, void* recv, int64_t bytes, void* stream) {
    LoopCtx* c = static_cast<LoopCtx*>(ctx);
    const char* mine = static_cast<const char*>(send) + (size_t)c->rank * bytes;     // what every "other rank" would send me
    for (int r = 0; r < c->world; ++r)
        CMH_CUDA(cudaMemcpyAsync(static_cast<char*>(recv) + (size_t)r * bytes, mine, (size_t)bytes, cudaMemcpyDeviceToDevice,
                                 (cudaStream_t)stream));
    return CMH_OK;
}

static cmh_comm* wrap(ncclComm_t comm, int rank, int world) {
    NcclCtx* ctx = new NcclCtx{comm, rank, world};
    cmh_comm* c = new cmh_comm;
    c->ctx = ctx;
    c->rank = rank;
    c->world = world;
    c->all_reduce_u32 = nccl_all_reduce_u32;
    c->all_gather = nccl_all_gather;
    c->all_to_all = nccl_all_to_all;
    return c;
}

}  // namespace cmh

using namespace cmh;

extern "C" int cmh_comm_unique_id(void* id128) {
    CMH_REQUIRE(id128, CMH_ERR_ARG, "cmh_comm_unique_id: NULL buffer");
    CMH_REQUIRE(nccl_ready(), CMH_ERR_UNSUPPORTED, "cmh_comm: libnccl.so.2 could not be loaded (set CMH_NCCL_LIB)");
    static_assert(sizeof(ncclUniqueId) == CMH_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    CMH_NCCL(g_nccl.GetUniqueId(&id));
    memcpy(id128, &id, sizeof(id));
    return CMH_OK;
}

extern "C" int cmh_comm_create_rank(const void* id128, int world, int rank, cmh_comm** out) {
    CMH_REQUIRE(id128 && out && world >= 1 && rank >= 0 && rank < world, CMH_ERR_ARG, "cmh_comm_create_rank: bad arguments");
    CMH_REQUIRE(nccl_ready(), CMH_ERR_UNSUPPORTED, "cmh_comm: libnccl.so.2 could not be loaded (set CMH_NCCL_LIB)");
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ncclComm_t comm;
    CMH_NCCL(g_nccl.CommInitRank(&comm, world, id, rank));
    *out = wrap(comm, rank, world);
    return CMH_OK;
}

extern "C" int cmh_comm_create(int ndev, const int* devs, cmh_comm** out) {
    CMH_REQUIRE(ndev >= 1 && out, CMH_ERR_ARG, "cmh_comm_create: bad arguments");
    CMH_REQUIRE(nccl_ready(), CMH_ERR_UNSUPPORTED, "cmh_comm: libnccl.so.2 could not be loaded (set CMH_NCCL_LIB)");
    std::vector<ncclComm_t> comms((size_t)ndev);
    CMH_NCCL(g_nccl.CommInitAll(comms.data(), ndev, devs));
    for (int i = 0; i < ndev; ++i) out[i] = wrap(comms[(size_t)i], i, ndev);
    return CMH_OK;
}

extern "C" int cmh_comm_create_loopback(int world, int rank, cmh_comm** out) {
    CMH_REQUIRE(out && world >= 1 && rank >= 0 && rank < world, CMH_ERR_ARG, "cmh_comm_create_loopback: bad arguments");
    cmh_comm* c = new cmh_comm;
    c->ctx = new LoopCtx{rank, world};
    c->rank = rank;
    c->world = world;
    c->all_reduce_u32 = loop_all_reduce_u32;
    c->all_gather = loop_all_gather;
    c->all_to_all = loop_all_to_all;
    *out = c;
    return CMH_OK;
}

extern "C" int cmh_comm_destroy(cmh_comm* comm) {
    if (!comm) return CMH_OK;
    if (comm->all_reduce_u32 == loop_all_reduce_u32) {
        delete static_cast<LoopCtx*>(comm->ctx);
        delete comm;
        return CMH_OK;
    }
    // only transports made by this file own an NcclCtx; a caller-supplied table is the caller's to free
    if (comm->all_reduce_u32 == nccl_all_reduce_u32 && comm->ctx) {
        NcclCtx* ctx = static_cast<NcclCtx*>(comm->ctx);
        if (g_nccl.ok) g_nccl.CommDestroy(ctx->comm);
        delete ctx;
        delete comm;
    }
    return CMH_OK;
}
