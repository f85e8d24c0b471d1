// comm.cu -- the multi-GPU data plane of the C-ABI (include/exotica_b200.h, "collectives"): NCCL over NVLink 5 / NVSwitch.
//
// The hot path shards with no exchange (every configuration / trajectory is independent); the only
// exchange steps are the all-gather of per-trajectory costs (what AbstractTimeIndexedProblem::GetCost returns per trajectory,
// exotica_core/src/problems/abstract_time_indexed_problem.cpp:299-307) and the min-reduction of the SphereCollision distance.
// Both run on the caller's stream right behind the kernel that produced their input, so one batched call is kernel(s) +
// one collective with no host synchronisation in between, and the whole sequence can be captured into a CUDA graph
// (exo_graph_*).
//
// NCCL is resolved at run time with dlopen("libnccl.so.2"): the library itself has no link-time dependency on it (single-GPU
// users never load it), and a process that already carries an NCCL (e.g. one that imported torch) shares that copy instead of
// loading a second one.  Only the stable core API is used (ncclGetUniqueId / ncclCommInitRank / ncclAllGather / ncclAllReduce).
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>
#include <stdlib.h>
#include <string.h>

#include <string>

#include "../../include/exotica_b200.h"

extern "C" void exo_internal_set_error(const char* msg);

namespace
{
struct NcclApi
{
    void* handle = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string error;
};

NcclApi& api()
{
    static NcclApi a;
    return a;
}

int fail(int code, const std::string& msg)
{
    exo_internal_set_error(msg.c_str());
    return code;
}

int load_nccl()
{
    NcclApi& a = api();
    if (a.handle) return EXO_OK;
    if (!a.error.empty()) return fail(EXO_ERR_UNSUPPORTED, a.error);
    const char* env = getenv("EXO_NCCL_LIB");
    const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    std::string tried;
    for (const char* n : names)
    {
        if (!n || !*n) continue;
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
        tried += std::string(tried.empty() ? "" : "; ") + dlerror();
    }
    if (!h)
    {
        a.error = "NCCL is not available (" + tried + "); set EXO_NCCL_LIB to a libnccl.so.2";
        return fail(EXO_ERR_UNSUPPORTED, a.error);
    }
#define EXO_NCCL_SYM(field, name)                                        \
    a.field = (decltype(a.field))dlsym(h, name);                        \
    if (!a.field)                                                        \
    {                                                                    \
        a.error = std::string("NCCL symbol missing: ") + name;          \
        dlclose(h);                                                      \
        return fail(EXO_ERR_UNSUPPORTED, a.error);                       \
    }
    EXO_NCCL_SYM(GetVersion, "ncclGetVersion")
    EXO_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
    EXO_NCCL_SYM(CommInitRank, "ncclCommInitRank")
    EXO_NCCL_SYM(CommDestroy, "ncclCommDestroy")
    EXO_NCCL_SYM(AllGather, "ncclAllGather")
    EXO_NCCL_SYM(AllReduce, "ncclAllReduce")
    EXO_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef EXO_NCCL_SYM
    a.handle = h;
    return EXO_OK;
}

int nccl_fail(ncclResult_t r, const char* what)
{
    return fail(EXO_ERR_CUDA, std::string(what) + ": " + api().GetErrorString(r));
}

int cuda_fail(cudaError_t e, const char* what)
{
    return fail(EXO_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

#define NCCL_TRY(expr)                                            \
    do                                                            \
    {                                                             \
        ncclResult_t _r = (expr);                                 \
        if (_r != ncclSuccess) return nccl_fail(_r, #expr);       \
    } while (0)
#define CUDA_TRY(expr)                                            \
    do                                                            \
    {                                                             \
        cudaError_t _e = (expr);                                  \
        if (_e != cudaSuccess) return cuda_fail(_e, #expr);       \
    } while (0)
}  // namespace

struct exo_comm
{
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    double* d_token = nullptr;  // one double for exo_comm_barrier
    int64_t collectives = 0;
};

struct exo_graph
{
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
};

extern "C" {

int exo_comm_nccl_version(int* version)
{
    if (!version) return fail(EXO_ERR_INVALID, "version is NULL");
    int rc = load_nccl();
    if (rc != EXO_OK) return rc;
    NCCL_TRY(api().GetVersion(version));
    return EXO_OK;
}

int exo_comm_get_unique_id(unsigned char* id)
{
    if (!id) return fail(EXO_ERR_INVALID, "id is NULL");
    int rc = load_nccl();
    if (rc != EXO_OK) return rc;
    static_assert(sizeof(ncclUniqueId) == EXO_COMM_ID_BYTES, "ncclUniqueId size");
    ncclUniqueId u;
    NCCL_TRY(api().GetUniqueId(&u));
    memcpy(id, &u, sizeof(u));
    return EXO_OK;
}

int exo_comm_create(int rank, int world, const unsigned char* id, exo_comm** out)
{
    if (!out) return fail(EXO_ERR_INVALID, "out is NULL");
    if (world < 1 || rank < 0 || rank >= world) return fail(EXO_ERR_INVALID, "bad rank / world size");
    exo_comm* c = new exo_comm();
    c->rank = rank;
    c->world = world;
    if (world > 1)
    {
        if (!id)
        {
            delete c;
            return fail(EXO_ERR_INVALID, "id is NULL");
        }
        int rc = load_nccl();
        if (rc != EXO_OK)
        {
            delete c;
            return rc;
        }
        ncclUniqueId u;
        memcpy(&u, id, sizeof(u));
        ncclResult_t r = api().CommInitRank(&c->comm, world, u, rank);
        if (r != ncclSuccess)
        {
            delete c;
            return nccl_fail(r, "ncclCommInitRank");
        }
    }
    cudaError_t e = cudaMalloc((void**)&c->d_token, sizeof(double));
    if (e != cudaSuccess)
    {
        if (c->comm) api().CommDestroy(c->comm);
        delete c;
        return cuda_fail(e, "cudaMalloc");
    }
    cudaMemset(c->d_token, 0, sizeof(double));
    *out = c;
    return EXO_OK;
}

int exo_comm_destroy(exo_comm* c)
{
    if (!c) return EXO_OK;
    if (c->comm) api().CommDestroy(c->comm);
    if (c->d_token) cudaFree(c->d_token);
    delete c;
    return EXO_OK;
}

int exo_comm_rank(const exo_comm* c)
{
    return c ? c->rank : -1;
}

int exo_comm_world_size(const exo_comm* c)
{
    return c ? c->world : -1;
}

int64_t exo_comm_collective_count(const exo_comm* c)
{
    return c ? c->collectives : -1;
}

int exo_comm_allgather(exo_comm* c, const double* d_send, double* d_recv, int64_t count_per_rank, void* stream)
{
    if (!c) return fail(EXO_ERR_INVALID, "NULL communicator");
    if (count_per_rank < 0) return fail(EXO_ERR_INVALID, "negative count");
    if (count_per_rank == 0) return EXO_OK;
    if (!d_send || !d_recv) return fail(EXO_ERR_INVALID, "NULL buffer");
    cudaStream_t s = (cudaStream_t)stream;
    if (c->world == 1)
    {
        if (d_send != d_recv + (size_t)c->rank * count_per_rank)
            CUDA_TRY(cudaMemcpyAsync(d_recv, d_send, (size_t)count_per_rank * sizeof(double), cudaMemcpyDeviceToDevice, s));
        return EXO_OK;
    }
    NCCL_TRY(api().AllGather(d_send, d_recv, (size_t)count_per_rank, ncclDouble, c->comm, s));
    c->collectives++;
    return EXO_OK;
}

int exo_comm_allreduce(exo_comm* c, const double* d_send, double* d_recv, int64_t count, int32_t op, void* stream)
{
    if (!c) return fail(EXO_ERR_INVALID, "NULL communicator");
    if (count < 0) return fail(EXO_ERR_INVALID, "negative count");
    if (op != EXO_REDUCE_MIN && op != EXO_REDUCE_SUM && op != EXO_REDUCE_MAX) return fail(EXO_ERR_INVALID, "unknown reduction");
    if (count == 0) return EXO_OK;
    if (!d_send || !d_recv) return fail(EXO_ERR_INVALID, "NULL buffer");
    cudaStream_t s = (cudaStream_t)stream;
    if (c->world == 1)
    {
        if (d_send != d_recv) CUDA_TRY(cudaMemcpyAsync(d_recv, d_send, (size_t)count * sizeof(double), cudaMemcpyDeviceToDevice, s));
        return EXO_OK;
    }
    const ncclRedOp_t nop = op == EXO_REDUCE_MIN ? ncclMin : (op == EXO_REDUCE_MAX ? ncclMax : ncclSum);
    NCCL_TRY(api().AllReduce(d_send, d_recv, (size_t)count, ncclDouble, nop, c->comm, s));
    c->collectives++;
    return EXO_OK;
}

int exo_comm_barrier(exo_comm* c, void* stream)
{
    if (!c) return fail(EXO_ERR_INVALID, "NULL communicator");
    cudaStream_t s = (cudaStream_t)stream;
    if (c->world > 1)
    {
        NCCL_TRY(api().AllReduce(c->d_token, c->d_token, 1, ncclDouble, ncclSum, c->comm, s));
        c->collectives++;
    }
    CUDA_TRY(cudaStreamSynchronize(s));
    return EXO_OK;
}

// ---- events and graphs (plumbing for callers without their own CUDA runtime binding) ------------------------------------
int exo_event_create(void** event)
{
    if (!event) return fail(EXO_ERR_INVALID, "event is NULL");
    cudaEvent_t e;
    CUDA_TRY(cudaEventCreate(&e));
    *event = (void*)e;
    return EXO_OK;
}

int exo_event_destroy(void* event)
{
    CUDA_TRY(cudaEventDestroy((cudaEvent_t)event));
    return EXO_OK;
}

int exo_event_record(void* event, void* stream)
{
    CUDA_TRY(cudaEventRecord((cudaEvent_t)event, (cudaStream_t)stream));
    return EXO_OK;
}

int exo_event_synchronize(void* event)
{
    CUDA_TRY(cudaEventSynchronize((cudaEvent_t)event));
    return EXO_OK;
}

int exo_event_elapsed_ms(void* start, void* stop, float* ms)
{
    if (!ms) return fail(EXO_ERR_INVALID, "ms is NULL");
    CUDA_TRY(cudaEventElapsedTime(ms, (cudaEvent_t)start, (cudaEvent_t)stop));
    return EXO_OK;
}

int exo_graph_begin_capture(void* stream)
{
    if (!stream) return fail(EXO_ERR_INVALID, "graph capture needs a stream created with exo_stream_create (not the default stream)");
    CUDA_TRY(cudaStreamBeginCapture((cudaStream_t)stream, cudaStreamCaptureModeThreadLocal));
    return EXO_OK;
}

int exo_graph_end_capture(void* stream, exo_graph** out)
{
    if (!out) return fail(EXO_ERR_INVALID, "out is NULL");
    exo_graph* g = new exo_graph();
    cudaError_t e = cudaStreamEndCapture((cudaStream_t)stream, &g->graph);
    if (e != cudaSuccess || !g->graph)
    {
        delete g;
        return cuda_fail(e, "cudaStreamEndCapture");
    }
    e = cudaGraphInstantiate(&g->exec, g->graph, 0);
    if (e != cudaSuccess)
    {
        cudaGraphDestroy(g->graph);
        delete g;
        return cuda_fail(e, "cudaGraphInstantiate");
    }
    *out = g;
    return EXO_OK;
}

int exo_graph_launch(exo_graph* g, void* stream)
{
    if (!g) return fail(EXO_ERR_INVALID, "NULL graph");
    CUDA_TRY(cudaGraphLaunch(g->exec, (cudaStream_t)stream));
    return EXO_OK;
}

int exo_graph_destroy(exo_graph* g)
{
    if (!g) return EXO_OK;
    if (g->exec) cudaGraphExecDestroy(g->exec);
    if (g->graph) cudaGraphDestroy(g->graph);
    delete g;
    return EXO_OK;
}

}  // extern "C"
