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
// Messages of at most 64 KB per rank -- every message of this path -- bypass NCCL: one kernel over peer-mapped windows (NVLink P2P
// stores + flags, below); NCCL remains the rendezvous, the carrier of large messages and the fallback (EXO_COMM_PEER=0).
#include <cuda_runtime.h>
#include <math.h>
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

// ---- small messages over peer memory ---------------------------------------------------------------------------------------------
// The messages of this path are tiny (512 costs per rank in BASELINE config 3 on eight GPUs) and NCCL's latency (22-26 us per
// all-gather, measured inside the step's CUDA graph) is then most of a 60 us step.  Every rank owns a WINDOW in its own HBM --
// two message buffers [parity][rank][WIN_SLOT] and one flag per rank -- exported with cudaIpcGetMemHandle and mapped by every
// peer (NVLink / NVSwitch peer access).  One kernel per collective: the rank stores its message into every peer's window over
// NVLink -- by default as 16-byte cells that carry their own flags (ll_store below: 512-double all-gather 3.7 us at two GPUs, 8.7 us
// at eight); with EXO_COMM_LL=0 as plain doubles followed by a fence and a flag store per peer (7.4 / 11.7 us) --, waits until every
// peer's message is complete in its own window, and copies / reduces the gathered slots into the caller's buffer.  No NCCL call, no proxy thread, capturable in a CUDA
// graph (the sequence number lives in device memory).  Two parities: a rank can run at most one collective ahead of the slowest
// peer, which may still be reading the previous one.
constexpr int WIN_SLOT = 8192;   // doubles per rank and message (64 KB): larger messages go through NCCL
constexpr int WIN_MAXW = 16;     // ranks per node

struct Window
{
    double* buf;                 // [2][world][WIN_SLOT]
    unsigned long long* flag;    // [WIN_MAXW] sequence number of the last message of every rank that has fully arrived here
    unsigned long long* seq;     // [1] collectives issued by this rank (device-side so that graph replays advance it)
    int* error;                  // [1] set when a peer did not show up within the spin limit
    uint4* ll;                   // [2][world][WIN_SLOT] flag-in-payload cells {value lo, tag, value hi, tag} (see ll_store)
};
struct PeerTable
{
    double* buf[WIN_MAXW];
    unsigned long long* flag[WIN_MAXW];
    uint4* ll[WIN_MAXW];
};

// Flag-in-payload exchange (the default; EXO_COMM_LL=0 restores store + fence + flag): every double travels as one 16-byte store
// {low word, tag, high word, tag} whose tag is derived from the collective's sequence number.  NVLink delivers at least 8 bytes
// atomically, so a cell whose two tags match the expected one holds both halves of the value: the receiver polls the cells of its own
// window and needs neither the sender's system-scope fence (a round trip over NVLink) nor a separate flag store behind it -- the
// message is complete one one-way latency after it was written.  Tags never repeat within 2^32 - 1 collectives and are never 0 (the
// window starts zeroed); the two parities keep a cell from being overwritten while its previous value may still be read.
__device__ __forceinline__ void ll_store(uint4* p, double v, unsigned tag)
{
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"((unsigned)b), "r"(tag), "r"((unsigned)(b >> 32)), "r"(tag) : "memory");
}
__device__ __forceinline__ double ll_wait(const uint4* p, unsigned tag, int* error)
{
    const long long t0 = clock64();
    for (;;)
    {
        unsigned a, b, c, d;
        asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p) : "memory");
        if (b == tag && d == tag) return __longlong_as_double((long long)(((unsigned long long)c << 32) | a));
        if (clock64() - t0 > 600000000000ll)  // minutes: a dead peer (ranks may legitimately be seconds apart on the host)
        {
            *error = 1;
            return 0.0;
        }
    }
}
__device__ __forceinline__ unsigned ll_tag(unsigned long long seq)
{
    return (unsigned)(seq % 0xFFFFFFFFull) + 1u;
}

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// op < 0: all-gather into recv[world][recv_stride]; op = EXO_REDUCE_*: all-reduce into recv[count]
__global__ void __launch_bounds__(512) exo_peer_collective_kernel(Window w, PeerTable peers, const double* send, double* recv, int count, int64_t recv_stride,
                                                                   int rank, int world, int op, int ll)
{
    __shared__ unsigned long long s_seq;
    const int tid = threadIdx.x;
    if (tid == 0) s_seq = ++(*w.seq);
    __syncthreads();
    const unsigned long long seq = s_seq;
    const size_t slot = ((size_t)(seq & 1) * world + rank) * WIN_SLOT;
    if (ll)
    {
        const unsigned tag = ll_tag(seq);
        for (int k = tid; k < world * count; k += blockDim.x)
        {
            const int p = k % world, i = k / world;
            ll_store(peers.ll[p] + slot + i, send[i], tag);
        }
        __syncthreads();  // send may alias recv: all reads of the message precede the first write of the result
        const uint4* mine = w.ll + (size_t)(seq & 1) * world * WIN_SLOT;
        if (op < 0)
        {
            for (int idx = tid; idx < world * count; idx += blockDim.x)
            {
                const int r = idx / count, i = idx - r * count;
                recv[(int64_t)r * recv_stride + i] = ll_wait(mine + (size_t)r * WIN_SLOT + i, tag, w.error);
            }
        }
        else
        {
            for (int i = tid; i < count; i += blockDim.x)
            {
                double a = ll_wait(mine + i, tag, w.error);
                for (int r = 1; r < world; ++r)
                {
                    const double b = ll_wait(mine + (size_t)r * WIN_SLOT + i, tag, w.error);
                    a = op == EXO_REDUCE_MIN ? fmin(a, b) : (op == EXO_REDUCE_MAX ? fmax(a, b) : a + b);
                }
                recv[i] = a;
            }
        }
        return;
    }
    // my message into every window (my own included): thread t serves peer t % world, so all NVLink destinations are in flight at once
    for (int k = tid; k < world * count; k += blockDim.x)
    {
        const int p = k % world, i = k / world;
        peers.buf[p][slot + i] = send[i];
    }
    __syncthreads();
    // every rank's message has arrived in MY window once its flag carries this sequence number.  The wait has no short timeout: ranks
    // may be seconds apart on the host (NCCL kernels wait the same way); only a wait of minutes -- a dead peer -- gives up and
    // reports through the window's error word.
    if (tid < world)
    {
        __threadfence_system();  // the block's stores (ordered before this by the barrier) are visible system-wide before the flag
        st_release_sys(peers.flag[tid] + rank, seq);
        const long long t0 = clock64();
        while (ld_acquire_sys(w.flag + tid) < seq)
            if (clock64() - t0 > 600000000000ll)
            {
                *w.error = 1;
                break;
            }
    }
    __syncthreads();
    const double* mine = w.buf + (size_t)(seq & 1) * world * WIN_SLOT;
    if (op < 0)
    {
        for (int idx = tid; idx < world * count; idx += blockDim.x)
        {
            const int r = idx / count, i = idx - r * count;
            recv[(int64_t)r * recv_stride + i] = __ldcg(mine + (size_t)r * WIN_SLOT + i);
        }
    }
    else
    {
        for (int i = tid; i < count; i += blockDim.x)
        {
            double a = __ldcg(mine + i);
            for (int r = 1; r < world; ++r)
            {
                const double b = __ldcg(mine + (size_t)r * WIN_SLOT + i);
                a = op == EXO_REDUCE_MIN ? fmin(a, b) : (op == EXO_REDUCE_MAX ? fmax(a, b) : a + b);
            }
            recv[i] = a;
        }
    }
}

// Exchange and global result of exo_min_distance_allgather in one kernel: this rank's per-trajectory minima (padded with +inf for
// the trajectories it does not hold) into every peer's window, flags, wait, then the gathered minima copied out and reduced to the
// global minimum -- every rank holds every trajectory's value, so no second collective is needed.  Replaces fill + gather + reduce.
// (The reduction over the timesteps stays a kernel of its own: done by this one CTA it took 215 us for 2048 x 100 evaluations.)
__global__ void __launch_bounds__(512) exo_peer_min_gather_kernel(Window w, PeerTable peers, const double* traj_in, int n_local, int count, double* recv_all,
                                                                  double* global_out, int rank, int world, int ll)  // traj_in may be a slot of recv_all
{
    __shared__ unsigned long long s_seq;
    __shared__ double s_part[16];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_seq = ++(*w.seq);
    __syncthreads();
    const unsigned long long seq = s_seq;
    const size_t slot = ((size_t)(seq & 1) * world + rank) * WIN_SLOT;
    double g = INFINITY;
    if (ll)
    {
        const unsigned tag = ll_tag(seq);
        for (int k = tid; k < world * count; k += blockDim.x)
        {
            const int p = k % world, i = k / world;
            ll_store(peers.ll[p] + slot + i, i < n_local ? traj_in[i] : INFINITY, tag);
        }
        __syncthreads();  // traj_in may be a slot of recv_all
        const uint4* mine = w.ll + (size_t)(seq & 1) * world * WIN_SLOT;
        for (int idx = tid; idx < world * count; idx += blockDim.x)
        {
            const int r = idx / count, i = idx - r * count;
            const double v = ll_wait(mine + (size_t)r * WIN_SLOT + i, tag, w.error);
            if (recv_all) recv_all[(int64_t)r * count + i] = v;
            g = fmin(g, v);
        }
    }
    else
    {
        for (int k = tid; k < world * count; k += blockDim.x)
        {
            const int p = k % world, i = k / world;
            peers.buf[p][slot + i] = i < n_local ? traj_in[i] : INFINITY;
        }
        __syncthreads();
        if (tid < world)
        {
            __threadfence_system();
            st_release_sys(peers.flag[tid] + rank, seq);
            const long long t0 = clock64();
            while (ld_acquire_sys(w.flag + tid) < seq)
                if (clock64() - t0 > 600000000000ll)
                {
                    *w.error = 1;
                    break;
                }
        }
        __syncthreads();
        const double* mine = w.buf + (size_t)(seq & 1) * world * WIN_SLOT;
        for (int idx = tid; idx < world * count; idx += blockDim.x)
        {
            const int r = idx / count, i = idx - r * count;
            const double v = __ldcg(mine + (size_t)r * WIN_SLOT + i);
            if (recv_all) recv_all[(int64_t)r * count + i] = v;
            g = fmin(g, v);
        }
    }
    if (global_out)
    {
        for (int o = 16; o > 0; o >>= 1) g = fmin(g, __shfl_xor_sync(0xffffffffu, g, o));
        if (lane == 0) s_part[warp] = g;
        __syncthreads();
        if (tid == 0)
        {
            for (int k = 1; k < 16; ++k) g = fmin(g, s_part[k]);
            global_out[0] = g;
        }
    }
}

struct exo_comm
{
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    double* d_token = nullptr;  // one double for exo_comm_barrier
    int64_t collectives = 0;    // NCCL calls
    int64_t peer_collectives = 0;
    // peer-memory window (world > 1, every rank could map every other rank's window)
    bool peer_ok = false;
    bool ll = true;             // flag-in-payload exchange (ll_store) instead of store + fence + flag
    void* d_window = nullptr;
    Window win = {};
    PeerTable peers = {};
    void* opened[WIN_MAXW] = {};
};

// Maps every rank's window into every other rank (collective over the communicator).  Any failure on any rank leaves the whole
// communicator on the NCCL path: the ranks agree through an all-reduce(min) of their success flags.
static void peer_window_teardown(exo_comm* c)
{
    for (int r = 0; r < WIN_MAXW; ++r)
        if (c->opened[r])
        {
            cudaIpcCloseMemHandle(c->opened[r]);
            c->opened[r] = nullptr;
        }
    if (c->d_window) cudaFree(c->d_window);
    c->d_window = nullptr;
    c->peer_ok = false;
}

static void peer_window_setup(exo_comm* c)
{
    const char* env = getenv("EXO_COMM_PEER");
    if (c->world < 2 || c->world > WIN_MAXW || (env && atoi(env) == 0)) return;
    const size_t buf_bytes = (size_t)2 * c->world * WIN_SLOT * sizeof(double);
    const size_t ll_off = buf_bytes + (WIN_MAXW + 2) * sizeof(unsigned long long);  // a multiple of 16
    const size_t bytes = ll_off + (size_t)2 * c->world * WIN_SLOT * sizeof(uint4);
    struct Wire
    {
        cudaIpcMemHandle_t handle;
        long long ok;
    } mine;
    memset(&mine, 0, sizeof(mine));
    bool ok = cudaMalloc(&c->d_window, bytes) == cudaSuccess && cudaMemset(c->d_window, 0, bytes) == cudaSuccess;
    ok = ok && cudaIpcGetMemHandle(&mine.handle, c->d_window) == cudaSuccess;
    mine.ok = ok ? 1 : 0;
    cudaGetLastError();
    // ship the handles: all-gather of raw bytes
    std::string all((size_t)c->world * sizeof(Wire), '\0');
    unsigned char* d_wire = nullptr;
    bool shipped = cudaMalloc((void**)&d_wire, all.size()) == cudaSuccess;
    shipped = shipped && cudaMemcpy(d_wire + (size_t)c->rank * sizeof(Wire), &mine, sizeof(Wire), cudaMemcpyHostToDevice) == cudaSuccess;
    shipped = shipped && api().AllGather(d_wire + (size_t)c->rank * sizeof(Wire), d_wire, sizeof(Wire), ncclChar, c->comm, nullptr) == ncclSuccess;
    shipped = shipped && cudaStreamSynchronize(nullptr) == cudaSuccess;
    shipped = shipped && cudaMemcpy(&all[0], d_wire, all.size(), cudaMemcpyDeviceToHost) == cudaSuccess;
    if (d_wire) cudaFree(d_wire);
    double good = shipped ? 1.0 : 0.0;
    if (shipped)
    {
        const Wire* w = reinterpret_cast<const Wire*>(all.data());
        for (int r = 0; r < c->world; ++r) good = w[r].ok ? good : 0.0;
        for (int r = 0; r < c->world && good > 0.0; ++r)
        {
            void* base = c->d_window;
            if (r != c->rank)
            {
                if (cudaIpcOpenMemHandle(&base, w[r].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess)
                {
                    cudaGetLastError();
                    good = 0.0;
                    break;
                }
                c->opened[r] = base;
            }
            c->peers.buf[r] = static_cast<double*>(base);
            c->peers.flag[r] = reinterpret_cast<unsigned long long*>(static_cast<char*>(base) + buf_bytes);
            c->peers.ll[r] = reinterpret_cast<uint4*>(static_cast<char*>(base) + ll_off);
        }
    }
    // every rank mapped every window?  (collective even when this rank failed early: the others are waiting in it)
    bool agreed = cudaMemcpy(c->d_token, &good, sizeof(double), cudaMemcpyHostToDevice) == cudaSuccess;
    agreed = agreed && api().AllReduce(c->d_token, c->d_token, 1, ncclDouble, ncclMin, c->comm, nullptr) == ncclSuccess;
    agreed = agreed && cudaStreamSynchronize(nullptr) == cudaSuccess;
    agreed = agreed && cudaMemcpy(&good, c->d_token, sizeof(double), cudaMemcpyDeviceToHost) == cudaSuccess;
    const double zero = 0.0;
    cudaMemcpy(c->d_token, &zero, sizeof(double), cudaMemcpyHostToDevice);
    if (!agreed || good < 1.0)
    {
        peer_window_teardown(c);
        return;
    }
    char* base = static_cast<char*>(c->d_window);
    c->win.buf = reinterpret_cast<double*>(base);
    c->win.flag = reinterpret_cast<unsigned long long*>(base + buf_bytes);
    c->win.seq = c->win.flag + WIN_MAXW;
    c->win.error = reinterpret_cast<int*>(c->win.flag + WIN_MAXW + 1);
    c->win.ll = reinterpret_cast<uint4*>(base + ll_off);
    const char* ll_env = getenv("EXO_COMM_LL");
    c->ll = !(ll_env && atoi(ll_env) == 0);
    c->peer_ok = true;
}

// One kernel over the peer windows instead of an NCCL call; false: the message does not qualify
static bool peer_collective(exo_comm* c, const double* d_send, double* d_recv, int64_t count, int64_t recv_stride, int op, cudaStream_t s, cudaError_t& err)
{
    if (!c->peer_ok || count < 1 || count > WIN_SLOT) return false;
    exo_peer_collective_kernel<<<1, 512, 0, s>>>(c->win, c->peers, d_send, d_recv, (int)count, recv_stride, c->rank, c->world, op, c->ll ? 1 : 0);
    err = cudaGetLastError();
    c->peer_collectives++;
    return true;
}

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
    if (world > 1) peer_window_setup(c);
    *out = c;
    return EXO_OK;
}

int exo_comm_destroy(exo_comm* c)
{
    if (!c) return EXO_OK;
    if (c->d_window)
    {
        cudaDeviceSynchronize();  // no peer collective of this rank is still writing into a window
        peer_window_teardown(c);
    }
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

// used by exo_min_distance_allgather (capi.cu) on this rank's per-trajectory minima: EXO_OK when the fused peer kernel took the call,
// EXO_ERR_UNSUPPORTED (no error text) when the communicator has no peer windows or the message does not fit them -- the caller then
// runs the unfused sequence
int exo_internal_min_gather(exo_comm* c, const double* d_min_traj_local, int64_t n_traj_local, int64_t n_traj_per_rank, double* d_min_traj_all,
                            double* d_min_global, void* stream)
{
    if (!c || !c->peer_ok || n_traj_per_rank < 1 || n_traj_per_rank > WIN_SLOT || n_traj_local > n_traj_per_rank) return EXO_ERR_UNSUPPORTED;
    exo_peer_min_gather_kernel<<<1, 512, 0, (cudaStream_t)stream>>>(c->win, c->peers, d_min_traj_local, (int)n_traj_local, (int)n_traj_per_rank, d_min_traj_all,
                                                                     d_min_global, c->rank, c->world, c->ll ? 1 : 0);
    CUDA_TRY(cudaGetLastError());
    c->peer_collectives++;
    return EXO_OK;
}

int64_t exo_comm_peer_collective_count(const exo_comm* c)
{
    if (!c) return -1;
    if (c->world > 1 && !c->peer_ok) return -1;
    return c->peer_collectives;
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
    cudaError_t ke = cudaSuccess;
    if (peer_collective(c, d_send, d_recv, count_per_rank, count_per_rank, -1, s, ke))
    {
        CUDA_TRY(ke);
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
    cudaError_t ke = cudaSuccess;
    if (peer_collective(c, d_send, d_recv, count, 0, op, s, ke))
    {
        CUDA_TRY(ke);
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
