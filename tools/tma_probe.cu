// tma_probe.cu -- can shared-memory staging + TMA tensor stores (cp.async.bulk.tensor.2d.global.shared::cta) beat per-lane
// st.global.cs for the SoA output pattern of the hot path?  Two patterns:
//   chain: 7 rows read, 48 rows [E] written per evaluation (cfg2a, E = 2^20)
//   tree : 38 rows read, 936 rows written of which 576 are structural zeros (cfg4, E = 2^18); the zero rows come from one
//          shared zero tile that is stored again and again, the others are staged
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tma_probe tools/tma_probe.cu ; run under gpurun.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x)                                                                                   \
    do                                                                                          \
    {                                                                                           \
        cudaError_t e_ = (x);                                                                   \
        if (e_ != cudaSuccess)                                                                  \
        {                                                                                       \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);     \
            exit(1);                                                                            \
        }                                                                                       \
    } while (0)

typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiled g_encode;

static CUtensorMap make_map(double* base, uint64_t n, uint64_t rows, uint32_t box_n, uint32_t box_rows)
{
    CUtensorMap m;
    cuuint64_t dims[2] = {n, rows};
    cuuint64_t strides[1] = {n * 8};
    cuuint32_t box[2] = {box_n, box_rows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = g_encode(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                          CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
    {
        printf("cuTensorMapEncodeTiled failed: %d\n", (int)r);
        exit(1);
    }
    return m;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(m), "r"(smem_u32(smem)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_2d_hint(const CUtensorMap* m, const void* smem, int c0, int c1, uint64_t pol)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3}], [%1], %4;" ::"l"(m), "r"(smem_u32(smem)), "r"(c0),
                 "r"(c1), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ uint64_t policy_evict_first()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}

// baseline: per-lane streaming stores
template <int ROWS_IN, int ROWS_OUT>
__global__ void soa_stg(const double* __restrict__ in, double* __restrict__ out, int64_t n)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    double s = 0;
#pragma unroll
    for (int k = 0; k < ROWS_IN; ++k) s += in[k * n + e];
#pragma unroll 8
    for (int k = 0; k < ROWS_OUT; ++k) __stcs(out + k * n + e, s + k);
}

// pull a whole input array into L2 (one prefetch per 128-byte line), keeping it in preference to streaming lines (KEEP)
template <int KEEP>
__global__ void prefetch_all(const char* __restrict__ in, int64_t bytes)
{
    int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 128;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x * 128;
    for (; i < bytes; i += stride)
    {
        if (KEEP) asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(in + i));
        else asm volatile("prefetch.global.L2 [%0];" ::"l"(in + i));
    }
}
// soa_stg whose first CTAs prefetch the whole input first (reads bunched at the start of the launch)
template <int ROWS_IN, int ROWS_OUT>
__global__ void soa_stg_prologue(const double* __restrict__ in, double* __restrict__ out, int64_t n, int pf_ctas)
{
    if ((int)blockIdx.x < pf_ctas)
    {
        const int64_t bytes = (int64_t)ROWS_IN * n * 8;
        const char* base = reinterpret_cast<const char*>(in);
        for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 128; i < bytes; i += (int64_t)pf_ctas * blockDim.x * 128)
            asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(base + i));
    }
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    double s = 0;
#pragma unroll
    for (int k = 0; k < ROWS_IN; ++k) s += in[k * n + e];
#pragma unroll 8
    for (int k = 0; k < ROWS_OUT; ++k) __stcs(out + k * n + e, s + k);
}

// chain pattern through shared memory: CTA of TPB lanes stages [ROWS_OUT][TPB] and one lane stores the tile with one TMA op
template <int TPB, int ROWS_IN, int ROWS_OUT, int HINT>
__global__ void __launch_bounds__(TPB) soa_tma(const double* __restrict__ in, const __grid_constant__ CUtensorMap map, int64_t n)
{
    extern __shared__ __align__(128) double tile[];
    const int64_t e0 = (int64_t)blockIdx.x * TPB;
    const int64_t e = e0 + threadIdx.x;
    double s = 0;
    if (e < n)
    {
#pragma unroll
        for (int k = 0; k < ROWS_IN; ++k) s += in[k * n + e];
    }
#pragma unroll 8
    for (int k = 0; k < ROWS_OUT; ++k) tile[k * TPB + threadIdx.x] = s + k;
    fence_async();
    __syncthreads();
    if (threadIdx.x == 0)
    {
        if (HINT)
            tma_store_2d_hint(&map, tile, (int)e0, 0, policy_evict_first());
        else
            tma_store_2d(&map, tile, (int)e0, 0);
        tma_commit();
        tma_wait_read0();
    }
}

// per-warp tiles: every warp stages [ROWS_OUT][32] in its own buffer and its lane 0 issues the store (no CTA barrier)
template <int WARPS, int ROWS_IN, int ROWS_OUT>
__global__ void __launch_bounds__(WARPS * 32) soa_tma_warp(const double* __restrict__ in, const __grid_constant__ CUtensorMap map, int64_t n)
{
    extern __shared__ __align__(128) double tile[];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* t = tile + w * ROWS_OUT * 32;
    const int64_t e0 = ((int64_t)blockIdx.x * WARPS + w) * 32;
    const int64_t e = e0 + lane;
    double s = 0;
    if (e < n)
    {
#pragma unroll
        for (int k = 0; k < ROWS_IN; ++k) s += in[k * n + e];
    }
#pragma unroll 8
    for (int k = 0; k < ROWS_OUT; ++k) t[k * 32 + lane] = s + k;
    fence_async();
    __syncwarp();
    if (lane == 0 && e0 < n)
    {
        tma_store_2d(&map, t, (int)e0, 0);
        tma_commit();
        tma_wait_read0();
    }
}

// tree pattern, per-lane stores: all 936 rows by st.global.cs
__global__ void tree_stg(const double* __restrict__ in, double* __restrict__ out, int64_t n, int rows_in, int rows_out)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    double s = 0;
    for (int k = 0; k < rows_in; ++k) s += in[k * n + e];
    for (int k = 0; k < rows_out; ++k) __stcs(out + (int64_t)k * n + e, (k % 13) < 5 ? s + k : 0.0);
}

// tree pattern with TMA: rows come in runs; run r covers rows [start, start + len): even runs are data (staged, CHUNK rows at a time),
// odd runs are zeros (stored from a shared zero tile of ZROWS rows, last box overlapping the previous one)
struct Runs
{
    int n_runs;
    int start[64], len[64];
};
template <int TPB, int CHUNK, int ZROWS>
__global__ void __launch_bounds__(TPB) tree_tma(const double* __restrict__ in, const __grid_constant__ CUtensorMap map_data,
                                                const __grid_constant__ CUtensorMap map_zero, const __grid_constant__ CUtensorMap map_one,
                                                const __grid_constant__ Runs runs, int64_t n, int rows_in)
{
    extern __shared__ __align__(128) double sm[];
    double* zero = sm;                  // [ZROWS][TPB]
    double* stage = sm + ZROWS * TPB;   // [CHUNK][TPB]
    const int64_t e0 = (int64_t)blockIdx.x * TPB;
    const int64_t e = e0 + threadIdx.x;
    for (int k = 0; k < ZROWS; ++k) zero[k * TPB + threadIdx.x] = 0.0;
    double s = 0;
    if (e < n)
        for (int k = 0; k < rows_in; ++k) s += in[k * n + e];
    fence_async();
    __syncthreads();
    for (int r = 0; r < runs.n_runs; ++r)
    {
        const int start = runs.start[r], len = runs.len[r];
        if (r & 1)
        {
            if (threadIdx.x == 0)
            {
                int k = 0;
                for (; k + ZROWS <= len; k += ZROWS) tma_store_2d(&map_zero, zero, (int)e0, start + k);
                if (k < len)
                {
                    if (len >= ZROWS)
                        tma_store_2d(&map_zero, zero, (int)e0, start + len - ZROWS);
                    else
                        for (; k < len; ++k) tma_store_2d(&map_one, zero, (int)e0, start + k);
                }
                tma_commit();
            }
        }
        else
        {
            // data run: stage CHUNK rows at a time (runs are multiples of CHUNK in this probe)
            for (int k = 0; k < len; k += CHUNK)
            {
                if (threadIdx.x == 0) tma_wait_read0();
                __syncthreads();
                for (int i = 0; i < CHUNK; ++i) stage[i * TPB + threadIdx.x] = s + (start + k + i);
                fence_async();
                __syncthreads();
                if (threadIdx.x == 0)
                {
                    tma_store_2d(&map_data, stage, (int)e0, start + k);
                    tma_commit();
                }
            }
        }
    }
    if (threadIdx.x == 0) tma_wait_read0();
}

template <class F>
static float time_us(F f, int reps)
{
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    for (int i = 0; i < 3; ++i) f();
    CK(cudaGetLastError());
    cudaEventRecord(a);
    for (int i = 0; i < reps; ++i) f();
    cudaEventRecord(b);
    CK(cudaEventSynchronize(b));
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    return ms * 1e3f / reps;
}

template <int TPB, int HINT>
static void run_chain_tma(const double* in, double* out, int64_t n)
{
    CUtensorMap m = make_map(out, n, 48, TPB, 48);
    const size_t smem = 48 * TPB * 8;
    CK(cudaFuncSetAttribute(soa_tma<TPB, 7, 48, HINT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = (int)((n + TPB - 1) / TPB);
    float us = time_us([&] { soa_tma<TPB, 7, 48, HINT><<<grid, TPB, smem>>>(in, m, n); }, 50);
    printf("chain tma   cta %3d hint %d        : %6.1f us  %5.0f GB/s\n", TPB, HINT, us, 55.0 * n * 8 / us / 1e3);
}
template <int WARPS>
static void run_chain_tma_warp(const double* in, double* out, int64_t n)
{
    CUtensorMap m = make_map(out, n, 48, 32, 48);
    const size_t smem = (size_t)WARPS * 48 * 32 * 8;
    CK(cudaFuncSetAttribute(soa_tma_warp<WARPS, 7, 48>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = (int)((n + WARPS * 32 - 1) / (WARPS * 32));
    float us = time_us([&] { soa_tma_warp<WARPS, 7, 48><<<grid, WARPS * 32, smem>>>(in, m, n); }, 50);
    printf("chain tma   per-warp tiles, %d warps : %6.1f us  %5.0f GB/s\n", WARPS, us, 55.0 * n * 8 / us / 1e3);
}
template <int TPB, int CHUNK, int ZROWS>
static void run_tree_tma(const double* in, double* out, int64_t n, const Runs& runs, int rows_out)
{
    CUtensorMap md = make_map(out, n, rows_out, TPB, CHUNK), mz = make_map(out, n, rows_out, TPB, ZROWS), m1 = make_map(out, n, rows_out, TPB, 1);
    const size_t smem = (size_t)(CHUNK + ZROWS) * TPB * 8;
    CK(cudaFuncSetAttribute(tree_tma<TPB, CHUNK, ZROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = (int)((n + TPB - 1) / TPB);
    float us = time_us([&] { tree_tma<TPB, CHUNK, ZROWS><<<grid, TPB, smem>>>(in, md, mz, m1, runs, n, 38); }, 20);
    printf("tree tma    cta %3d chunk %2d zrows %2d : %6.1f us  %5.0f GB/s\n", TPB, CHUNK, ZROWS, us, (38.0 + rows_out) * n * 8 / us / 1e3);
}

int main()
{
    cudaDriverEntryPointQueryResult q;
    void* fn = nullptr;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    g_encode = (EncodeTiled)fn;
    {
        const int64_t n = 1 << 20;
        double *in, *out;
        CK(cudaMalloc(&in, 7 * n * 8));
        CK(cudaMalloc(&out, 48 * n * 8));
        CK(cudaMemset(in, 0, 7 * n * 8));
        for (int t : {64, 128})
        {
            const int g = (int)(n / t);
            float us = time_us([&] { soa_stg<7, 48><<<g, t>>>(in, out, n); }, 50);
            printf("chain st.cs cta %3d               : %6.1f us  %5.0f GB/s\n", t, us, 55.0 * n * 8 / us / 1e3);
        }
        {
            // read / write separation: inputs pulled into L2 before the store-heavy phase; 4 rotating inputs defeat residency
            double* ins[4];
            for (int i = 0; i < 4; ++i)
            {
                CK(cudaMalloc(&ins[i], 7 * n * 8));
                CK(cudaMemset(ins[i], 0, 7 * n * 8));
            }
            int it = 0;
            float us = time_us([&] { soa_stg<7, 48><<<(int)(n / 128), 128>>>(ins[it++ & 3], out, n); }, 48);
            printf("rotating inputs, st.cs cta 128                 : %6.1f us\n", us);
            it = 0;
            us = time_us([&] { prefetch_all<1><<<148 * 4, 256>>>((const char*)ins[it & 3], 7 * n * 8); soa_stg<7, 48><<<(int)(n / 128), 128>>>(ins[it++ & 3], out, n); }, 48);
            printf("rotating inputs, prefetch kernel (evict_last) + st.cs : %6.1f us\n", us);
            it = 0;
            us = time_us([&] { prefetch_all<0><<<148 * 4, 256>>>((const char*)ins[it & 3], 7 * n * 8); soa_stg<7, 48><<<(int)(n / 128), 128>>>(ins[it++ & 3], out, n); }, 48);
            printf("rotating inputs, prefetch kernel (normal) + st.cs     : %6.1f us\n", us);
            it = 0;
            us = time_us([&] { prefetch_all<1><<<148 * 4, 256>>>((const char*)ins[it++ & 3], 7 * n * 8); }, 48);
            printf("prefetch kernel alone                                  : %6.1f us\n", us);
            for (int pf : {148, 592, 2368})
            {
                it = 0;
                us = time_us([&] { soa_stg_prologue<7, 48><<<(int)(n / 128), 128>>>(ins[it++ & 3], out, n, pf); }, 48);
                printf("rotating inputs, prologue prefetch by %4d CTAs + st.cs : %6.1f us\n", pf, us);
            }
            for (int i = 0; i < 4; ++i) cudaFree(ins[i]);
        }
        run_chain_tma<64, 0>(in, out, n);
        run_chain_tma<64, 1>(in, out, n);
        run_chain_tma<128, 0>(in, out, n);
        run_chain_tma<128, 1>(in, out, n);
        run_chain_tma<256, 0>(in, out, n);
        run_chain_tma_warp<2>(in, out, n);
        run_chain_tma_warp<4>(in, out, n);
        // check one tile
        double h[4];
        CK(cudaMemcpy(h, out + 5 * n + 12345, 8, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(h + 1, out + 47 * n + (n - 1), 8, cudaMemcpyDeviceToHost));
        printf("check: out[5][12345] = %g (want 5), out[47][n-1] = %g (want 47)\n", h[0], h[1]);
        float us = time_us([&] { cudaMemsetAsync(out, 0, 48 * n * 8); }, 50);
        printf("cudaMemset 403 MB: %.1f us  %.0f GB/s\n", us, 48.0 * n * 8 / us / 1e3);
        cudaFree(in);
        cudaFree(out);
    }
    {
        const int64_t n = 1 << 18;
        const int rows_out = 936;
        double *in, *out;
        CK(cudaMalloc(&in, 38 * n * 8));
        CK(cudaMalloc(&out, (size_t)rows_out * n * 8));
        CK(cudaMemset(in, 0, 38 * n * 8));
        float us = time_us([&] { tree_stg<<<(int)(n / 64), 64>>>(in, out, n, 38, rows_out); }, 20);
        printf("tree st.cs  cta  64               : %6.1f us  %5.0f GB/s\n", us, (38.0 + rows_out) * n * 8 / us / 1e3);
        us = time_us([&] { tree_stg<<<(int)(n / 128), 128>>>(in, out, n, 38, rows_out); }, 20);
        printf("tree st.cs  cta 128               : %6.1f us  %5.0f GB/s\n", us, (38.0 + rows_out) * n * 8 / us / 1e3);
        // 24 phi rows (data), then per task row of J (24 rows x 38 columns): 16 data columns + 22 zero columns
        Runs runs;
        runs.n_runs = 0;
        int row = 0;
        runs.start[runs.n_runs] = row; runs.len[runs.n_runs++] = 24; row += 24;   // data (phi)
        runs.start[runs.n_runs] = row; runs.len[runs.n_runs++] = 0;               // empty zero run keeps the parity
        for (int r = 0; r < 24; ++r)
        {
            runs.start[runs.n_runs] = row; runs.len[runs.n_runs++] = 16; row += 16;
            runs.start[runs.n_runs] = row; runs.len[runs.n_runs++] = 22; row += 22;
        }
        printf("tree: %d runs, %d rows\n", runs.n_runs, row);
        run_tree_tma<64, 8, 22>(in, out, n, runs, rows_out);
        run_tree_tma<64, 8, 11>(in, out, n, runs, rows_out);
        run_tree_tma<128, 8, 22>(in, out, n, runs, rows_out);
        run_tree_tma<128, 8, 11>(in, out, n, runs, rows_out);
        run_tree_tma<32, 8, 22>(in, out, n, runs, rows_out);
        double h[2];
        CK(cudaMemcpy(h, out + (int64_t)30 * n + 777, 8, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(h + 1, out + (int64_t)45 * n + 777, 8, cudaMemcpyDeviceToHost));
        printf("check: out[30][777] = %g (want 30), out[45][777] = %g (want 0)\n", h[0], h[1]);
        us = time_us([&] { cudaMemsetAsync(out, 0, (size_t)rows_out * n * 8); }, 20);
        printf("cudaMemset %.0f MB: %.1f us  %.0f GB/s\n", rows_out * n * 8 / 1e6, us, (double)rows_out * n * 8 / us / 1e3);
    }
    return 0;
}
