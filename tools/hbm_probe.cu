// hbm_probe.cu -- what HBM bandwidth can a store-dominated streaming kernel reach on this B200?
// The hot path writes 384 B and reads 56 B per evaluation (87 % stores); MEASURED_PEAKS.json's denominator is a 1:1 copy.
// Variants: plain / streaming (st.global.cs) / write-through stores; SoA rows [48][E] vs tiles [E/32][48][32].
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/hbm_probe tools/hbm_probe.cu ; run under gpurun.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__global__ void fill64(double* __restrict__ out, int64_t n, double v)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = v + (double)i;
}
template <int MODE>
__device__ __forceinline__ void store(double* p, double v)
{
    if (MODE == 0) *p = v;
    else if (MODE == 1) __stcs(p, v);
    else if (MODE == 2) __stwt(p, v);
    else __stcg(p, v);
}
// 48 output rows of n each, one thread writes element e of every row (the SoA pattern of the Jacobian) and reads 7 rows
template <int MODE>
__global__ void soa_like(const double* __restrict__ in, double* __restrict__ out, int64_t n)
{
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; e < n; e += stride)
    {
        double s = 0;
#pragma unroll
        for (int k = 0; k < 7; ++k) s += in[k * n + e];
#pragma unroll
        for (int k = 0; k < 48; ++k) store<MODE>(out + k * n + e, s + k);
    }
}
// same as soa_like but the rows are `pitch` doubles apart (pitch >= n): does a power-of-two row stride hurt?
template <int MODE>
__global__ void soa_pitch(const double* __restrict__ in, double* __restrict__ out, int64_t n, int64_t pitch)
{
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; e < n; e += stride)
    {
        double s = 0;
#pragma unroll
        for (int k = 0; k < 7; ++k) s += in[k * n + e];
#pragma unroll
        for (int k = 0; k < 48; ++k) store<MODE>(out + k * pitch + e, s + k);
    }
}
// persistent, but every warp starts its 48 stores at a different row (de-synchronises the rows the GPU writes at any moment)
__global__ void soa_rotated(const double* __restrict__ in, double* __restrict__ out, int64_t n)
{
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int k0 = (int)(((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 7 % 48);
    for (; e < n; e += stride)
    {
        double s = 0;
#pragma unroll
        for (int k = 0; k < 7; ++k) s += in[k * n + e];
        for (int k = 0; k < 48; ++k)
        {
            const int kk = (k + k0) % 48;
            out[kk * n + e] = s + kk;
        }
    }
}
// persistent, contiguous block of evaluations per CTA
__global__ void soa_chunked(const double* __restrict__ in, double* __restrict__ out, int64_t n)
{
    const int64_t per = (n + gridDim.x - 1) / gridDim.x;
    const int64_t e1 = min(n, (int64_t)(blockIdx.x + 1) * per);
    for (int64_t e = (int64_t)blockIdx.x * per + threadIdx.x; e < e1; e += blockDim.x)
    {
        double s = 0;
#pragma unroll
        for (int k = 0; k < 7; ++k) s += in[k * n + e];
#pragma unroll
        for (int k = 0; k < 48; ++k) out[k * n + e] = s + k;
    }
}
// persistent with the next iteration's loads issued before this iteration's stores (register prefetch)
__global__ void soa_prefetch(const double* __restrict__ in, double* __restrict__ out, int64_t n)
{
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    double q[7];
    if (e < n)
        for (int k = 0; k < 7; ++k) q[k] = in[k * n + e];
    for (; e < n; e += stride)
    {
        double s = 0;
#pragma unroll
        for (int k = 0; k < 7; ++k) s += q[k];
        if (e + stride < n)
#pragma unroll
            for (int k = 0; k < 7; ++k) q[k] = in[k * n + e + stride];
#pragma unroll
        for (int k = 0; k < 48; ++k) out[k * n + e] = s + k;
    }
}
// persistent CTAs that fetch tiles of TILE evaluations from a global counter (dynamic load balancing), with the next
// tile's loads in flight while the current tile is stored
__device__ unsigned int g_counter;
template <int TILE>
__global__ void soa_dynamic(const double* __restrict__ in, double* __restrict__ out, int64_t n)
{
    __shared__ unsigned int s_tile[2];
    const int n_tiles = (int)((n + TILE - 1) / TILE);
    if (threadIdx.x == 0) s_tile[0] = atomicAdd(&g_counter, 1u);
    __syncthreads();
    unsigned int tile = s_tile[0];
    int buf = 0;
    while (tile < (unsigned)n_tiles)
    {
        if (threadIdx.x == 0) s_tile[buf ^ 1] = atomicAdd(&g_counter, 1u);
        for (int64_t e = (int64_t)tile * TILE + threadIdx.x; e < min(n, (int64_t)(tile + 1) * TILE); e += blockDim.x)
        {
            double s = 0;
#pragma unroll
            for (int k = 0; k < 7; ++k) s += in[k * n + e];
#pragma unroll
            for (int k = 0; k < 48; ++k) out[k * n + e] = s + k;
        }
        __syncthreads();
        buf ^= 1;
        tile = s_tile[buf];
        __syncthreads();
    }
}
// tiles of 32 evaluations: out[e/32][48][32]
template <int MODE>
__global__ void tile_like(const double* __restrict__ in, double* __restrict__ out, int64_t n)
{
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; e < n; e += stride)
    {
        double s = 0;
#pragma unroll
        for (int k = 0; k < 7; ++k) s += in[k * n + e];
        double* o = out + (e >> 5) * (48 * 32) + (e & 31);
#pragma unroll
        for (int k = 0; k < 48; ++k) store<MODE>(o + k * 32, s + k);
    }
}
__global__ void copy64(const double* __restrict__ in, double* __restrict__ out, int64_t n)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = in[i];
}

template <class F>
static float time_ms(F f, int reps)
{
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    for (int i = 0; i < 3; ++i) f();
    cudaEventRecord(a);
    for (int i = 0; i < reps; ++i) f();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    return ms / reps;
}

int main()
{
    const int64_t n = 1 << 20;
    double *in, *out;
    cudaMalloc(&in, 7 * n * 8);
    cudaMalloc(&out, 64 * n * 8);
    cudaMemset(in, 0, 7 * n * 8);
    const int grids[] = {148 * 3, 148 * 12, 8192};
    const int tpbs[] = {128, 128, 128, 128, 128};
    for (int i = 0; i < 3; ++i)
    {
        const int g = grids[i], t = tpbs[i];
        float ms = time_ms([&] { fill64<<<g, t>>>(out, 48 * n, 1.0); }, 50);
        printf("fill64      grid %5d x%d: %6.1f us  %5.0f GB/s (write only, 403 MB)\n", g, t, ms * 1e3, 48.0 * n * 8 / ms / 1e6);
        ms = time_ms([&] { soa_like<0><<<g, t>>>(in, out, n); }, 50);
        printf("soa plain   grid %5d x%d: %6.1f us  %5.0f GB/s\n", g, t, ms * 1e3, 55.0 * n * 8 / ms / 1e6);
        ms = time_ms([&] { soa_like<1><<<g, t>>>(in, out, n); }, 50);
        printf("soa st.cs   grid %5d x%d: %6.1f us  %5.0f GB/s\n", g, t, ms * 1e3, 55.0 * n * 8 / ms / 1e6);
        ms = time_ms([&] { soa_like<2><<<g, t>>>(in, out, n); }, 50);
        printf("soa st.wt   grid %5d x%d: %6.1f us  %5.0f GB/s\n", g, t, ms * 1e3, 55.0 * n * 8 / ms / 1e6);
        ms = time_ms([&] { soa_like<3><<<g, t>>>(in, out, n); }, 50);
        printf("soa st.cg   grid %5d x%d: %6.1f us  %5.0f GB/s\n", g, t, ms * 1e3, 55.0 * n * 8 / ms / 1e6);
        ms = time_ms([&] { tile_like<0><<<g, t>>>(in, out, n); }, 50);
        printf("tile plain  grid %5d x%d: %6.1f us  %5.0f GB/s\n", g, t, ms * 1e3, 55.0 * n * 8 / ms / 1e6);
        ms = time_ms([&] { tile_like<1><<<g, t>>>(in, out, n); }, 50);
        printf("tile st.cs  grid %5d x%d: %6.1f us  %5.0f GB/s\n", g, t, ms * 1e3, 55.0 * n * 8 / ms / 1e6);
        ms = time_ms([&] { copy64<<<g, t>>>(out, out + 24 * n, 24 * n); }, 50);
        printf("copy64      grid %5d x%d: %6.1f us  %5.0f GB/s (1:1 copy, 403 MB)\n", g, t, ms * 1e3, 48.0 * n * 8 / ms / 1e6);
    }
    for (int g : {148 * 3, 148 * 7, 148 * 16, 148 * 32, 148 * 55, 8192})
        for (int t : {64, 128})
        {
            const int gg = g * 128 / t;
            float ms = time_ms([&] { soa_like<0><<<gg, t>>>(in, out, n); }, 50);
            printf("soa plain    grid %5d x%3d: %6.1f us  %5.0f GB/s\n", gg, t, ms * 1e3, 55.0 * n * 8 / ms / 1e6);
            ms = time_ms([&] { soa_rotated<<<gg, t>>>(in, out, n); }, 50);
            printf("soa rotated  grid %5d x%3d: %6.1f us  %5.0f GB/s\n", gg, t, ms * 1e3, 55.0 * n * 8 / ms / 1e6);
            ms = time_ms([&] { soa_chunked<<<gg, t>>>(in, out, n); }, 50);
            printf("soa chunked  grid %5d x%3d: %6.1f us  %5.0f GB/s\n", gg, t, ms * 1e3, 55.0 * n * 8 / ms / 1e6);
            ms = time_ms([&] { soa_prefetch<<<gg, t>>>(in, out, n); }, 50);
            printf("soa prefetch grid %5d x%3d: %6.1f us  %5.0f GB/s\n", gg, t, ms * 1e3, 55.0 * n * 8 / ms / 1e6);
        }
    for (int g : {148 * 3, 148 * 7, 148 * 16})
        for (int t : {64, 128})
        {
            const int gg = g * 128 / t;
            unsigned int zero = 0;
            float ms = time_ms([&] { cudaMemcpyToSymbolAsync(g_counter, &zero, 4, 0, cudaMemcpyHostToDevice, 0); soa_dynamic<128><<<gg, t>>>(in, out, n); }, 50);
            printf("soa dynamic tile 128 grid %5d x%3d: %6.1f us  %5.0f GB/s\n", gg, t, ms * 1e3, 55.0 * n * 8 / ms / 1e6);
            ms = time_ms([&] { cudaMemcpyToSymbolAsync(g_counter, &zero, 4, 0, cudaMemcpyHostToDevice, 0); soa_dynamic<256><<<gg, t>>>(in, out, n); }, 50);
            printf("soa dynamic tile 256 grid %5d x%3d: %6.1f us  %5.0f GB/s\n", gg, t, ms * 1e3, 55.0 * n * 8 / ms / 1e6);
            ms = time_ms([&] { cudaMemcpyToSymbolAsync(g_counter, &zero, 4, 0, cudaMemcpyHostToDevice, 0); soa_dynamic<1024><<<gg, t>>>(in, out, n); }, 50);
            printf("soa dynamic tile 1024 grid %5d x%3d: %6.1f us  %5.0f GB/s\n", gg, t, ms * 1e3, 55.0 * n * 8 / ms / 1e6);
        }
    const int64_t pads[] = {0};
    for (int g : {444, 1776, 8192})
        for (int64_t pad : pads)
        {
            float ms = time_ms([&] { soa_pitch<0><<<g, 128>>>(in, out, n, n + pad); }, 50);
            printf("soa pitch n+%-6lld grid %5d: %6.1f us  %5.0f GB/s\n", (long long)pad, g, ms * 1e3, 55.0 * n * 8 / ms / 1e6);
        }
    float ms = time_ms([&] { cudaMemsetAsync(out, 0, 48 * n * 8); }, 50);
    printf("cudaMemset: %.1f us  %.0f GB/s\n", ms * 1e3, 48.0 * n * 8 / ms / 1e6);
    return 0;
}
