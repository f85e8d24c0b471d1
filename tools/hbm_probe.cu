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
    const int64_t pads[] = {0, 16, 32, 48, 64, 96, 128, 160, 256, 512, 1040, 4112, 65552};
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
