// hbm_probe.cu -- what HBM bandwidth can a store-dominated streaming kernel reach on this B200?
// The hot path writes 384 B and reads 56 B per evaluation (87 % stores); MEASURED_PEAKS.json's denominator is a 1:1 copy.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/hbm_probe tools/hbm_probe.cu ; run under gpurun.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

__global__ void fill64(double* __restrict__ out, int64_t n, double v)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = v + (double)i;
}
// 48 output rows of n each, one thread writes element e of every row (the SoA pattern of the Jacobian) and reads 7 rows
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
        for (int k = 0; k < 48; ++k) out[k * n + e] = s + k;
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
    const int grids[] = {148 * 4, 148 * 8, 148 * 16, 8192};
    for (int g : grids)
    {
        float ms = time_ms([&] { fill64<<<g, 128>>>(out, 48 * n, 1.0); }, 50);
        printf("fill64   grid %5d x128: %.1f us  %.0f GB/s (write only, 403 MB)\n", g, ms * 1e3, 48.0 * n * 8 / ms / 1e6);
        ms = time_ms([&] { soa_like<<<g, 128>>>(in, out, n); }, 50);
        printf("soa_like grid %5d x128: %.1f us  %.0f GB/s (7 rows read, 48 rows written, 461 MB)\n", g, ms * 1e3, 55.0 * n * 8 / ms / 1e6);
        ms = time_ms([&] { copy64<<<g, 128>>>(out, out + 24 * n, 24 * n); }, 50);
        printf("copy64   grid %5d x128: %.1f us  %.0f GB/s (1:1 copy, 403 MB)\n", g, ms * 1e3, 48.0 * n * 8 / ms / 1e6);
    }
    float ms = time_ms([&] { cudaMemsetAsync(out, 0, 48 * n * 8); }, 50);
    printf("cudaMemset: %.1f us  %.0f GB/s\n", ms * 1e3, 48.0 * n * 8 / ms / 1e6);
    return 0;
}
