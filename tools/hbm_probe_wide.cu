// hbm_probe_wide.cu -- store pattern of the 38-DoF configuration (cfg4): 974 output rows per evaluation, E = 2^18.
// SoA rows [974][E] against tiles [E/32][974][32]; answers whether the row-major SoA layout itself limits HBM efficiency
// when the set of rows being written at one time no longer fits in L2.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

template <int ROWS>
__global__ void soa_rows(const double* __restrict__ in, double* __restrict__ out, int64_t n)
{
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; e < n; e += stride)
    {
        double s = 0;
        for (int k = 0; k < 38; ++k) s += in[k * n + e];
#pragma unroll 8
        for (int k = 0; k < ROWS; ++k) out[(int64_t)k * n + e] = s + k;
    }
}
template <int ROWS>
__global__ void tile_rows(const double* __restrict__ in, double* __restrict__ out, int64_t n)
{
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; e < n; e += stride)
    {
        double s = 0;
        for (int k = 0; k < 38; ++k) s += in[k * n + e];
        double* o = out + (e >> 5) * (int64_t)(ROWS * 32) + (e & 31);
#pragma unroll 8
        for (int k = 0; k < ROWS; ++k) o[k * 32] = s + k;
    }
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
    const int64_t n = 1 << 18;
    double *in, *out;
    cudaMalloc(&in, 38 * n * 8);
    cudaMalloc(&out, 974 * n * 8);
    cudaMemset(in, 0, 38 * n * 8);
    const double bytes = (974.0 + 38.0) * n * 8;
    for (int g : {148 * 3, 148 * 12, 2048})
    {
        float ms = time_ms([&] { soa_rows<974><<<g, 128>>>(in, out, n); }, 10);
        printf("soa  974 rows grid %5d x128: %7.1f us  %5.0f GB/s\n", g, ms * 1e3, bytes / ms / 1e6);
        ms = time_ms([&] { tile_rows<974><<<g, 128>>>(in, out, n); }, 10);
        printf("tile 974 rows grid %5d x128: %7.1f us  %5.0f GB/s\n", g, ms * 1e3, bytes / ms / 1e6);
    }
    return 0;
}
