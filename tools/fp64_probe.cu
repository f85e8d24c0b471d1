// Measures the FP64 issue rate one SM sustains (DFMA per clock per SM) -- the denominator of the FP64 roofline DESIGN.md quotes
// for the SphereCollision kernel.  Variants: pure DFMA chains at several occupancies / ILP, and DFMA interleaved 1:1 with
// integer ALU work (does a non-FP64 instruction cost an FP64 issue slot?).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_probe tools/fp64_probe.cu && tools/fp64_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

// KIND 0: IMAD (half-rate FMA-heavy pipe), 1: LOP3 (ALU pipe), 2: FFMA (full-rate FP32), 3: FSEL on a predicate (ALU)
template <int ILP, int MIX, int KIND = 0>
__global__ void probe(double* out, long long* cycles, int iters, double b, double c)
{
    double a[ILP];
    unsigned int x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i)
    {
        a[i] = threadIdx.x * 1e-3 + i;
        x[i] = threadIdx.x + i;
    }
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it)
    {
#pragma unroll
        for (int i = 0; i < ILP; ++i)
        {
            a[i] = fma(a[i], b, c);
            if (MIX)
            {
#pragma unroll
                for (int m = 0; m < MIX; ++m)
                {
                    if (KIND == 0) x[i] = x[i] * 3u + (unsigned int)it;  // IMAD
                    else if (KIND == 1) x[i] = (x[i] ^ (unsigned int)it) & 0x7fffffffu | (unsigned int)m;  // LOP3
                    else if (KIND == 2) x[i] = __float_as_uint(fmaf(__uint_as_float(x[i]), 1.0001f, 0.5f));  // FFMA
                    else x[i] = (it & (1 << m)) ? x[i] : __float_as_uint(__uint_as_float(x[i]) + 1.0f);  // FADD + select
                }
            }
        }
    }
    const long long t1 = clock64();
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += a[i] + (double)x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

__constant__ double kC[2] = {1.0000001, 1e-9};

// OP 0 DFMA, 1 DADD, 2 DMUL, 3 DFMA with constant-bank operands, 4 a < b ? a : b (DSETP + selects), 5 rsqrt.approx.ftz.f64 (MUFU.RSQ64H)
template <int OP>
__global__ void probe_op(double* out, long long* cycles, int iters, double b, double c)
{
    constexpr int ILP = 8;
    double a[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = 1.0 + threadIdx.x * 1e-3 + i;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it)
    {
#pragma unroll
        for (int i = 0; i < ILP; ++i)
        {
            if (OP == 0) a[i] = fma(a[i], b, c);
            else if (OP == 1) a[i] = a[i] + c;
            else if (OP == 2) a[i] = a[i] * b;
            else if (OP == 3) a[i] = fma(a[i], kC[0], kC[1]);
            else if (OP == 4) a[i] = a[i] < b ? a[i] : b + a[(i + 1) % ILP];
            else asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(a[i]) : "d"(a[i]));
        }
    }
    const long long t1 = clock64();
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// three distinct, changing register operands per DFMA (register-file bandwidth): a <- a*b+c, b <- b*c+a, c <- c*a+b
template <int ILP>
__global__ void probe_rf(double* out, long long* cycles, int iters)
{
    double a[ILP], b[ILP], c[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i)
    {
        a[i] = 1.0 + threadIdx.x * 1e-9 + i * 1e-6;
        b[i] = 1.0 - threadIdx.x * 1e-9;
        c[i] = 1e-9 * i;
    }
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it)
    {
#pragma unroll
        for (int i = 0; i < ILP; ++i) a[i] = fma(a[i], b[i], c[i]);
#pragma unroll
        for (int i = 0; i < ILP; ++i) b[i] = fma(b[i], c[i], a[i]);
#pragma unroll
        for (int i = 0; i < ILP; ++i) c[i] = fma(c[i], a[i], b[i]);
    }
    const long long t1 = clock64();
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += a[i] + b[i] + c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int ILP>
void run_rf(int ctas, int threads, int sms, double* d_out, long long* d_cyc)
{
    const int iters = 5000, blocks = sms * ctas;
    probe_rf<ILP><<<blocks, threads>>>(d_out, d_cyc, 100);
    cudaDeviceSynchronize();
    probe_rf<ILP><<<blocks, threads>>>(d_out, d_cyc, iters);
    cudaDeviceSynchronize();
    static long long h[4096];
    cudaMemcpy(h, d_cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < blocks; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("DFMA, 3 varying register operands, ILP %d, %2d warps/SM: %6.2f DFMA/clk/SM\n", ILP, ctas * threads / 32,
           (double)blocks * threads * ILP * 3 * iters / sms / (double)mx);
}

template <int OP>
void run_op(const char* name, int sms, double* d_out, long long* d_cyc)
{
    const int iters = 10000, blocks = sms * 4, threads = 256;
    probe_op<OP><<<blocks, threads>>>(d_out, d_cyc, 100, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    probe_op<OP><<<blocks, threads>>>(d_out, d_cyc, iters, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    static long long h[4096];
    cudaMemcpy(h, d_cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < blocks; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("%-28s 32 warps/SM ILP 8: %6.2f thread-ops/clk/SM\n", name, (double)blocks * threads * 8 * iters / sms / (double)mx);
}

template <int ILP, int MIX, int KIND = 0>
void run(int ctas_per_sm, int threads, int sms, double* d_out, long long* d_cyc)
{
    const int iters = 20000;
    const int blocks = sms * ctas_per_sm;
    probe<ILP, MIX, KIND><<<blocks, threads>>>(d_out, d_cyc, 100, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe<ILP, MIX, KIND><<<blocks, threads>>>(d_out, d_cyc, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    static long long h[4096];
    cudaMemcpy(h, d_cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < blocks; ++i) mx = h[i] > mx ? h[i] : mx;
    const double dfma = (double)blocks * threads * ILP * iters;
    printf("ILP %2d MIX %d KIND %d  warps/SM %2d : %7.3f ms  %6.2f DFMA/clk/SM (kernel clocks)  %6.2f TFLOP/s (events)\n", ILP, MIX, KIND,
           ctas_per_sm * threads / 32, ms, dfma / sms / (double)mx, 2.0 * dfma / (ms * 1e-3) / 1e12);
}

int main()
{
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double* d_out;
    long long* d_cyc;
    cudaMalloc(&d_out, sizeof(double) * 4096 * 1024);
    cudaMalloc(&d_cyc, sizeof(long long) * 4096);
    printf("SMs %d\n", sms);
    run<1, 0>(1, 128, sms, d_out, d_cyc);
    run<1, 0>(2, 256, sms, d_out, d_cyc);
    run<1, 0>(4, 256, sms, d_out, d_cyc);
    run<1, 0>(8, 256, sms, d_out, d_cyc);
    run<4, 0>(1, 128, sms, d_out, d_cyc);
    run<4, 0>(3, 128, sms, d_out, d_cyc);
    run<4, 0>(6, 128, sms, d_out, d_cyc);
    run<8, 0>(2, 256, sms, d_out, d_cyc);
    run<8, 0>(4, 256, sms, d_out, d_cyc);
    run<4, 1>(3, 128, sms, d_out, d_cyc);
    run<4, 1>(6, 128, sms, d_out, d_cyc);
    run<4, 2>(6, 128, sms, d_out, d_cyc);
    run<8, 1>(4, 256, sms, d_out, d_cyc);
    run<8, 2>(4, 256, sms, d_out, d_cyc);
    run<8, 3>(4, 256, sms, d_out, d_cyc);
    run<8, 1, 1>(4, 256, sms, d_out, d_cyc);
    run<8, 2, 1>(4, 256, sms, d_out, d_cyc);
    run<8, 3, 1>(4, 256, sms, d_out, d_cyc);
    run<8, 1, 2>(4, 256, sms, d_out, d_cyc);
    run<8, 2, 2>(4, 256, sms, d_out, d_cyc);
    run<8, 3, 2>(4, 256, sms, d_out, d_cyc);
    run<8, 1, 3>(4, 256, sms, d_out, d_cyc);
    run<8, 2, 3>(4, 256, sms, d_out, d_cyc);
    run<4, 1, 1>(5, 128, sms, d_out, d_cyc);
    run<4, 1, 2>(5, 128, sms, d_out, d_cyc);
    run_op<0>("DFMA", sms, d_out, d_cyc);
    run_op<1>("DADD", sms, d_out, d_cyc);
    run_op<2>("DMUL", sms, d_out, d_cyc);
    run_op<3>("DFMA const-bank operands", sms, d_out, d_cyc);
    run_op<4>("DSETP + select (+DADD)", sms, d_out, d_cyc);
    run_op<5>("MUFU.RSQ64H", sms, d_out, d_cyc);
    run_rf<4>(4, 128, sms, d_out, d_cyc);
    run_rf<4>(5, 128, sms, d_out, d_cyc);
    run_rf<8>(4, 128, sms, d_out, d_cyc);
    run_rf<8>(4, 256, sms, d_out, d_cyc);
    return 0;
}
