// cost_kernel.cu -- batched scalar task cost and per-trajectory reductions (SURVEY.md section 8, rows e and f1).
//
// Reference path replaced, for B trajectories at once:
//   TimeIndexedTask::Update           ydiff[t] = Phi[t] - y[t]                       (exotica_core/src/tasks.cpp:306-314)
//   TaskSpaceVector::operator-        SO(3) log difference on rotation sub-blocks  (exotica_core/src/task_space_vector.cpp:64-91)
//   GetScalarTaskCost / GetScalarTransitionCost / GetCost                           (abstract_time_indexed_problem.cpp:299-342)
//   UnconstrainedEndPoseProblem::GetScalarCost                                      (unconstrained_end_pose_problem.cpp:79-82)
// The result, one double per trajectory, is what the multi-GPU path gathers with NCCL; exo_reduce_kernel does the same
// for the per-evaluation minimum sphere distance (min over the timesteps of a trajectory).
//
// One warp per trajectory, lanes stride over its timesteps (coalesced reads of Phi in the SoA layout), warp-shuffle sum.
// Reads 8 * length_Phi + 8 * N bytes per evaluation: HBM-bound, ~1/10 of the traffic of the kernel that produced Phi and J.
#include <math.h>
#include <string.h>

#include "cost_device.cuh"

namespace exo
{
namespace
{
using namespace dev;

__global__ void __launch_bounds__(128) exo_cost_kernel(const __grid_constant__ CostParams P)
{
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t b = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; b < P.n_traj; b += warps)
    {
        double acc = 0.0;
        for (int t = lane; t < P.T; t += 32)
        {
            const int64_t e = b * P.T + t;
            const double c = t >= P.t_first ? eval_cost(P, e, t) : 0.0;
            if (P.cost_eval) P.cost_eval[e] = c;
            acc += c;
        }
        acc = warp_sum_d(acc);
        if (lane == 0 && P.cost_traj) P.cost_traj[b] = acc;
    }
}

// Longer trajectories (T > 32): one CTA per trajectory, one lane per timestep -- four times the loads in flight of the warp-per-
// trajectory form, which was latency-bound (59 us for 4096 x 100 timesteps, long-scoreboard stalls 7 per issue); fixed summation
// order (lanes, then the four warps), so the result does not depend on scheduling.  What bounds it (ncu, 4096 x 100 timesteps): the
// ~1000 warp-instructions of one evaluation are a serial stream (dependent FP64 chains of the rotation log, ~10 exposed memory waits)
// at 14 cycles per issue, so the time follows the resident warps: 4 / 5 / 7 / 8 CTAs per SM (118 / 96 / 72 / 64 registers) take
// 39.3 / 35.7 / 30.2 / 29.8 us -- hence the 64-register build.  Measured and dropped: staging the trajectory's Phi rows, states and
// the goal / rho tables in shared memory first (every load independent, none left in the evaluation; 44 us -- the load and
// evaluation phases of a CTA no longer overlap) and L1 / L2 prefetches of the evaluation's entries (38 us).
__global__ void __launch_bounds__(128, 8) exo_cost_block_kernel(const __grid_constant__ CostParams P)
{
    __shared__ double part[4];
    for (int64_t b = blockIdx.x; b < P.n_traj; b += gridDim.x)
    {
        double acc = 0.0;
        for (int t = threadIdx.x; t < P.T; t += 128)
        {
            const int64_t e = b * P.T + t;
            const double c = t >= P.t_first ? eval_cost(P, e, t) : 0.0;
            if (P.cost_eval) P.cost_eval[e] = c;
            acc += c;
        }
        acc = warp_sum_d(acc);
        if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
        __syncthreads();
        if (threadIdx.x == 0 && P.cost_traj) P.cost_traj[b] = (part[0] + part[1]) + (part[2] + part[3]);
        __syncthreads();
    }
}

// End-pose problems (T = 1): one lane per configuration instead of one warp per single-timestep "trajectory"
__global__ void __launch_bounds__(128) exo_cost_endpose_kernel(const __grid_constant__ CostParams P)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < P.n_traj; e += stride)
    {
        const double c = P.t_first <= 0 ? eval_cost(P, e, 0) : 0.0;
        if (P.cost_eval) P.cost_eval[e] = c;
        if (P.cost_traj) P.cost_traj[e] = c;
    }
}

// AbstractTimeIndexedProblem::GetCostJacobian (abstract_time_indexed_problem.cpp:309-318) for every (trajectory, timestep):
//   grad[t] = 2 ct J[t]^T S[t] ydiff[t] + 2 ct W xdiff[t] - 2 ct W xdiff[t + 1]      (t >= t_first; zero before)
// and UnconstrainedEndPoseProblem::GetScalarJacobian (unconstrained_end_pose_problem.cpp:84-87) when T = 1.  One lane per
// evaluation: ydiff (SO(3) log included) goes to shared memory [row][lane], then one pass over the Jacobian (coalesced in SoA).
// Occupancy: the ydiff phase is a latency-bound serial stream like exo_cost_block_kernel's, so resident warps count -- 4 / 6 / 7 / 8
// CTAs per SM (108 / 80 / 72 / 64 registers) take 115 / 95 / 95 / 85 us on 4096 x 100 timesteps (LWR, EffFrame + JointLimit).
template <int NT>  // NT = 8: at most 8 controlled joints (register accumulators); NT = 0: any N
__global__ void __launch_bounds__(128, NT > 0 ? 8 : 4) exo_cost_gradient_kernel(const __grid_constant__ CostGradParams P)
{
    extern __shared__ double yd_s[];  // [length_jac][128]
    double* yd = yd_s + threadIdx.x;
    const CostParams& C = P.c;
    const int N = C.n_ctrl, T = C.T;
    const int64_t n_evals = C.n_traj * T, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_evals; e += stride)
    {
        const int t = (int)(e % T);
        double* g = P.grad + e * P.g_se;
        if (t < C.t_first)
        {
            for (int n = 0; n < N; ++n) g[n * P.g_sk] = 0.0;
            continue;
        }
        eval_ydiff(C, e, t, yd, 128);
        const double* rho = C.rho + (int64_t)t * C.n_maps;
        for (int k = 0; k < C.n_maps; ++k)
        {
            const CostMapP& m = C.map[k];
            for (int i = 0; i < m.length_jacobian; ++i) yd[(m.start_jacobian + i) * 128] *= rho[k];  // S ydiff
        }
        const double* J = P.jac + e * P.j_se;
        const double* x = C.x + e * C.x_se;
        const double two_ct = 2.0 * C.ct;
        double acc[NT > 0 ? NT : 1];
        if (NT > 0)
        {
            // N <= NT: rows outermost, N independent accumulators in registers (one shared-memory read per row, N loads in flight)
#pragma unroll
            for (int n = 0; n < NT; ++n) acc[n] = 0.0;
            // two rows per iteration: 2 N loads in flight per lane (the kernel is a stream of J: 365 MB per 409 600 evaluations;
            // measured 120 us with one row per iteration, 109 us with two, 117 us with four)
            int r = 0;
            for (; r + 1 < P.length_jac; r += 2)
            {
                double j0[NT > 0 ? NT : 1], j1[NT > 0 ? NT : 1];
#pragma unroll
                for (int n = 0; n < NT; ++n)
                {
                    j0[n] = n < N ? J[r * P.j_sr + n * P.j_sc] : 0.0;
                    j1[n] = n < N ? J[(r + 1) * P.j_sr + n * P.j_sc] : 0.0;
                }
                const double y0 = yd[r * 128], y1 = yd[(r + 1) * 128];
#pragma unroll
                for (int n = 0; n < NT; ++n) acc[n] += j0[n] * y0 + j1[n] * y1;
            }
            for (; r < P.length_jac; ++r)
            {
                const double y = yd[r * 128];
#pragma unroll
                for (int n = 0; n < NT; ++n)
                    if (n < N) acc[n] += J[r * P.j_sr + n * P.j_sc] * y;
            }
        }
#pragma unroll
        for (int n = 0; n < (NT > 0 ? NT : N); ++n)
        {
            if (n >= N) continue;
            double s = 0.0;
            if (NT > 0)
                s = acc[n];
            else
                for (int r = 0; r < P.length_jac; ++r) s += J[r * P.j_sr + n * P.j_sc] * yd[r * 128];
            s *= two_ct;
            if (C.W && t > 0)
            {
                const double xn = x[n * C.x_sk];
                s += two_ct * C.W[n] * (xn - x[n * C.x_sk - C.x_se]);
                if (t + 1 < T) s -= two_ct * C.W[n] * (x[n * C.x_sk + C.x_se] - xn);
            }
            g[n * P.g_sk] = s;
        }
    }
}

// AICOSolver::GetTaskCosts (aico_solver.cpp:364-386) for every (trajectory, timestep), on the Phi / J of prob_->Update(qhat[t], t):
// tasks with precision rho > 0 add  C += rho |ydiff|^2,  R += rho J^T J,  v = -ydiff + J qhat,  r += rho J^T v,  rhat += rho |v|^2.
// One lane per evaluation; per row the precision and v sit in shared memory [row][lane]; J is re-read from L1 / L2 for the
// N (N + 1) / 2 distinct entries of the symmetric R.
// (NT = 8 holds N (N + 3) / 2 accumulators: 4 CTAs per SM at 128 registers 160 us, 3 at 162 registers 171 us, 5 / 6 with spills 235 / 267 us)
template <int NT>  // NT = 8: at most 8 controlled joints, R and r accumulate in registers; NT = 0: any N
__global__ void __launch_bounds__(128, NT > 0 ? 4 : 1) exo_aico_messages_kernel(const __grid_constant__ AicoParams P)
{
    extern __shared__ double sh[];  // [2][length_jac][128]
    const CostParams& C = P.c;
    const int N = C.n_ctrl, T = C.T, LJ = P.length_jac;
    double* v = sh + threadIdx.x;
    double* a = sh + (size_t)LJ * 128 + threadIdx.x;
    const int64_t n_evals = C.n_traj * T, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_evals; e += stride)
    {
        const int t = (int)(e % T);
        eval_ydiff(C, e, t, v, 128);
        const double* rho = C.rho + (int64_t)t * C.n_maps;
        for (int k = 0; k < C.n_maps; ++k)
        {
            const CostMapP& m = C.map[k];
            const double prec = rho[k] > 0 ? rho[k] : 0.0;  // "if (prec > 0)"
            for (int i = 0; i < m.length_jacobian; ++i) a[(m.start_jacobian + i) * 128] = prec;
        }
        const double* J = P.jac + e * P.j_se;
        const double* q = C.x + e * C.x_se;
        double cost = 0.0, rhat = 0.0;
        double* Re = P.R + e * P.R_se;
        double* re = P.r + e * P.r_se;
        if (NT > 0)
        {
            // N <= NT: one pass over J -- each row is loaded once into registers and added to R and r as a rank-one update
            constexpr int NR = NT > 0 ? NT : 1;  // (the NT = 0 instantiation never runs this block)
            double qn[NR], Racc[NR * (NR + 1) / 2], racc[NR];
#pragma unroll
            for (int n = 0; n < NT; ++n)
            {
                qn[n] = n < N ? q[n * C.x_sk] : 0.0;
                racc[n] = 0.0;
            }
#pragma unroll
            for (int k = 0; k < NT * (NT + 1) / 2; ++k) Racc[k] = 0.0;
#ifndef EXO_AICO_UNROLL
#define EXO_AICO_UNROLL 1  // 2 or 3 rows cost 20 registers (178: two CTAs per SM instead of three) for no more loads in flight per SM
#endif
            constexpr int kRowsInFlight = EXO_AICO_UNROLL;  // rows of J whose loads are issued before the first is consumed
#pragma unroll kRowsInFlight
            for (int r = 0; r < LJ; ++r)
            {
                const double yd = v[r * 128], prec = a[r * 128];
                double Jr[NR];
#pragma unroll
                for (int n = 0; n < NT; ++n) Jr[n] = n < N ? J[r * P.j_sr + n * P.j_sc] : 0.0;
                double s = -yd;
#pragma unroll
                for (int n = 0; n < NT; ++n) s += Jr[n] * qn[n];
                cost += prec * yd * yd;
                rhat += prec * s * s;
#pragma unroll
                for (int i = 0; i < NT; ++i)
                {
                    const double w = prec * Jr[i];
                    racc[i] += w * s;
#pragma unroll
                    for (int j = 0; j <= i; ++j) Racc[i * (i + 1) / 2 + j] += w * Jr[j];
                }
            }
#pragma unroll
            for (int i = 0; i < NT; ++i)
            {
                if (i >= N) continue;
                __stcs(&re[i * P.r_sk], racc[i]);
#pragma unroll
                for (int j = 0; j <= i; ++j)
                {
                    __stcs(&Re[i * P.R_si + j * P.R_sj], Racc[i * (i + 1) / 2 + j]);
                    __stcs(&Re[j * P.R_si + i * P.R_sj], Racc[i * (i + 1) / 2 + j]);
                }
            }
            P.cost[e] = C.ct * cost;
            P.rhat[e] = rhat;
            continue;
        }
        for (int r = 0; r < LJ; ++r)
        {
            const double yd = v[r * 128], prec = a[r * 128];
            double s = -yd;
            for (int n = 0; n < N; ++n) s += J[r * P.j_sr + n * P.j_sc] * q[n * C.x_sk];
            cost += prec * yd * yd;
            rhat += prec * s * s;
            v[r * 128] = s;
        }
        P.cost[e] = C.ct * cost;
        P.rhat[e] = rhat;
        for (int i = 0; i < N; ++i)
        {
            double ri = 0.0;
            for (int r = 0; r < LJ; ++r) ri += a[r * 128] * J[r * P.j_sr + i * P.j_sc] * v[r * 128];
            __stcs(&re[i * P.r_sk], ri);
            for (int j = 0; j <= i; ++j)
            {
                double s = 0.0;
                for (int r = 0; r < LJ; ++r) s += a[r * 128] * J[r * P.j_sr + i * P.j_sc] * J[r * P.j_sr + j * P.j_sc];
                __stcs(&Re[i * P.R_si + j * P.R_sj], s);
                __stcs(&Re[j * P.R_si + i * P.R_sj], s);
            }
        }
    }
}

// out[b] = min / sum over t in [t_first, T) of in[b * T + t]
__global__ void __launch_bounds__(128) exo_reduce_kernel(const double* __restrict__ in, int64_t n_traj, int T, int t_first, int op, double* __restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t b = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; b < n_traj; b += warps)
    {
        double acc = op == 0 ? CUDART_INF : 0.0;
        for (int t = t_first + lane; t < T; t += 32)
        {
            const double v = in[b * T + t];
            acc = op == 0 ? fmin(acc, v) : acc + v;
        }
        acc = op == 0 ? warp_min_d(acc) : warp_sum_d(acc);
        if (lane == 0) out[b] = acc;
    }
}
__global__ void __launch_bounds__(128) exo_fill_kernel(double* __restrict__ out, int64_t n, double value)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = value;
}
}  // namespace

cudaError_t launch_fill(double* out, int64_t n, double value, cudaStream_t stream)
{
    if (n <= 0) return cudaSuccess;
    const int64_t blocks = (n + 127) / 128;
    exo_fill_kernel<<<(int)(blocks < 148 * 8 ? blocks : 148 * 8), 128, 0, stream>>>(out, n, value);
    return cudaGetLastError();
}

cudaError_t launch_cost(const CostParams& P, cudaStream_t stream)
{
    if (P.n_traj <= 0) return cudaSuccess;
    if (P.T == 1)
    {
        const int64_t b1 = (P.n_traj + 127) / 128;
        exo_cost_endpose_kernel<<<(int)(b1 < 148 * 16 ? b1 : 148 * 16), 128, 0, stream>>>(P);
        return cudaGetLastError();
    }
    if (P.T > 32)
    {
        exo_cost_block_kernel<<<(int)(P.n_traj < 148 * 32 ? P.n_traj : 148 * 32), 128, 0, stream>>>(P);
        return cudaGetLastError();
    }
    const int64_t blocks = (P.n_traj + 3) / 4;  // 4 warps (trajectories) per CTA
    exo_cost_kernel<<<(int)(blocks < 148 * 16 ? blocks : 148 * 16), 128, 0, stream>>>(P);
    return cudaGetLastError();
}

cudaError_t launch_cost_gradient(const CostGradParams& P, cudaStream_t stream)
{
    const int64_t n_evals = P.c.n_traj * P.c.T;
    if (n_evals <= 0) return cudaSuccess;
    const size_t smem = (size_t)(P.length_jac > 0 ? P.length_jac : 1) * 128 * sizeof(double);
    if (smem > 200 * 1024) return cudaErrorInvalidConfiguration;
    if (smem > 48 * 1024)
    {
        cudaError_t e = cudaFuncSetAttribute(exo_cost_gradient_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(exo_cost_gradient_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    const int64_t blocks = (n_evals + 127) / 128;
    const int grid = (int)(blocks < 148 * 16 ? blocks : 148 * 16);
    if (P.c.n_ctrl <= 8)
        exo_cost_gradient_kernel<8><<<grid, 128, smem, stream>>>(P);
    else
        exo_cost_gradient_kernel<0><<<grid, 128, smem, stream>>>(P);
    return cudaGetLastError();
}

cudaError_t launch_aico_messages(const AicoParams& P, cudaStream_t stream)
{
    const int64_t n_evals = P.c.n_traj * P.c.T;
    if (n_evals <= 0) return cudaSuccess;
    const size_t smem = (size_t)2 * (P.length_jac > 0 ? P.length_jac : 1) * 128 * sizeof(double);
    if (smem > 200 * 1024) return cudaErrorInvalidConfiguration;
    if (smem > 48 * 1024)
    {
        cudaError_t e = cudaFuncSetAttribute(exo_aico_messages_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(exo_aico_messages_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    const int64_t blocks = (n_evals + 127) / 128;
    const int grid = (int)(blocks < 148 * 16 ? blocks : 148 * 16);
    if (P.c.n_ctrl <= 8)
        exo_aico_messages_kernel<8><<<grid, 128, smem, stream>>>(P);
    else
        exo_aico_messages_kernel<0><<<grid, 128, smem, stream>>>(P);
    return cudaGetLastError();
}

cudaError_t launch_reduce(const double* in, int64_t n_traj, int T, int t_first, int op, double* out, cudaStream_t stream)
{
    if (n_traj <= 0) return cudaSuccess;
    const int64_t blocks = (n_traj + 3) / 4;
    exo_reduce_kernel<<<(int)(blocks < 148 * 16 ? blocks : 148 * 16), 128, 0, stream>>>(in, n_traj, T, t_first, op, out);
    return cudaGetLastError();
}

// ---- host: exotica::GetRotation for the goal's rotation sub-blocks (same formulas as dev_get_rotation) ------------------
void host_get_rotation(const double* d, int type, double* out)
{
    static const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    memcpy(out, I, sizeof(I));
    switch (type)
    {
        case EXO_ROT_QUATERNION:
        {
            const double x = d[0], y = d[1], z = d[2], w = d[3];
            const double x2 = x * x, y2 = y * y, z2 = z * z, w2 = w * w;
            out[0] = w2 + x2 - y2 - z2; out[1] = 2 * x * y - 2 * w * z; out[2] = 2 * x * z + 2 * w * y;
            out[3] = 2 * x * y + 2 * w * z; out[4] = w2 - x2 + y2 - z2; out[5] = 2 * y * z - 2 * w * x;
            out[6] = 2 * x * z - 2 * w * y; out[7] = 2 * y * z + 2 * w * x; out[8] = w2 - x2 - y2 + z2;
            break;
        }
        case EXO_ROT_RPY:
        case EXO_ROT_ZYX:
        {
            const bool zyx = type == EXO_ROT_ZYX;
            const double roll = zyx ? d[2] : d[0], pitch = d[1], yaw = zyx ? d[0] : d[2];
            const double sa = sin(yaw), ca = cos(yaw), sb = sin(pitch), cb = cos(pitch), sc = sin(roll), cc = cos(roll);
            out[0] = ca * cb; out[1] = ca * sb * sc - sa * cc; out[2] = ca * sb * cc + sa * sc;
            out[3] = sa * cb; out[4] = sa * sb * sc + ca * cc; out[5] = sa * sb * cc - ca * sc;
            out[6] = -sb; out[7] = cb * sc; out[8] = cb * cc;
            break;
        }
        case EXO_ROT_ZYZ:
        {
            const double sa = sin(d[0]), ca = cos(d[0]), sb = sin(d[1]), cb = cos(d[1]), sg = sin(d[2]), cg = cos(d[2]);
            out[0] = ca * cb * cg - sa * sg; out[1] = -ca * cb * sg - sa * cg; out[2] = ca * sb;
            out[3] = sa * cb * cg + ca * sg; out[4] = -sa * cb * sg + ca * cg; out[5] = sa * sb;
            out[6] = -sb * cg; out[7] = sb * sg; out[8] = cb;
            break;
        }
        case EXO_ROT_ANGLE_AXIS:
        {
            const double angle = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
            if (fabs(angle) > 1e-10)
            {
                const double ux = d[0] / angle, uy = d[1] / angle, uz = d[2] / angle;
                const double st = sin(angle), ct = cos(angle), vt = 1 - ct;
                out[0] = ct + vt * ux * ux; out[1] = -uz * st + vt * ux * uy; out[2] = uy * st + vt * ux * uz;
                out[3] = uz * st + vt * ux * uy; out[4] = ct + vt * uy * uy; out[5] = -ux * st + vt * uy * uz;
                out[6] = -uy * st + vt * ux * uz; out[7] = ux * st + vt * uy * uz; out[8] = ct + vt * uz * uz;
            }
            break;
        }
        default:
            memcpy(out, d, 9 * sizeof(double));
            break;
    }
}

}  // namespace exo

// ---- task-map Hessians ---------------------------------------------------------------------------
// KinematicTree::ComputeH (exotica_core/src/kinematic_tree.cpp:947-978) is a pure function of a frame's 6 x N Jacobian:
//     H(0..2)(i, j) = H(0..2)(j, i) = w_i x v_j        (i <= j;  v = linear rows, w = angular rows of the column)
//     H(3..5)(j, i) = w_i x w_j                         (i <  j only: the upper triangle and the diagonal stay zero)
// and the task maps copy blocks of it (eff_frame.cpp:75-90, eff_position.cpp:57-70, eff_orientation.cpp:73-88,
// eff_position_xy.cpp:59-75), combine them (distance.cpp:59-91, including its integer-division exponent pow(|p|^2, 3/2 = 1))
// or, without an overload of their own, fall back to J_row^T J_row (exotica_core/src/task_map.cpp:99-108: JointLimit,
// SphereCollision; JointPose overrides with zero).  One lane = one evaluation; the frame Jacobians come from a raw-frames
// shadow problem evaluated on the same states (capi.cu), the map Jacobian from the caller's J.
namespace exo
{
namespace
{
// Hessian entries are written once and never re-read by this library: streaming stores (EXO_HESS_PLAIN_STORES restores plain ones)
#ifndef EXO_HESS_PLAIN_STORES
#define HST(ptr, v) __stcs((ptr), (v))
#else
#define HST(ptr, v) (*(ptr) = (v))
#endif
__global__ void __launch_bounds__(128) exo_hessian_kernel(const __grid_constant__ HessParams P)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < P.n_evals; e += stride)
    {
        const double* fj = P.frames_jac + e * P.fj_se;  // frames' Jacobians: row (6 f + r), column c at fj[(6 f + r) * fj_sr + c * fj_sc]
        const double* fp = P.frames_phi + e * P.fp_se;  // frames: entry k of frame f at fp[(12 f + k) * fp_sk]
        const double* mj = P.map_jac + e * P.mj_se;     // the problem's own Jacobian
        double* h = P.hess + e * P.h_se;                 // H(row)(i, j) at h[row * h_sr + i * h_si + j * h_sj]
        const int N = P.n_ctrl;
        for (int k = 0; k < P.n_maps; ++k)
        {
            const HessMapP& m = P.map[k];
            double* hm = h + (int64_t)m.start_jacobian * P.h_sr;
            if (m.kind == EXO_MAP_JOINT_POSE || m.kind == EXO_MAP_JOINT_LIMIT || m.kind == EXO_MAP_SPHERE_COLLISION ||
                m.kind == EXO_MAP_EFF_AXIS_ALIGNMENT || m.kind == EXO_MAP_CENTER_OF_MASS || !m.is_used)
            {
                // zero (JointPose, unused maps) or the base-class fallback J_row^T J_row
                const bool outer = m.is_used && m.kind != EXO_MAP_JOINT_POSE;
                for (int r = 0; r < m.length_jacobian; ++r)
                {
                    const double* row = mj + (int64_t)(m.start_jacobian + r) * P.mj_sr;
                    for (int i = 0; i < N; ++i)
                    {
                        const double a = outer ? row[i * P.mj_sc] : 0.0;
                        for (int j = 0; j < N; ++j) HST(&hm[r * P.h_sr + i * P.h_si + j * P.h_sj], outer ? a * row[j * P.mj_sc] : 0.0);
                    }
                }
                continue;
            }
            for (int f = 0; f < m.n_frames; ++f)
            {
                const double* J = fj + (int64_t)(6 * (m.frame_start + f)) * P.fj_sr;
                double* hf = hm + (int64_t)(f * m.rows_per_frame) * P.h_sr;
                double px = 0.0, py = 0.0, pz = 0.0, nrm = 1.0;
                if (m.kind == EXO_MAP_DISTANCE)
                {
                    const double* ph = fp + (int64_t)(12 * (m.frame_start + f)) * P.fp_sk;
                    px = ph[9 * P.fp_sk]; py = ph[10 * P.fp_sk]; pz = ph[11 * P.fp_sk];
                    nrm = norm3_scaled(px, py, pz);
                }
                for (int i = 0; i < N; ++i)
                {
                    const double wix = J[3 * P.fj_sr + i * P.fj_sc], wiy = J[4 * P.fj_sr + i * P.fj_sc], wiz = J[5 * P.fj_sr + i * P.fj_sc];
                    const double vix = J[i * P.fj_sc], viy = J[P.fj_sr + i * P.fj_sc], viz = J[2 * P.fj_sr + i * P.fj_sc];
                    for (int j = i; j < N; ++j)
                    {
                        const double vjx = J[j * P.fj_sc], vjy = J[P.fj_sr + j * P.fj_sc], vjz = J[2 * P.fj_sr + j * P.fj_sc];
                        const double wjx = J[3 * P.fj_sr + j * P.fj_sc], wjy = J[4 * P.fj_sr + j * P.fj_sc], wjz = J[5 * P.fj_sr + j * P.fj_sc];
                        // Hij = axis * column_j with axis = (0, w_i): vel = w_i x v_j, rot = w_i x w_j
                        const double l0 = wiy * vjz - wiz * vjy, l1 = wiz * vjx - wix * vjz, l2 = wix * vjy - wiy * vjx;
                        const double a0 = wiy * wjz - wiz * wjy, a1 = wiz * wjx - wix * wjz, a2 = wix * wjy - wiy * wjx;
                        double* hij = hf + i * P.h_si + j * P.h_sj;
                        double* hji = hf + j * P.h_si + i * P.h_sj;
                        const bool diag = i == j;
                        switch (m.kind)
                        {
                            case EXO_MAP_EFF_FRAME:
                                HST(&hij[0], l0); HST(&hij[P.h_sr], l1); HST(&hij[2 * P.h_sr], l2);
                                HST(&hji[0], l0); HST(&hji[P.h_sr], l1); HST(&hji[2 * P.h_sr], l2);
                                HST(&hij[3 * P.h_sr], 0.0); HST(&hij[4 * P.h_sr], 0.0); HST(&hij[5 * P.h_sr], 0.0);
                                if (!diag)
                                {
                                    HST(&hji[3 * P.h_sr], a0); HST(&hji[4 * P.h_sr], a1); HST(&hji[5 * P.h_sr], a2);
                                }
                                break;
                            case EXO_MAP_EFF_POSITION:
                                HST(&hij[0], l0); HST(&hij[P.h_sr], l1); HST(&hij[2 * P.h_sr], l2);
                                HST(&hji[0], l0); HST(&hji[P.h_sr], l1); HST(&hji[2 * P.h_sr], l2);
                                break;
                            case EXO_MAP_EFF_POSITION_XY:
                                HST(&hij[0], l0); HST(&hij[P.h_sr], l1);
                                HST(&hji[0], l0); HST(&hji[P.h_sr], l1);
                                break;
                            case EXO_MAP_EFF_ORIENTATION:
                                HST(&hij[0], 0.0); HST(&hij[P.h_sr], 0.0); HST(&hij[2 * P.h_sr], 0.0);
                                if (!diag)
                                {
                                    HST(&hji[0], a0); HST(&hji[P.h_sr], a1); HST(&hji[2 * P.h_sr], a2);
                                }
                                break;
                            default:  // EXO_MAP_DISTANCE (distance.cpp:76-90)
                            {
                                const double ni = px * vix + py * viy + pz * viz, nj = px * vjx + py * vjy + pz * vjz;
                                const double part1 = -(ni * nj) / (px * px + py * py + pz * pz);  // pow(., 3 / 2) with integer 3 / 2 == 1
                                const double part2 = ((px * l0 + py * l1 + pz * l2) + (vix * vjx + viy * vjy + viz * vjz)) / nrm;
                                HST(&hij[0], part1 + part2);
                                HST(&hji[0], part1 + part2);
                                break;
                            }
                        }
                    }
                }
            }
        }
    }
}
}  // namespace

cudaError_t launch_hessian(const HessParams& P, cudaStream_t stream)
{
    if (P.n_evals <= 0) return cudaSuccess;
    const int64_t blocks = (P.n_evals + 127) / 128;
    exo_hessian_kernel<<<(int)(blocks < 148 * 16 ? blocks : 148 * 16), 128, 0, stream>>>(P);
    return cudaGetLastError();
}
}  // namespace exo
