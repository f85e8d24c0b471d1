// ik_kernel.cu -- batched IK step for multi-start inverse kinematics.
//
// Reference path replaced, for E independent start states at once: the inner loop of IKSolver::Solve
// (exotations/solvers/exotica_ik_solver/src/ik_solver.cpp:92-197) on an UnconstrainedEndPoseProblem:
//     error = ydiff^T S ydiff;  yd = S ydiff;  cJ = S J
//     qd = W^-1 cJ^T (cJ W^-1 cJ^T + lambda I)^-1 yd           (Cholesky; lambda x10 on failure, :110-133)
//     backtracking line search over alpha = linspace(1, 0.1, 10): accept the first q - alpha qd whose cost drops (:145-161)
//     stop on cost < Tolerance, |qd|^2 < StepTolerance, |cJ|_F < GradientTolerance, lambda > MaximumRegularization
// The forward kinematics / Jacobian evaluations are the batched kernels of this library (one launch per iteration for (Phi, J),
// one FK-only launch per line-search step); the kernels here are the per-configuration linear algebra and bookkeeping between
// them.  One lane = one configuration, everything SoA ([k][E]); converged configurations keep their state and idle.
#include "ik_device.cuh"

namespace exo
{
namespace
{
using namespace dev;

// cost of the trial state, per configuration
__device__ __forceinline__ double ik_cost(const CostParams& C, int64_t e)
{
    return eval_cost(C, e, 0);
}

__global__ void __launch_bounds__(128) exo_ik_step_kernel(const __grid_constant__ IkParams K, const __grid_constant__ CostParams C)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= K.n_evals || !K.active[e]) return;
    if (K.length_jac == 6)
        ik_step_one<6>(K, C, e);
    else
        ik_step_one<0>(K, C, e);
}

// The same step for task spaces of up to IK_BIG_MAXLJ rows and IK_MAXN joints (whole-body problems: four end-effector frames of a
// humanoid are 24 rows, 38 joints), ONE WARP PER CONFIGURATION on the compacted list of active configurations
// (exo_ik_compact_kernel below).  One lane per configuration with the system in a global workspace was bound by the latency of one
// lane's ~13 000 dependent accesses: 2.0 ms per step for 65 536 configurations and still 0.49 ms for a single one, whatever the
// number of active configurations -- most starts converge after two or three iterations.  Here the warp keeps cJ = S J and
// A = cJ W^-1 cJ^T in shared memory: the entries of A are spread over the lanes (each one still sums its columns in order), the
// Cholesky factorisation goes column by column with the rows of a column spread over the lanes (every entry subtracts its products
// in the same order as Eigen::LLT's), the two triangular solves and |cJ|_F run on one lane (their summation order is sequential by
// definition), and W^-1 cJ^T z is spread over the columns.  A's lower triangle holds A, the strict upper triangle the Cholesky factor
// transposed, so a failed factorisation restarts with a larger lambda without rebuilding A.  Same numbers as the one-lane form.
__global__ void __launch_bounds__(128) exo_ik_step_warp_kernel(const __grid_constant__ IkParams K, const __grid_constant__ CostParams C, int per_warp)
{
    extern __shared__ double ik_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
    const int LJ = K.length_jac, N = K.n_ctrl;
    const int64_t E = K.n_evals;
    double* const cj = ik_smem + (size_t)warp * per_warp;  // [LJ][N]
    double* const A = cj + LJ * N;                         // [LJ][LJ]
    double* const yd = A + LJ * LJ;                        // [LJ]
    double* const diag = yd + LJ;                          // [LJ]
    const int n_list = *K.n_list;
    for (int slot = blockIdx.x * wpc + warp; slot < n_list; slot += gridDim.x * wpc)
    {
        const int64_t e = K.list[slot];
        __syncwarp();
        int go = 1;
        if (lane == 0)
        {
            eval_ydiff(C, e, 0, yd, 1);
            double error = 0.0;
            for (int r = 0; r < LJ; ++r)
            {
                const double s = K.row_rho[r];
                error += yd[r] * s * yd[r];
                yd[r] *= s;
            }
            K.cost[e] = error;
            K.err_prev[e] = error;
            K.ls_done[e] = 0;
            if (error < K.tolerance)
            {
                K.termination[e] = IK_TERM_FUNCTION_TOLERANCE;
                K.active[e] = 0;
                go = 0;
            }
        }
        go = __shfl_sync(0xffffffffu, go, 0);
        if (!go) continue;
        const double* J = K.jac + e;
        for (int idx = lane; idx < LJ * N; idx += 32) cj[idx] = K.row_rho[idx / N] * J[(int64_t)idx * E];
        __syncwarp();
        // A(r, s) = sum_c (cJ(r, c) W^-1(c)) cJ(s, c), s <= r: entries over the lanes, columns in order
        for (int t = lane; t < LJ * (LJ + 1) / 2; t += 32)
        {
            int r = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
            while (r * (r + 1) / 2 > t) --r;
            while ((r + 1) * (r + 2) / 2 <= t) ++r;
            const int s2 = t - r * (r + 1) / 2;
            const double *a = cj + r * N, *b = cj + s2 * N;
            double acc = 0.0;
            for (int c = 0; c < N; ++c) acc += (a[c] * K.w_inv[c]) * b[c];
            A[r * LJ + s2] = acc;
        }
        __syncwarp();
        double lambda = K.lambda[e];
        bool ok = false, diverged = false;
        while (!ok && !diverged)
        {
            // Eigen::LLT of A + lambda I; L(i, j), i > j, is kept at the transposed position (j, i).  Every lane forms the pivot
            // (the same reads, the same order), the rows below it are spread over the lanes.
            ok = true;
            for (int j = 0; j < LJ && ok; ++j)
            {
                double d = A[j * LJ + j] + lambda;
                for (int k = 0; k < j; ++k)
                {
                    const double l = A[k * LJ + j];
                    d -= l * l;
                }
                if (!(d > 0.0))
                {
                    ok = false;
                    break;
                }
                const double ljj = sqrt(d);
                if (lane == 0) diag[j] = ljj;
                for (int i = j + 1 + lane; i < LJ; i += 32)
                {
                    double v = A[i * LJ + j];
                    for (int k = 0; k < j; ++k) v -= A[k * LJ + i] * A[k * LJ + j];
                    A[j * LJ + i] = v / ljj;
                }
                __syncwarp();
            }
            if (!ok)
            {
                lambda *= 10.0;  // IncreaseRegularization (ik_solver.h: regfactor_ = 10)
                if (lambda > K.regmax) diverged = true;
            }
        }
        if (diverged)
        {
            if (lane == 0)
            {
                K.lambda[e] = K.regmax;
                K.termination[e] = IK_TERM_DIVERGENCE;
                K.active[e] = 0;
            }
            continue;
        }
        if (lane == 0)
        {
            K.lambda[e] = lambda;
            // z = (A + lambda I)^-1 yd: forward / backward substitution with L (z overwrites yd)
            for (int i = 0; i < LJ; ++i)
            {
                double v = yd[i];
                for (int k = 0; k < i; ++k) v -= A[k * LJ + i] * yd[k];
                yd[i] = v / diag[i];
            }
            for (int i = LJ - 1; i >= 0; --i)
            {
                double v = yd[i];
                for (int k = i + 1; k < LJ; ++k) v -= A[i * LJ + k] * yd[k];
                yd[i] = v / diag[i];
            }
            double stop2 = 0.0;  // |cJ|_F^2, columns outermost
            for (int c = 0; c < N; ++c)
                for (int r = 0; r < LJ; ++r) stop2 += cj[r * N + c] * cj[r * N + c];
            K.stop[e] = sqrt(stop2);
        }
        __syncwarp();
        // qd = W^-1 cJ^T z
        double max_vel = 0.0;
        for (int c = lane; c < N; c += 32)
        {
            double acc = 0.0;
            for (int r = 0; r < LJ; ++r) acc += cj[r * N + c] * yd[r];
            const double v = K.w_inv[c] * acc;
            K.qd[(int64_t)c * E + e] = v;
            max_vel = fmax(max_vel, fabs(v));
        }
        if (K.single_iteration && K.max_step != 0.0)
        {
            // ScaleToStepSize + q_ -= qd_ (ik_solver.cpp:139-143,235-243)
            for (int o = 16; o > 0; o >>= 1) max_vel = fmax(max_vel, __shfl_xor_sync(0xffffffffu, max_vel, o));
            for (int c = lane; c < N; c += 32)
            {
                double v = K.qd[(int64_t)c * E + e];
                if (max_vel > K.max_step) v = v * K.max_step / max_vel;
                K.qd[(int64_t)c * E + e] = v;
                K.q[(int64_t)c * E + e] -= v;
            }
            if (lane == 0) K.ls_done[e] = 1;
        }
    }
}

// the active configurations, compacted (warp-aggregated append; the order does not matter: every configuration's step is its own)
__global__ void __launch_bounds__(128) exo_ik_compact_kernel(const __grid_constant__ IkParams K)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool on = e < K.n_evals && K.active[e];
    const unsigned m = __ballot_sync(0xffffffffu, on);
    if (m == 0) return;
    const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(K.n_list, __popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (on) K.list[base + __popc(m & ((1u << lane) - 1u))] = (int32_t)e;
}

// q_trial = q - alpha qd for the configurations still searching; the others keep q (their evaluation is ignored)
__global__ void __launch_bounds__(128) exo_ik_trial_kernel(const __grid_constant__ IkParams K, double alpha, int first)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= K.n_evals) return;
    const bool search = K.active[e] && !K.ls_done[e];
    // configurations that are not searching keep q: written with the first step length of a line search, unchanged afterwards (a
    // configuration that accepts a step copies q_trial into q), so the later step lengths touch the searching ones only
    if (!search && !first) return;
    for (int c = 0; c < K.n_ctrl; ++c)
    {
        const int64_t i = (int64_t)c * K.n_evals + e;
        K.q_trial[i] = search ? K.q[i] - alpha * K.qd[i] : K.q[i];
    }
}

// accept the trial state if its cost dropped (ik_solver.cpp:153-158)
__global__ void __launch_bounds__(128) exo_ik_accept_kernel(const __grid_constant__ IkParams K, const __grid_constant__ CostParams C, double alpha)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= K.n_evals || !K.active[e] || K.ls_done[e]) return;
    const double error = ik_cost(C, e);
    K.cost[e] = error;  // error_: the cost at the last state evaluated
    K.steplen[e] = alpha;
    if (error < K.err_prev[e])
    {
        for (int c = 0; c < K.n_ctrl; ++c)
        {
            const int64_t i = (int64_t)c * K.n_evals + e;
            K.q[i] = K.q_trial[i];
            K.qd[i] *= alpha;
        }
        K.ls_done[e] = 1;
    }
}

// step / gradient tolerance checks and the iteration counter (ik_solver.cpp:164-186)
__global__ void __launch_bounds__(128) exo_ik_finish_kernel(const __grid_constant__ IkParams K, int last_iteration)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= K.n_evals || !K.active[e]) return;
    if (ik_finish_one(K, e, last_iteration != 0)) atomicAdd(K.n_active, 1);  // read by the host every few iterations to stop early
}

__global__ void __launch_bounds__(128) exo_ik_init_kernel(const __grid_constant__ IkParams K, double lambda0)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= K.n_evals) return;
    K.active[e] = 1;
    K.ls_done[e] = 0;
    K.iterations[e] = 0;
    K.termination[e] = IK_TERM_ITERATION_LIMIT;
    K.lambda[e] = lambda0;
    K.cost[e] = 0.0;
    K.stop[e] = 0.0;
    K.steplen[e] = 0.0;
}
}  // namespace

static int ik_blocks(int64_t n)
{
    return (int)((n + 127) / 128);
}

cudaError_t launch_ik_init(const IkParams& K, double lambda0, cudaStream_t s)
{
    exo_ik_init_kernel<<<ik_blocks(K.n_evals), 128, 0, s>>>(K, lambda0);
    return cudaGetLastError();
}
cudaError_t launch_ik_step(const IkParams& K, const CostParams& C, cudaStream_t s)
{
    if (K.warp_step)
    {
        cudaError_t e = cudaMemsetAsync(K.n_list, 0, sizeof(int32_t), s);
        if (e != cudaSuccess) return e;
        exo_ik_compact_kernel<<<ik_blocks(K.n_evals), 128, 0, s>>>(K);
        // shared memory per warp: cJ, A, yd, diag; as many warps per CTA (at most four) as 200 KB hold
        const int per_warp = K.length_jac * K.n_ctrl + K.length_jac * K.length_jac + 2 * K.length_jac;
        int wpc = 4;
        while (wpc > 1 && (size_t)wpc * per_warp * sizeof(double) > 200 * 1024) --wpc;
        const size_t smem = (size_t)wpc * per_warp * sizeof(double);
        static size_t smem_set = 48 * 1024;
        if (smem > smem_set)
        {
            e = cudaFuncSetAttribute(exo_ik_step_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            smem_set = smem;
        }
        // one warp per active configuration, walking the compacted list: enough CTAs to fill the machine a few times over
        const int64_t want = (K.n_evals + wpc - 1) / wpc;
        exo_ik_step_warp_kernel<<<(int)(want < 148 * 16 ? want : 148 * 16), wpc * 32, smem, s>>>(K, C, per_warp);
    }
    else
        exo_ik_step_kernel<<<ik_blocks(K.n_evals), 128, 0, s>>>(K, C);
    return cudaGetLastError();
}
cudaError_t launch_ik_trial(const IkParams& K, double alpha, bool first, cudaStream_t s)
{
    exo_ik_trial_kernel<<<ik_blocks(K.n_evals), 128, 0, s>>>(K, alpha, first ? 1 : 0);
    return cudaGetLastError();
}
cudaError_t launch_ik_accept(const IkParams& K, const CostParams& C, double alpha, cudaStream_t s)
{
    exo_ik_accept_kernel<<<ik_blocks(K.n_evals), 128, 0, s>>>(K, C, alpha);
    return cudaGetLastError();
}
cudaError_t launch_ik_finish(const IkParams& K, int last_iteration, cudaStream_t s)
{
    exo_ik_finish_kernel<<<ik_blocks(K.n_evals), 128, 0, s>>>(K, last_iteration);
    return cudaGetLastError();
}

}  // namespace exo
