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
// humanoid are 24 rows, 38 joints): one lane per configuration as above, but A = cJ W^-1 cJ^T lives in a global workspace laid out
// [entry][E] (every access of a warp is one coalesced segment) -- its lower triangle holds A, the strict upper triangle the Cholesky
// factor transposed, so a failed factorisation restarts with a larger lambda without rebuilding A.
__global__ void __launch_bounds__(128) exo_ik_step_big_kernel(const __grid_constant__ IkParams K, const __grid_constant__ CostParams C)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= K.n_evals || !K.active[e]) return;
    const int LJ = K.length_jac, N = K.n_ctrl;
    const int64_t E = K.n_evals;
    double yd[IK_BIG_MAXLJ], diag[IK_BIG_MAXLJ];
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
        return;
    }
    const double* J = K.jac + e;
    double* A = K.big_A + e;  // entry (r, s) at A[(r * LJ + s) * E]
    // lower triangle of A in 8 x 8 register tiles: per tile and column of cJ 16 loads feed 64 fused multiply-adds and the tile is
    // stored once (a rank-one update of A in the workspace per column moved 25 times the data: 2 x 300 accesses of A per column
    // against 16 loads of J per tile and column); every entry still sums its columns in order, so the numbers are unchanged
    for (int r0 = 0; r0 < LJ; r0 += 8)
        for (int s0 = 0; s0 <= r0; s0 += 8)
        {
            double acc[8][8];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = 0.0;
            for (int c = 0; c < N; ++c)
            {
                const double wi = K.w_inv[c];
                double a[8], b[8];
#pragma unroll
                for (int i = 0; i < 8; ++i)
                {
                    const int r = r0 + i, s2 = s0 + i;
                    const double cr = r < LJ ? K.row_rho[r] * J[(int64_t)(r * N + c) * E] : 0.0;
                    b[i] = s0 == r0 ? cr : K.row_rho[s2] * J[(int64_t)(s2 * N + c) * E];  // s0 < r0: a full tile of rows below LJ
                    a[i] = cr * wi;
                }
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[i][j] += a[i] * b[j];
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (r0 + i < LJ && s0 + j <= r0 + i) A[(int64_t)((r0 + i) * LJ + s0 + j) * E] = acc[i][j];
        }
    double lambda = K.lambda[e];
    bool ok = false;
    while (!ok)
    {
        // Eigen::LLT of A + lambda I; L(i, j), i > j, is kept at the transposed position (j, i)
        ok = true;
        for (int j = 0; j < LJ && ok; ++j)
        {
            double d = A[(int64_t)(j * LJ + j) * E] + lambda;
            for (int k = 0; k < j; ++k)
            {
                const double l = A[(int64_t)(k * LJ + j) * E];
                d -= l * l;
            }
            if (!(d > 0.0))
            {
                ok = false;
                break;
            }
            const double ljj = sqrt(d);
            diag[j] = ljj;
            for (int i = j + 1; i < LJ; ++i)
            {
                double v = A[(int64_t)(i * LJ + j) * E];
                for (int k = 0; k < j; ++k) v -= A[(int64_t)(k * LJ + i) * E] * A[(int64_t)(k * LJ + j) * E];
                A[(int64_t)(j * LJ + i) * E] = v / ljj;
            }
        }
        if (!ok)
        {
            lambda *= 10.0;  // IncreaseRegularization (ik_solver.h: regfactor_ = 10)
            if (lambda > K.regmax)
            {
                K.lambda[e] = K.regmax;
                K.termination[e] = IK_TERM_DIVERGENCE;
                K.active[e] = 0;
                return;
            }
        }
    }
    K.lambda[e] = lambda;
    // z = (A + lambda I)^-1 yd: forward / backward substitution with L (z overwrites yd)
    for (int i = 0; i < LJ; ++i)
    {
        double v = yd[i];
        for (int k = 0; k < i; ++k) v -= A[(int64_t)(k * LJ + i) * E] * yd[k];
        yd[i] = v / diag[i];
    }
    for (int i = LJ - 1; i >= 0; --i)
    {
        double v = yd[i];
        for (int k = i + 1; k < LJ; ++k) v -= A[(int64_t)(i * LJ + k) * E] * yd[k];
        yd[i] = v / diag[i];
    }
    // qd = W^-1 cJ^T z
    double stop2 = 0.0;  // |cJ|_F^2, columns outermost
    for (int c = 0; c < N; ++c)
    {
        double acc = 0.0;
        for (int r = 0; r < LJ; ++r)
        {
            const double cj = K.row_rho[r] * J[(int64_t)(r * N + c) * E];
            stop2 += cj * cj;
            acc += cj * yd[r];
        }
        K.qd[(int64_t)c * E + e] = K.w_inv[c] * acc;
    }
    K.stop[e] = sqrt(stop2);
    if (K.single_iteration && K.max_step != 0.0)
    {
        // ScaleToStepSize + q_ -= qd_ (ik_solver.cpp:139-143,235-243)
        double max_vel = 0.0;
        for (int c = 0; c < N; ++c) max_vel = fmax(max_vel, fabs(K.qd[(int64_t)c * E + e]));
        for (int c = 0; c < N; ++c)
        {
            double v = K.qd[(int64_t)c * E + e];
            if (max_vel > K.max_step) v = v * K.max_step / max_vel;
            K.qd[(int64_t)c * E + e] = v;
            K.q[(int64_t)c * E + e] -= v;
        }
        K.ls_done[e] = 1;
    }
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
    if (K.big_A)
        exo_ik_step_big_kernel<<<ik_blocks(K.n_evals), 128, 0, s>>>(K, C);
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
