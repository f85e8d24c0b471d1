// ik_device.cuh -- per-configuration pieces of IKSolver::Solve's inner loop (exotations/solvers/exotica_ik_solver/src/ik_solver.cpp:
// 92-197), shared by the one-launch-per-stage kernels (ik_kernel.cu) and the single-launch solve of chain programs
// (exo_chain_ik_solve_kernel, chain_kernel.cu).  One lane = one configuration, everything SoA ([k][E]).
#pragma once
#include "cost_device.cuh"

namespace exo
{
namespace dev
{
// alpha_space_ = Eigen::VectorXd::LinSpaced(10, 1.0, 0.1) (ik_solver.cpp:56).  Eigen 3.3 (external; the version Ubuntu 18.04 / 20.04
// ship) evaluates a descending LinSpaced from the high end, (i == 0) ? low : high - (size - 1 - i) * step with step = (high - low) / 9;
// separate multiply and subtract (no FMA) like the host code.
static __device__ __forceinline__ double ik_alpha(int ai)
{
    const double step = (0.1 - 1.0) / 9.0;
    return ai == 0 ? 1.0 : __dsub_rn(0.1, __dmul_rn((double)(9 - ai), step));
}

// error, damped weighted pseudo-inverse step qd (Cholesky; lambda x10 on failure), |cJ|_F -- from the Phi / J of the current state.
// Returns false when the configuration stops here (FunctionTolerance or Divergence).
// LJT > 0: the task space has exactly LJT rows (6: one EffFrame, the usual IK problem) -- every loop unrolls and the LJT x LJT
// system lives in registers; LJT = 0: any size up to IK_MAXLJ, arrays in local memory.
template <int LJT>
static __device__ __forceinline__ bool ik_step_one(const IkParams& K, const CostParams& C, int64_t e)
{
    constexpr int CAP = LJT > 0 ? LJT : IK_MAXLJ;
    const int LJ = LJT > 0 ? LJT : K.length_jac, N = K.n_ctrl;
    double yd[CAP], A[CAP * CAP], Ainv[CAP * CAP], z[CAP];
    {
        double yd_any[IK_MAXLJ];  // eval_ydiff indexes by map: runtime subscripts
        eval_ydiff(C, e, 0, yd_any, 1);
#pragma unroll
        for (int r = 0; r < CAP; ++r) yd[r] = r < LJ ? yd_any[r] : 0.0;
    }
    // S = diag(rho of the row's map); error = ydiff^T S ydiff
    double error = 0.0;
#pragma unroll
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
        return false;
    }
    // cJ = S J (row r, column c at jac[(r * N + c) * E + e]);  A = cJ W^-1 cJ^T + lambda I;  stop = |cJ|_F
    const double* J = K.jac + e;
    const int64_t E = K.n_evals;
    double stop2 = 0.0;
#pragma unroll
    for (int r = 0; r < LJ; ++r)
#pragma unroll
        for (int s2 = 0; s2 <= r; ++s2)
        {
            double acc = 0.0;
            for (int c = 0; c < N; ++c)
            {
                const double a = K.row_rho[r] * J[(int64_t)(r * N + c) * E], b = K.row_rho[s2] * J[(int64_t)(s2 * N + c) * E];
                acc += a * K.w_inv[c] * b;
                if (s2 == r) stop2 += a * a;
            }
            A[r * LJ + s2] = acc;
            A[s2 * LJ + r] = acc;
        }
    double lambda = K.lambda[e];
    bool ok = false;
    while (!ok)
    {
        // Eigen::LLT (lower Cholesky of A + lambda I); fails on a non-positive pivot
        ok = true;
#pragma unroll
        for (int j = 0; j < LJ; ++j)
        {
            if (!ok) continue;  // (a break would keep the loop from unrolling)
            double d = A[j * LJ + j] + lambda;
#pragma unroll
            for (int k = 0; k < j; ++k) d -= Ainv[j * LJ + k] * Ainv[j * LJ + k];
            if (!(d > 0.0))
            {
                ok = false;
                continue;
            }
            const double ljj = sqrt(d);
            Ainv[j * LJ + j] = ljj;  // Ainv holds L for now
#pragma unroll
            for (int i = j + 1; i < LJ; ++i)
            {
                double v = A[i * LJ + j];
#pragma unroll
                for (int k = 0; k < j; ++k) v -= Ainv[i * LJ + k] * Ainv[j * LJ + k];
                Ainv[i * LJ + j] = v / ljj;
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
                return false;
            }
        }
    }
    K.lambda[e] = lambda;
    // z = (A + lambda I)^-1 yd by forward / backward substitution with L (the reference forms the explicit inverse and
    // multiplies; same result up to rounding)
#pragma unroll
    for (int i = 0; i < LJ; ++i)
    {
        double v = yd[i];
#pragma unroll
        for (int k = 0; k < i; ++k) v -= Ainv[i * LJ + k] * z[k];
        z[i] = v / Ainv[i * LJ + i];
    }
#pragma unroll
    for (int i = LJ - 1; i >= 0; --i)
    {
        double v = z[i];
#pragma unroll
        for (int k = i + 1; k < LJ; ++k) v -= Ainv[k * LJ + i] * z[k];
        z[i] = v / Ainv[i * LJ + i];
    }
    // qd = W^-1 cJ^T z
    for (int c = 0; c < N; ++c)
    {
        double acc = 0.0;
#pragma unroll
        for (int r = 0; r < LJ; ++r) acc += K.row_rho[r] * J[(int64_t)(r * N + c) * E] * z[r];
        K.qd[(int64_t)c * E + e] = K.w_inv[c] * acc;
    }
    K.stop[e] = sqrt(stop2);
    if (K.single_iteration && K.max_step != 0.0)
    {
        // ScaleToStepSize + q_ -= qd_ (ik_solver.cpp:139-143,235-243): no line search in the interactive one-step mode
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
    return true;
}

// step / gradient tolerance checks and the iteration counter (ik_solver.cpp:164-186); false when the configuration stops
static __device__ __forceinline__ bool ik_finish_one(const IkParams& K, int64_t e, bool last_iteration)
{
    K.iterations[e] += 1;
    double step = 0.0;
    for (int c = 0; c < K.n_ctrl; ++c)
    {
        const double v = K.qd[(int64_t)c * K.n_evals + e];
        step += v * v;
    }
    if (step < K.step_tolerance)
    {
        K.termination[e] = IK_TERM_STEP_TOLERANCE;
        K.active[e] = 0;
        return false;
    }
    if (K.stop[e] < K.gradient_tolerance)
    {
        K.termination[e] = IK_TERM_GRADIENT_TOLERANCE;
        K.active[e] = 0;
        return false;
    }
    if (K.single_iteration && K.max_step == 0.0)
    {
        // adapt the regularisation to the step length the line search ended on (ik_solver.cpp:187-199; ik_solver.h:81-97)
        const double steplength = K.steplen[e];
        double lambda = K.lambda[e];
        if (steplength > K.th_stepdec)
        {
            lambda /= 10.0;
            if (lambda < K.regmin) lambda = K.regmin;
        }
        if (steplength <= K.th_stepinc)
        {
            lambda *= 10.0;
            if (lambda > K.regmax) lambda = K.regmax;
            if (lambda == K.regmax)
            {
                K.lambda[e] = lambda;
                K.termination[e] = IK_TERM_DIVERGENCE;
                K.active[e] = 0;
                return false;
            }
        }
        K.lambda[e] = lambda;
    }
    if (last_iteration)
    {
        K.termination[e] = IK_TERM_ITERATION_LIMIT;
        return false;
    }
    return true;
}
}  // namespace dev
}  // namespace exo
