// chain_kernel.cu -- branch-free Denavit-Hartenberg kernel for serial-chain programs (sm_100a).
//
// The general interpreter (kernels.cu) decodes ops from shared memory.  For the most common robot shape -- a serial
// manipulator whose requested frames all hang off the last link (the LWR of BASELINE.json configs 1-3) -- the host
// lowers the compiled tree program once more (chain_lower.cpp) into modified-DH form,
//
//     D_j = D_{j-1} * Rx(alpha_j) * Tx(a_j) * Rz(theta_j + q_j) * Tz(d_j),
//
// and passes all of its constants BY VALUE as the kernel parameter: they live in constant bank 0 at compile-time
// offsets, so the fully unrolled chain is straight-line code -- 30 FP64 instructions per joint, no per-joint case
// analysis, no shared-memory program, no __syncthreads.  The joint axis is the z column of the frame in front of Rz
// and its origin lies on the axis, which is all a Jacobian column needs (KinematicTree::ComputeJ,
// exotica_core/src/kinematic_tree.cpp:919-945).
//
// One lane = one evaluation, one 64-thread CTA per 64 evaluations (the hardware CTA scheduler balances the two dies better
// than a grid of long-lived CTAs with a static split, tools/hbm_probe.cu); every CTA pulls the joint values of a CTA 3/4 of a
// resident generation behind it into L2 (prefetch.global.L2), so the only exposed memory operations are coalesced
// streaming stores of Phi and J.  sin/cos use a Cody-Waite reduction and fdlibm's minimax kernels with their coefficients in __constant__
// memory (CUDA's sincos spends most of its instructions materialising 64-bit immediates and on int<->double
// conversions).  HBM-bound FP64 streaming work: no tensor cores.
//
// Variants (exo_problem_set_kernel; measured in DESIGN.md):
//   CHAIN_REGS      every joint's world axis and origin stay in registers until the end-effector point is known
//   CHAIN_TWO_PASS  the chain is composed twice; the second pass emits each joint's column on the fly (one requested
//                   frame; fewer registers, 30 more FP64 instructions per joint)
//
// Reference path replaced: KinematicTree::Update/UpdateTree/FK/ComputeJ (exotica_core/src/kinematic_tree.cpp:678-737,
// 829-867,919-945) + EffFrame/EffPosition/JointLimit::Update (exotations/exotica_core_task_maps/src/*.cpp) +
// Phi.SetZero/jacobian.setZero (exotica_core/src/problems/abstract_time_indexed_problem.cpp:243-244).
#include <stdlib.h>
#include <string.h>

#define EXO_PHI_STREAMING 1  // this translation unit writes Phi and the generic Jacobian rows with streaming stores too (device_math.cuh):
#define EXO_STORE_CS 1       // all stores streaming: cfg2a 86.3 -> 85.0 us, cfg2b 55.2 -> 53.7 us, cfg3 75.6 -> 68.4 us

#include "chain_device.cuh"

namespace exo
{
namespace
{
using namespace dev;

// ALLREV is a bit set: 1 = every joint revolute, 2 = signed-swap link step (dh_link), 4 = 128-register build of the signed-swap
// kernel (16 instead of 14 warps per SM).  Measured: the extra warps pay where arithmetic bounds the kernel (cfg2b, EffPosition
// only: 60.5 -> 55.2 us) and cost 1.5-2 us where the store pattern does (cfg2a 86.1 -> 87.6 us, cfg3 76 -> 78.5 us), so the launcher
// picks bit 4 only for position-only frames.  Denser still does not pay: 112 / 96 registers (18 / 20 warps per SM, 20 / 100 bytes of
// spills) take cfg2b from 51.7 to 53.5 / 55.8 us.
#ifndef EXO_CHAIN_DENSE_MAXNREG
#define EXO_CHAIN_DENSE_MAXNREG 128
#endif
// GEN: some requested frame belongs to a map that keeps other rows than (v, w) / (v) -- EffOrientation, EffPositionXY,
// Distance; the plain instantiations do not carry that code (it costs the hot kernel registers)
template <int NJ, int ALLREV, int VARIANT, bool GEN>
__global__ void __maxnreg__((ALLREV & 4) ? EXO_CHAIN_DENSE_MAXNREG : ChainOcc<VARIANT>::maxnreg) exo_chain_kernel(const __grid_constant__ ChainParams P)
{
    const int64_t stride = (int64_t)gridDim.x * CTPB;
    int64_t e = (int64_t)blockIdx.x * CTPB + threadIdx.x;
    // running byte pointers of this lane's evaluation
    const char* xe = reinterpret_cast<const char*>(P.x) + e * P.x_se_b;
    char* phi_e = reinterpret_cast<char*>(P.phi) + e * P.phi_se_b;
    char* jac_e = P.jac ? reinterpret_cast<char*>(P.jac) + e * P.j_se_b : nullptr;
    const int64_t x_step = stride * P.x_se_b, phi_step = stride * P.phi_se_b, jac_step = P.jac ? stride * P.j_se_b : 0;
    const int64_t sr_b = P.j_sr_b;
    const int64_t pf_dist = stride < P.n_evals ? stride : P.prefetch_ahead, pf_step = pf_dist * P.x_se_b;
    for (; e < P.n_evals; e += stride, xe += x_step, phi_e += phi_step, jac_e += jac_step)
    {
        double q[NJ];
#pragma unroll
        for (int j = 0; j < NJ; ++j) q[j] = *reinterpret_cast<const double*>(xe + P.joint[j].x_byte_off);
        // pull joint values that will be needed soon into L2 (no registers held): this lane's next evaluation, or -- when
        // the grid covers the batch and every lane has one evaluation -- those of a CTA a few generations behind
        if (e + pf_dist < P.n_evals)
        {
#pragma unroll
            for (int j = 0; j < NJ; ++j) prefetch_l2(xe + pf_step + P.joint[j].x_byte_off);
        }
        if (P.has_prefix)
        {
            Sink o;
            o.phi = reinterpret_cast<double*>(phi_e);
            o.jac = reinterpret_cast<double*>(jac_e);
            o.phi_sk = P.phi_sk;
            o.j_sr = P.j_sr;
            o.j_sc = P.j_sc;
            chain_prefix(P, o, reinterpret_cast<const double*>(xe), P.x_sk);
        }
        // all joint angles first: NJ independent sin/cos evaluations keep the FP64 pipe busy
        double sn[NJ], cs[NJ];
        bool big = false;
#pragma unroll
        for (int j = 0; j < NJ; ++j)
        {
            const double th = fma(P.joint[j].qmult, q[j], P.joint[j].th0);
            big = big || !(fabs(th) <= 1.0e5);
            sn[j] = 0.0;
            cs[j] = 1.0;
            if (ALLREV || !P.joint[j].prismatic) fast_sincos(th, sn[j], cs[j]);
        }
        if (big)
        {
            chain_slow_eval(P, xe, phi_e, jac_e);
            continue;
        }
        Fr cur = load_frame(P.base);
        if (VARIANT == CHAIN_TWO_PASS && !GEN && P.n_frames == 1)
        {
            ChainUnroll<NJ, ALLREV, 0>::compose(P, cur, q, sn, cs);
            const ChainFrameP& fp = P.frame[0];
            const Fr A = chain_frame_a(fp, cur);
            if (jac_e)
            {
                cur = load_frame(P.base);
                const bool six = fp.row_mode == ROWS_VW;
                if (fp.b_const)
                {
                    if (six) ChainUnroll<NJ, ALLREV, 0>::template emit<true, true>(P, jac_e, sr_b, cur, q, sn, cs, A.x, A.y, A.z);
                    else ChainUnroll<NJ, ALLREV, 0>::template emit<true, false>(P, jac_e, sr_b, cur, q, sn, cs, A.x, A.y, A.z);
                }
                else
                {
                    if (six) ChainUnroll<NJ, ALLREV, 0>::template emit<false, true>(P, jac_e, sr_b, cur, q, sn, cs, A.x, A.y, A.z);
                    else ChainUnroll<NJ, ALLREV, 0>::template emit<false, false>(P, jac_e, sr_b, cur, q, sn, cs, A.x, A.y, A.z);
                }
            }
            chain_frame_phi(fp, A, phi_e, P.phi_sk);
        }
        else
        {
            double W[NJ][3], PJ[NJ][3];
            ChainUnroll<NJ, ALLREV, 0>::record(P, cur, q, sn, cs, W, PJ);
            // Jacobian columns first (they release the joint data), the rotation encodings of Phi after them
            for (int f = 0; f < P.n_frames; ++f)
            {
                const ChainFrameP& fp = P.frame[f];
                const Fr A = chain_frame_a(fp, cur);
                if (GEN && jac_e && fp.row_mode > ROWS_V)
                {
                    // EffOrientation / EffPositionXY / Distance: the generic emitter; Distance needs Phi.p / |Phi.p|
                    double ux = 0.0, uy = 0.0, uz = 0.0;
                    if (fp.row_mode == ROWS_DIST)
                    {
                        const Fr Ph = fp.b_const ? fr_mul(load_frame(fp.binv), A) : A;
                        const double nrm = norm3_scaled(Ph.x, Ph.y, Ph.z);
                        ux = Ph.x / nrm; uy = Ph.y / nrm; uz = Ph.z / nrm;
                    }
                    ChainUnroll<NJ, ALLREV, 0>::columns_g(P, f, jac_e, sr_b, W, PJ, A.x, A.y, A.z, ux, uy, uz);
                }
                else if (jac_e)
                {
                    const bool six = fp.row_mode == ROWS_VW;
                    if (fp.b_const)
                    {
                        if (six) ChainUnroll<NJ, ALLREV, 0>::template columns<true, true>(P, f, jac_e, sr_b, W, PJ, A.x, A.y, A.z);
                        else ChainUnroll<NJ, ALLREV, 0>::template columns<true, false>(P, f, jac_e, sr_b, W, PJ, A.x, A.y, A.z);
                    }
                    else
                    {
                        if (six) ChainUnroll<NJ, ALLREV, 0>::template columns<false, true>(P, f, jac_e, sr_b, W, PJ, A.x, A.y, A.z);
                        else ChainUnroll<NJ, ALLREV, 0>::template columns<false, false>(P, f, jac_e, sr_b, W, PJ, A.x, A.y, A.z);
                    }
                }
            }
            for (int f = 0; f < P.n_frames; ++f)
            {
                const ChainFrameP& fp = P.frame[f];
                chain_frame_phi(fp, chain_frame_a(fp, cur), phi_e, P.phi_sk);
            }
        }
    }
}

// ---- fused forward kinematics + task cost + cost gradient --------------------------------------------------------------------
// grad[e][n] = 2 ct ( J^T S ydiff )[n] + 2 ct W (xdiff[t] - xdiff[t+1])[n]   (AbstractTimeIndexedProblem::GetCostJacobian,
// abstract_time_indexed_problem.cpp:309-342; T = 1: UnconstrainedEndPoseProblem::GetScalarJacobian) without ever materialising
// J: every joint's column (KinematicTree::ComputeJ) is dotted with S ydiff of its frame the moment it exists.  For a frame with
// a constant base, column = Binv [v; w], so the dot product is taken against Binv^T (S ydiff) instead -- one 3x3 product per frame,
// not per column.  HBM traffic per evaluation: N doubles in (+ 2 N neighbours from L2 for the transition term), N + 1 out.
template <int J, int ALLREV>
__device__ __forceinline__ double chain_dot(const ChainParams& P, double wx, double wy, double wz, double px, double py, double pz, double ax, double ay,
                                            double az, const double (&yd)[6], bool six)
{
    const ChainJointP& jp = P.joint[J];
    if (jp.col < 0) return 0.0;
    if (!ALLREV && jp.prismatic) return wx * yd[0] + wy * yd[1] + wz * yd[2];
    const double dx = ax - px, dy = ay - py, dz = az - pz;
    double s = (wy * dz - wz * dy) * yd[0] + (wz * dx - wx * dz) * yd[1] + (wx * dy - wy * dx) * yd[2];
    if (six) s += wx * yd[3] + wy * yd[4] + wz * yd[5];
    return s;
}

template <int NJ, int ALLREV, int J>
struct ChainDots
{
    static __device__ __forceinline__ void run(const ChainParams& P, const double (&W)[NJ][3], const double (&PJ)[NJ][3], double ax, double ay, double az,
                                               const double (&yd)[6], bool six, double (&g)[NJ])
    {
        g[J] += chain_dot<J, ALLREV>(P, W[J][0], W[J][1], W[J][2], PJ[J][0], PJ[J][1], PJ[J][2], ax, ay, az, yd, six);
        ChainDots<NJ, ALLREV, J + 1>::run(P, W, PJ, ax, ay, az, yd, six, g);
    }
};
template <int NJ, int ALLREV>
struct ChainDots<NJ, ALLREV, NJ>
{
    static __device__ __forceinline__ void run(const ChainParams&, const double (&)[NJ][3], const double (&)[NJ][3], double, double, double,
                                               const double (&)[6], bool, double (&)[NJ]) {}
};

// JointLimit and transition terms of column n (both diagonal in the joint index): returns their share of the gradient before
// the factor 2 ct and adds their share of the cost to c
__device__ __forceinline__ double costgrad_column_extra(const ChainParams& P, const ChainGradParams& G, const char* xe, int t, int n, double rho_jl,
                                                        const double* goal, double& c)
{
    const double xn = *reinterpret_cast<const double*>(xe + (int64_t)n * P.x_sk * 8);
    double s = 0.0;
    if (P.jl_enabled && G.jl_map >= 0)
    {
        const double lo_t = P.jl_lo_tau[n], hi_t = P.jl_hi_tau[n];
        double v = 0.0;
        if (xn < lo_t) v = xn - P.jl_lo[n] - P.jl_tau[n];
        if (xn > hi_t) v = xn - P.jl_hi[n] + P.jl_tau[n];
        const double yd = v - goal[P.jl_phi_off + n];
        c += rho_jl * yd * yd;
        if (!(xn >= lo_t && xn <= hi_t)) s += rho_jl * yd;
    }
    if (G.W && t > 0)
    {
        const double w = G.W[n];
        const double d0 = xn - *reinterpret_cast<const double*>(xe + (int64_t)n * P.x_sk * 8 - P.x_se_b);
        c += w * d0 * d0;
        s += w * d0;
        if (t + 1 < G.T) s -= w * (*reinterpret_cast<const double*>(xe + (int64_t)n * P.x_sk * 8 + P.x_se_b) - xn);
    }
    return s;
}

template <int NJ, int ALLREV, int J>
struct ChainGradStore
{
    static __device__ __forceinline__ void run(const ChainParams& P, const ChainGradParams& G, const char* xe, int t, double rho_jl, const double* goal,
                                               const double (&g)[NJ], double* ge, double& c)
    {
        const int n = P.joint[J].col;
        if (n >= 0) ge[(int64_t)n * G.g_sk] = 2.0 * G.ct * (g[J] + costgrad_column_extra(P, G, xe, t, n, rho_jl, goal, c));
        ChainGradStore<NJ, ALLREV, J + 1>::run(P, G, xe, t, rho_jl, goal, g, ge, c);
    }
};
template <int NJ, int ALLREV>
struct ChainGradStore<NJ, ALLREV, NJ>
{
    static __device__ __forceinline__ void run(const ChainParams&, const ChainGradParams&, const char*, int, double, const double*, const double (&)[NJ],
                                               double*, double&) {}
};

#ifndef EXO_COSTGRAD_MAXNREG
#define EXO_COSTGRAD_MAXNREG 168
#endif
template <int NJ, int ALLREV>
__global__ void __maxnreg__(EXO_COSTGRAD_MAXNREG) exo_chain_costgrad_kernel(const __grid_constant__ ChainParams P, const __grid_constant__ ChainGradParams G)
{
    const int64_t stride = (int64_t)gridDim.x * CTPB;
    for (int64_t e = (int64_t)blockIdx.x * CTPB + threadIdx.x; e < P.n_evals; e += stride)
    {
        const char* xe = reinterpret_cast<const char*>(P.x) + e * P.x_se_b;
        double* ge = G.grad + e * G.g_se;
        const int t = (int)(e % G.T);
        if (t < G.t_first)
        {
            for (int n = 0; n < P.n_ctrl; ++n) ge[(int64_t)n * G.g_sk] = 0.0;
            if (G.cost_eval) G.cost_eval[e] = 0.0;
            continue;
        }
        double q[NJ], sn[NJ], cs[NJ];
#pragma unroll
        for (int j = 0; j < NJ; ++j) q[j] = *reinterpret_cast<const double*>(xe + P.joint[j].x_byte_off);
        // as in exo_chain_kernel: pull the joint values of a CTA most of a resident generation ahead into L2
        if (e + P.prefetch_ahead < P.n_evals)
        {
#pragma unroll
            for (int j = 0; j < NJ; ++j) prefetch_l2(xe + P.prefetch_ahead * P.x_se_b + P.joint[j].x_byte_off);
        }
#pragma unroll
        for (int j = 0; j < NJ; ++j)
        {
            const double th = fma(P.joint[j].qmult, q[j], P.joint[j].th0);
            sn[j] = 0.0;
            cs[j] = 1.0;
            if (ALLREV || !P.joint[j].prismatic)
            {
                if (fabs(th) <= 1.0e5)
                    fast_sincos(th, sn[j], cs[j]);
                else
                    big_sincos(P.joint[j].qmult * q[j], P.joint[j].c0, P.joint[j].s0, sn[j], cs[j]);
            }
        }
        Fr cur = load_frame(P.base);
        double W[NJ][3], PJ[NJ][3];
        ChainUnroll<NJ, ALLREV, 0>::record(P, cur, q, sn, cs, W, PJ);
        const double* rho = G.rho + (int64_t)t * G.n_maps;
        const int tg = G.T_goal > 1 ? t : 0;
        const double* goal = G.goal + (int64_t)tg * G.length_phi;
        const double* grot = G.goal_rot + (int64_t)tg * G.n_rot * 9;
        double g[NJ];
#pragma unroll
        for (int j = 0; j < NJ; ++j) g[j] = 0.0;
        double c = 0.0;
        for (int f = 0; f < P.n_frames; ++f)
        {
            const ChainFrameP& fp = P.frame[f];
            const Fr A = chain_frame_a(fp, cur);
            const Fr Ph = fp.b_const ? fr_mul(load_frame(fp.binv), A) : A;
            const double r = rho[G.frame_map[f]];
            const double* gl = goal + fp.phi_off;
            double yd[6];
            yd[0] = Ph.x - gl[0];
            yd[1] = Ph.y - gl[1];
            yd[2] = Ph.z - gl[2];
            yd[3] = yd[4] = yd[5] = 0.0;
            const bool six = G.frame_rot[f] >= 0;
            if (six)
            {
                // rotvec = M1 * (M2^-1 M1).GetRot()  (TaskSpaceVector::operator-, task_space_vector.cpp:64-91)
                const double* M2 = grot + 9 * G.frame_rot[f];
                Fr M;
                M.m0 = M2[0] * Ph.m0 + M2[3] * Ph.m3 + M2[6] * Ph.m6;
                M.m1 = M2[0] * Ph.m1 + M2[3] * Ph.m4 + M2[6] * Ph.m7;
                M.m2 = M2[0] * Ph.m2 + M2[3] * Ph.m5 + M2[6] * Ph.m8;
                M.m3 = M2[1] * Ph.m0 + M2[4] * Ph.m3 + M2[7] * Ph.m6;
                M.m4 = M2[1] * Ph.m1 + M2[4] * Ph.m4 + M2[7] * Ph.m7;
                M.m5 = M2[1] * Ph.m2 + M2[4] * Ph.m5 + M2[7] * Ph.m8;
                M.m6 = M2[2] * Ph.m0 + M2[5] * Ph.m3 + M2[8] * Ph.m6;
                M.m7 = M2[2] * Ph.m1 + M2[5] * Ph.m4 + M2[8] * Ph.m7;
                M.m8 = M2[2] * Ph.m2 + M2[5] * Ph.m5 + M2[8] * Ph.m8;
                M.x = M.y = M.z = 0.0;
                double vx, vy, vz;
                get_angle_axis(M, vx, vy, vz);
                yd[3] = Ph.m0 * vx + Ph.m1 * vy + Ph.m2 * vz;
                yd[4] = Ph.m3 * vx + Ph.m4 * vy + Ph.m5 * vz;
                yd[5] = Ph.m6 * vx + Ph.m7 * vy + Ph.m8 * vz;
            }
            c += r * (yd[0] * yd[0] + yd[1] * yd[1] + yd[2] * yd[2] + yd[3] * yd[3] + yd[4] * yd[4] + yd[5] * yd[5]);
#pragma unroll
            for (int k = 0; k < 6; ++k) yd[k] *= r;  // S ydiff
            if (fp.b_const)
            {
                // column = Binv [v; w]  ->  dot against Binv^T (S ydiff)
                const double a0 = fp.binv[0] * yd[0] + fp.binv[3] * yd[1] + fp.binv[6] * yd[2];
                const double a1 = fp.binv[1] * yd[0] + fp.binv[4] * yd[1] + fp.binv[7] * yd[2];
                const double a2 = fp.binv[2] * yd[0] + fp.binv[5] * yd[1] + fp.binv[8] * yd[2];
                const double b0 = fp.binv[0] * yd[3] + fp.binv[3] * yd[4] + fp.binv[6] * yd[5];
                const double b1 = fp.binv[1] * yd[3] + fp.binv[4] * yd[4] + fp.binv[7] * yd[5];
                const double b2 = fp.binv[2] * yd[3] + fp.binv[5] * yd[4] + fp.binv[8] * yd[5];
                yd[0] = a0; yd[1] = a1; yd[2] = a2; yd[3] = b0; yd[4] = b1; yd[5] = b2;
            }
            ChainDots<NJ, ALLREV, 0>::run(P, W, PJ, A.x, A.y, A.z, yd, six, g);
        }
        const double rho_jl = G.jl_map >= 0 ? rho[G.jl_map] : 0.0;
        ChainGradStore<NJ, ALLREV, 0>::run(P, G, xe, t, rho_jl, goal, g, ge, c);
        // columns no joint of the chain owns: JointLimit and transition terms only
        for (int n = 0; n < P.n_ctrl; ++n)
            if ((P.zero_cols >> n) & 1u) ge[(int64_t)n * G.g_sk] = 2.0 * G.ct * costgrad_column_extra(P, G, xe, t, n, rho_jl, goal, c);
        if (G.cost_eval) G.cost_eval[e] = G.ct * c;
    }
}

// ct * (ydiff^T S ydiff + transition term) of the state at xe, formed in registers (the cost half of the kernel above)
template <int NJ, int ALLREV>
__device__ __forceinline__ double chain_cost_only(const ChainParams& P, const ChainGradParams& G, const char* xe, int t)
{
    double q[NJ], sn[NJ], cs[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) q[j] = *reinterpret_cast<const double*>(xe + P.joint[j].x_byte_off);
#pragma unroll
    for (int j = 0; j < NJ; ++j)
    {
        const double th = fma(P.joint[j].qmult, q[j], P.joint[j].th0);
        sn[j] = 0.0;
        cs[j] = 1.0;
        if (ALLREV || !P.joint[j].prismatic)
        {
            if (fabs(th) <= 1.0e5)
                fast_sincos(th, sn[j], cs[j]);
            else
                big_sincos(P.joint[j].qmult * q[j], P.joint[j].c0, P.joint[j].s0, sn[j], cs[j]);
        }
    }
    Fr cur = load_frame(P.base);
    ChainUnroll<NJ, ALLREV, 0>::compose(P, cur, q, sn, cs);
    const double* rho = G.rho + (int64_t)t * G.n_maps;
    const int tg = G.T_goal > 1 ? t : 0;
    const double* goal = G.goal + (int64_t)tg * G.length_phi;
    const double* grot = G.goal_rot + (int64_t)tg * G.n_rot * 9;
    double c = 0.0;
    for (int f = 0; f < P.n_frames; ++f)
    {
        const ChainFrameP& fp = P.frame[f];
        const Fr A = chain_frame_a(fp, cur);
        const Fr Ph = fp.b_const ? fr_mul(load_frame(fp.binv), A) : A;
        const double* gl = goal + fp.phi_off;
        const double d0 = Ph.x - gl[0], d1 = Ph.y - gl[1], d2 = Ph.z - gl[2];
        double s = d0 * d0 + d1 * d1 + d2 * d2;
        if (G.frame_rot[f] >= 0)
        {
            const double* M2 = grot + 9 * G.frame_rot[f];
            Fr M;
            M.m0 = M2[0] * Ph.m0 + M2[3] * Ph.m3 + M2[6] * Ph.m6;
            M.m1 = M2[0] * Ph.m1 + M2[3] * Ph.m4 + M2[6] * Ph.m7;
            M.m2 = M2[0] * Ph.m2 + M2[3] * Ph.m5 + M2[6] * Ph.m8;
            M.m3 = M2[1] * Ph.m0 + M2[4] * Ph.m3 + M2[7] * Ph.m6;
            M.m4 = M2[1] * Ph.m1 + M2[4] * Ph.m4 + M2[7] * Ph.m7;
            M.m5 = M2[1] * Ph.m2 + M2[4] * Ph.m5 + M2[7] * Ph.m8;
            M.m6 = M2[2] * Ph.m0 + M2[5] * Ph.m3 + M2[8] * Ph.m6;
            M.m7 = M2[2] * Ph.m1 + M2[5] * Ph.m4 + M2[8] * Ph.m7;
            M.m8 = M2[2] * Ph.m2 + M2[5] * Ph.m5 + M2[8] * Ph.m8;
            M.x = M.y = M.z = 0.0;
            double vx, vy, vz;
            get_angle_axis(M, vx, vy, vz);
            const double r0 = Ph.m0 * vx + Ph.m1 * vy + Ph.m2 * vz;
            const double r1 = Ph.m3 * vx + Ph.m4 * vy + Ph.m5 * vz;
            const double r2 = Ph.m6 * vx + Ph.m7 * vy + Ph.m8 * vz;
            s += r0 * r0 + r1 * r1 + r2 * r2;
        }
        c += rho[G.frame_map[f]] * s;
    }
    const double rho_jl = G.jl_map >= 0 ? rho[G.jl_map] : 0.0;
    for (int n = 0; n < P.n_ctrl; ++n) costgrad_column_extra(P, G, xe, t, n, rho_jl, goal, c);
    return G.ct * c;
}

// Task cost only (GetScalarTaskCost + GetScalarTransitionCost per timestep / GetScalarCost): forward kinematics and cost in
// registers, one double per evaluation leaves the SM
template <int NJ, int ALLREV>
__global__ void __maxnreg__(128) exo_chain_cost_kernel(const __grid_constant__ ChainParams P, const __grid_constant__ ChainGradParams G)
{
    const int64_t e = (int64_t)blockIdx.x * CTPB + threadIdx.x;
    if (e >= P.n_evals) return;
    const int t = (int)(e % G.T);
    const char* xe = reinterpret_cast<const char*>(P.x) + e * P.x_se_b;
    G.cost_eval[e] = t < G.t_first ? 0.0 : chain_cost_only<NJ, ALLREV>(P, G, xe, t);
}

template <int NJ, int ALLREV>
cudaError_t chain_cost_launch_t(const ChainParams& P, const ChainGradParams& G, cudaStream_t stream)
{
    const int64_t blocks = (P.n_evals + CTPB - 1) / CTPB;
    exo_chain_cost_kernel<NJ, ALLREV><<<(int)(blocks > 0 ? blocks : 1), CTPB, 0, stream>>>(P, G);
    return cudaGetLastError();
}

template <int ALLREV>
cudaError_t chain_cost_launch_nj(const ChainParams& P, const ChainGradParams& G, cudaStream_t stream)
{
    switch (P.n_joints)
    {
        case 1: return chain_cost_launch_t<1, ALLREV>(P, G, stream);
        case 2: return chain_cost_launch_t<2, ALLREV>(P, G, stream);
        case 3: return chain_cost_launch_t<3, ALLREV>(P, G, stream);
        case 4: return chain_cost_launch_t<4, ALLREV>(P, G, stream);
        case 5: return chain_cost_launch_t<5, ALLREV>(P, G, stream);
        case 6: return chain_cost_launch_t<6, ALLREV>(P, G, stream);
        case 7: return chain_cost_launch_t<7, ALLREV>(P, G, stream);
        case 8: return chain_cost_launch_t<8, ALLREV>(P, G, stream);
        default: return cudaErrorInvalidConfiguration;
    }
}

// IKSolver::Solve's backtracking line search (ik_solver.cpp:146-161) for every active start state: the first of the ten steps
// alpha = 1, 0.9, ..., 0.1 whose cost is below the current one is taken (q <- q - alpha qd, qd <- alpha qd).
template <int NJ, int ALLREV>
__global__ void __maxnreg__(128) exo_chain_ik_linesearch_kernel(const __grid_constant__ ChainParams P, const __grid_constant__ ChainGradParams G,
                                                                const __grid_constant__ ChainLineSearchParams L)
{
    const int64_t e = (int64_t)blockIdx.x * CTPB + threadIdx.x;
    if (e >= P.n_evals || !L.active[e] || L.ls_done[e]) return;
    const int N = P.n_ctrl;
    const int64_t E = P.n_evals;
    const char* xe = reinterpret_cast<const char*>(L.q_trial) + e * P.x_se_b;
    const double err_prev = L.err_prev[e];
    for (int ai = 0; ai < 10; ++ai)
    {
        const double alpha = ik_alpha(ai);  // Eigen::VectorXd::LinSpaced(10, 1.0, 0.1)
        for (int c = 0; c < N; ++c) L.q_trial[(int64_t)c * E + e] = L.q[(int64_t)c * E + e] - alpha * L.qd[(int64_t)c * E + e];
        const double error = chain_cost_only<NJ, ALLREV>(P, G, xe, 0);
        L.cost[e] = error;  // error_: the cost at the last state evaluated
        L.steplen[e] = alpha;
        if (error < err_prev)
        {
            for (int c = 0; c < N; ++c)
            {
                L.q[(int64_t)c * E + e] = L.q_trial[(int64_t)c * E + e];
                L.qd[(int64_t)c * E + e] *= alpha;
            }
            L.ls_done[e] = 1;
            break;
        }
    }
}

// One evaluation (Phi and J of every map of the program) with the generic row emitter: what exo_chain_kernel computes, without
// its streaming machinery -- used inside the single-launch IK solve, where Phi and J go to an L2-resident workspace
template <int NJ, int ALLREV>
__device__ __forceinline__ void chain_eval_one(const ChainParams& P, const char* xe, char* phi_e, char* jac_e)
{
    if (P.has_prefix)
    {
        Sink o;
        o.phi = reinterpret_cast<double*>(phi_e);
        o.jac = reinterpret_cast<double*>(jac_e);
        o.phi_sk = P.phi_sk;
        o.j_sr = P.j_sr;
        o.j_sc = P.j_sc;
        chain_prefix(P, o, reinterpret_cast<const double*>(xe), P.x_sk);
    }
    double q[NJ], sn[NJ], cs[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) q[j] = *reinterpret_cast<const double*>(xe + P.joint[j].x_byte_off);
#pragma unroll
    for (int j = 0; j < NJ; ++j)
    {
        const double th = fma(P.joint[j].qmult, q[j], P.joint[j].th0);
        sn[j] = 0.0;
        cs[j] = 1.0;
        if (ALLREV || !P.joint[j].prismatic)
        {
            if (fabs(th) <= 1.0e5)
                fast_sincos(th, sn[j], cs[j]);
            else
                big_sincos(P.joint[j].qmult * q[j], P.joint[j].c0, P.joint[j].s0, sn[j], cs[j]);
        }
    }
    Fr cur = load_frame(P.base);
    double W[NJ][3], PJ[NJ][3];
    ChainUnroll<NJ, ALLREV, 0>::record(P, cur, q, sn, cs, W, PJ);
    for (int f = 0; f < P.n_frames; ++f)
    {
        const ChainFrameP& fp = P.frame[f];
        const Fr A = chain_frame_a(fp, cur);
        if (jac_e)
        {
            double ux = 0.0, uy = 0.0, uz = 0.0;
            if (fp.row_mode == ROWS_DIST)
            {
                const Fr Ph = fp.b_const ? fr_mul(load_frame(fp.binv), A) : A;
                const double nrm = norm3_scaled(Ph.x, Ph.y, Ph.z);
                ux = Ph.x / nrm; uy = Ph.y / nrm; uz = Ph.z / nrm;
            }
            ChainUnroll<NJ, ALLREV, 0>::columns_g(P, f, jac_e, P.j_sr_b, W, PJ, A.x, A.y, A.z, ux, uy, uz);
        }
        chain_frame_phi(fp, A, phi_e, P.phi_sk);
    }
}

// IKSolver::Solve (ik_solver.cpp:73-240) of one start state per lane, start to finish: prob_->Update(q), the damped
// pseudo-inverse step, the backtracking line search and the termination tests, iterated in registers / on the lane's own column
// of the SoA workspace.  Lanes that stop wait for the slowest lane of their warp; nothing waits for the host.
template <int NJ, int ALLREV>
__global__ void __maxnreg__(168) exo_chain_ik_solve_kernel(const __grid_constant__ ChainParams P, const __grid_constant__ ChainGradParams G,
                                                           const __grid_constant__ ChainLineSearchParams L, const __grid_constant__ IkParams K,
                                                           const __grid_constant__ CostParams C, int max_iterations)
{
    const int64_t e = (int64_t)blockIdx.x * CTPB + threadIdx.x;
    if (e >= P.n_evals) return;
    const int N = P.n_ctrl;
    const int64_t E = P.n_evals;
    const char* xe = reinterpret_cast<const char*>(P.x) + e * P.x_se_b;
    const char* xt = reinterpret_cast<const char*>(L.q_trial) + e * P.x_se_b;
    char* phi_e = reinterpret_cast<char*>(P.phi) + e * P.phi_se_b;
    char* jac_e = reinterpret_cast<char*>(P.jac) + e * P.j_se_b;
    for (int it = 0; it < max_iterations; ++it)
    {
        chain_eval_one<NJ, ALLREV>(P, xe, phi_e, jac_e);  // prob_->Update(q_)
        if (!ik_step_one<0>(K, C, e)) break;
        const double err_prev = L.err_prev[e];
        for (int ai = 0; ai < 10 && !L.ls_done[e]; ++ai)  // (the MaxStep mode has taken its step already)
        {
            const double alpha = ik_alpha(ai);  // Eigen::VectorXd::LinSpaced(10, 1.0, 0.1)
            for (int c = 0; c < N; ++c) L.q_trial[(int64_t)c * E + e] = L.q[(int64_t)c * E + e] - alpha * L.qd[(int64_t)c * E + e];
            const double error = chain_cost_only<NJ, ALLREV>(P, G, xt, 0);
            L.cost[e] = error;
            L.steplen[e] = alpha;
            if (error < err_prev)
            {
                for (int c = 0; c < N; ++c)
                {
                    L.q[(int64_t)c * E + e] = L.q_trial[(int64_t)c * E + e];
                    L.qd[(int64_t)c * E + e] *= alpha;
                }
                L.ls_done[e] = 1;
                break;
            }
        }
        if (!ik_finish_one(K, e, it + 1 == max_iterations)) break;
    }
}

template <int NJ, int ALLREV>
cudaError_t chain_ik_solve_launch_t(const ChainParams& P, const ChainGradParams& G, const ChainLineSearchParams& L, const IkParams& K,
                                    const CostParams& C, int max_iterations, cudaStream_t stream)
{
    const int64_t blocks = (P.n_evals + CTPB - 1) / CTPB;
    exo_chain_ik_solve_kernel<NJ, ALLREV><<<(int)(blocks > 0 ? blocks : 1), CTPB, 0, stream>>>(P, G, L, K, C, max_iterations);
    return cudaGetLastError();
}

template <int ALLREV>
cudaError_t chain_ik_solve_launch_nj(const ChainParams& P, const ChainGradParams& G, const ChainLineSearchParams& L, const IkParams& K,
                                     const CostParams& C, int max_iterations, cudaStream_t stream)
{
    switch (P.n_joints)
    {
        case 1: return chain_ik_solve_launch_t<1, ALLREV>(P, G, L, K, C, max_iterations, stream);
        case 2: return chain_ik_solve_launch_t<2, ALLREV>(P, G, L, K, C, max_iterations, stream);
        case 3: return chain_ik_solve_launch_t<3, ALLREV>(P, G, L, K, C, max_iterations, stream);
        case 4: return chain_ik_solve_launch_t<4, ALLREV>(P, G, L, K, C, max_iterations, stream);
        case 5: return chain_ik_solve_launch_t<5, ALLREV>(P, G, L, K, C, max_iterations, stream);
        case 6: return chain_ik_solve_launch_t<6, ALLREV>(P, G, L, K, C, max_iterations, stream);
        case 7: return chain_ik_solve_launch_t<7, ALLREV>(P, G, L, K, C, max_iterations, stream);
        case 8: return chain_ik_solve_launch_t<8, ALLREV>(P, G, L, K, C, max_iterations, stream);
        default: return cudaErrorInvalidConfiguration;
    }
}

template <int NJ, int ALLREV>
cudaError_t chain_ik_linesearch_launch_t(const ChainParams& P, const ChainGradParams& G, const ChainLineSearchParams& L, cudaStream_t stream)
{
    const int64_t blocks = (P.n_evals + CTPB - 1) / CTPB;
    exo_chain_ik_linesearch_kernel<NJ, ALLREV><<<(int)(blocks > 0 ? blocks : 1), CTPB, 0, stream>>>(P, G, L);
    return cudaGetLastError();
}

template <int ALLREV>
cudaError_t chain_ik_linesearch_launch_nj(const ChainParams& P, const ChainGradParams& G, const ChainLineSearchParams& L, cudaStream_t stream)
{
    switch (P.n_joints)
    {
        case 1: return chain_ik_linesearch_launch_t<1, ALLREV>(P, G, L, stream);
        case 2: return chain_ik_linesearch_launch_t<2, ALLREV>(P, G, L, stream);
        case 3: return chain_ik_linesearch_launch_t<3, ALLREV>(P, G, L, stream);
        case 4: return chain_ik_linesearch_launch_t<4, ALLREV>(P, G, L, stream);
        case 5: return chain_ik_linesearch_launch_t<5, ALLREV>(P, G, L, stream);
        case 6: return chain_ik_linesearch_launch_t<6, ALLREV>(P, G, L, stream);
        case 7: return chain_ik_linesearch_launch_t<7, ALLREV>(P, G, L, stream);
        case 8: return chain_ik_linesearch_launch_t<8, ALLREV>(P, G, L, stream);
        default: return cudaErrorInvalidConfiguration;
    }
}

static int chain_prefetch_generations();

template <int NJ, int ALLREV>
cudaError_t chain_costgrad_launch_t(const ChainParams& P, const ChainGradParams& G, cudaStream_t stream)
{
    static int resident = 0;  // threads resident on the device for this instantiation: query once
    if (resident == 0)
    {
        int dev = 0, sms = 0, ctas = 0;
        cudaError_t err = cudaGetDevice(&dev);
        if (err == cudaSuccess) err = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (err == cudaSuccess) err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, exo_chain_costgrad_kernel<NJ, ALLREV>, CTPB, 0);
        if (err != cudaSuccess) return err;
        resident = sms * (ctas > 0 ? ctas : 1) * CTPB;
    }
    const int64_t want = (P.n_evals + CTPB - 1) / CTPB;
    ChainParams Q = P;
    Q.prefetch_ahead = (int64_t)resident * chain_prefetch_generations() / 4;
    exo_chain_costgrad_kernel<NJ, ALLREV><<<(int)(want < (1 << 20) ? (want > 0 ? want : 1) : (1 << 20)), CTPB, 0, stream>>>(Q, G);
    return cudaGetLastError();
}

template <int ALLREV>
cudaError_t chain_costgrad_launch_nj(const ChainParams& P, const ChainGradParams& G, cudaStream_t stream)
{
    switch (P.n_joints)
    {
        case 1: return chain_costgrad_launch_t<1, ALLREV>(P, G, stream);
        case 2: return chain_costgrad_launch_t<2, ALLREV>(P, G, stream);
        case 3: return chain_costgrad_launch_t<3, ALLREV>(P, G, stream);
        case 4: return chain_costgrad_launch_t<4, ALLREV>(P, G, stream);
        case 5: return chain_costgrad_launch_t<5, ALLREV>(P, G, stream);
        case 6: return chain_costgrad_launch_t<6, ALLREV>(P, G, stream);
        case 7: return chain_costgrad_launch_t<7, ALLREV>(P, G, stream);
        case 8: return chain_costgrad_launch_t<8, ALLREV>(P, G, stream);
        default: return cudaErrorInvalidConfiguration;
    }
}

// Grid size in units of "resident CTAs per GPU" (tuning knob EXO_CHAIN_WAVES).  Measured on B200 (cfg2a, 2^20 evaluations):
// 1 generation (persistent, static grid-stride split) 98.6 us, 4 generations 90.1 us, 8 generations 88.2 us, and one CTA per
// 64 evaluations -- the default -- 84.8 us once every CTA pulls the joint values of the CTA 3/4 of a generation behind it
// into L2 (EXO_CHAIN_PF_QGENS, in quarter generations: 1 -> 86.1, 2 -> 85.9, 3 -> 84.8, 4 -> 85.4, 6 -> 87.9 us).  The
// hardware CTA scheduler balances the two dies better than any static split, and tiles that advance together keep the
// DRAM pages of the 48 output rows hot (tools/hbm_probe.cu); a software tile counter (atomicAdd) was slower (94-105 us).
static int chain_grid_waves()
{
    static int waves = 0;
    if (waves == 0)
    {
        const char* env = getenv("EXO_CHAIN_WAVES");
        waves = env ? atoi(env) : 1 << 16;
        if (waves < 1) waves = 1;
    }
    return waves;
}

static int chain_prefetch_generations()
{
    static int gens = 0;
    if (gens == 0)
    {
        const char* env = getenv("EXO_CHAIN_PF_QGENS");  // in quarters of a generation of resident CTAs
        gens = env ? atoi(env) : 3;
        if (gens < 1) gens = 1;
    }
    return gens;
}

template <int NJ, int ALLREV, int VARIANT, bool GEN>
cudaError_t chain_launch_t(const ChainParams& P, int sms, cudaStream_t stream)
{
    static int ctas_per_sm = 0;  // occupancy is a property of the instantiation: query once
    if (ctas_per_sm == 0)
    {
        int ctas = 0;
        cudaError_t err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, exo_chain_kernel<NJ, ALLREV, VARIANT, GEN>, CTPB, 0);
        if (err != cudaSuccess) return err;
        if (ctas < 1) return cudaErrorInvalidConfiguration;
        ctas_per_sm = ctas;
    }
    const int64_t want = (P.n_evals + CTPB - 1) / CTPB;
    const int64_t resident = (int64_t)sms * ctas_per_sm * chain_grid_waves();
    const int blocks = (int)(want < resident ? (want > 0 ? want : 1) : resident);
    ChainParams Q = P;
    Q.prefetch_ahead = (int64_t)sms * ctas_per_sm * CTPB * chain_prefetch_generations() / 4;
    exo_chain_kernel<NJ, ALLREV, VARIANT, GEN><<<blocks, CTPB, 0, stream>>>(Q);
    return cudaGetLastError();
}

template <int ALLREV, int VARIANT, bool GEN>
cudaError_t chain_launch_nj(const ChainParams& P, int sms, cudaStream_t stream)
{
    switch (P.n_joints)
    {
        case 1: return chain_launch_t<1, ALLREV, VARIANT, GEN>(P, sms, stream);
        case 2: return chain_launch_t<2, ALLREV, VARIANT, GEN>(P, sms, stream);
        case 3: return chain_launch_t<3, ALLREV, VARIANT, GEN>(P, sms, stream);
        case 4: return chain_launch_t<4, ALLREV, VARIANT, GEN>(P, sms, stream);
        case 5: return chain_launch_t<5, ALLREV, VARIANT, GEN>(P, sms, stream);
        case 6: return chain_launch_t<6, ALLREV, VARIANT, GEN>(P, sms, stream);
        case 7: return chain_launch_t<7, ALLREV, VARIANT, GEN>(P, sms, stream);
        case 8: return chain_launch_t<8, ALLREV, VARIANT, GEN>(P, sms, stream);
        default: return cudaErrorInvalidConfiguration;
    }
}

template <int VARIANT>
cudaError_t chain_launch_v(const ChainParams& P, int sms, cudaStream_t stream)
{
    bool gen = false;
    for (int f = 0; f < P.n_frames; ++f) gen = gen || P.frame[f].row_mode > ROWS_V;
    if (gen)  // the generic row emitter exists in the one-pass variant only
        return P.all_revolute ? chain_launch_nj<true, CHAIN_REGS, true>(P, sms, stream) : chain_launch_nj<false, CHAIN_REGS, true>(P, sms, stream);
    if (P.all_revolute && P.ortho && VARIANT == CHAIN_REGS)
    {
        bool position_only = !P.jl_enabled;
        for (int f = 0; f < P.n_frames; ++f) position_only = position_only && P.frame[f].row_mode == ROWS_V;
        return position_only ? chain_launch_nj<7, CHAIN_REGS, false>(P, sms, stream) : chain_launch_nj<3, CHAIN_REGS, false>(P, sms, stream);
    }
    return P.all_revolute ? chain_launch_nj<true, VARIANT, false>(P, sms, stream) : chain_launch_nj<false, VARIANT, false>(P, sms, stream);
}
}  // namespace

void chain_bind(ChainParams& P, const double* x, double* phi, double* jac, int64_t n_evals, int64_t x_se, int64_t x_sk, int64_t phi_se,
                int64_t phi_sk, int64_t j_se, int64_t j_sr, int64_t j_sc)
{
    P.x = x; P.phi = phi; P.jac = jac; P.n_evals = n_evals;
    P.x_sk = x_sk; P.phi_sk = phi_sk; P.j_sr = j_sr; P.j_sc = j_sc;
    P.x_se_b = x_se * 8; P.phi_se_b = phi_se * 8; P.j_se_b = j_se * 8; P.j_sr_b = j_sr * 8;
    for (int j = 0; j < P.n_joints; ++j) P.joint[j].x_byte_off = (P.joint[j].qidx >= 0 ? P.joint[j].qidx : 0) * x_sk * 8;
    for (int f = 0; f < P.n_frames; ++f)
    {
        P.frame[f].phi_byte_off = P.frame[f].phi_off * phi_sk * 8;
        for (int j = 0; j < P.n_joints; ++j)
            P.jf_byte_off[f][j] = (P.frame[f].jrow_off * j_sr + (P.joint[j].col >= 0 ? P.joint[j].col : 0) * j_sc) * 8;
    }
}

cudaError_t chain_costgrad_launch(const ChainParams& P, const ChainGradParams& G, cudaStream_t stream)
{
    return P.all_revolute ? chain_costgrad_launch_nj<true>(P, G, stream) : chain_costgrad_launch_nj<false>(P, G, stream);
}

cudaError_t chain_cost_launch(const ChainParams& P, const ChainGradParams& G, cudaStream_t stream)
{
    return P.all_revolute ? chain_cost_launch_nj<true>(P, G, stream) : chain_cost_launch_nj<false>(P, G, stream);
}

cudaError_t chain_ik_linesearch_launch(const ChainParams& P, const ChainGradParams& G, const ChainLineSearchParams& L, cudaStream_t stream)
{
    return P.all_revolute ? chain_ik_linesearch_launch_nj<true>(P, G, L, stream) : chain_ik_linesearch_launch_nj<false>(P, G, L, stream);
}

cudaError_t chain_ik_solve_launch(const ChainParams& P, const ChainGradParams& G, const ChainLineSearchParams& L, const IkParams& K,
                                  const CostParams& C, int max_iterations, cudaStream_t stream)
{
    return P.all_revolute ? chain_ik_solve_launch_nj<true>(P, G, L, K, C, max_iterations, stream)
                          : chain_ik_solve_launch_nj<false>(P, G, L, K, C, max_iterations, stream);
}

cudaError_t chain_launch(const ChainParams& P, int variant, cudaStream_t stream)
{
    static int sms = 0;
    if (sms == 0)
    {
        int dev = 0;
        cudaError_t err = cudaGetDevice(&dev);
        if (err != cudaSuccess) return err;
        err = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (err != cudaSuccess) return err;
    }
    switch (variant)
    {
        case CHAIN_REGS: return chain_launch_v<CHAIN_REGS>(P, sms, stream);
        case CHAIN_TWO_PASS: return chain_launch_v<CHAIN_TWO_PASS>(P, sms, stream);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace exo
