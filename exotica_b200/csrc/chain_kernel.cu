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
// One lane = one evaluation; the kernel is persistent (grid = SMs x resident CTAs) and prefetches the next evaluation's
// joint values while it works on the current one, so the only exposed memory operations are coalesced 8-byte stores of
// Phi and J.  sin/cos use a Cody-Waite reduction and fdlibm's minimax kernels with their coefficients in __constant__
// memory (CUDA's sincos spends most of its instructions materialising 64-bit immediates and on int<->double
// conversions).  HBM-bound FP64 streaming work: no tensor cores.
//
// Variants (exo_problem_set_kernel; measured in DESIGN.md):
//   CHAIN_REGS      every joint's world axis and origin stay in registers until the end-effector point is known
//   CHAIN_STASH     they are parked in shared memory, [slot][thread] as double2 (conflict-free): fewer registers
//   CHAIN_TWO_PASS  the chain is composed twice; the second pass emits each joint's column on the fly
//
// Reference path replaced: KinematicTree::Update/UpdateTree/FK/ComputeJ (exotica_core/src/kinematic_tree.cpp:678-737,
// 829-867,919-945) + EffFrame/EffPosition/JointLimit::Update (exotations/exotica_core_task_maps/src/*.cpp) +
// Phi.SetZero/jacobian.setZero (exotica_core/src/problems/abstract_time_indexed_problem.cpp:243-244).
#include <string.h>

#include "compiler.h"
#include "device_math.cuh"
#include "kernels.cuh"

namespace exo
{
namespace
{
using namespace dev;

constexpr int CTPB = 128;

struct Sink
{
    double* phi;  // &Phi[e][0]
    double* jac;  // &J[e][0][0] or nullptr
    int64_t phi_sk, j_sr, j_sc;
};

// Rx(alpha_J) * Tx(a_J) applied on the right of cur (joint 0: already part of the base frame)
template <int J>
__device__ __forceinline__ void dh_link(const ChainParams& P, Fr& cur)
{
    if (J == 0) return;
    const double ca = P.joint[J].ca, sa = P.joint[J].sa, a = P.joint[J].a;
    cur.x = fma(a, cur.m0, cur.x);
    cur.y = fma(a, cur.m3, cur.y);
    cur.z = fma(a, cur.m6, cur.z);
    const double y0 = cur.m1, y1 = cur.m4, y2 = cur.m7;
    cur.m1 = fma(ca, y0, sa * cur.m2); cur.m2 = fma(ca, cur.m2, -(sa * y0));
    cur.m4 = fma(ca, y1, sa * cur.m5); cur.m5 = fma(ca, cur.m5, -(sa * y1));
    cur.m7 = fma(ca, y2, sa * cur.m8); cur.m8 = fma(ca, cur.m8, -(sa * y2));
}

// Rz(theta) * Tz(d) applied on the right of cur; xq is the state entry feeding the joint
template <int J, bool ALLREV>
__device__ __forceinline__ void dh_joint(const ChainParams& P, Fr& cur, double sn, double cs, double xq)
{
    const ChainJointP& jp = P.joint[J];
    double d = jp.d;
    if (!ALLREV && jp.prismatic)
    {
        sn = jp.s0;
        cs = jp.c0;
        d = fma(jp.qmult, xq, d);
    }
    const double a0 = cur.m0, a3 = cur.m3, a6 = cur.m6;
    cur.m0 = fma(cs, a0, sn * cur.m1); cur.m1 = fma(cs, cur.m1, -(sn * a0));
    cur.m3 = fma(cs, a3, sn * cur.m4); cur.m4 = fma(cs, cur.m4, -(sn * a3));
    cur.m6 = fma(cs, a6, sn * cur.m7); cur.m7 = fma(cs, cur.m7, -(sn * a6));
    cur.x = fma(d, cur.m2, cur.x);
    cur.y = fma(d, cur.m5, cur.y);
    cur.z = fma(d, cur.m8, cur.z);
}

// Jacobian column of joint J for every requested frame (KinematicTree::ComputeJ, kinematic_tree.cpp:919-945):
//   column = temp_B.M^-1 * [ w x (temp_A.p - p_joint) ; w ]   (revolute)   |   temp_B.M^-1 * [ w ; 0 ]   (prismatic)
template <int J, bool ALLREV, int NF>
__device__ __forceinline__ void chain_column(const ChainParams& P, const Sink& o, double wx, double wy, double wz, double px, double py,
                                             double pz, const double (&pa)[NF][3])
{
    const ChainJointP& jp = P.joint[J];
    if (jp.col < 0) return;
    const bool prismatic = !ALLREV && jp.prismatic;
#pragma unroll
    for (int f = 0; f < NF; ++f)
    {
        if (NF > 1 && f >= P.n_frames) break;
        const ChainFrameP& fp = P.frame[f];
        double v0, v1, v2, w0, w1, w2;
        if (prismatic)
        {
            v0 = wx; v1 = wy; v2 = wz;
            w0 = w1 = w2 = 0.0;
        }
        else
        {
            const double dx = pa[f][0] - px, dy = pa[f][1] - py, dz = pa[f][2] - pz;
            v0 = wy * dz - wz * dy;
            v1 = wz * dx - wx * dz;
            v2 = wx * dy - wy * dx;
            w0 = wx; w1 = wy; w2 = wz;
        }
        if (fp.b_const)
        {
            const double t0 = fp.binv[0] * v0 + fp.binv[1] * v1 + fp.binv[2] * v2;
            const double t1 = fp.binv[3] * v0 + fp.binv[4] * v1 + fp.binv[5] * v2;
            const double t2 = fp.binv[6] * v0 + fp.binv[7] * v1 + fp.binv[8] * v2;
            v0 = t0; v1 = t1; v2 = t2;
            const double u0 = fp.binv[0] * w0 + fp.binv[1] * w1 + fp.binv[2] * w2;
            const double u1 = fp.binv[3] * w0 + fp.binv[4] * w1 + fp.binv[5] * w2;
            const double u2 = fp.binv[6] * w0 + fp.binv[7] * w1 + fp.binv[8] * w2;
            w0 = u0; w1 = u1; w2 = u2;
        }
        double* jc = o.jac + (fp.jrow_elem_off + jp.jcol_elem_off);
        jc[0] = v0;
        jc += o.j_sr; jc[0] = v1;
        jc += o.j_sr; jc[0] = v2;
        if (fp.n_rows == 6)
        {
            jc += o.j_sr; jc[0] = w0;
            jc += o.j_sr; jc[0] = w1;
            jc += o.j_sr; jc[0] = w2;
        }
    }
}

// temp_A = D_{n-1} * K_f ; Phi = temp_B^-1 * temp_A ; write Phi ; remember temp_A.p for the Jacobian columns
template <int NF>
__device__ __forceinline__ void chain_frames(const ChainParams& P, const Sink& o, const Fr& tip, double (&pa)[NF][3])
{
#pragma unroll
    for (int f = 0; f < NF; ++f)
    {
        if (NF > 1 && f >= P.n_frames) break;
        const ChainFrameP& fp = P.frame[f];
        Fr A = tip;
        if (fp.k_kind == F_OFFA_TRANS)
            fr_translate(A, fp.k[9], fp.k[10], fp.k[11]);
        else if (fp.k_kind == F_OFFA_FULL)
        {
            Fr c;
            c.m0 = fp.k[0]; c.m1 = fp.k[1]; c.m2 = fp.k[2]; c.m3 = fp.k[3]; c.m4 = fp.k[4]; c.m5 = fp.k[5];
            c.m6 = fp.k[6]; c.m7 = fp.k[7]; c.m8 = fp.k[8]; c.x = fp.k[9]; c.y = fp.k[10]; c.z = fp.k[11];
            A = fr_mul(A, c);
        }
        pa[f][0] = A.x; pa[f][1] = A.y; pa[f][2] = A.z;
        if (fp.b_const)
        {
            Fr b;
            b.m0 = fp.binv[0]; b.m1 = fp.binv[1]; b.m2 = fp.binv[2]; b.m3 = fp.binv[3]; b.m4 = fp.binv[4]; b.m5 = fp.binv[5];
            b.m6 = fp.binv[6]; b.m7 = fp.binv[7]; b.m8 = fp.binv[8]; b.x = fp.binv[9]; b.y = fp.binv[10]; b.z = fp.binv[11];
            A = fr_mul(b, A);
        }
        write_phi(fp.map_kind, fp.rot_type, A, o.phi + fp.phi_elem_off, o.phi_sk);
    }
}

// JointLimit::Update (joint_limit.cpp:63-99), unused-map fills (task_space_vector.cpp:54-62) and the Jacobian entries no
// joint of the chain writes (the reference's jacobian.setZero()).
__device__ __noinline__ void chain_prefix(const ChainParams& P, const Sink& o, const double* xe, int64_t x_sk)
{
    const int N = P.n_ctrl;
    if (P.jl_enabled)
    {
        double* ph = o.phi + (int64_t)P.jl_phi_off * o.phi_sk;
        for (int i = 0; i < N; ++i)
        {
            const double q = xe[i * x_sk];
            const double lo_t = P.jl_lo_tau[i], hi_t = P.jl_hi_tau[i];
            double v = 0.0;
            if (q < lo_t) v = q - P.jl_lo[i] - P.jl_tau[i];
            if (q > hi_t) v = q - P.jl_hi[i] + P.jl_tau[i];
            ph[0] = v;
            ph += o.phi_sk;
            if (o.jac)
            {
                const double d = (q >= lo_t && q <= hi_t) ? 0.0 : 1.0;
                double* row = o.jac + (int64_t)(P.jl_jrow_off + i) * o.j_sr;
                for (int col = 0; col < N; ++col)
                {
                    row[0] = (col == i) ? d : 0.0;
                    row += o.j_sc;
                }
            }
        }
    }
    for (int k = 0; k < P.n_fill; ++k)
    {
        const ChainFillP& fl = P.fill[k];
        double* ph = o.phi + (int64_t)fl.phi_off * o.phi_sk;
        for (int i = 0; i < fl.n_items; ++i)
        {
            ph[0] = fl.value[i];
            ph += o.phi_sk;
        }
        if (o.jac)
            for (int r = 0; r < fl.n_rows; ++r)
            {
                double* row = o.jac + (int64_t)(fl.jrow_off + r) * o.j_sr;
                for (int col = 0; col < N; ++col)
                {
                    row[0] = 0.0;
                    row += o.j_sc;
                }
            }
    }
    if (o.jac && P.zero_cols)
    {
        // columns of the frame rows that no chain joint owns
        for (int f = 0; f < P.n_frames; ++f)
        {
            const ChainFrameP& fp = P.frame[f];
            for (int col = 0; col < N; ++col)
            {
                if (!((P.zero_cols >> col) & 1u)) continue;
                double* jc = o.jac + (int64_t)fp.jrow_off * o.j_sr + (int64_t)col * o.j_sc;
                for (int r = 0; r < fp.n_rows; ++r)
                {
                    jc[0] = 0.0;
                    jc += o.j_sr;
                }
            }
        }
    }
}

// Whole evaluation with CUDA's sincos (Payne-Hanek reduction) and runtime loops: taken only when some joint angle exceeds
// the range of the Cody-Waite reduction (|theta| > 1e5 rad).  Cold code, kept out of line so that it does not constrain
// the register allocation of the hot loop.
__device__ __noinline__ void chain_slow_eval(const ChainParams& P, const Sink& o, const double* xe)
{
    Fr cur;
    cur.m0 = P.base[0]; cur.m1 = P.base[1]; cur.m2 = P.base[2]; cur.m3 = P.base[3]; cur.m4 = P.base[4]; cur.m5 = P.base[5];
    cur.m6 = P.base[6]; cur.m7 = P.base[7]; cur.m8 = P.base[8]; cur.x = P.base[9]; cur.y = P.base[10]; cur.z = P.base[11];
    double W[CH_MAXJ][3], PJ[CH_MAXJ][3];
    for (int j = 0; j < P.n_joints; ++j)
    {
        const ChainJointP& jp = P.joint[j];
        const double xq = xe[jp.x_elem_off];
        if (j > 0)
        {
            Fr l = fr_identity();
            l.m4 = jp.ca; l.m5 = -jp.sa; l.m7 = jp.sa; l.m8 = jp.ca;
            l.x = jp.a;
            cur = fr_mul(cur, l);
        }
        W[j][0] = cur.m2; W[j][1] = cur.m5; W[j][2] = cur.m8;
        PJ[j][0] = cur.x; PJ[j][1] = cur.y; PJ[j][2] = cur.z;
        double sn = jp.s0, cs = jp.c0, d = jp.d;
        if (jp.prismatic)
            d = fma(jp.qmult, xq, d);
        else
            sincos(fma(jp.qmult, xq, jp.th0), &sn, &cs);
        Fr r = fr_identity();
        r.m0 = cs; r.m1 = -sn; r.m3 = sn; r.m4 = cs;
        r.z = d;
        cur = fr_mul(cur, r);
    }
    double pa[CH_MAXF][3];
    chain_frames<CH_MAXF>(P, o, cur, pa);
    if (!o.jac) return;
    for (int j = 0; j < P.n_joints; ++j)
    {
        const ChainJointP& jp = P.joint[j];
        if (jp.col < 0) continue;
        for (int f = 0; f < P.n_frames; ++f)
        {
            const ChainFrameP& fp = P.frame[f];
            double v[3], w[3] = {0.0, 0.0, 0.0};
            if (jp.prismatic)
            {
                v[0] = W[j][0]; v[1] = W[j][1]; v[2] = W[j][2];
            }
            else
            {
                const double dx = pa[f][0] - PJ[j][0], dy = pa[f][1] - PJ[j][1], dz = pa[f][2] - PJ[j][2];
                v[0] = W[j][1] * dz - W[j][2] * dy;
                v[1] = W[j][2] * dx - W[j][0] * dz;
                v[2] = W[j][0] * dy - W[j][1] * dx;
                w[0] = W[j][0]; w[1] = W[j][1]; w[2] = W[j][2];
            }
            double* jc = o.jac + (fp.jrow_elem_off + jp.jcol_elem_off);
            for (int r = 0; r < fp.n_rows; ++r)
            {
                const double* src = r < 3 ? v : w;
                double val = src[r % 3];
                if (fp.b_const) val = fp.binv[3 * (r % 3)] * src[0] + fp.binv[3 * (r % 3) + 1] * src[1] + fp.binv[3 * (r % 3) + 2] * src[2];
                jc[0] = val;
                jc += o.j_sr;
            }
        }
    }
}

template <int NJ, bool ALLREV, int VARIANT, int NF, int J>
struct ChainUnroll
{
    // compose only (first pass of CHAIN_TWO_PASS)
    static __device__ __forceinline__ void compose(const ChainParams& P, Fr& cur, const double (&q)[NJ], const double (&sn)[NJ],
                                                   const double (&cs)[NJ])
    {
        dh_link<J>(P, cur);
        dh_joint<J, ALLREV>(P, cur, sn[J], cs[J], q[J]);
        ChainUnroll<NJ, ALLREV, VARIANT, NF, J + 1>::compose(P, cur, q, sn, cs);
    }
    // compose and emit the column of each joint (second pass of CHAIN_TWO_PASS)
    static __device__ __forceinline__ void emit(const ChainParams& P, const Sink& o, Fr& cur, const double (&q)[NJ], const double (&sn)[NJ],
                                                const double (&cs)[NJ], const double (&pa)[NF][3])
    {
        dh_link<J>(P, cur);
        chain_column<J, ALLREV, NF>(P, o, cur.m2, cur.m5, cur.m8, cur.x, cur.y, cur.z, pa);
        dh_joint<J, ALLREV>(P, cur, sn[J], cs[J], q[J]);
        ChainUnroll<NJ, ALLREV, VARIANT, NF, J + 1>::emit(P, o, cur, q, sn, cs, pa);
    }
    // one pass, joint data to registers (W, PJ) or to shared memory (stash)
    static __device__ __forceinline__ void record(const ChainParams& P, Fr& cur, const double (&q)[NJ], const double (&th)[NJ], double (&W)[NJ][3],
                                                  double (&PJ)[NJ][3], double2* stash)
    {
        double sn = 0.0, cs = 1.0;
        if (ALLREV || !P.joint[J].prismatic) fast_sincos(th[J], sn, cs);
        dh_link<J>(P, cur);
        if (VARIANT == CHAIN_STASH)
        {
            stash[(3 * J + 0) * CTPB] = make_double2(cur.m2, cur.m5);
            stash[(3 * J + 1) * CTPB] = make_double2(cur.m8, cur.x);
            stash[(3 * J + 2) * CTPB] = make_double2(cur.y, cur.z);
        }
        else
        {
            W[J][0] = cur.m2; W[J][1] = cur.m5; W[J][2] = cur.m8;
            PJ[J][0] = cur.x; PJ[J][1] = cur.y; PJ[J][2] = cur.z;
        }
        dh_joint<J, ALLREV>(P, cur, sn, cs, q[J]);
        ChainUnroll<NJ, ALLREV, VARIANT, NF, J + 1>::record(P, cur, q, th, W, PJ, stash);
    }
    static __device__ __forceinline__ void columns(const ChainParams& P, const Sink& o, const double (&W)[NJ][3], const double (&PJ)[NJ][3],
                                                   const double2* stash, const double (&pa)[NF][3])
    {
        if (VARIANT == CHAIN_STASH)
        {
            const double2 a = stash[(3 * J + 0) * CTPB], b = stash[(3 * J + 1) * CTPB], c = stash[(3 * J + 2) * CTPB];
            chain_column<J, ALLREV, NF>(P, o, a.x, a.y, b.x, b.y, c.x, c.y, pa);
        }
        else
            chain_column<J, ALLREV, NF>(P, o, W[J][0], W[J][1], W[J][2], PJ[J][0], PJ[J][1], PJ[J][2], pa);
        ChainUnroll<NJ, ALLREV, VARIANT, NF, J + 1>::columns(P, o, W, PJ, stash, pa);
    }
};
template <int NJ, bool ALLREV, int VARIANT, int NF>
struct ChainUnroll<NJ, ALLREV, VARIANT, NF, NJ>
{
    static __device__ __forceinline__ void compose(const ChainParams&, Fr&, const double (&)[NJ], const double (&)[NJ], const double (&)[NJ]) {}
    static __device__ __forceinline__ void emit(const ChainParams&, const Sink&, Fr&, const double (&)[NJ], const double (&)[NJ],
                                                const double (&)[NJ], const double (&)[NF][3]) {}
    static __device__ __forceinline__ void record(const ChainParams&, Fr&, const double (&)[NJ], const double (&)[NJ], double (&)[NJ][3],
                                                  double (&)[NJ][3], double2*) {}
    static __device__ __forceinline__ void columns(const ChainParams&, const Sink&, const double (&)[NJ][3], const double (&)[NJ][3], const double2*,
                                                   const double (&)[NF][3]) {}
};

template <int VARIANT>
struct ChainOcc
{
#ifdef EXO_CHAIN_MINB
    static constexpr int min_blocks = EXO_CHAIN_MINB;
#else
    static constexpr int min_blocks = VARIANT == CHAIN_REGS ? 3 : (VARIANT == CHAIN_STASH ? 5 : 4);
#endif
};

__device__ __forceinline__ void prefetch_l2(const void* p)
{
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

template <int NJ, bool ALLREV, int VARIANT, int NF>
__global__ void __launch_bounds__(CTPB, ChainOcc<VARIANT>::min_blocks) exo_chain_kernel(const __grid_constant__ ChainParams P)
{
    extern __shared__ double2 stash_all[];
    double2* stash = stash_all + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * CTPB;
    const int64_t x_step = stride * P.x_se;
    int64_t e = (int64_t)blockIdx.x * CTPB + threadIdx.x;
    const double* xe = P.x + e * P.x_se;
    for (; e < P.n_evals; e += stride, xe += x_step)
    {
        double q[NJ], th[NJ];
#pragma unroll
        for (int j = 0; j < NJ; ++j) q[j] = xe[P.joint[j].x_elem_off];
        // pull the next evaluation's joint values into L2 while this one is computed (no registers held)
        if (e + stride < P.n_evals)
        {
#pragma unroll
            for (int j = 0; j < NJ; ++j) prefetch_l2(xe + x_step + P.joint[j].x_elem_off);
        }
        Sink o;
        o.phi = P.phi + e * P.phi_se;
        o.jac = P.jac ? P.jac + e * P.j_se : nullptr;
        o.phi_sk = P.phi_sk;
        o.j_sr = P.j_sr;
        o.j_sc = P.j_sc;
        if (P.has_prefix) chain_prefix(P, o, xe, P.x_sk);

        bool big = false;
#pragma unroll
        for (int j = 0; j < NJ; ++j)
        {
            th[j] = fma(P.joint[j].qmult, q[j], P.joint[j].th0);
            big = big || !(fabs(th[j]) <= 1.0e5);
        }
        if (big)
        {
            chain_slow_eval(P, o, xe);
            continue;
        }

        Fr cur;
        cur.m0 = P.base[0]; cur.m1 = P.base[1]; cur.m2 = P.base[2]; cur.m3 = P.base[3]; cur.m4 = P.base[4]; cur.m5 = P.base[5];
        cur.m6 = P.base[6]; cur.m7 = P.base[7]; cur.m8 = P.base[8]; cur.x = P.base[9]; cur.y = P.base[10]; cur.z = P.base[11];
        double pa[NF][3];
        if (VARIANT == CHAIN_TWO_PASS)
        {
            double sn[NJ], cs[NJ];
#pragma unroll
            for (int j = 0; j < NJ; ++j)
            {
                sn[j] = 0.0;
                cs[j] = 1.0;
                if (ALLREV || !P.joint[j].prismatic) fast_sincos(th[j], sn[j], cs[j]);
            }
            ChainUnroll<NJ, ALLREV, VARIANT, NF, 0>::compose(P, cur, q, sn, cs);
            chain_frames<NF>(P, o, cur, pa);
            if (o.jac)
            {
                cur.m0 = P.base[0]; cur.m1 = P.base[1]; cur.m2 = P.base[2]; cur.m3 = P.base[3]; cur.m4 = P.base[4]; cur.m5 = P.base[5];
                cur.m6 = P.base[6]; cur.m7 = P.base[7]; cur.m8 = P.base[8]; cur.x = P.base[9]; cur.y = P.base[10]; cur.z = P.base[11];
                ChainUnroll<NJ, ALLREV, VARIANT, NF, 0>::emit(P, o, cur, q, sn, cs, pa);
            }
        }
        else
        {
            double W[NJ][3], PJ[NJ][3];
            ChainUnroll<NJ, ALLREV, VARIANT, NF, 0>::record(P, cur, q, th, W, PJ, stash);
            chain_frames<NF>(P, o, cur, pa);
            if (o.jac) ChainUnroll<NJ, ALLREV, VARIANT, NF, 0>::columns(P, o, W, PJ, stash, pa);
        }
    }
}

template <int NJ, bool ALLREV, int VARIANT, int NF>
cudaError_t chain_launch_t(const ChainParams& P, int sms, cudaStream_t stream)
{
    static int ctas_per_sm = 0;  // occupancy is a property of the instantiation: query once
    const size_t smem = VARIANT == CHAIN_STASH ? (size_t)3 * NJ * CTPB * sizeof(double2) : 0;
    if (ctas_per_sm == 0)
    {
        cudaError_t err = cudaFuncSetAttribute(exo_chain_kernel<NJ, ALLREV, VARIANT, NF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return err;
        int ctas = 0;
        err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, exo_chain_kernel<NJ, ALLREV, VARIANT, NF>, CTPB, smem);
        if (err != cudaSuccess) return err;
        if (ctas < 1) return cudaErrorInvalidConfiguration;
        ctas_per_sm = ctas;
    }
    const int64_t want = (P.n_evals + CTPB - 1) / CTPB;
    const int64_t resident = (int64_t)sms * ctas_per_sm;
    const int blocks = (int)(want < resident ? (want > 0 ? want : 1) : resident);
    exo_chain_kernel<NJ, ALLREV, VARIANT, NF><<<blocks, CTPB, smem, stream>>>(P);
    return cudaGetLastError();
}

template <bool ALLREV, int VARIANT, int NF>
cudaError_t chain_launch_nj(const ChainParams& P, int sms, cudaStream_t stream)
{
    switch (P.n_joints)
    {
        case 1: return chain_launch_t<1, ALLREV, VARIANT, NF>(P, sms, stream);
        case 2: return chain_launch_t<2, ALLREV, VARIANT, NF>(P, sms, stream);
        case 3: return chain_launch_t<3, ALLREV, VARIANT, NF>(P, sms, stream);
        case 4: return chain_launch_t<4, ALLREV, VARIANT, NF>(P, sms, stream);
        case 5: return chain_launch_t<5, ALLREV, VARIANT, NF>(P, sms, stream);
        case 6: return chain_launch_t<6, ALLREV, VARIANT, NF>(P, sms, stream);
        case 7: return chain_launch_t<7, ALLREV, VARIANT, NF>(P, sms, stream);
        case 8: return chain_launch_t<8, ALLREV, VARIANT, NF>(P, sms, stream);
        default: return cudaErrorInvalidConfiguration;
    }
}

template <int VARIANT>
cudaError_t chain_launch_v(const ChainParams& P, int sms, cudaStream_t stream)
{
    // the common shapes (all joints revolute; one requested frame) get their own instantiations
    if (P.all_revolute)
        return P.n_frames <= 1 ? chain_launch_nj<true, VARIANT, 1>(P, sms, stream) : chain_launch_nj<true, VARIANT, CH_MAXF>(P, sms, stream);
    return chain_launch_nj<false, VARIANT, CH_MAXF>(P, sms, stream);
}
}  // namespace

void chain_bind(ChainParams& P, const double* x, double* phi, double* jac, int64_t n_evals, int64_t x_se, int64_t x_sk, int64_t phi_se,
                int64_t phi_sk, int64_t j_se, int64_t j_sr, int64_t j_sc)
{
    P.x = x; P.phi = phi; P.jac = jac; P.n_evals = n_evals;
    P.x_se = x_se; P.x_sk = x_sk; P.phi_se = phi_se; P.phi_sk = phi_sk; P.j_se = j_se; P.j_sr = j_sr; P.j_sc = j_sc;
    for (int j = 0; j < P.n_joints; ++j)
    {
        P.joint[j].x_elem_off = (P.joint[j].qidx >= 0 ? P.joint[j].qidx : 0) * x_sk;
        P.joint[j].jcol_elem_off = (P.joint[j].col >= 0 ? P.joint[j].col : 0) * j_sc;
    }
    for (int f = 0; f < P.n_frames; ++f)
    {
        P.frame[f].phi_elem_off = P.frame[f].phi_off * phi_sk;
        P.frame[f].jrow_elem_off = P.frame[f].jrow_off * j_sr;
    }
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
        case CHAIN_STASH: return chain_launch_v<CHAIN_STASH>(P, sms, stream);
        case CHAIN_TWO_PASS: return chain_launch_v<CHAIN_TWO_PASS>(P, sms, stream);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace exo
