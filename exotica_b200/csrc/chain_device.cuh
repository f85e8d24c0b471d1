// chain_device.cuh -- device code shared by the serial-chain kernels (chain_kernel.cu: per-lane streaming stores;
// chain_tma_kernel.cu: shared-memory staging + TMA tensor stores): the modified-DH link / joint steps, the unrolled chain,
// Jacobian columns (KinematicTree::ComputeJ, exotica_core/src/kinematic_tree.cpp:919-945), the JointLimit / fill prefix and the
// cold big-angle path.  Include after defining EXO_PHI_STREAMING / EXO_STORE_CS as the translation unit wants them.
#pragma once
#include "compiler.h"
#include "cost_device.cuh"
#include "ik_device.cuh"
#include "device_math.cuh"
#include "kernels.cuh"

namespace exo
{
namespace
{
using namespace dev;

#ifndef EXO_CHAIN_CTPB
#define EXO_CHAIN_CTPB 64
#endif
constexpr int CTPB = EXO_CHAIN_CTPB;

struct Sink
{
    double* phi;  // &Phi[e][0]
    double* jac;  // &J[e][0][0] or nullptr
    int64_t phi_sk, j_sr, j_sc;
};

__device__ __forceinline__ Fr load_frame(const double* c)
{
    Fr f;
    f.m0 = c[0]; f.m1 = c[1]; f.m2 = c[2]; f.m3 = c[3]; f.m4 = c[4]; f.m5 = c[5];
    f.m6 = c[6]; f.m7 = c[7]; f.m8 = c[8]; f.x = c[9]; f.y = c[10]; f.z = c[11];
    return f;
}

__device__ __forceinline__ void st(char* p, double v)
{
#ifndef EXO_CHAIN_NO_STCS
    __stcs(reinterpret_cast<double*>(p), v);  // streaming store: cfg3 75.6 -> 73.3 us, cfg2a 86.3 -> 86.0 us (outputs are not re-read here)
#else
    *reinterpret_cast<double*>(p) = v;
#endif
}

// sin / cos of (theta0 + qm) for |angle| > 1e5 rad: adding the constant first would round it away at that magnitude (ulp(1e9) =
// 1.2e-7), so the two angles are combined after the (Payne-Hanek) sincos of the state-dependent part
__device__ __forceinline__ void big_sincos(double qm, double c0, double s0, double& sn, double& cs)
{
    double sq, cq;
    sincos(qm, &sq, &cq);
    sn = fma(sq, c0, cq * s0);
    cs = fma(cq, c0, -(sq * s0));
}

// Rx(alpha_J) * Tx(a_J) applied on the right of cur (joint 0: already part of the base frame).
// ORTHO: every link of the chain has alpha = +-90 degrees and a = 0 (consecutive axes perpendicular and intersecting: the LWR,
// most 7-DoF arms) -- the rotation is a signed swap of the y and z columns, 6 multiplications instead of 15 FP64 operations.
// Same bits as the general form: fma(0, y, s * z) = s * z and a * x contributes an exact zero.
template <int J, bool ORTHO>
__device__ __forceinline__ void dh_link(const ChainParams& P, Fr& cur)
{
    if (J == 0) return;
    const double sa = P.joint[J].sa;
    const double y0 = cur.m1, y1 = cur.m4, y2 = cur.m7;
    if (ORTHO)
    {
        cur.m1 = sa * cur.m2; cur.m2 = -(sa * y0);
        cur.m4 = sa * cur.m5; cur.m5 = -(sa * y1);
        cur.m7 = sa * cur.m8; cur.m8 = -(sa * y2);
        return;
    }
    const double ca = P.joint[J].ca, a = P.joint[J].a;
    cur.x = fma(a, cur.m0, cur.x);
    cur.y = fma(a, cur.m3, cur.y);
    cur.z = fma(a, cur.m6, cur.z);
    cur.m1 = fma(ca, y0, sa * cur.m2); cur.m2 = fma(ca, cur.m2, -(sa * y0));
    cur.m4 = fma(ca, y1, sa * cur.m5); cur.m5 = fma(ca, cur.m5, -(sa * y1));
    cur.m7 = fma(ca, y2, sa * cur.m8); cur.m8 = fma(ca, cur.m8, -(sa * y2));
}

// Rz(theta) * Tz(d) applied on the right of cur; xq is the state entry feeding the joint
template <int J, int ALLREV>
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

// Jacobian column of joint J for requested frame F (KinematicTree::ComputeJ, kinematic_tree.cpp:919-945):
//   column = temp_B.M^-1 * [ w x (temp_A.p - p_joint) ; w ]   (revolute)   |   temp_B.M^-1 * [ w ; 0 ]   (prismatic)
// WITHB / SIX are per-frame facts hoisted out of the joint loop by the caller.
template <int J, int ALLREV, bool WITHB, bool SIX>
__device__ __forceinline__ void chain_column(const ChainParams& P, int f, char* jac_e, int64_t sr_b, double wx, double wy, double wz, double px,
                                             double py, double pz, double ax, double ay, double az)
{
    const ChainJointP& jp = P.joint[J];
    if (jp.col < 0) return;
    const ChainFrameP& fp = P.frame[f];
    double v0, v1, v2, w0, w1, w2;
    if (!ALLREV && jp.prismatic)
    {
        v0 = wx; v1 = wy; v2 = wz;
        w0 = w1 = w2 = 0.0;
    }
    else
    {
        const double dx = ax - px, dy = ay - py, dz = az - pz;
        v0 = wy * dz - wz * dy;
        v1 = wz * dx - wx * dz;
        v2 = wx * dy - wy * dx;
        w0 = wx; w1 = wy; w2 = wz;
    }
    if (WITHB)
    {
        const double t0 = fp.binv[0] * v0 + fp.binv[1] * v1 + fp.binv[2] * v2;
        const double t1 = fp.binv[3] * v0 + fp.binv[4] * v1 + fp.binv[5] * v2;
        const double t2 = fp.binv[6] * v0 + fp.binv[7] * v1 + fp.binv[8] * v2;
        v0 = t0; v1 = t1; v2 = t2;
        if (SIX)
        {
            const double u0 = fp.binv[0] * w0 + fp.binv[1] * w1 + fp.binv[2] * w2;
            const double u1 = fp.binv[3] * w0 + fp.binv[4] * w1 + fp.binv[5] * w2;
            const double u2 = fp.binv[6] * w0 + fp.binv[7] * w1 + fp.binv[8] * w2;
            w0 = u0; w1 = u1; w2 = u2;
        }
    }
    char* jc = jac_e + P.jf_byte_off[f][J];
    st(jc, v0); jc += sr_b;
    st(jc, v1); jc += sr_b;
    st(jc, v2);
    if (SIX)
    {
        jc += sr_b; st(jc, w0);
        jc += sr_b; st(jc, w1);
        jc += sr_b; st(jc, w2);
    }
}

// Same for the maps that keep other rows of the column (EffOrientation, EffPositionXY, Distance): runtime row mode
template <int J, int ALLREV>
__device__ __forceinline__ void chain_column_g(const ChainParams& P, int f, char* jac_e, int64_t sr_b, double wx, double wy, double wz, double px,
                                               double py, double pz, double ax, double ay, double az, double ux, double uy, double uz)
{
    const ChainJointP& jp = P.joint[J];
    if (jp.col < 0) return;
    const ChainFrameP& fp = P.frame[f];
    double v0, v1, v2, w0, w1, w2;
    if (!ALLREV && jp.prismatic)
    {
        v0 = wx; v1 = wy; v2 = wz;
        w0 = w1 = w2 = 0.0;
    }
    else
    {
        const double dx = ax - px, dy = ay - py, dz = az - pz;
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
    store_rows(fp.row_mode, reinterpret_cast<double*>(jac_e + P.jf_byte_off[f][J]), sr_b >> 3, v0, v1, v2, w0, w1, w2, ux, uy, uz);
}

// temp_A = D_{n-1} * K_f
__device__ __forceinline__ Fr chain_frame_a(const ChainFrameP& fp, const Fr& tip)
{
    Fr A = tip;
    if (fp.k_kind == F_OFFA_TRANS)
        fr_translate(A, fp.k[9], fp.k[10], fp.k[11]);
    else if (fp.k_kind == F_OFFA_FULL)
        A = fr_mul(A, load_frame(fp.k));
    return A;
}

// Phi = temp_B^-1 * temp_A, encoded as the task map wants it
__device__ __forceinline__ void chain_frame_phi(const ChainFrameP& fp, Fr A, char* phi_e, int64_t phi_sk)
{
    if (fp.b_const) A = fr_mul(load_frame(fp.binv), A);
    write_phi(fp.map_kind, fp.rot_type, A, reinterpret_cast<double*>(phi_e + fp.phi_byte_off), phi_sk);
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
            __stcs(ph, v);
            ph += o.phi_sk;
            if (o.jac)
            {
                const double d = (q >= lo_t && q <= hi_t) ? 0.0 : 1.0;
                double* row = o.jac + (int64_t)(P.jl_jrow_off + i) * o.j_sr;
                for (int col = 0; col < N; ++col)
                {
                    __stcs(row, (col == i) ? d : 0.0);
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
            __stcs(ph, fl.value[i]);
            ph += o.phi_sk;
        }
        if (o.jac)
            for (int r = 0; r < fl.n_rows; ++r)
            {
                double* row = o.jac + (int64_t)(fl.jrow_off + r) * o.j_sr;
                for (int col = 0; col < N; ++col)
                {
                    __stcs(row, 0.0);
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
                    __stcs(jc, 0.0);
                    jc += o.j_sr;
                }
            }
        }
    }
}

// Whole evaluation with CUDA's sincos (Payne-Hanek reduction) and runtime loops: taken only when some joint angle exceeds
// the range of the Cody-Waite reduction (|theta| > 1e5 rad).  Cold code, kept out of line so that it does not constrain
// the register allocation of the hot loop.
__device__ __noinline__ void chain_slow_eval(const ChainParams& P, const char* xe, char* phi_e, char* jac_e)
{
    Fr cur = load_frame(P.base);
    double W[CH_MAXJ][3], PJ[CH_MAXJ][3];
    for (int j = 0; j < P.n_joints; ++j)
    {
        const ChainJointP& jp = P.joint[j];
        const double xq = *reinterpret_cast<const double*>(xe + jp.x_byte_off);
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
            big_sincos(jp.qmult * xq, jp.c0, jp.s0, sn, cs);
        Fr r = fr_identity();
        r.m0 = cs; r.m1 = -sn; r.m3 = sn; r.m4 = cs;
        r.z = d;
        cur = fr_mul(cur, r);
    }
    for (int f = 0; f < P.n_frames; ++f)
    {
        const ChainFrameP& fp = P.frame[f];
        const Fr A = chain_frame_a(fp, cur);
        chain_frame_phi(fp, A, phi_e, P.phi_sk);
        if (!jac_e) continue;
        for (int j = 0; j < P.n_joints; ++j)
        {
            const ChainJointP& jp = P.joint[j];
            if (jp.col < 0) continue;
            double v[3], w[3] = {0.0, 0.0, 0.0};
            if (jp.prismatic)
            {
                v[0] = W[j][0]; v[1] = W[j][1]; v[2] = W[j][2];
            }
            else
            {
                const double dx = A.x - PJ[j][0], dy = A.y - PJ[j][1], dz = A.z - PJ[j][2];
                v[0] = W[j][1] * dz - W[j][2] * dy;
                v[1] = W[j][2] * dx - W[j][0] * dz;
                v[2] = W[j][0] * dy - W[j][1] * dx;
                w[0] = W[j][0]; w[1] = W[j][1]; w[2] = W[j][2];
            }
            if (fp.b_const)
            {
                const double t[3] = {v[0], v[1], v[2]}, u[3] = {w[0], w[1], w[2]};
                for (int r = 0; r < 3; ++r)
                {
                    v[r] = fp.binv[3 * r] * t[0] + fp.binv[3 * r + 1] * t[1] + fp.binv[3 * r + 2] * t[2];
                    w[r] = fp.binv[3 * r] * u[0] + fp.binv[3 * r + 1] * u[1] + fp.binv[3 * r + 2] * u[2];
                }
            }
            double ux = 0.0, uy = 0.0, uz = 0.0;
            if (fp.row_mode == ROWS_DIST)
            {
                const Fr Ph = fp.b_const ? fr_mul(load_frame(fp.binv), A) : A;
                const double nrm = norm3_scaled(Ph.x, Ph.y, Ph.z);
                ux = Ph.x / nrm; uy = Ph.y / nrm; uz = Ph.z / nrm;
            }
            store_rows(fp.row_mode, reinterpret_cast<double*>(jac_e + P.jf_byte_off[f][j]), P.j_sr, v[0], v[1], v[2], w[0], w[1], w[2], ux, uy, uz);
        }
    }
}

template <int NJ, int ALLREV, int J>
struct ChainUnroll
{
    // compose only (first pass of CHAIN_TWO_PASS)
    static __device__ __forceinline__ void compose(const ChainParams& P, Fr& cur, const double (&q)[NJ], const double (&sn)[NJ],
                                                   const double (&cs)[NJ])
    {
        dh_link<J, (ALLREV & 2) != 0>(P, cur);
        dh_joint<J, ALLREV>(P, cur, sn[J], cs[J], q[J]);
        ChainUnroll<NJ, ALLREV, J + 1>::compose(P, cur, q, sn, cs);
    }
    // compose and emit the column of each joint for frame 0 (second pass of CHAIN_TWO_PASS, one requested frame)
    template <bool WITHB, bool SIX>
    static __device__ __forceinline__ void emit(const ChainParams& P, char* jac_e, int64_t sr_b, Fr& cur, const double (&q)[NJ],
                                                const double (&sn)[NJ], const double (&cs)[NJ], double ax, double ay, double az)
    {
        dh_link<J, (ALLREV & 2) != 0>(P, cur);
        chain_column<J, ALLREV, WITHB, SIX>(P, 0, jac_e, sr_b, cur.m2, cur.m5, cur.m8, cur.x, cur.y, cur.z, ax, ay, az);
        dh_joint<J, ALLREV>(P, cur, sn[J], cs[J], q[J]);
        ChainUnroll<NJ, ALLREV, J + 1>::template emit<WITHB, SIX>(P, jac_e, sr_b, cur, q, sn, cs, ax, ay, az);
    }
    // one pass: compose and keep every joint's axis and origin
    static __device__ __forceinline__ void record(const ChainParams& P, Fr& cur, const double (&q)[NJ], const double (&sn)[NJ],
                                                  const double (&cs)[NJ], double (&W)[NJ][3], double (&PJ)[NJ][3])
    {
        dh_link<J, (ALLREV & 2) != 0>(P, cur);
        W[J][0] = cur.m2; W[J][1] = cur.m5; W[J][2] = cur.m8;
        PJ[J][0] = cur.x; PJ[J][1] = cur.y; PJ[J][2] = cur.z;
        dh_joint<J, ALLREV>(P, cur, sn[J], cs[J], q[J]);
        ChainUnroll<NJ, ALLREV, J + 1>::record(P, cur, q, sn, cs, W, PJ);
    }
    static __device__ __forceinline__ void columns_g(const ChainParams& P, int f, char* jac_e, int64_t sr_b, const double (&W)[NJ][3],
                                                     const double (&PJ)[NJ][3], double ax, double ay, double az, double ux, double uy, double uz)
    {
        chain_column_g<J, ALLREV>(P, f, jac_e, sr_b, W[J][0], W[J][1], W[J][2], PJ[J][0], PJ[J][1], PJ[J][2], ax, ay, az, ux, uy, uz);
        ChainUnroll<NJ, ALLREV, J + 1>::columns_g(P, f, jac_e, sr_b, W, PJ, ax, ay, az, ux, uy, uz);
    }
    template <bool WITHB, bool SIX>
    static __device__ __forceinline__ void columns(const ChainParams& P, int f, char* jac_e, int64_t sr_b, const double (&W)[NJ][3],
                                                   const double (&PJ)[NJ][3], double ax, double ay, double az)
    {
        chain_column<J, ALLREV, WITHB, SIX>(P, f, jac_e, sr_b, W[J][0], W[J][1], W[J][2], PJ[J][0], PJ[J][1], PJ[J][2], ax, ay, az);
        ChainUnroll<NJ, ALLREV, J + 1>::template columns<WITHB, SIX>(P, f, jac_e, sr_b, W, PJ, ax, ay, az);
    }
};
template <int NJ, int ALLREV>
struct ChainUnroll<NJ, ALLREV, NJ>
{
    static __device__ __forceinline__ void compose(const ChainParams&, Fr&, const double (&)[NJ], const double (&)[NJ], const double (&)[NJ]) {}
    template <bool WITHB, bool SIX>
    static __device__ __forceinline__ void emit(const ChainParams&, char*, int64_t, Fr&, const double (&)[NJ], const double (&)[NJ],
                                                const double (&)[NJ], double, double, double) {}
    static __device__ __forceinline__ void record(const ChainParams&, Fr&, const double (&)[NJ], const double (&)[NJ], const double (&)[NJ],
                                                  double (&)[NJ][3], double (&)[NJ][3]) {}
    static __device__ __forceinline__ void columns_g(const ChainParams&, int, char*, int64_t, const double (&)[NJ][3], const double (&)[NJ][3], double,
                                                     double, double, double, double, double) {}
    template <bool WITHB, bool SIX>
    static __device__ __forceinline__ void columns(const ChainParams&, int, char*, int64_t, const double (&)[NJ][3], const double (&)[NJ][3], double,
                                                   double, double) {}
};

// Register budget per variant (B200: 64 K registers per SM, CTAs of 64 threads): 144 registers -> 7 CTAs = 14 warps per
// SM for CHAIN_REGS (no spills), 128 -> 8 CTAs = 16 warps for CHAIN_TWO_PASS.  Measured alternatives are in DESIGN.md.
template <int VARIANT>
struct ChainOcc
{
#ifdef EXO_CHAIN_MAXNREG
    static constexpr int maxnreg = EXO_CHAIN_MAXNREG;
#else
    static constexpr int maxnreg = VARIANT == CHAIN_REGS ? 144 : 128;
#endif
};

__device__ __forceinline__ void prefetch_l2(const void* p)
{
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

}  // namespace
}  // namespace exo
