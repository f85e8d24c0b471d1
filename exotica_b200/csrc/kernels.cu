// kernels.cu -- CUDA interpreter for the compiled tree program (sm_100a).
//
// Replaces, for a whole batch at once, the reference's per-configuration
//   KinematicTree::Update -> UpdateTree -> UpdateFK -> UpdateJ   (exotica_core/src/kinematic_tree.cpp:678-737,829-867,919-988)
//   EffFrame/EffPosition/JointLimit/SphereCollision::Update      (exotations/exotica_core_task_maps/src/*.cpp)
//   Phi.SetZero / jacobian.setZero of the problem Update           (abstract_time_indexed_problem.cpp:243-244)
//
// Mapping: one lane = one evaluation (configuration or (trajectory,timestep)); a warp walks the op list in
// lock-step, so every branch on op fields is warp-uniform and every program constant is a shared-memory
// broadcast.  The running frame lives in registers (FP64); frames/joint data that must outlive the next op
// sit in per-thread shared-memory slots laid out [slot word][thread] (conflict-free).  The kernel is
// persistent (grid = SMs x resident CTAs, grid-stride over the batch) and writes every entry of Phi and J
// exactly once (zeros included), so no memset pass is needed.  HBM-bound FP64 streaming work: no tensor cores.
#include "kernels.cuh"

#include <math_constants.h>

#include "../../include/exotica_b200.h"

namespace exo
{
namespace
{

struct Fr
{
    double m0, m1, m2, m3, m4, m5, m6, m7, m8;  // row-major rotation
    double x, y, z;
};

__device__ __forceinline__ Fr fr_identity()
{
    Fr f;
    f.m0 = 1.0; f.m1 = 0.0; f.m2 = 0.0;
    f.m3 = 0.0; f.m4 = 1.0; f.m5 = 0.0;
    f.m6 = 0.0; f.m7 = 0.0; f.m8 = 1.0;
    f.x = f.y = f.z = 0.0;
    return f;
}

template <int TPB>
__device__ __forceinline__ Fr fr_load_scratch(const double* s)
{
    Fr f;
    f.m0 = s[0 * TPB]; f.m1 = s[1 * TPB]; f.m2 = s[2 * TPB];
    f.m3 = s[3 * TPB]; f.m4 = s[4 * TPB]; f.m5 = s[5 * TPB];
    f.m6 = s[6 * TPB]; f.m7 = s[7 * TPB]; f.m8 = s[8 * TPB];
    f.x = s[9 * TPB]; f.y = s[10 * TPB]; f.z = s[11 * TPB];
    return f;
}

template <int TPB>
__device__ __forceinline__ void fr_store_scratch(double* s, const Fr& f)
{
    s[0 * TPB] = f.m0; s[1 * TPB] = f.m1; s[2 * TPB] = f.m2;
    s[3 * TPB] = f.m3; s[4 * TPB] = f.m4; s[5 * TPB] = f.m5;
    s[6 * TPB] = f.m6; s[7 * TPB] = f.m7; s[8 * TPB] = f.m8;
    s[9 * TPB] = f.x; s[10 * TPB] = f.y; s[11 * TPB] = f.z;
}

__device__ __forceinline__ Fr fr_load_const(const double* c)
{
    Fr f;
    f.m0 = c[0]; f.m1 = c[1]; f.m2 = c[2];
    f.m3 = c[3]; f.m4 = c[4]; f.m5 = c[5];
    f.m6 = c[6]; f.m7 = c[7]; f.m8 = c[8];
    f.x = c[9]; f.y = c[10]; f.z = c[11];
    return f;
}

// a * b  (Frame * Frame = (Ma Mb, Ma pb + pa))
__device__ __forceinline__ Fr fr_mul(const Fr& a, const Fr& b)
{
    Fr r;
    r.m0 = a.m0 * b.m0 + a.m1 * b.m3 + a.m2 * b.m6;
    r.m1 = a.m0 * b.m1 + a.m1 * b.m4 + a.m2 * b.m7;
    r.m2 = a.m0 * b.m2 + a.m1 * b.m5 + a.m2 * b.m8;
    r.m3 = a.m3 * b.m0 + a.m4 * b.m3 + a.m5 * b.m6;
    r.m4 = a.m3 * b.m1 + a.m4 * b.m4 + a.m5 * b.m7;
    r.m5 = a.m3 * b.m2 + a.m4 * b.m5 + a.m5 * b.m8;
    r.m6 = a.m6 * b.m0 + a.m7 * b.m3 + a.m8 * b.m6;
    r.m7 = a.m6 * b.m1 + a.m7 * b.m4 + a.m8 * b.m7;
    r.m8 = a.m6 * b.m2 + a.m7 * b.m5 + a.m8 * b.m8;
    r.x = a.m0 * b.x + a.m1 * b.y + a.m2 * b.z + a.x;
    r.y = a.m3 * b.x + a.m4 * b.y + a.m5 * b.z + a.y;
    r.z = a.m6 * b.x + a.m7 * b.y + a.m8 * b.z + a.z;
    return r;
}

// a * (I, t)
__device__ __forceinline__ void fr_translate(Fr& a, double tx, double ty, double tz)
{
    a.x += a.m0 * tx + a.m1 * ty + a.m2 * tz;
    a.y += a.m3 * tx + a.m4 * ty + a.m5 * tz;
    a.z += a.m6 * tx + a.m7 * ty + a.m8 * tz;
}

__device__ __forceinline__ Fr fr_inverse(const Fr& a)
{
    Fr r;
    r.m0 = a.m0; r.m1 = a.m3; r.m2 = a.m6;
    r.m3 = a.m1; r.m4 = a.m4; r.m5 = a.m7;
    r.m6 = a.m2; r.m7 = a.m5; r.m8 = a.m8;
    r.x = -(a.m0 * a.x + a.m3 * a.y + a.m6 * a.z);
    r.y = -(a.m1 * a.x + a.m4 * a.y + a.m7 * a.z);
    r.z = -(a.m2 * a.x + a.m5 * a.y + a.m8 * a.z);
    return r;
}

// ---- rotation encodings: exotica::SetRotation (exotica_core/src/tools/conversions.cpp:150-179) over KDL / Eigen ----
__device__ __forceinline__ void get_rpy(const Fr& f, double& roll, double& pitch, double& yaw)
{
    // KDL::Rotation::GetRPY, including the gimbal-lock branch at |pitch| > pi/2 - 1e-12
    pitch = atan2(-f.m6, sqrt(f.m0 * f.m0 + f.m3 * f.m3));
    if (fabs(pitch) > (CUDART_PIO2 - 1e-12))
    {
        yaw = atan2(-f.m1, f.m4);
        roll = 0.0;
    }
    else
    {
        roll = atan2(f.m7, f.m8);
        yaw = atan2(f.m3, f.m0);
    }
}

__device__ __forceinline__ void get_zyz(const Fr& f, double& alpha, double& beta, double& gamma)
{
    if (fabs(f.m8) > 1 - 1e-12)
    {
        gamma = 0.0;
        if (f.m8 > 0)
        {
            beta = 0.0;
            alpha = atan2(f.m3, f.m0);
        }
        else
        {
            beta = CUDART_PI;
            alpha = atan2(-f.m3, -f.m0);
        }
    }
    else
    {
        alpha = atan2(f.m5, f.m2);
        beta = atan2(sqrt(f.m6 * f.m6 + f.m7 * f.m7), f.m8);
        gamma = atan2(f.m7, -f.m6);
    }
}

__device__ __forceinline__ double norm3_scaled(double a, double b, double c)
{
    // KDL::Vector::Norm (scaled form)
    double t1 = fabs(a), t2 = fabs(b);
    if (t1 >= t2)
    {
        t2 = fabs(c);
        if (t1 >= t2)
        {
            if (t1 == 0) return 0;
            return t1 * sqrt(1 + (b / a) * (b / a) + (c / a) * (c / a));
        }
        return t2 * sqrt(1 + (a / c) * (a / c) + (b / c) * (b / c));
    }
    t1 = fabs(c);
    if (t2 > t1) return t2 * sqrt(1 + (a / b) * (a / b) + (c / b) * (c / b));
    return t1 * sqrt(1 + (a / c) * (a / c) + (b / c) * (b / c));
}

__device__ __noinline__ void get_angle_axis(const Fr& f, double& ox, double& oy, double& oz)
{
    // KDL::Rotation::GetRot = axis * angle with GetRotAngle's symmetric-matrix handling
    const double eps = 1e-6, eps2 = 1e-5;
    double x, y, z, angle;
    if ((fabs(f.m1 - f.m3) < eps) && (fabs(f.m2 - f.m6) < eps) && (fabs(f.m5 - f.m7) < eps))
    {
        if ((fabs(f.m1 + f.m3) < eps2) && (fabs(f.m2 + f.m6) < eps2) && (fabs(f.m5 + f.m7) < eps2) && (fabs(f.m0 + f.m4 + f.m8 - 3) < eps2))
        {
            ox = oy = oz = 0.0;
            return;
        }
        angle = CUDART_PI;
        const double xx = (f.m0 + 1) / 2, yy = (f.m4 + 1) / 2, zz = (f.m8 + 1) / 2;
        const double xy = (f.m1 + f.m3) / 4, xz = (f.m2 + f.m6) / 4, yz = (f.m5 + f.m7) / 4;
        if ((xx > yy) && (xx > zz))
        {
            x = sqrt(xx);
            y = xy / x;
            z = xz / x;
        }
        else if (yy > zz)
        {
            y = sqrt(yy);
            x = xy / y;
            z = yz / y;
        }
        else
        {
            z = sqrt(zz);
            x = xz / z;
            y = yz / z;
        }
    }
    else
    {
        double c = (f.m0 + f.m4 + f.m8 - 1) / 2;
        c = fmin(1.0, fmax(-1.0, c));
        angle = acos(c);
        x = f.m7 - f.m5;
        y = f.m2 - f.m6;
        z = f.m3 - f.m1;
        const double nrm = norm3_scaled(x, y, z);
        x /= nrm;
        y /= nrm;
        z /= nrm;
    }
    ox = x * angle;
    oy = y * angle;
    oz = z * angle;
}

__device__ __noinline__ void get_quaternion(const Fr& f, double* q /* x y z w */)
{
    // Eigen::Quaterniond(Matrix3d).coeffs()
    const double m[9] = {f.m0, f.m1, f.m2, f.m3, f.m4, f.m5, f.m6, f.m7, f.m8};
    double t = m[0] + m[4] + m[8];
    if (t > 0.0)
    {
        t = sqrt(t + 1.0);
        q[3] = 0.5 * t;
        t = 0.5 / t;
        q[0] = (m[7] - m[5]) * t;
        q[1] = (m[2] - m[6]) * t;
        q[2] = (m[3] - m[1]) * t;
    }
    else
    {
        int i = 0;
        if (m[4] > m[0]) i = 1;
        if (m[8] > m[4 * i]) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        t = sqrt(m[4 * i] - m[4 * j] - m[4 * k] + 1.0);
        double qi = 0.5 * t;
        t = 0.5 / t;
        const double qw = (m[3 * k + j] - m[3 * j + k]) * t;
        const double qj = (m[3 * j + i] + m[3 * i + j]) * t;
        const double qk = (m[3 * k + i] + m[3 * i + k]) * t;
        // scatter without dynamic register indexing
        q[0] = (i == 0) ? qi : ((j == 0) ? qj : qk);
        q[1] = (i == 1) ? qi : ((j == 1) ? qj : qk);
        q[2] = (i == 2) ? qi : ((j == 2) ? qj : qk);
        q[3] = qw;
    }
}

template <int TPB>
__global__ void __launch_bounds__(TPB) exo_eval_kernel(const KernelParams P)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x;
    // ---- stage the program (topology, axes, offsets, tables) once per CTA ---------------------------------
    {
        const int n_words = P.program_bytes >> 2;
        const int32_t* src = reinterpret_cast<const int32_t*>(P.program);
        int32_t* dst = reinterpret_cast<int32_t*>(smem);
        for (int i = tid; i < n_words; i += TPB) dst[i] = src[i];
    }
    __syncthreads();
    const ProgramHeader* H = reinterpret_cast<const ProgramHeader*>(smem);
    const Op* ops = reinterpret_cast<const Op*>(smem + sizeof(ProgramHeader));
    const int32_t* ints = reinterpret_cast<const int32_t*>(ops + H->n_ops);
    const double* dbls = reinterpret_cast<const double*>(smem + P.dbl_offset);
    double* scr = reinterpret_cast<double*>(smem + P.scratch_offset) + tid;  // word k at scr[k * TPB]
    const int N = H->n_ctrl;
    const int joint_base = 12 * H->n_frame_slots;
    const int acc_base = joint_base + 6 * H->n_joint_slots;
    double* qs = scr + (size_t)H->scratch_doubles * TPB;
    const bool want_j = P.jac != nullptr;

    for (int64_t e = (int64_t)blockIdx.x * TPB + tid; e < P.n_evals; e += (int64_t)gridDim.x * TPB)
    {
        // ---- stage this evaluation's state (N independent loads in flight) --------------------------------
        {
            const double* xe = P.x + e * P.x_se;
            for (int i = 0; i < N; ++i) qs[i * TPB] = xe[i * P.x_sk];
        }
        double* phi_e = P.phi + e * P.phi_se;
        double* jac_e = want_j ? P.jac + e * P.j_se : nullptr;
        Fr cur = fr_identity();

        for (int pc = 0;; ++pc)
        {
            const Op op = ops[pc];
            if (op.kind == OP_END) break;
            const double* c = dbls + op.cst;
            if (op.kind == OP_JOINT || op.kind == OP_TABLE)
            {
                if (op.src == SRC_ROOT)
                    cur = fr_identity();
                else if (op.src >= 0)
                    cur = fr_load_scratch<TPB>(scr + 12 * op.src * TPB);
                const int pre = op.flags & F_PRE_MASK;
                if (pre == F_PRE_TRANS)
                    fr_translate(cur, c[9], c[10], c[11]);
                else if (pre == F_PRE_FULL)
                    cur = fr_mul(cur, fr_load_const(c));
                if (op.kind == OP_TABLE)
                {
                    const int t = (int)(e % P.T);
                    const double* g = P.generated + ((size_t)t * H->n_generated + op.qidx) * 12;
                    Fr gf;
                    gf.m0 = g[0]; gf.m1 = g[1]; gf.m2 = g[2]; gf.m3 = g[3]; gf.m4 = g[4]; gf.m5 = g[5];
                    gf.m6 = g[6]; gf.m7 = g[7]; gf.m8 = g[8]; gf.x = g[9]; gf.y = g[10]; gf.z = g[11];
                    cur = fr_mul(cur, gf);
                }
                else
                {
                    double q = c[CJ_QOFF];
                    if (op.qidx >= 0) q = fma(c[CJ_QMULT], qs[op.qidx * TPB], q);
                    const int ax = op.flags & F_AXIS_MASK;
                    double wx, wy, wz;  // axis in world coordinates
                    if (op.flags & F_PRISMATIC)
                    {
                        if (ax == F_AXIS_GENERAL)
                        {
                            const double ux = c[CJ_AXIS], uy = c[CJ_AXIS + 1], uz = c[CJ_AXIS + 2];
                            wx = cur.m0 * ux + cur.m1 * uy + cur.m2 * uz;
                            wy = cur.m3 * ux + cur.m4 * uy + cur.m5 * uz;
                            wz = cur.m6 * ux + cur.m7 * uy + cur.m8 * uz;
                        }
                        else
                        {
                            const double sg = (op.flags & F_AXIS_NEG) ? -1.0 : 1.0;
                            if (ax == F_AXIS_X) { wx = sg * cur.m0; wy = sg * cur.m3; wz = sg * cur.m6; }
                            else if (ax == F_AXIS_Y) { wx = sg * cur.m1; wy = sg * cur.m4; wz = sg * cur.m7; }
                            else { wx = sg * cur.m2; wy = sg * cur.m5; wz = sg * cur.m8; }
                        }
                        cur.x = fma(wx, q, cur.x);
                        cur.y = fma(wy, q, cur.y);
                        cur.z = fma(wz, q, cur.z);
                    }
                    else
                    {
                        double s, cs;
                        sincos(q, &s, &cs);
                        if (ax == F_AXIS_GENERAL)
                        {
                            const double ux = c[CJ_AXIS], uy = c[CJ_AXIS + 1], uz = c[CJ_AXIS + 2];
                            const double vt = 1.0 - cs;
                            Fr r;
                            r.m0 = cs + vt * ux * ux;
                            r.m1 = -uz * s + vt * ux * uy;
                            r.m2 = uy * s + vt * ux * uz;
                            r.m3 = uz * s + vt * ux * uy;
                            r.m4 = cs + vt * uy * uy;
                            r.m5 = -ux * s + vt * uy * uz;
                            r.m6 = -uy * s + vt * ux * uz;
                            r.m7 = ux * s + vt * uy * uz;
                            r.m8 = cs + vt * uz * uz;
                            r.x = r.y = r.z = 0.0;
                            const double px = cur.x, py = cur.y, pz = cur.z;
                            cur = fr_mul(cur, r);
                            cur.x = px; cur.y = py; cur.z = pz;
                            wx = cur.m0 * ux + cur.m1 * uy + cur.m2 * uz;
                            wy = cur.m3 * ux + cur.m4 * uy + cur.m5 * uz;
                            wz = cur.m6 * ux + cur.m7 * uy + cur.m8 * uz;
                        }
                        else
                        {
                            const bool neg = (op.flags & F_AXIS_NEG) != 0;
                            if (neg) s = -s;
                            const double sg = neg ? -1.0 : 1.0;
                            if (ax == F_AXIS_X)
                            {
                                // M * Rx: col1' = c col1 + s col2, col2' = -s col1 + c col2
                                const double a1 = cur.m1, a4 = cur.m4, a7 = cur.m7;
                                cur.m1 = cs * a1 + s * cur.m2; cur.m2 = cs * cur.m2 - s * a1;
                                cur.m4 = cs * a4 + s * cur.m5; cur.m5 = cs * cur.m5 - s * a4;
                                cur.m7 = cs * a7 + s * cur.m8; cur.m8 = cs * cur.m8 - s * a7;
                                wx = sg * cur.m0; wy = sg * cur.m3; wz = sg * cur.m6;
                            }
                            else if (ax == F_AXIS_Y)
                            {
                                // M * Ry: col0' = c col0 - s col2, col2' = s col0 + c col2
                                const double a0 = cur.m0, a3 = cur.m3, a6 = cur.m6;
                                cur.m0 = cs * a0 - s * cur.m2; cur.m2 = s * a0 + cs * cur.m2;
                                cur.m3 = cs * a3 - s * cur.m5; cur.m5 = s * a3 + cs * cur.m5;
                                cur.m6 = cs * a6 - s * cur.m8; cur.m8 = s * a6 + cs * cur.m8;
                                wx = sg * cur.m1; wy = sg * cur.m4; wz = sg * cur.m7;
                            }
                            else
                            {
                                // M * Rz: col0' = c col0 + s col1, col1' = -s col0 + c col1
                                const double a0 = cur.m0, a3 = cur.m3, a6 = cur.m6;
                                cur.m0 = cs * a0 + s * cur.m1; cur.m1 = cs * cur.m1 - s * a0;
                                cur.m3 = cs * a3 + s * cur.m4; cur.m4 = cs * cur.m4 - s * a3;
                                cur.m6 = cs * a6 + s * cur.m7; cur.m7 = cs * cur.m7 - s * a6;
                                wx = sg * cur.m2; wy = sg * cur.m5; wz = sg * cur.m8;
                            }
                        }
                    }
                    if (op.jslot >= 0)
                    {
                        double* js = scr + (joint_base + 6 * op.jslot) * TPB;
                        js[0] = wx; js[TPB] = wy; js[2 * TPB] = wz;
                        js[3 * TPB] = cur.x; js[4 * TPB] = cur.y; js[5 * TPB] = cur.z;
                    }
                }
                if (op.dst >= 0) fr_store_scratch<TPB>(scr + 12 * op.dst * TPB, cur);
            }
            else if (op.kind == OP_EMIT_FRAME)
            {
                // temp_A = frame_A * offA ; temp_B = frame_B * offB ; Phi = temp_B^-1 * temp_A  (kinematic_tree.cpp:829-835)
                Fr A;
                if (op.src == SRC_CUR) A = cur;
                else if (op.src == SRC_ROOT) A = fr_identity();
                else A = fr_load_scratch<TPB>(scr + 12 * op.src * TPB);
                const int offa = op.flags & F_OFFA_MASK;
                if (offa == F_OFFA_TRANS) fr_translate(A, c[CE_OFFA_P], c[CE_OFFA_P + 1], c[CE_OFFA_P + 2]);
                else if (offa == F_OFFA_FULL) A = fr_mul(A, fr_load_const(c + CE_OFFA_M));
                const int bk = op.flags & F_B_MASK;
                Fr Binv, Phi;
                if (bk == F_B_IDENTITY)
                    Phi = A;
                else
                {
                    if (bk == F_B_CONST)
                        Binv = fr_load_const(c + CE_B_M);
                    else
                    {
                        Fr B;
                        if (op.src_b == SRC_CUR) B = cur;
                        else B = fr_load_scratch<TPB>(scr + 12 * op.src_b * TPB);
                        B = fr_mul(B, fr_load_const(c + CE_B_M));
                        Binv = fr_inverse(B);
                    }
                    Phi = fr_mul(Binv, A);
                }
                double* ph = phi_e + (int64_t)op.phi_off * P.phi_sk;
                const int64_t sk = P.phi_sk;
                if (op.map_kind == EXO_MAP_FRAMES)
                {
                    ph[0] = Phi.m0; ph[sk] = Phi.m1; ph[2 * sk] = Phi.m2; ph[3 * sk] = Phi.m3; ph[4 * sk] = Phi.m4;
                    ph[5 * sk] = Phi.m5; ph[6 * sk] = Phi.m6; ph[7 * sk] = Phi.m7; ph[8 * sk] = Phi.m8;
                    ph[9 * sk] = Phi.x; ph[10 * sk] = Phi.y; ph[11 * sk] = Phi.z;
                }
                else
                {
                    ph[0] = Phi.x; ph[sk] = Phi.y; ph[2 * sk] = Phi.z;
                    if (op.map_kind == EXO_MAP_EFF_FRAME)
                    {
                        ph += 3 * sk;
                        switch (op.rot_type)
                        {
                            case EXO_ROT_RPY:
                            {
                                double r, p, y;
                                get_rpy(Phi, r, p, y);
                                ph[0] = r; ph[sk] = p; ph[2 * sk] = y;
                                break;
                            }
                            case EXO_ROT_ZYX:
                            {
                                double r, p, y;
                                get_rpy(Phi, r, p, y);  // GetEulerZYX(a,b,c) == GetRPY(c,b,a)
                                ph[0] = y; ph[sk] = p; ph[2 * sk] = r;
                                break;
                            }
                            case EXO_ROT_ZYZ:
                            {
                                double a, b, g;
                                get_zyz(Phi, a, b, g);
                                ph[0] = a; ph[sk] = b; ph[2 * sk] = g;
                                break;
                            }
                            case EXO_ROT_ANGLE_AXIS:
                            {
                                double a, b, g;
                                get_angle_axis(Phi, a, b, g);
                                ph[0] = a; ph[sk] = b; ph[2 * sk] = g;
                                break;
                            }
                            case EXO_ROT_QUATERNION:
                            {
                                double q[4];
                                get_quaternion(Phi, q);
                                ph[0] = q[0]; ph[sk] = q[1]; ph[2 * sk] = q[2]; ph[3 * sk] = q[3];
                                break;
                            }
                            default:  // EXO_ROT_MATRIX
                                ph[0] = Phi.m0; ph[sk] = Phi.m1; ph[2 * sk] = Phi.m2; ph[3 * sk] = Phi.m3; ph[4 * sk] = Phi.m4;
                                ph[5 * sk] = Phi.m5; ph[6 * sk] = Phi.m6; ph[7 * sk] = Phi.m7; ph[8 * sk] = Phi.m8;
                                break;
                        }
                    }
                }
                if (want_j)
                {
                    // ComputeJ (kinematic_tree.cpp:919-945): column c = temp_B.M^-1 * [ w x (temp_A.p - p_joint) ; w ]
                    const int32_t* tab = ints + op.tab;
                    double* jr = jac_e + (int64_t)op.jrow_off * P.j_sr;
                    const int64_t sr = P.j_sr;
                    const bool three = op.n_rows == 3;
                    for (int col = 0; col < N; ++col)
                    {
                        const int ent = tab[col];
                        double v0 = 0.0, v1 = 0.0, v2 = 0.0, w0 = 0.0, w1 = 0.0, w2 = 0.0;
                        if (ent >= 0)
                        {
                            const double* js = scr + (joint_base + 6 * (ent >> 2)) * TPB;
                            const double ax = js[0], ay = js[TPB], az = js[2 * TPB];
                            if (ent & 2)
                            {
                                v0 = ax; v1 = ay; v2 = az;
                            }
                            else
                            {
                                const double dx = A.x - js[3 * TPB], dy = A.y - js[4 * TPB], dz = A.z - js[5 * TPB];
                                v0 = ay * dz - az * dy;
                                v1 = az * dx - ax * dz;
                                v2 = ax * dy - ay * dx;
                                w0 = ax; w1 = ay; w2 = az;
                            }
                            if (bk != F_B_IDENTITY)
                            {
                                const double t0 = Binv.m0 * v0 + Binv.m1 * v1 + Binv.m2 * v2;
                                const double t1 = Binv.m3 * v0 + Binv.m4 * v1 + Binv.m5 * v2;
                                const double t2 = Binv.m6 * v0 + Binv.m7 * v1 + Binv.m8 * v2;
                                v0 = t0; v1 = t1; v2 = t2;
                                const double u0 = Binv.m0 * w0 + Binv.m1 * w1 + Binv.m2 * w2;
                                const double u1 = Binv.m3 * w0 + Binv.m4 * w1 + Binv.m5 * w2;
                                const double u2 = Binv.m6 * w0 + Binv.m7 * w1 + Binv.m8 * w2;
                                w0 = u0; w1 = u1; w2 = u2;
                            }
                            if (ent & 1)
                            {
                                // frame-B chain: 0 - column (kinematic_tree.cpp:941)
                                v0 = 0.0 - v0; v1 = 0.0 - v1; v2 = 0.0 - v2;
                                w0 = 0.0 - w0; w1 = 0.0 - w1; w2 = 0.0 - w2;
                            }
                        }
                        double* jc = jr + (int64_t)col * P.j_sc;
                        jc[0] = v0; jc[sr] = v1; jc[2 * sr] = v2;
                        if (!three)
                        {
                            jc[3 * sr] = w0; jc[4 * sr] = w1; jc[5 * sr] = w2;
                        }
                    }
                }
            }
            else if (op.kind == OP_EMIT_JLIMIT)
            {
                // JointLimit::Update (joint_limit.cpp:63-99); phi starts from the SetZero value 0
                const double safe = c[0];
                const double* lo = c + 1;
                const double* hi = c + 1 + N;
                double* ph = phi_e + (int64_t)op.phi_off * P.phi_sk;
                double* jr = want_j ? jac_e + (int64_t)op.jrow_off * P.j_sr : nullptr;
                for (int i = 0; i < N; ++i)
                {
                    const double q = qs[i * TPB];
                    const double tau = (0.5 * safe) * (hi[i] - lo[i]);
                    double v = 0.0;
                    if (q < lo[i] + tau) v = q - lo[i] - tau;
                    if (q > hi[i] - tau) v = q - hi[i] + tau;
                    ph[i * P.phi_sk] = v;
                    if (want_j)
                    {
                        const double d = (q >= lo[i] + tau && q <= hi[i] - tau) ? 0.0 : 1.0;
                        double* row = jr + (int64_t)i * P.j_sr;
                        for (int col = 0; col < N; ++col) row[(int64_t)col * P.j_sc] = (col == i) ? d : 0.0;
                    }
                }
            }
            else if (op.kind == OP_EMIT_FILL)
            {
                double* ph = phi_e + (int64_t)op.phi_off * P.phi_sk;
                for (int i = 0; i < op.n_items; ++i) ph[i * P.phi_sk] = c[i];
                if (want_j)
                {
                    double* jr = jac_e + (int64_t)op.jrow_off * P.j_sr;
                    for (int r = 0; r < op.n_rows; ++r)
                        for (int col = 0; col < N; ++col) jr[(int64_t)r * P.j_sr + (int64_t)col * P.j_sc] = 0.0;
                }
            }
            else if (op.kind == OP_EMIT_SPHERES)
            {
                // SphereCollision::Update (sphere_collision.cpp:123-162).  The reference accumulates, per sphere pair
                // and per column, -(J_i - J_j)^T d / |d|; here the pair loop accumulates unit "forces" per carrying
                // link (F, M = p x F) and the Jacobian row is recovered per joint as  -w . (sum M - p_joint x sum F)
                // (same sum, regrouped: O(pairs + joints) instead of O(pairs * joints)).
                const int32_t* tab = ints + op.tab;
                const int G = tab[0], n_acc = tab[1];
                const int32_t* gs = tab + 2;
                const int32_t* sph = gs + G + 1;
                const int32_t* cols = sph + SPH_INTS * op.n_items;
                const double k5 = c[0];
                const double* sc = c + 1;  // per sphere: ox oy oz radius
                double* acc = scr + (size_t)acc_base * TPB;
                int out_idx = 0;
                for (int ga = 0; ga < G; ++ga)
                {
                    for (int gb = ga + 1; gb < G; ++gb, ++out_idx)
                    {
                        for (int a = 0; a < 6 * n_acc; ++a) acc[a * TPB] = 0.0;
                        double phi_sum = 0.0;
                        for (int i = gs[ga]; i < gs[ga + 1]; ++i)
                        {
                            const int src_i = sph[SPH_INTS * i], acc_i = sph[SPH_INTS * i + 1];
                            double pix = sc[4 * i], piy = sc[4 * i + 1], piz = sc[4 * i + 2];
                            const double ri = sc[4 * i + 3];
                            if (src_i != SRC_ROOT)
                            {
                                const Fr f = (src_i == SRC_CUR) ? cur : fr_load_scratch<TPB>(scr + 12 * src_i * TPB);
                                const double tx = f.m0 * pix + f.m1 * piy + f.m2 * piz + f.x;
                                const double ty = f.m3 * pix + f.m4 * piy + f.m5 * piz + f.y;
                                const double tz = f.m6 * pix + f.m7 * piy + f.m8 * piz + f.z;
                                pix = tx; piy = ty; piz = tz;
                            }
                            double gx = 0.0, gy = 0.0, gz = 0.0;
                            for (int j = gs[gb]; j < gs[gb + 1]; ++j)
                            {
                                const int src_j = sph[SPH_INTS * j], acc_j = sph[SPH_INTS * j + 1];
                                double pjx = sc[4 * j], pjy = sc[4 * j + 1], pjz = sc[4 * j + 2];
                                if (src_j != SRC_ROOT)
                                {
                                    const Fr f = (src_j == SRC_CUR) ? cur : fr_load_scratch<TPB>(scr + 12 * src_j * TPB);
                                    const double tx = f.m0 * pjx + f.m1 * pjy + f.m2 * pjz + f.x;
                                    const double ty = f.m3 * pjx + f.m4 * pjy + f.m5 * pjz + f.y;
                                    const double tz = f.m6 * pjx + f.m7 * pjy + f.m8 * pjz + f.z;
                                    pjx = tx; pjy = ty; pjz = tz;
                                }
                                const double dx = pix - pjx, dy = piy - pjy, dz = piz - pjz;
                                const double nrm = sqrt(dx * dx + dy * dy + dz * dz);
                                phi_sum += 1.0 / (1.0 + exp(k5 * (nrm - ri - sc[4 * j + 3])));
                                if (want_j)
                                {
                                    const double ux = dx / nrm, uy = dy / nrm, uz = dz / nrm;
                                    gx += ux; gy += uy; gz += uz;
                                    if (acc_j >= 0)
                                    {
                                        double* aj = acc + 6 * acc_j * TPB;
                                        aj[0] -= ux; aj[TPB] -= uy; aj[2 * TPB] -= uz;
                                        aj[3 * TPB] -= pjy * uz - pjz * uy;
                                        aj[4 * TPB] -= pjz * ux - pjx * uz;
                                        aj[5 * TPB] -= pjx * uy - pjy * ux;
                                    }
                                }
                            }
                            if (want_j && acc_i >= 0)
                            {
                                double* ai = acc + 6 * acc_i * TPB;
                                ai[0] += gx; ai[TPB] += gy; ai[2 * TPB] += gz;
                                ai[3 * TPB] += piy * gz - piz * gy;
                                ai[4 * TPB] += piz * gx - pix * gz;
                                ai[5 * TPB] += pix * gy - piy * gx;
                            }
                        }
                        phi_e[(int64_t)(op.phi_off + out_idx) * P.phi_sk] = phi_sum;
                        if (want_j)
                        {
                            double* row = jac_e + (int64_t)(op.jrow_off + out_idx) * P.j_sr;
                            for (int col = 0; col < N; ++col)
                            {
                                const int ent = cols[3 * col];
                                double val = 0.0;
                                if (ent >= 0)
                                {
                                    const uint32_t mlo = (uint32_t)cols[3 * col + 1], mhi = (uint32_t)cols[3 * col + 2];
                                    double fx = 0, fy = 0, fz = 0, mx = 0, my = 0, mz = 0;
                                    for (int a = 0; a < n_acc; ++a)
                                    {
                                        const bool in = a < 32 ? ((mlo >> a) & 1u) : ((mhi >> (a - 32)) & 1u);
                                        if (!in) continue;
                                        const double* aa = acc + 6 * a * TPB;
                                        fx += aa[0]; fy += aa[TPB]; fz += aa[2 * TPB];
                                        mx += aa[3 * TPB]; my += aa[4 * TPB]; mz += aa[5 * TPB];
                                    }
                                    const double* js = scr + (joint_base + 6 * (ent >> 2)) * TPB;
                                    const double ax = js[0], ay = js[TPB], az = js[2 * TPB];
                                    if (ent & 2)
                                        val = -(ax * fx + ay * fy + az * fz);
                                    else
                                    {
                                        const double px = js[3 * TPB], py = js[4 * TPB], pz = js[5 * TPB];
                                        const double tx = mx - (py * fz - pz * fy);
                                        const double ty = my - (pz * fx - px * fz);
                                        const double tz = mz - (px * fy - py * fx);
                                        val = -(ax * tx + ay * ty + az * tz);
                                    }
                                }
                                row[(int64_t)col * P.j_sc] = val;
                            }
                        }
                    }
                }
            }
        }
    }
}
}  // namespace

template <int TPB>
cudaError_t try_config(size_t prog, size_t per_thread, int max_smem, int& ctas_per_sm, size_t& smem_bytes)
{
    smem_bytes = prog + per_thread * TPB;
    ctas_per_sm = 0;
    if (smem_bytes > (size_t)max_smem) return cudaSuccess;
    cudaError_t err = cudaFuncSetAttribute(exo_eval_kernel<TPB>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
    if (err != cudaSuccess) return err;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, exo_eval_kernel<TPB>, TPB, smem_bytes);
}

cudaError_t plan_launch(const ProgramHeader& hdr, int program_bytes, int64_t n_evals, LaunchConfig& cfg, KernelParams& params)
{
    int dev = 0, sms = 0, max_smem = 0;
    cudaError_t err = cudaGetDevice(&dev);
    if (err != cudaSuccess) return err;
    err = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (err != cudaSuccess) return err;
    err = cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (err != cudaSuccess) return err;
    const size_t prog = ((size_t)program_bytes + 15) & ~(size_t)15;
    const size_t per_thread = (size_t)(hdr.scratch_doubles + hdr.n_ctrl) * sizeof(double);
    params.scratch_offset = (int32_t)prog;
    // pick the CTA size that keeps the most threads resident per SM (per-thread scratch lives in shared memory)
    int best_threads = 0, best_tpb = 0, best_ctas = 0;
    size_t best_smem = 0;
    const int sizes[3] = {128, 64, 32};
    for (int k = 0; k < 3; ++k)
    {
        int ctas = 0;
        size_t smem_bytes = 0;
        if (sizes[k] == 128) err = try_config<128>(prog, per_thread, max_smem, ctas, smem_bytes);
        else if (sizes[k] == 64) err = try_config<64>(prog, per_thread, max_smem, ctas, smem_bytes);
        else err = try_config<32>(prog, per_thread, max_smem, ctas, smem_bytes);
        if (err != cudaSuccess) return err;
        if (ctas * sizes[k] > best_threads)
        {
            best_threads = ctas * sizes[k];
            best_tpb = sizes[k];
            best_ctas = ctas;
            best_smem = smem_bytes;
        }
    }
    if (best_threads == 0) return cudaErrorInvalidConfiguration;
    cfg.threads = best_tpb;
    cfg.smem_bytes = best_smem;
    const int64_t want = (n_evals + best_tpb - 1) / best_tpb;
    const int64_t resident = (int64_t)sms * best_ctas;
    cfg.blocks = (int)(want < resident ? (want > 0 ? want : 1) : resident);
    return cudaSuccess;
}

cudaError_t launch_eval(const LaunchConfig& cfg, const KernelParams& params, cudaStream_t stream)
{
    if (cfg.threads == 128) exo_eval_kernel<128><<<cfg.blocks, 128, cfg.smem_bytes, stream>>>(params);
    else if (cfg.threads == 64) exo_eval_kernel<64><<<cfg.blocks, 64, cfg.smem_bytes, stream>>>(params);
    else exo_eval_kernel<32><<<cfg.blocks, 32, cfg.smem_bytes, stream>>>(params);
    return cudaGetLastError();
}

}  // namespace exo
