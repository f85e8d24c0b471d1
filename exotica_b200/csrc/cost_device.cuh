// cost_device.cuh -- device helpers shared by the cost kernel and the batched IK kernels: exotica::GetRotation,
// TaskSpaceVector::operator- (ydiff) and the scalar task cost of one evaluation.
#pragma once
#include "device_math.cuh"
#include "kernels.cuh"

namespace exo
{
namespace dev
{
__device__ __forceinline__ double warp_sum_d(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ double warp_min_d(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// sin / cos of an Euler angle read back from Phi: the kernels' own reduction (device_math.cuh) inside its range -- Phi holds atan2
// results, so always --, CUDA's beyond it
__device__ __forceinline__ void euler_sincos(double a, double& s, double& c)
{
    if (fabs(a) <= 1.0e5)
        fast_sincos(a, s, c);
    else
        sincos(a, &s, &c);
}

// exotica::GetRotation (tools/conversions.cpp:114-148) over KDL::Rotation::{Quaternion, RPY, EulerZYX, EulerZYZ, Rot}
static __device__ Fr dev_get_rotation(const double* d, int64_t sk, int type)
{
    Fr r = fr_identity();
    switch (type)
    {
        case EXO_ROT_QUATERNION:
        {
            const double x = d[0], y = d[sk], z = d[2 * sk], w = d[3 * sk];
            const double x2 = x * x, y2 = y * y, z2 = z * z, w2 = w * w;
            r.m0 = w2 + x2 - y2 - z2; r.m1 = 2 * x * y - 2 * w * z; r.m2 = 2 * x * z + 2 * w * y;
            r.m3 = 2 * x * y + 2 * w * z; r.m4 = w2 - x2 + y2 - z2; r.m5 = 2 * y * z - 2 * w * x;
            r.m6 = 2 * x * z - 2 * w * y; r.m7 = 2 * y * z + 2 * w * x; r.m8 = w2 - x2 - y2 + z2;
            break;
        }
        case EXO_ROT_RPY:
        case EXO_ROT_ZYX:
        {
            const bool zyx = type == EXO_ROT_ZYX;  // EulerZYX(a, b, c) = RPY(c, b, a)
            const double roll = zyx ? d[2 * sk] : d[0], pitch = d[sk], yaw = zyx ? d[0] : d[2 * sk];
            double sa, ca, sb, cb, sc, cc;
            euler_sincos(yaw, sa, ca);
            euler_sincos(pitch, sb, cb);
            euler_sincos(roll, sc, cc);
            r.m0 = ca * cb; r.m1 = ca * sb * sc - sa * cc; r.m2 = ca * sb * cc + sa * sc;
            r.m3 = sa * cb; r.m4 = sa * sb * sc + ca * cc; r.m5 = sa * sb * cc - ca * sc;
            r.m6 = -sb; r.m7 = cb * sc; r.m8 = cb * cc;
            break;
        }
        case EXO_ROT_ZYZ:
        {
            double sa, ca, sb, cb, sg, cg;
            euler_sincos(d[0], sa, ca);
            euler_sincos(d[sk], sb, cb);
            euler_sincos(d[2 * sk], sg, cg);
            r.m0 = ca * cb * cg - sa * sg; r.m1 = -ca * cb * sg - sa * cg; r.m2 = ca * sb;
            r.m3 = sa * cb * cg + ca * sg; r.m4 = -sa * cb * sg + ca * cg; r.m5 = sa * sb;
            r.m6 = -sb * cg; r.m7 = sb * sg; r.m8 = cb;
            break;
        }
        case EXO_ROT_ANGLE_AXIS:
        {
            const double ax = d[0], ay = d[sk], az = d[2 * sk];
            const double angle = norm3_scaled(ax, ay, az);
            if (fabs(angle) > 1e-10)
            {
                const double ux = ax / angle, uy = ay / angle, uz = az / angle;
                double st, ct;
                sincos(angle, &st, &ct);
                const double vt = 1 - ct;
                r.m0 = ct + vt * ux * ux; r.m1 = -uz * st + vt * ux * uy; r.m2 = uy * st + vt * ux * uz;
                r.m3 = uz * st + vt * ux * uy; r.m4 = ct + vt * uy * uy; r.m5 = -ux * st + vt * uy * uz;
                r.m6 = -uy * st + vt * ux * uz; r.m7 = ux * st + vt * uy * uz; r.m8 = ct + vt * uz * uz;
            }
            break;
        }
        default:  // EXO_ROT_MATRIX
            r.m0 = d[0]; r.m1 = d[sk]; r.m2 = d[2 * sk]; r.m3 = d[3 * sk]; r.m4 = d[4 * sk];
            r.m5 = d[5 * sk]; r.m6 = d[6 * sk]; r.m7 = d[7 * sk]; r.m8 = d[8 * sk];
            break;
    }
    return r;
}

// ct * ( ydiff^T S ydiff + xdiff^T W xdiff ) of evaluation (b, t)
static __device__ double eval_cost(const CostParams& P, int64_t e, int t)
{
    const double* ph = P.phi + e * P.phi_se;
    const int tg = P.T_goal > 1 ? t : 0;
    const double* goal = P.goal + (int64_t)tg * P.length_phi;
    const double* grot = P.goal_rot + (int64_t)tg * P.n_rot * 9;
    const double* rho = P.rho + (int64_t)t * P.n_maps;
    double c = 0.0;
    int rot_i = 0;
    for (int k = 0; k < P.n_maps; ++k)
    {
        const CostMapP& m = P.map[k];
        const double w = rho[k];
        const bool lie = m.kind == EXO_MAP_EFF_FRAME || m.kind == EXO_MAP_EFF_ORIENTATION;
        if (w == 0.0)
        {
            if (lie) rot_i += m.n_frames;
            continue;
        }
        double s = 0.0;
        if (lie)
        {
            const int lead = m.kind == EXO_MAP_EFF_FRAME ? 3 : 0;  // EffOrientation has no position entries
            for (int f = 0; f < m.n_frames; ++f, ++rot_i)
            {
                const double* a = ph + (int64_t)(m.start + f * m.stride) * P.phi_sk;
                const double* g = goal + m.start + f * m.stride;
                for (int i = 0; i < lead; ++i)
                {
                    const double d = a[i * P.phi_sk] - g[i];
                    s += d * d;
                }
                // rotvec = M1 * (M2^-1 M1).GetRot()
                const Fr M1 = dev_get_rotation(a + lead * P.phi_sk, P.phi_sk, m.rot_type);
                const double* M2 = grot + 9 * rot_i;
                Fr M;
                M.m0 = M2[0] * M1.m0 + M2[3] * M1.m3 + M2[6] * M1.m6;
                M.m1 = M2[0] * M1.m1 + M2[3] * M1.m4 + M2[6] * M1.m7;
                M.m2 = M2[0] * M1.m2 + M2[3] * M1.m5 + M2[6] * M1.m8;
                M.m3 = M2[1] * M1.m0 + M2[4] * M1.m3 + M2[7] * M1.m6;
                M.m4 = M2[1] * M1.m1 + M2[4] * M1.m4 + M2[7] * M1.m7;
                M.m5 = M2[1] * M1.m2 + M2[4] * M1.m5 + M2[7] * M1.m8;
                M.m6 = M2[2] * M1.m0 + M2[5] * M1.m3 + M2[8] * M1.m6;
                M.m7 = M2[2] * M1.m1 + M2[5] * M1.m4 + M2[8] * M1.m7;
                M.m8 = M2[2] * M1.m2 + M2[5] * M1.m5 + M2[8] * M1.m8;
                M.x = M.y = M.z = 0.0;
                double vx, vy, vz;
                get_angle_axis(M, vx, vy, vz);
                const double r0 = M1.m0 * vx + M1.m1 * vy + M1.m2 * vz;
                const double r1 = M1.m3 * vx + M1.m4 * vy + M1.m5 * vz;
                const double r2 = M1.m6 * vx + M1.m7 * vy + M1.m8 * vz;
                s += r0 * r0 + r1 * r1 + r2 * r2;
            }
        }
        else
        {
            const double* a = ph + (int64_t)m.start * P.phi_sk;
            const double* g = goal + m.start;
            for (int i = 0; i < m.length; ++i)
            {
                const double d = a[i * P.phi_sk] - g[i];
                s += d * d;
            }
        }
        c += w * s;
    }
    if (P.W && t > 0)
    {
        const double* x1 = P.x + e * P.x_se;
        const double* x0 = x1 - P.x_se;  // previous timestep of the same trajectory
        for (int i = 0; i < P.n_ctrl; ++i)
        {
            const double d = x1[i * P.x_sk] - x0[i * P.x_sk];
            c += P.W[i] * d * d;
        }
    }
    return c * P.ct;
}


// ydiff = Phi - y as TaskSpaceVector::operator- defines it (task_space_vector.cpp:64-91): out[length_jacobian], entry k at
// out[k * so]; rows of maps listed in CostParams keep their Jacobian indexing (start_jacobian).
static __device__ void eval_ydiff(const CostParams& P, int64_t e, int t, double* out, int so)
{
    const double* ph = P.phi + e * P.phi_se;
    const int tg = P.T_goal > 1 ? t : 0;
    const double* goal = P.goal + (int64_t)tg * P.length_phi;
    const double* grot = P.goal_rot + (int64_t)tg * P.n_rot * 9;
    int rot_i = 0;
    for (int k = 0; k < P.n_maps; ++k)
    {
        const CostMapP& m = P.map[k];
        const bool lie = m.kind == EXO_MAP_EFF_FRAME || m.kind == EXO_MAP_EFF_ORIENTATION;
        if (lie)
        {
            const int lead = m.kind == EXO_MAP_EFF_FRAME ? 3 : 0;
            for (int f = 0; f < m.n_frames; ++f, ++rot_i)
            {
                const double* a = ph + (int64_t)(m.start + f * m.stride) * P.phi_sk;
                const double* g = goal + m.start + f * m.stride;
                double* o = out + (m.start_jacobian + f * (lead + 3)) * so;
                for (int i = 0; i < lead; ++i) o[i * so] = a[i * P.phi_sk] - g[i];
                const Fr M1 = dev_get_rotation(a + lead * P.phi_sk, P.phi_sk, m.rot_type);
                const double* M2 = grot + 9 * rot_i;
                Fr M;
                M.m0 = M2[0] * M1.m0 + M2[3] * M1.m3 + M2[6] * M1.m6;
                M.m1 = M2[0] * M1.m1 + M2[3] * M1.m4 + M2[6] * M1.m7;
                M.m2 = M2[0] * M1.m2 + M2[3] * M1.m5 + M2[6] * M1.m8;
                M.m3 = M2[1] * M1.m0 + M2[4] * M1.m3 + M2[7] * M1.m6;
                M.m4 = M2[1] * M1.m1 + M2[4] * M1.m4 + M2[7] * M1.m7;
                M.m5 = M2[1] * M1.m2 + M2[4] * M1.m5 + M2[7] * M1.m8;
                M.m6 = M2[2] * M1.m0 + M2[5] * M1.m3 + M2[8] * M1.m6;
                M.m7 = M2[2] * M1.m1 + M2[5] * M1.m4 + M2[8] * M1.m7;
                M.m8 = M2[2] * M1.m2 + M2[5] * M1.m5 + M2[8] * M1.m8;
                M.x = M.y = M.z = 0.0;
                double vx, vy, vz;
                get_angle_axis(M, vx, vy, vz);
                o[lead * so] = M1.m0 * vx + M1.m1 * vy + M1.m2 * vz;
                o[(lead + 1) * so] = M1.m3 * vx + M1.m4 * vy + M1.m5 * vz;
                o[(lead + 2) * so] = M1.m6 * vx + M1.m7 * vy + M1.m8 * vz;
            }
        }
        else
        {
            const double* a = ph + (int64_t)m.start * P.phi_sk;
            const double* g = goal + m.start;
            for (int i = 0; i < m.length; ++i) out[(m.start_jacobian + i) * so] = a[i * P.phi_sk] - g[i];
        }
    }
}
}  // namespace dev
}  // namespace exo
