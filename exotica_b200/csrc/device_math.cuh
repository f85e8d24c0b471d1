// device_math.cuh -- register-resident FP64 frame algebra and rotation encodings shared by the CUDA kernels.
// Restates (for the device) the orocos_kdl / Eigen formulas behind KinematicTree::FK/ComputeJ and exotica::SetRotation
// (exotica_core/src/kinematic_tree.cpp:829-835,919-945; exotica_core/src/tools/conversions.cpp:150-179).
#pragma once
#include <math_constants.h>
#include <stdint.h>

#include "../../include/exotica_b200.h"
#include "program.h"

namespace exo
{
namespace dev
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

// ---- sin/cos ---------------------------------------------------------------------------------------------------------
// Cody-Waite reduction by pi/2 (three FMAs, exact products) + fdlibm's __kernel_sin/__kernel_cos minimax polynomials on
// [-pi/4, pi/4] (< 1 ulp).  Coefficients sit in __constant__ memory so that they reach the FP64 pipe through uniform
// registers (one 128-bit constant load per pair) instead of two 32-bit immediates per double.
__constant__ double kTrig[16] = {
    6.36619772367581382433e-01,   // 0: 2/pi
    1.57079632679489655800e+00,   // 1: pi/2 rounded to double
    6.123233995736757e-17,        // 2: 0x3c91a62633145c00, leading bits of pi/2 - [1]
    8.478427660368898e-32,        // 3: 0x397b839a252049c0, pi/2 - [1] - [2]
    -1.66666666666666324348e-01,  // 4: S1
    8.33333333332248946124e-03,   // 5: S2
    -1.98412698298579493134e-04,  // 6: S3
    2.75573137070700676789e-06,   // 7: S4
    -2.50507602534068634195e-08,  // 8: S5
    1.58969099521155010221e-10,   // 9: S6
    4.16666666666666019037e-02,   // 10: C1
    -1.38888888888741095749e-03,  // 11: C2
    2.48015872894767294178e-05,   // 12: C3
    -2.75573143513906633035e-07,  // 13: C4
    2.08757232129817482790e-09,   // 14: C5
    -1.13596475577881948265e-11   // 15: C6
};

// Valid for |x| <= 1e5 (the reduction is exact there); callers route larger arguments to CUDA's sincos.
__device__ __forceinline__ void fast_sincos(double x, double& s, double& c)
{

    const double magic = 6755399441055744.0;  // 1.5 * 2^52: adding it rounds to the nearest integer
    const double t = fma(x, kTrig[0], magic);
    const int k = __double2loint(t);
    const double kd = t - magic;
    double r = fma(-kd, kTrig[1], x);
    r = fma(-kd, kTrig[2], r);
    r = fma(-kd, kTrig[3], r);
    const double z = r * r;
    double ps = fma(z, kTrig[9], kTrig[8]);
    double pc = fma(z, kTrig[15], kTrig[14]);
    ps = fma(z, ps, kTrig[7]);
    pc = fma(z, pc, kTrig[13]);
    ps = fma(z, ps, kTrig[6]);
    pc = fma(z, pc, kTrig[12]);
    ps = fma(z, ps, kTrig[5]);
    pc = fma(z, pc, kTrig[11]);
    ps = fma(z, ps, kTrig[4]);
    pc = fma(z, pc, kTrig[10]);
    const double sr = fma(z * r, ps, r);
    const double cr = fma(z * z, pc, fma(z, -0.5, 1.0));
    // quadrant k mod 4: (sin, cos) = (sr, cr), (cr, -sr), (-sr, -cr), (-cr, sr)
    const bool swap = (k & 1) != 0;
    double ss = swap ? cr : sr;
    double cc = swap ? sr : cr;
    // flip signs through the high word: sin negative for k mod 4 in {2,3}, cos negative for {1,2}
    const int sflip = (k & 2) << 30;
    const int cflip = ((k + 1) & 2) << 30;
    s = __hiloint2double(__double2hiint(ss) ^ sflip, __double2loint(ss));
    c = __hiloint2double(__double2hiint(cc) ^ cflip, __double2loint(cc));
}

// ---- atan2 ---------------------------------------------------------------------------------------------------------------
// atan2 for the Euler-angle encodings (KDL::Rotation::GetRPY / GetEulerZYZ call std::atan2).  One division, octant
// reduction to |u| <= tan(pi/8) via atan(t) = pi/4 + atan((t-1)/(t+1)), fdlibm's atan minimax kernel (valid to 7/16,
// < 1 ulp), coefficients in __constant__ memory.  Signed zeros are honoured (atan2(+0,-0) = pi) because products of
// rotation-matrix entries do produce -0.0.  Absolute error ~1e-16; inputs are rotation-matrix entries (no
// overflow/denormal handling beyond the 0/0 case).
__constant__ double kAtan[12] = {
    3.33333333333329318027e-01,  -1.99999999998764832476e-01, 1.42857142725034663711e-01,  -1.11111104054623557880e-01,
    9.09088713343650656196e-02,  -7.69187620504482999495e-02, 6.66107313738753120669e-02,  -5.83357013379057348645e-02,
    4.97687799461593236017e-02,  -3.65315727442169155270e-02, 1.62858201153657823623e-02,  0.0};

__device__ __forceinline__ double fast_div(double num, double den)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(den));
    double e = fma(-den, r, 1.0);
    r = fma(r, e, r);
    e = fma(-den, r, 1.0);
    r = fma(r, e, r);
    const double q = num * r;
    const double rem = fma(-den, q, num);
    return fma(rem, r, q);
}

__device__ __forceinline__ double fast_atan2(double y, double x)
{
    const double ax = fabs(x), ay = fabs(y);
    const double mx = fmax(ax, ay), mn = fmin(ax, ay);
    const bool upper = mn > 0.41421356237309503 * mx;  // octant beyond tan(pi/8)
    const double num = upper ? mn - mx : mn;
    double den = upper ? mn + mx : mx;
    den = (den == 0.0) ? 1.0 : den;  // atan2(0, 0): u = 0
    const double u = fast_div(num, den);
    const double z = u * u, w = z * z;
    double s1 = fma(w, kAtan[10], kAtan[8]);
    double s2 = fma(w, kAtan[9], kAtan[7]);
    s1 = fma(w, s1, kAtan[6]);
    s2 = fma(w, s2, kAtan[5]);
    s1 = fma(w, s1, kAtan[4]);
    s2 = fma(w, s2, kAtan[3]);
    s1 = fma(w, s1, kAtan[2]);
    s2 = fma(w, s2, kAtan[1]);
    s1 = fma(w, s1, kAtan[0]);
    const double poly = fma(z, s1, w * s2);
    double a = fma(-u, poly, u);                       // atan(u)
    a = upper ? a + 7.85398163397448279e-01 : a;       // + pi/4
    a = (ay > ax) ? 1.57079632679489656e+00 - a : a;   // reflect about the diagonal
    a = (__double2hiint(x) < 0) ? 3.14159265358979312e+00 - a : a;
    return __hiloint2double((__double2hiint(a) & 0x7fffffff) | (__double2hiint(y) & 0x80000000), __double2loint(a));
}

// ---- rotation encodings: exotica::SetRotation (exotica_core/src/tools/conversions.cpp:150-179) over KDL / Eigen ----
__device__ __forceinline__ void get_rpy(const Fr& f, double& roll, double& pitch, double& yaw)
{
    // KDL::Rotation::GetRPY, including the gimbal-lock branch at |pitch| > pi/2 - 1e-12.
    // Written with three atan2 call sites (arguments selected first) to keep the code small.
    pitch = fast_atan2(-f.m6, sqrt(f.m0 * f.m0 + f.m3 * f.m3));
    const bool lock = fabs(pitch) > (CUDART_PIO2 - 1e-12);
    const double r = fast_atan2(f.m7, f.m8);
    yaw = fast_atan2(lock ? -f.m1 : f.m3, lock ? f.m4 : f.m0);
    roll = lock ? 0.0 : r;
}

__device__ __forceinline__ void get_zyz(const Fr& f, double& alpha, double& beta, double& gamma)
{
    // KDL::Rotation::GetEulerZYZ with its |R22| > 1 - 1e-12 branch, three atan2 call sites
    const bool sing = fabs(f.m8) > 1 - 1e-12;
    const bool pos = f.m8 > 0;
    const double ay = sing ? (pos ? f.m3 : -f.m3) : f.m5;
    const double ax = sing ? (pos ? f.m0 : -f.m0) : f.m2;
    alpha = fast_atan2(ay, ax);
    const double b = fast_atan2(sqrt(f.m6 * f.m6 + f.m7 * f.m7), f.m8);
    const double g = fast_atan2(f.m7, -f.m6);
    beta = sing ? (pos ? 0.0 : CUDART_PI) : b;
    gamma = sing ? 0.0 : g;
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

// The two encoders below are cold (one of six rotation types each) and stay out of line; their arguments and results travel in
// registers (nine doubles by value in, a small struct by value out) -- a `const Fr&` / pointer interface made every caller spill a
// 96-byte argument block to local memory in front of the rotation-type switch, whatever type it then took.
#ifdef EXO_INLINE_ENCODERS  // the run-time generated kernels call them from straight-line code with many live registers: a call
#define EXO_ENCODER_LINKAGE __forceinline__  // would save and restore most of them, so there they are expanded in place
#else
#define EXO_ENCODER_LINKAGE __noinline__
#endif
struct Rot3
{
    double x, y, z;
};
struct Rot4
{
    double x, y, z, w;
};
struct Mat9
{
    double m0, m1, m2, m3, m4, m5, m6, m7, m8;
};
static __device__ EXO_ENCODER_LINKAGE Rot3 angle_axis_of(double m0, double m1, double m2, double m3, double m4, double m5, double m6, double m7, double m8)
{
    const Mat9 f = {m0, m1, m2, m3, m4, m5, m6, m7, m8};
    double ox, oy, oz;
    // KDL::Rotation::GetRot = axis * angle with GetRotAngle's symmetric-matrix handling
    const double eps = 1e-6, eps2 = 1e-5;
    double x, y, z, angle;
    if ((fabs(f.m1 - f.m3) < eps) && (fabs(f.m2 - f.m6) < eps) && (fabs(f.m5 - f.m7) < eps))
    {
        if ((fabs(f.m1 + f.m3) < eps2) && (fabs(f.m2 + f.m6) < eps2) && (fabs(f.m5 + f.m7) < eps2) && (fabs(f.m0 + f.m4 + f.m8 - 3) < eps2))
        {
            const Rot3 zero = {0.0, 0.0, 0.0};
            return zero;
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
    const Rot3 out = {ox, oy, oz};
    return out;
}
__device__ __forceinline__ void get_angle_axis(const Fr& f, double& ox, double& oy, double& oz)
{
    const Rot3 r = angle_axis_of(f.m0, f.m1, f.m2, f.m3, f.m4, f.m5, f.m6, f.m7, f.m8);
    ox = r.x; oy = r.y; oz = r.z;
}

static __device__ EXO_ENCODER_LINKAGE Rot4 quaternion_of(double m0, double m1, double m2, double m3, double m4, double m5, double m6, double m7, double m8)
{
    // Eigen::Quaterniond(Matrix3d).coeffs()
    const double m[9] = {m0, m1, m2, m3, m4, m5, m6, m7, m8};
    double q[4];
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
    const Rot4 out = {q[0], q[1], q[2], q[3]};
    return out;
}
__device__ __forceinline__ void get_quaternion(const Fr& f, double* q /* x y z w */)
{
    const Rot4 r = quaternion_of(f.m0, f.m1, f.m2, f.m3, f.m4, f.m5, f.m6, f.m7, f.m8);
    q[0] = r.x; q[1] = r.y; q[2] = r.z; q[3] = r.w;
}


// Output stores: streaming (st.global.cs) when EXO_STORE_CS is defined -- Phi / J are written once and not re-read by the kernel
#ifdef EXO_STORE_CS
#define EXO_ST(ptr, v) __stcs((ptr), (v))
#else
#define EXO_ST(ptr, v) (*(ptr) = (v))
#endif

#ifdef EXO_PHI_STREAMING
#define EXO_ST_PHI(ptr, v) __stcs((ptr), (v))
#else
#define EXO_ST_PHI(ptr, v) (*(ptr) = (v))
#endif

// Rows of one Jacobian column the map keeps (program.h RowMode); (v, w) = linear / angular part of the frame's column,
// (ux, uy, uz) = Phi.p / |Phi.p| for ROWS_DIST (distance.cpp:52), (Phi.M a) x direction for ROWS_WDOT
// (eff_axis_alignment.cpp:140-144), ux = mass share for the accumulating CenterOfMass modes (center_of_mass.cpp:70)
__device__ __forceinline__ void store_rows(int mode, double* jc, int64_t sr, double v0, double v1, double v2, double w0, double w1, double w2,
                                           double ux, double uy, double uz)
{
    switch (mode)
    {
        case ROWS_VW:
            EXO_ST(&jc[0], v0); EXO_ST(&jc[sr], v1); EXO_ST(&jc[2 * sr], v2); EXO_ST(&jc[3 * sr], w0); EXO_ST(&jc[4 * sr], w1); EXO_ST(&jc[5 * sr], w2);
            break;
        case ROWS_V:
            EXO_ST(&jc[0], v0); EXO_ST(&jc[sr], v1); EXO_ST(&jc[2 * sr], v2);
            break;
        case ROWS_W:
            EXO_ST(&jc[0], w0); EXO_ST(&jc[sr], w1); EXO_ST(&jc[2 * sr], w2);
            break;
        case ROWS_VXY:
            EXO_ST(&jc[0], v0); EXO_ST(&jc[sr], v1);
            break;
        case ROWS_DIST:
            EXO_ST(&jc[0], ux * v0 + uy * v1 + uz * v2);
            break;
        case ROWS_WDOT:  // EffAxisAlignment: direction . (w x u) = w . (u x direction); (ux, uy, uz) = u x direction
            EXO_ST(&jc[0], ux * w0 + uy * w1 + uz * w2);
            break;
        case ROWS_V_ACC:  // CenterOfMass: ux = mass_i / M
            jc[0] += ux * v0; jc[sr] += ux * v1; jc[2 * sr] += ux * v2;
            break;
        default:  // ROWS_VXY_ACC
            jc[0] += ux * v0; jc[sr] += ux * v1;
            break;
    }
}

// JointPose::Update (joint_pose.cpp:36-57); the reference writes only J(i, map_i) = 1 into the caller-zeroed matrix
__device__ __forceinline__ void emit_joint_pose(const double* ref, const int32_t* map, int n, int N, const double* qs, int qstride, double* ph,
                                                int64_t phi_sk, double* jr, int64_t j_sr, int64_t j_sc)
{
    for (int i = 0; i < n; ++i)
    {
        ph[i * phi_sk] = qs[map[i] * qstride] - ref[i];
        if (jr)
        {
            double* row = jr + (int64_t)i * j_sr;
            for (int col = 0; col < N; ++col) row[(int64_t)col * j_sc] = (col == map[i]) ? 1.0 : 0.0;
        }
    }
}

// Phi entries of one requested frame for EffFrame / EffPosition / raw frames (eff_frame.cpp:63-73, eff_position.cpp:46-55)
__device__ __forceinline__ void write_phi(int map_kind, int rot_type, const Fr& Phi, double* ph, int64_t sk)
{
    if (map_kind == EXO_MAP_FRAMES)
    {
        EXO_ST_PHI(&ph[0], Phi.m0); EXO_ST_PHI(&ph[sk], Phi.m1); EXO_ST_PHI(&ph[2 * sk], Phi.m2); EXO_ST_PHI(&ph[3 * sk], Phi.m3); EXO_ST_PHI(&ph[4 * sk], Phi.m4);
        EXO_ST_PHI(&ph[5 * sk], Phi.m5); EXO_ST_PHI(&ph[6 * sk], Phi.m6); EXO_ST_PHI(&ph[7 * sk], Phi.m7); EXO_ST_PHI(&ph[8 * sk], Phi.m8);
        EXO_ST_PHI(&ph[9 * sk], Phi.x); EXO_ST_PHI(&ph[10 * sk], Phi.y); EXO_ST_PHI(&ph[11 * sk], Phi.z);
        return;
    }
    if (map_kind == EXO_MAP_DISTANCE)
    {
        EXO_ST_PHI(&ph[0], norm3_scaled(Phi.x, Phi.y, Phi.z));  // distance.cpp:40: Phi.p.Norm()
        return;
    }
    if (map_kind == EXO_MAP_EFF_POSITION_XY)
    {
        EXO_ST_PHI(&ph[0], Phi.x); EXO_ST_PHI(&ph[sk], Phi.y);
        return;
    }
    if (map_kind != EXO_MAP_EFF_ORIENTATION)
    {
        EXO_ST_PHI(&ph[0], Phi.x); EXO_ST_PHI(&ph[sk], Phi.y); EXO_ST_PHI(&ph[2 * sk], Phi.z);
        if (map_kind != EXO_MAP_EFF_FRAME) return;
        ph += 3 * sk;
    }
    switch (rot_type)
    {
        case EXO_ROT_RPY:
        case EXO_ROT_ZYX:
        {
            double r, p, y;
            get_rpy(Phi, r, p, y);  // GetEulerZYX(a,b,c) == GetRPY(c,b,a)
            const bool zyx = rot_type == EXO_ROT_ZYX;
            EXO_ST_PHI(&ph[0], zyx ? y : r); EXO_ST_PHI(&ph[sk], p); EXO_ST_PHI(&ph[2 * sk], zyx ? r : y);
            break;
        }
        case EXO_ROT_ZYZ:
        {
            double a, b, g;
            get_zyz(Phi, a, b, g);
            EXO_ST_PHI(&ph[0], a); EXO_ST_PHI(&ph[sk], b); EXO_ST_PHI(&ph[2 * sk], g);
            break;
        }
        case EXO_ROT_ANGLE_AXIS:
        {
            double a, b, g;
            get_angle_axis(Phi, a, b, g);
            EXO_ST_PHI(&ph[0], a); EXO_ST_PHI(&ph[sk], b); EXO_ST_PHI(&ph[2 * sk], g);
            break;
        }
        case EXO_ROT_QUATERNION:
        {
            double q[4];
            get_quaternion(Phi, q);
            EXO_ST_PHI(&ph[0], q[0]); EXO_ST_PHI(&ph[sk], q[1]); EXO_ST_PHI(&ph[2 * sk], q[2]); EXO_ST_PHI(&ph[3 * sk], q[3]);
            break;
        }
        default:  // EXO_ROT_MATRIX
            EXO_ST_PHI(&ph[0], Phi.m0); EXO_ST_PHI(&ph[sk], Phi.m1); EXO_ST_PHI(&ph[2 * sk], Phi.m2); EXO_ST_PHI(&ph[3 * sk], Phi.m3); EXO_ST_PHI(&ph[4 * sk], Phi.m4);
            EXO_ST_PHI(&ph[5 * sk], Phi.m5); EXO_ST_PHI(&ph[6 * sk], Phi.m6); EXO_ST_PHI(&ph[7 * sk], Phi.m7); EXO_ST_PHI(&ph[8 * sk], Phi.m8);
            break;
    }
}

// JointLimit::Update (joint_limit.cpp:63-99); phi starts from the SetZero value 0.  qs: staged state, stride qstride.
__device__ __forceinline__ void emit_joint_limit(const double* c, int N, const double* qs, int64_t qstride, double* ph, int64_t phi_sk, double* jr,
                                                 int64_t j_sr, int64_t j_sc)
{
    const double safe = c[0];
    const double* lo = c + 1;
    const double* hi = c + 1 + N;
    for (int i = 0; i < N; ++i)
    {
        const double q = qs[i * qstride];
        const double tau = (0.5 * safe) * (hi[i] - lo[i]);
        double v = 0.0;
        if (q < lo[i] + tau) v = q - lo[i] - tau;
        if (q > hi[i] - tau) v = q - hi[i] + tau;
        ph[i * phi_sk] = v;
        if (jr)
        {
            const double d = (q >= lo[i] + tau && q <= hi[i] - tau) ? 0.0 : 1.0;
            double* row = jr + (int64_t)i * j_sr;
            for (int col = 0; col < N; ++col) row[(int64_t)col * j_sc] = (col == i) ? d : 0.0;
        }
    }
}

// Unused map: Phi.SetZero() values (task_space_vector.cpp:54-62) and zero Jacobian rows
__device__ __forceinline__ void emit_fill(const double* c, int n_items, int n_rows, int N, double* ph, int64_t phi_sk, double* jr, int64_t j_sr, int64_t j_sc)
{
    for (int i = 0; i < n_items; ++i) ph[i * phi_sk] = c[i];
    if (jr)
        for (int r = 0; r < n_rows; ++r)
            for (int col = 0; col < N; ++col) jr[(int64_t)r * j_sr + (int64_t)col * j_sc] = 0.0;
}

}  // namespace dev
}  // namespace exo
