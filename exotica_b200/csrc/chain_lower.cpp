// chain_lower.cpp -- host-side lowering of a serial-chain tree program to the Denavit-Hartenberg form the chain
// kernel executes (chain_kernel.cu).
//
// The compiled program (compiler.cpp) describes joint j of a chain as  G_j = G_{j-1} * pre_j * Jt(u_j, q_j)  with an
// arbitrary constant frame pre_j and an arbitrary unit axis u_j; executing that directly needs per-joint case
// analysis (axis-aligned or not, translation-only or not).  Here every joint frame is re-chosen, once per problem,
// so that its z axis IS the joint axis and its x axis is the common normal to the previous joint axis (modified DH):
//
//     D_j = D_{j-1} * Rx(alpha_j) * Tx(a_j) * Rz(theta_j + q_j) * Tz(d_j)        D_{-1} * Rx * Tx (j = 0) = base
//
// which costs a fixed 30 FP64 operations per joint with no branches, whatever the URDF's axes and origins are.  The
// joint axis in world coordinates is the z column of the frame before Rz, any point of that frame's z axis serves as
// the reference point of the Jacobian column (KinematicTree::ComputeJ, exotica_core/src/kinematic_tree.cpp:919-945),
// and each requested frame is D_{n-1} * K_f with a constant K_f.
//
// The derivation runs in long double; its result is then checked in plain double against the program's own form at
// random configurations, and the lowering is rejected (the interpreter runs the program instead) if they differ by
// more than 1e-13 -- e.g. for almost-but-not-exactly parallel consecutive axes, where DH parameters blow up.
#include <math.h>
#include <string.h>

#include <random>

#include "compiler.h"
#include "kernels.cuh"

namespace exo
{
namespace
{
typedef long double real;

struct LFrame
{
    real M[9];
    real p[3];
};

LFrame l_identity()
{
    LFrame f;
    for (int i = 0; i < 9; ++i) f.M[i] = (i % 4 == 0) ? 1.0L : 0.0L;
    f.p[0] = f.p[1] = f.p[2] = 0.0L;
    return f;
}

LFrame l_from12(const double* v)
{
    LFrame f;
    for (int i = 0; i < 9; ++i) f.M[i] = v[i];
    for (int i = 0; i < 3; ++i) f.p[i] = v[9 + i];
    return f;
}

LFrame l_mul(const LFrame& a, const LFrame& b)
{
    LFrame r;
    for (int i = 0; i < 3; ++i)
    {
        for (int j = 0; j < 3; ++j) r.M[3 * i + j] = a.M[3 * i] * b.M[j] + a.M[3 * i + 1] * b.M[3 + j] + a.M[3 * i + 2] * b.M[6 + j];
        r.p[i] = a.M[3 * i] * b.p[0] + a.M[3 * i + 1] * b.p[1] + a.M[3 * i + 2] * b.p[2] + a.p[i];
    }
    return r;
}

LFrame l_inverse(const LFrame& a)
{
    LFrame r;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) r.M[3 * i + j] = a.M[3 * j + i];
    for (int i = 0; i < 3; ++i) r.p[i] = -(a.M[i] * a.p[0] + a.M[3 + i] * a.p[1] + a.M[6 + i] * a.p[2]);
    return r;
}

LFrame l_rotz(real th)
{
    LFrame f = l_identity();
    const real c = cosl(th), s = sinl(th);
    f.M[0] = c; f.M[1] = -s; f.M[3] = s; f.M[4] = c;
    return f;
}

LFrame l_rotx_cs(real c, real s)
{
    LFrame f = l_identity();
    f.M[4] = c; f.M[5] = -s; f.M[7] = s; f.M[8] = c;
    return f;
}

LFrame l_trans(real x, real y, real z)
{
    LFrame f = l_identity();
    f.p[0] = x; f.p[1] = y; f.p[2] = z;
    return f;
}

// Re-orthonormalise a rotation given in double (URDF-derived, orthonormal to ~1e-16) in long double: Gram-Schmidt on
// the columns, keeping the third column's direction (it becomes the joint axis).
void l_orthonormalise_keep_z(LFrame& f)
{
    real x[3] = {f.M[0], f.M[3], f.M[6]}, z[3] = {f.M[2], f.M[5], f.M[8]};
    real nz = sqrtl(z[0] * z[0] + z[1] * z[1] + z[2] * z[2]);
    for (int i = 0; i < 3; ++i) z[i] /= nz;
    real d = x[0] * z[0] + x[1] * z[1] + x[2] * z[2];
    for (int i = 0; i < 3; ++i) x[i] -= d * z[i];
    real nx = sqrtl(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
    for (int i = 0; i < 3; ++i) x[i] /= nx;
    const real y[3] = {z[1] * x[2] - z[2] * x[1], z[2] * x[0] - z[0] * x[2], z[0] * x[1] - z[1] * x[0]};
    for (int i = 0; i < 3; ++i)
    {
        f.M[3 * i] = x[i];
        f.M[3 * i + 1] = y[i];
        f.M[3 * i + 2] = z[i];
    }
}

// Rotation-only frame whose z axis is the unit vector u (any x axis perpendicular to it).
LFrame l_align_z(const double* u)
{
    real z[3] = {u[0], u[1], u[2]};
    real n = sqrtl(z[0] * z[0] + z[1] * z[1] + z[2] * z[2]);
    for (int i = 0; i < 3; ++i) z[i] /= n;
    LFrame f = l_identity();
    // exact axis-aligned cases first so that common robots get exact permutation matrices
    if (z[0] == 0 && z[1] == 0) return z[2] > 0 ? f : l_rotx_cs(-1.0L, 0.0L);
    int k = 0;  // seed the x axis with the basis vector least aligned with z
    if (fabsl(z[1]) < fabsl(z[k])) k = 1;
    if (fabsl(z[2]) < fabsl(z[k])) k = 2;
    real x[3] = {0, 0, 0};
    x[k] = 1.0L;
    for (int i = 0; i < 3; ++i)
    {
        f.M[3 * i] = x[i];
        f.M[3 * i + 2] = z[i];
    }
    l_orthonormalise_keep_z(f);
    return f;
}

// Y = Rz(th1) Tz(d1) Rx(alpha) Tx(a) Rz(th2) Tz(d2): the screw about z, screw about x, screw about z decomposition of the
// transform between two joint-axis frames.  Returns false for nearly (but not exactly enough) parallel axes.
bool l_decompose(const LFrame& Y, real& th1, real& d1, real& ca, real& sa, real& a, real& th2, real& d2)
{
    const real* R = Y.M;
    ca = R[8];
    sa = sqrtl(R[2] * R[2] + R[5] * R[5]);  // >= 0: alpha in [0, pi]
    const real par_eps = 1e-12L;
    if (sa < par_eps)
    {
        // parallel axes: the common normal is not unique; put it along the offset between the axes
        sa = 0.0L;
        ca = ca > 0 ? 1.0L : -1.0L;
        const real h = hypotl(Y.p[0], Y.p[1]);
        th1 = h > 0 ? atan2l(Y.p[1], Y.p[0]) : 0.0L;
        a = h;
        d1 = Y.p[2];
        d2 = 0.0L;
        // Rz(th2) = Rx(alpha)^-1 Rz(th1)^-1 R
        const LFrame rest = l_mul(l_mul(l_inverse(l_rotx_cs(ca, sa)), l_inverse(l_rotz(th1))), Y);
        th2 = atan2l(rest.M[3], rest.M[0]);
        return true;
    }
    if (sa < 1e-6L) return false;  // ill-conditioned (d2 ~ offset / sin(alpha)): let the interpreter run this robot
    th1 = atan2l(R[2], -R[5]);
    th2 = atan2l(R[6], R[7]);
    // v = Rz(-th1) t = (a, -sa d2, d1 + ca d2)
    const real c1 = cosl(th1), s1 = sinl(th1);
    const real vx = c1 * Y.p[0] + s1 * Y.p[1], vy = -s1 * Y.p[0] + c1 * Y.p[1], vz = Y.p[2];
    a = vx;
    d2 = -vy / sa;
    d1 = vz - ca * d2;
    return true;
}

// ---- double-precision emulation of both forms (verification) ------------------------------------------------------
Frame joint_motion(const Op& op, const double* c, double q)
{
    Frame r = frame_identity();
    const double* u = c + CJ_AXIS;
    if (op.flags & F_PRISMATIC)
    {
        for (int k = 0; k < 3; ++k) r.p[k] = u[k] * q;
        return r;
    }
    const double ct = cos(q), st = sin(q), vt = 1 - ct;
    r.M[0] = ct + vt * u[0] * u[0];
    r.M[1] = -u[2] * st + vt * u[0] * u[1];
    r.M[2] = u[1] * st + vt * u[0] * u[2];
    r.M[3] = u[2] * st + vt * u[0] * u[1];
    r.M[4] = ct + vt * u[1] * u[1];
    r.M[5] = -u[0] * st + vt * u[1] * u[2];
    r.M[6] = -u[1] * st + vt * u[0] * u[2];
    r.M[7] = u[0] * st + vt * u[1] * u[2];
    r.M[8] = ct + vt * u[2] * u[2];
    return r;
}

Frame dh_step(const Frame& in, const ChainJointP& jp, bool first, double q, double* w, double* p)
{
    Frame cur = in;
    if (!first)
    {
        Frame rx = frame_identity();
        rx.M[4] = jp.ca; rx.M[5] = -jp.sa; rx.M[7] = jp.sa; rx.M[8] = jp.ca;
        rx.p[0] = jp.a;  // Rx(alpha) * Tx(a) = (Rx, (a,0,0))
        cur = frame_mul(cur, rx);
    }
    for (int k = 0; k < 3; ++k)
    {
        w[k] = cur.M[3 * k + 2];
        p[k] = cur.p[k];
    }
    const double v = jp.qmult * q;
    Frame rz = frame_identity();
    if (jp.prismatic)
    {
        rz.M[0] = jp.c0; rz.M[1] = -jp.s0; rz.M[3] = jp.s0; rz.M[4] = jp.c0;
        rz.p[2] = jp.d + v;
    }
    else
    {
        const double th = jp.th0 + v;
        rz.M[0] = cos(th); rz.M[1] = -sin(th); rz.M[3] = sin(th); rz.M[4] = cos(th);
        rz.p[2] = jp.d;
    }
    return frame_mul(cur, rz);
}
}  // namespace

bool chain_lower(const Program& prog, ChainParams& P)
{
    const ProgramHeader& h = prog.hdr;
    memset(&P, 0, sizeof(P));
    if (h.chain_joints < 1 || h.chain_joints > CH_MAXJ) return false;
    if (h.n_ctrl > CH_MAXN) return false;
    const int n = h.chain_joints;
    P.n_joints = n;
    P.n_ctrl = h.n_ctrl;
    P.all_revolute = 1;
    P.ortho = 1;

    std::vector<const Op*> jop, fop;
    for (const Op& op : prog.ops)
    {
        const double* c = prog.dbls.data() + op.cst;
        switch (op.kind)
        {
            case OP_END:
                break;
            case OP_JOINT:
                jop.push_back(&op);
                break;
            case OP_EMIT_FRAME:
            {
                const int bk = op.flags & F_B_MASK;
                if (bk != F_B_IDENTITY && bk != F_B_CONST) return false;
                fop.push_back(&op);
                break;
            }
            case OP_EMIT_JLIMIT:
            {
                if (P.jl_enabled) return false;
                P.jl_enabled = 1;
                P.jl_phi_off = op.phi_off;
                P.jl_jrow_off = op.jrow_off;
                const double safe = c[0];
                const double* lo = c + 1;
                const double* hi = c + 1 + h.n_ctrl;
                for (int i = 0; i < h.n_ctrl; ++i)
                {
                    // same operation order as JointLimit::Initialize/Update (joint_limit.cpp:52-55,69-75)
                    const double tau = 0.5 * safe * (hi[i] - lo[i]);
                    P.jl_lo[i] = lo[i];
                    P.jl_hi[i] = hi[i];
                    P.jl_tau[i] = tau;
                    P.jl_lo_tau[i] = lo[i] + tau;
                    P.jl_hi_tau[i] = hi[i] - tau;
                }
                break;
            }
            case OP_EMIT_FILL:
            {
                if (P.n_fill >= CH_MAXFILLS || op.n_items > CH_MAXFILL) return false;
                ChainFillP& fl = P.fill[P.n_fill++];
                fl.phi_off = op.phi_off;
                fl.jrow_off = op.jrow_off;
                fl.n_items = op.n_items;
                fl.n_rows = op.n_rows;
                memcpy(fl.value, c, op.n_items * sizeof(double));
                break;
            }
            default:
                return false;
        }
    }
    if ((int)jop.size() != n || (int)fop.size() > CH_MAXF) return false;
    P.n_frames = (int)fop.size();

    // ---- DH derivation (long double) --------------------------------------------------------------------------
    // S_j: constant frame (in the coordinates of F_j = G_{j-1} pre_j) whose z axis is the joint axis through F_j's origin.
    LFrame S = l_identity();
    real lead_th1[CH_MAXJ], lead_d1[CH_MAXJ];  // leading screw about z of the transform to the NEXT joint
    double qoff[CH_MAXJ];
    for (int j = 0; j < n; ++j) lead_th1[j] = lead_d1[j] = 0.0L;
    for (int j = 0; j < n; ++j)
    {
        const Op& op = *jop[j];
        const double* c = prog.dbls.data() + op.cst;
        ChainJointP& jp = P.joint[j];
        LFrame pre = l_from12(c + CJ_PRE_M);
        const LFrame S0 = l_align_z(c + CJ_AXIS);
        jp.qidx = op.qidx;
        jp.qmult = op.qidx >= 0 ? c[CJ_QMULT] : 0.0;
        jp.prismatic = (op.flags & F_PRISMATIC) ? 1 : 0;
        if (jp.prismatic) P.all_revolute = 0;
        jp.col = -1;
        real th2 = 0, d2 = 0;
        if (j == 0)
        {
            // E_0 = pre_0 * S_0 is the base frame
            S = S0;
            const LFrame base = l_mul(pre, S);
            for (int i = 0; i < 9; ++i) P.base[i] = (double)base.M[i];
            for (int i = 0; i < 3; ++i) P.base[9 + i] = (double)base.p[i];
            jp.ca = 1.0; jp.sa = 0.0; jp.a = 0.0;
        }
        else
        {
            // Y = S_{j-1}^-1 pre_j S0_j = Rz(th1) Tz(d1) Rx(alpha) Tx(a) Rz(th2) Tz(d2)
            const LFrame Y = l_mul(l_mul(l_inverse(S), pre), S0);
            real th1, d1, ca, sa, a;
            if (!l_decompose(Y, th1, d1, ca, sa, a, th2, d2)) return false;
            // the leading screw about z joins the previous joint's motion
            lead_th1[j - 1] = th1;
            lead_d1[j - 1] = d1;
            jp.ca = (double)ca; jp.sa = (double)sa; jp.a = (double)a;
            // perpendicular, intersecting axes up to rounding: make it exact (the verification below still has the last word)
            if (fabs(jp.ca) < 1e-15)
            {
                jp.ca = 0.0;
                jp.sa = jp.sa > 0 ? 1.0 : -1.0;
            }
            if (fabs(jp.a) < 1e-15) jp.a = 0.0;
            if (jp.ca != 0.0 || jp.a != 0.0) P.ortho = 0;
            // S_j = S0_j * (Rz(th2) Tz(d2))^-1
            LFrame tail = l_rotz(th2);
            tail.p[2] = d2;
            S = l_mul(S0, l_inverse(tail));
        }
        qoff[j] = c[CJ_QOFF];
    }
    // constant part of every joint value: revolute theta0 = th1 + qoff, prismatic d = d1 + qoff with a constant Rz(th1)
    for (int j = 0; j < n; ++j)
    {
        ChainJointP& jp = P.joint[j];
        const real th1 = lead_th1[j], d1 = lead_d1[j];
        if (jp.prismatic)
        {
            jp.th0 = (double)th1;
            jp.c0 = (double)cosl(th1);
            jp.s0 = (double)sinl(th1);
            jp.d = (double)(d1 + (real)qoff[j]);
        }
        else
        {
            jp.th0 = (double)(th1 + (real)qoff[j]);
            jp.c0 = (double)cosl((real)jp.th0);
            jp.s0 = (double)sinl((real)jp.th0);
            jp.d = (double)d1;
        }
    }
    // frames: A_f = G_{n-1} offa_f = D_{n-1} * (S_{n-1}^-1 offa_f)
    uint32_t owned = 0;
    for (int f = 0; f < P.n_frames; ++f)
    {
        const Op& op = *fop[f];
        const double* c = prog.dbls.data() + op.cst;
        ChainFrameP& fp = P.frame[f];
        const LFrame K = l_mul(l_inverse(S), l_from12(c + CE_OFFA_M));
        for (int i = 0; i < 9; ++i) fp.k[i] = (double)K.M[i];
        for (int i = 0; i < 3; ++i) fp.k[9 + i] = (double)K.p[i];
        // snap exact identities so that the kernel can skip the multiplication
        bool rot_id = true, trans0 = true;
        for (int i = 0; i < 9; ++i)
            if (fabs(fp.k[i] - ((i % 4 == 0) ? 1.0 : 0.0)) > 1e-17) rot_id = false;
        for (int i = 0; i < 3; ++i)
            if (fp.k[9 + i] != 0.0) trans0 = false;
        if (rot_id)
            for (int i = 0; i < 9; ++i) fp.k[i] = (i % 4 == 0) ? 1.0 : 0.0;
        fp.k_kind = rot_id ? (trans0 ? F_OFFA_NONE : F_OFFA_TRANS) : F_OFFA_FULL;
        memcpy(fp.binv, c + CE_B_M, 12 * sizeof(double));
        fp.b_const = ((op.flags & F_B_MASK) == F_B_CONST) ? 1 : 0;
        fp.map_kind = op.map_kind;
        fp.rot_type = op.rot_type;
        fp.phi_off = op.phi_off;
        fp.jrow_off = op.jrow_off;
        fp.n_rows = op.n_rows;
        fp.row_mode = (op.flags & F_ROWS_MASK) >> F_ROWS_SHIFT;
        // contributors by joint ordinal: (column << 2) | (prismatic << 1) | negative
        const int32_t* otab = prog.ints.data() + op.tab + h.n_ctrl;
        for (int j = 0; j < n; ++j)
        {
            const int ent = otab[j];
            const int col = ent < 0 ? -1 : (ent >> 2);
            if (ent >= 0 && (ent & 1)) return false;  // frame-B chain contributions never occur with a constant B
            if (f == 0)
                P.joint[j].col = col;
            else if (P.joint[j].col != col)
                return false;  // all frames hang off the chain tip, so they share the contributor set
            if (col >= 0) owned |= 1u << col;
        }
    }
    P.zero_cols = (h.n_ctrl >= 32 ? 0xffffffffu : ((1u << h.n_ctrl) - 1u)) & ~owned;
    if (P.n_frames == 0) P.zero_cols = 0;
    P.has_prefix = (P.jl_enabled || P.n_fill > 0 || P.zero_cols != 0) ? 1 : 0;

    // ---- verification in double against the program's own form ------------------------------------------------------
    std::mt19937_64 rng(12345);
    std::uniform_real_distribution<double> uni(-3.2, 3.2);
    for (int trial = 0; trial < 24; ++trial)
    {
        double x[CH_MAXN];
        for (int i = 0; i < h.n_ctrl; ++i) x[i] = trial == 0 ? 0.0 : uni(rng);
        Frame G = frame_identity(), D = frame_from12(P.base);
        double scale = 1.0;
        for (int j = 0; j < n; ++j)
        {
            const Op& op = *jop[j];
            const double* c = prog.dbls.data() + op.cst;
            const double xq = op.qidx >= 0 ? x[op.qidx] : 0.0;
            const double q = op.qidx >= 0 ? c[CJ_QMULT] * xq + c[CJ_QOFF] : c[CJ_QOFF];
            G = frame_mul(frame_mul(G, frame_from12(c + CJ_PRE_M)), joint_motion(op, c, q));
            double w[3], p[3];
            D = dh_step(D, P.joint[j], j == 0, xq, w, p);
            // joint axis direction must match: G.M * u
            for (int k = 0; k < 3; ++k)
            {
                const double wk = G.M[3 * k] * c[CJ_AXIS] + G.M[3 * k + 1] * c[CJ_AXIS + 1] + G.M[3 * k + 2] * c[CJ_AXIS + 2];
                if (fabs(wk - w[k]) > 1e-13) return false;
                scale = fmax(scale, fabs(G.p[k]));
            }
            // the DH reference point must lie on the joint axis: (p - G.p) x w = 0
            const double dx = p[0] - G.p[0], dy = p[1] - G.p[1], dz = p[2] - G.p[2];
            const double cx = dy * w[2] - dz * w[1], cy = dz * w[0] - dx * w[2], cz = dx * w[1] - dy * w[0];
            if (fabs(cx) + fabs(cy) + fabs(cz) > 1e-13 * scale) return false;
        }
        for (int f = 0; f < P.n_frames; ++f)
        {
            const Op& op = *fop[f];
            const Frame A = frame_mul(G, frame_from12(prog.dbls.data() + op.cst + CE_OFFA_M));
            const Frame A2 = frame_mul(D, frame_from12(P.frame[f].k));
            for (int i = 0; i < 9; ++i)
                if (fabs(A.M[i] - A2.M[i]) > 1e-13) return false;
            for (int i = 0; i < 3; ++i)
                if (fabs(A.p[i] - A2.p[i]) > 1e-13 * scale) return false;
        }
    }
    return true;
}

}  // namespace exo
