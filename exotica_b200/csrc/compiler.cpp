// compiler.cpp -- flattened KinematicTree + task maps -> straight-line tree program (see program.h).
//
// What the reference does per configuration at run time (pointer-tree BFS over every element, two
// parent-chain walks per requested frame; exotica_core/src/kinematic_tree.cpp:695-737,919-945) is resolved
// here once per problem: which elements matter, which transforms are constant and can be folded, which
// joint feeds which Jacobian column of which frame with which sign, and how few scratch slots are needed.
#include "compiler.h"

#include <math.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <set>

#include "../../include/exotica_b200.h"

namespace exo
{
// ---- small host-side frame algebra ----------------------------------------------------------------
Frame frame_identity()
{
    Frame f;
    memset(&f, 0, sizeof(f));
    f.M[0] = f.M[4] = f.M[8] = 1.0;
    return f;
}

Frame frame_from12(const double* v)
{
    Frame f;
    memcpy(f.M, v, 9 * sizeof(double));
    memcpy(f.p, v + 9, 3 * sizeof(double));
    return f;
}

Frame frame_mul(const Frame& a, const Frame& b)
{
    Frame r;
    for (int i = 0; i < 3; ++i)
    {
        for (int j = 0; j < 3; ++j) r.M[3 * i + j] = a.M[3 * i] * b.M[j] + a.M[3 * i + 1] * b.M[3 + j] + a.M[3 * i + 2] * b.M[6 + j];
        r.p[i] = (a.M[3 * i] * b.p[0] + a.M[3 * i + 1] * b.p[1] + a.M[3 * i + 2] * b.p[2]) + a.p[i];
    }
    return r;
}

Frame frame_inverse(const Frame& a)
{
    Frame r;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) r.M[3 * i + j] = a.M[3 * j + i];
    for (int i = 0; i < 3; ++i) r.p[i] = -(a.M[i] * a.p[0] + a.M[3 + i] * a.p[1] + a.M[6 + i] * a.p[2]);
    return r;
}

bool rot_is_identity(const Frame& f)
{
    static const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    for (int i = 0; i < 9; ++i)
        if (f.M[i] != I[i]) return false;
    return true;
}

bool frame_is_identity(const Frame& f)
{
    return rot_is_identity(f) && f.p[0] == 0.0 && f.p[1] == 0.0 && f.p[2] == 0.0;
}

static Frame rot_axis_angle(const double* u, double q)
{
    // Rodrigues, same closed form as KDL::Rotation::Rot2
    Frame r = frame_identity();
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

// Rotation-only frame S with S.M * z = u (u unit): signed permutation matrices for axis-aligned u (exact), Gram-Schmidt
// otherwise.  With it  Rot(u, q) = S Rz(q) S^-1  and  Trans(u q) = S Tz(q) S^-1.
static Frame align_z(const double* u)
{
    Frame f = frame_identity();
    double z[3] = {u[0], u[1], u[2]};
    const double n = sqrt(z[0] * z[0] + z[1] * z[1] + z[2] * z[2]);
    if (n > 0)
        for (int i = 0; i < 3; ++i) z[i] /= n;
    if (z[0] == 0.0 && z[1] == 0.0)
    {
        if (z[2] < 0.0) f.M[4] = f.M[8] = -1.0;  // Rx(pi)
        return f;
    }
    int k = 0;  // seed the x axis with the basis vector least aligned with z
    if (fabs(z[1]) < fabs(z[k])) k = 1;
    if (fabs(z[2]) < fabs(z[k])) k = 2;
    double x[3] = {0, 0, 0};
    x[k] = 1.0;
    const double d = x[0] * z[0] + x[1] * z[1] + x[2] * z[2];
    for (int i = 0; i < 3; ++i) x[i] -= d * z[i];
    const double nx = sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
    for (int i = 0; i < 3; ++i) x[i] /= nx;
    const double y[3] = {z[1] * x[2] - z[2] * x[1], z[2] * x[0] - z[0] * x[2], z[0] * x[1] - z[1] * x[0]};
    for (int i = 0; i < 3; ++i)
    {
        f.M[3 * i] = x[i];
        f.M[3 * i + 1] = y[i];
        f.M[3 * i + 2] = z[i];
    }
    return f;
}

int rotation_length(int rot_type)
{
    static const int len[6] = {4, 3, 3, 3, 3, 9};  // conversions.h:94-98
    return (rot_type >= 0 && rot_type < 6) ? len[rot_type] : -1;
}

// joint description of one element reduced to (origin, axis, kind)
struct JointGeom
{
    bool rotational = false, translational = false;
    double origin[3] = {0, 0, 0};
    double axis[3] = {0, 0, 0};
};

static JointGeom joint_geom(const TreeModel& t, int e)
{
    JointGeom g;
    const int jt = t.joint_type[e];
    switch (jt)
    {
        case EXO_JOINT_ROT_AXIS:
        case EXO_JOINT_TRANS_AXIS:
            for (int k = 0; k < 3; ++k)
            {
                g.origin[k] = t.origin[3 * e + k];
                g.axis[k] = t.axis[3 * e + k];
            }
            g.rotational = jt == EXO_JOINT_ROT_AXIS;
            g.translational = !g.rotational;
            break;
        case EXO_JOINT_ROT_X:
        case EXO_JOINT_ROT_Y:
        case EXO_JOINT_ROT_Z:
            g.rotational = true;
            g.axis[jt - EXO_JOINT_ROT_X] = 1.0;
            break;
        case EXO_JOINT_TRANS_X:
        case EXO_JOINT_TRANS_Y:
        case EXO_JOINT_TRANS_Z:
            g.translational = true;
            g.axis[jt - EXO_JOINT_TRANS_X] = 1.0;
            break;
        default:
            break;
    }
    return g;
}

// Segment::pose(q) = Joint::pose(q) * f_tip for a constant q
static Frame segment_pose_const(const TreeModel& t, int e, double q)
{
    const JointGeom g = joint_geom(t, e);
    Frame jp = frame_identity();
    if (g.rotational)
    {
        jp = rot_axis_angle(g.axis, q);
        for (int k = 0; k < 3; ++k) jp.p[k] = g.origin[k];
    }
    else if (g.translational)
    {
        for (int k = 0; k < 3; ++k) jp.p[k] = g.origin[k] + g.axis[k] * q;
    }
    return frame_mul(jp, frame_from12(&t.tip[12 * e]));
}

bool assign_indexing(const TreeModel& tree, std::vector<MapSpec>& maps, int& length_phi, int& length_jac, std::string& err)
{
    length_phi = length_jac = 0;
    for (MapSpec& m : maps)
    {
        switch (m.kind)
        {
            case EXO_MAP_EFF_FRAME:
                if (rotation_length(m.rot_type) < 0)
                {
                    err = "Unknown rotation representation type!";
                    return false;
                }
                m.length = m.n_frames * (3 + rotation_length(m.rot_type));
                m.length_jacobian = m.n_frames * 6;
                break;
            case EXO_MAP_EFF_POSITION:
                m.length = m.length_jacobian = m.n_frames * 3;
                break;
            case EXO_MAP_JOINT_LIMIT:
                m.length = m.length_jacobian = tree.n_ctrl;
                break;
            case EXO_MAP_SPHERE_COLLISION:
            {
                int g = 0;
                for (int v : m.group) g = std::max(g, v + 1);
                m.n_groups = g;
                m.length = m.length_jacobian = g * (g - 1) / 2;
                break;
            }
            case EXO_MAP_FRAMES:
                m.length = m.n_frames * 12;
                m.length_jacobian = m.n_frames * 6;
                break;
            case EXO_MAP_EFF_ORIENTATION:  // eff_orientation.cpp:36-52, :100-108
                if (rotation_length(m.rot_type) < 0)
                {
                    err = "Unknown rotation representation type!";
                    return false;
                }
                m.length = m.n_frames * rotation_length(m.rot_type);
                m.length_jacobian = m.n_frames * 3;
                break;
            case EXO_MAP_EFF_POSITION_XY:  // eff_position_xy.cpp
                m.length = m.length_jacobian = m.n_frames * 2;
                break;
            case EXO_MAP_DISTANCE:  // distance.cpp
                m.length = m.length_jacobian = m.n_frames;
                break;
            case EXO_MAP_JOINT_POSE:  // joint_pose.cpp:66-103
                m.length = m.length_jacobian = (int)m.joint_map.size();
                break;
            case EXO_MAP_EFF_AXIS_ALIGNMENT:  // eff_axis_alignment.cpp:171-174
                m.length = m.length_jacobian = m.n_frames;
                break;
            case EXO_MAP_CENTER_OF_MASS:  // center_of_mass.cpp:134-137
                m.length = m.length_jacobian = m.enable_z ? 3 : 2;
                break;
            default:
                err = "unknown task map kind";
                return false;
        }
        m.start = length_phi;
        m.start_jacobian = length_jac;
        length_phi += m.length;
        length_jac += m.length_jacobian;
    }
    return true;
}

namespace
{
struct Emit
{
    int kind;       // OpKind
    int map;        // index into maps
    int frame;      // frame index inside the map (EMIT_FRAME)
    int ready;      // op index after which it can run (-1: before any op)
    int seq = -1;   // position in the instruction sequence
};

struct Producer
{
    int elem;
    int seq = -1;
    bool needs_frame_slot = false;
    bool needs_joint_slot = false;
    int frame_last_use = -1;
    int joint_last_use = -1;
    int frame_slot = -1;
    int joint_slot = -1;
};
}  // namespace

bool compile_program(const TreeModel& t, std::vector<MapSpec>& maps, Program& out, std::string& err, int mode)
{
    const int n = t.n;
    const int N = t.n_ctrl;
    int length_phi = 0, length_jac = 0;
    if (!assign_indexing(t, maps, length_phi, length_jac, err)) return false;
    for (int e = 0; e < n; ++e)
    {
        if (t.parent[e] >= e || (e > 0 && t.parent[e] < 0))
        {
            err = "tree elements must be ordered parents-first with a single root at index 0";
            return false;
        }
    }

    // ---- 1. classify elements ---------------------------------------------------------------------
    enum
    {
        K_CONST = 0,
        K_JOINT = 1,
        K_TABLE = 2
    };
    std::vector<int> kind(n, K_CONST), q_ctrl(n, -1);
    std::vector<double> q_mult(n, 1.0), q_off(n, 0.0), q_const(n, 0.0);
    for (int e = 0; e < n; ++e)
    {
        if (t.generated[e] >= 0)
        {
            kind[e] = K_TABLE;  // KinematicElement::GetPose: generated_offset wins (kinematic_element.h:95-98)
            continue;
        }
        if (t.joint_type[e] == EXO_JOINT_NONE || t.id[e] < 0) continue;
        const int m = t.mimic_elem[e];
        if (m >= 0)
        {
            // UpdateTree: pose(mimic_offset + mimic_multiplier * tree_state_(mimic_joint_id)) (kinematic_tree.cpp:715-717)
            if (m < n && t.control_id[m] >= 0)
            {
                kind[e] = K_JOINT;
                q_ctrl[e] = t.control_id[m];
                q_mult[e] = t.mimic_mult[e];
                q_off[e] = t.mimic_off[e];
            }
            else
                q_const[e] = t.mimic_off[e] + t.mimic_mult[e] * (m < n ? t.tree_state[m] : 0.0);
        }
        else if (t.control_id[e] >= 0)
        {
            kind[e] = K_JOINT;
            q_ctrl[e] = t.control_id[e];
        }
        else
            q_const[e] = t.tree_state[e];
        if (kind[e] == K_CONST && t.control_id[e] >= 0)
        {
            // controlled element whose pose follows an uncontrolled mimic target: it still owns a Jacobian column
            kind[e] = K_JOINT;
            q_ctrl[e] = -1;
            q_mult[e] = 0.0;
            q_off[e] = q_const[e];
        }
    }

    // ---- 2. needed elements -------------------------------------------------------------------------
    std::vector<char> needed(n, 0);
    for (const MapSpec& m : maps)
    {
        if (!m.is_used) continue;
        for (int f = 0; f < m.n_frames; ++f)
        {
            const int ends[2] = {m.elem_a[f], m.elem_b[f]};
            for (int k = 0; k < 2; ++k)
            {
                int e = ends[k];
                if (e < 0 || e >= n)
                {
                    err = "frame refers to an element outside the tree";
                    return false;
                }
                while (e >= 0 && !needed[e])
                {
                    needed[e] = 1;
                    e = t.parent[e];
                }
            }
        }
    }

    // ---- 3. anchors, folded constants, joint ops --------------------------------------------------------
    std::vector<int> anchor(n, -1);                 // nearest movable ancestor-or-self (element index) or -1 = root
    std::vector<Frame> C(n, frame_identity());      // frame[e] = G[anchor[e]] * C[e]
    std::vector<Frame> pre(n, frame_identity());    // movable: G[e] = G[anchor(parent)] * pre[e] * J(u,q)
    std::vector<int> src_anchor(n, -1);             // movable: anchor of the parent
    std::vector<double> u(3 * (size_t)n, 0.0);
    std::vector<int> order;                         // movable needed elements, parents first
    for (int e = 0; e < n; ++e)
    {
        if (!needed[e]) continue;
        const int p = t.parent[e];
        const Frame Cp = p >= 0 ? C[p] : frame_identity();
        const int ap = p >= 0 ? anchor[p] : -1;
        if (kind[e] == K_TABLE)
        {
            pre[e] = Cp;
            src_anchor[e] = ap;
            anchor[e] = e;
            C[e] = frame_identity();
            order.push_back(e);
        }
        else if (kind[e] == K_JOINT)
        {
            const JointGeom g = joint_geom(t, e);
            const Frame tip = frame_from12(&t.tip[12 * e]);
            Frame jo = frame_identity();  // (I, origin) * (Rt, 0)
            memcpy(jo.M, tip.M, sizeof(jo.M));
            for (int k = 0; k < 3; ++k) jo.p[k] = g.origin[k];
            // u = Rt^T a,  trailing translation t' = Rt^T pt
            Frame trail = frame_identity();
            double ue[3];
            for (int k = 0; k < 3; ++k)
            {
                ue[k] = tip.M[k] * g.axis[0] + tip.M[3 + k] * g.axis[1] + tip.M[6 + k] * g.axis[2];
                trail.p[k] = tip.M[k] * tip.p[0] + tip.M[3 + k] * tip.p[1] + tip.M[6 + k] * tip.p[2];
            }
            // z-canonical form: the stored frame is H = G * S with S.M z = u, so that the joint moves about / along its own
            // z axis whatever the URDF axis is:  H_m = H_a * (pre * S) * Rz(q),  and everything that hangs off m sees S^-1 first
            const Frame S = align_z(ue);
            pre[e] = frame_mul(frame_mul(Cp, jo), S);
            trail = frame_mul(frame_inverse(S), trail);
            u[3 * e] = 0.0; u[3 * e + 1] = 0.0; u[3 * e + 2] = 1.0;
            src_anchor[e] = ap;
            anchor[e] = e;
            C[e] = trail;
            order.push_back(e);
        }
        else
        {
            // constant: root GetPose() (identity unless the segment says otherwise) or fixed / frozen joint
            const Frame pose = segment_pose_const(t, e, q_const[e]);
            anchor[e] = ap;
            C[e] = frame_mul(Cp, pose);
        }
    }
    const int n_ops = (int)order.size();
    std::map<int, int> op_of;  // element -> op index
    for (int i = 0; i < n_ops; ++i) op_of[order[i]] = i;
    auto ready_of = [&](int elem_anchor) { return elem_anchor < 0 ? -1 : op_of[elem_anchor]; };

    // ---- 4. emits -----------------------------------------------------------------------------------
    std::vector<Emit> emits;
    for (int mi = 0; mi < (int)maps.size(); ++mi)
    {
        const MapSpec& m = maps[mi];
        if (!m.is_used)
        {
            emits.push_back({OP_EMIT_FILL, mi, 0, -1});
            continue;
        }
        switch (m.kind)
        {
            case EXO_MAP_JOINT_LIMIT:
                emits.push_back({OP_EMIT_JLIMIT, mi, 0, -1});
                break;
            case EXO_MAP_JOINT_POSE:
                emits.push_back({OP_EMIT_JPOSE, mi, 0, -1});
                break;
            case EXO_MAP_SPHERE_COLLISION:
            {
                int r = -1;
                for (int f = 0; f < m.n_frames; ++f)
                {
                    r = std::max(r, ready_of(anchor[m.elem_a[f]]));
                    if (m.elem_b[f] != 0 || !frame_is_identity(frame_from12(&m.off_b[12 * f])))
                    {
                        err = "SphereCollision on the GPU path requires Base = root with identity BaseOffset";
                        return false;
                    }
                }
                emits.push_back({OP_EMIT_SPHERES, mi, 0, r});
                break;
            }
            case EXO_MAP_CENTER_OF_MASS:
                // every frame adds its share to the same rows: zero them first (Phi and J), then accumulate
                emits.push_back({OP_EMIT_FILL, mi, 0, -1});
                // fall through
            default:
                for (int f = 0; f < m.n_frames; ++f)
                    emits.push_back({OP_EMIT_FRAME, mi, f, std::max(ready_of(anchor[m.elem_a[f]]), ready_of(anchor[m.elem_b[f]]))});
        }
    }

    // ---- 5. instruction sequence -----------------------------------------------------------------------
    // seq item: (is_emit, index)
    std::vector<std::pair<int, int>> seq;
    for (int k = -1; k < n_ops; ++k)
    {
        if (k >= 0) seq.push_back({0, k});
        for (int i = 0; i < (int)emits.size(); ++i)
            if (emits[i].ready == k)
            {
                emits[i].seq = (int)seq.size();
                seq.push_back({1, i});
            }
    }
    std::vector<Producer> prod(n_ops);
    for (int i = 0; i < n_ops; ++i) prod[i].elem = order[i];
    for (int s = 0; s < (int)seq.size(); ++s)
        if (!seq[s].first) prod[seq[s].second].seq = s;
    // "current" producer visible in registers at sequence position s = the last op before s
    std::vector<int> cur_at(seq.size(), -1);
    {
        int cur = -1;
        for (int s = 0; s < (int)seq.size(); ++s)
        {
            cur_at[s] = cur;
            if (!seq[s].first) cur = seq[s].second;
        }
    }
    auto use_frame = [&](int elem_anchor, int s, bool force_slot = false) -> int {
        // returns the source code for reading G[elem_anchor] at sequence position s, registering slot needs
        // (force_slot: the reader is another lane of the warp, which cannot see this lane's registers)
        if (elem_anchor < 0) return SRC_ROOT;
        const int k = op_of[elem_anchor];
        if (cur_at[s] == k && !force_slot) return SRC_CUR;
        prod[k].needs_frame_slot = true;
        prod[k].frame_last_use = std::max(prod[k].frame_last_use, s);
        return k;  // provisional: op index, patched to the slot number after allocation
    };
    auto use_joint = [&](int elem, int s) -> int {
        const int k = op_of[elem];
        prod[k].needs_joint_slot = true;
        prod[k].joint_last_use = std::max(prod[k].joint_last_use, s);
        return k;
    };

    // controlled ancestors of an element (self included)
    auto controlled_chain = [&](int e, std::vector<int>& outv) {
        outv.clear();
        while (e >= 0)
        {
            if (t.control_id[e] >= 0 && kind[e] == K_JOINT) outv.push_back(e);
            e = t.parent[e];
        }
    };

    // first pass over the sequence: record sources (provisional) and liveness
    struct EmitPlan
    {
        int src_a = SRC_ROOT, src_b = SRC_ROOT;
        Frame off_a, off_b;
        std::vector<int> contrib;              // per column: -1 or (op index << 2 | pris << 1 | neg)
        std::vector<int> sphere_src;           // per sphere (sorted order): provisional source
        std::vector<int> sphere_anchor_elem;   // per sphere: anchor element or -1
        std::vector<int> sphere_order;         // sorted sphere indices
        std::vector<int> acc_elems;            // distinct non-root anchors
    };
    std::vector<EmitPlan> plans(emits.size());
    std::vector<int> op_src(n_ops, SRC_ROOT);
    for (int s = 0; s < (int)seq.size(); ++s)
    {
        if (!seq[s].first)
        {
            const int k = seq[s].second;
            op_src[k] = use_frame(src_anchor[order[k]], s);
            continue;
        }
        const int ei = seq[s].second;
        Emit& em = emits[ei];
        EmitPlan& pl = plans[ei];
        const MapSpec& m = maps[em.map];
        if (em.kind == OP_EMIT_FRAME)
        {
            const int ea = m.elem_a[em.frame], eb = m.elem_b[em.frame];
            pl.src_a = use_frame(anchor[ea], s);
            pl.src_b = use_frame(anchor[eb], s);
            pl.off_a = frame_mul(C[ea], frame_from12(&m.off_a[12 * em.frame]));
            pl.off_b = frame_mul(C[eb], frame_from12(&m.off_b[12 * em.frame]));
            std::vector<int> ca, cb;
            controlled_chain(ea, ca);
            controlled_chain(eb, cb);
            pl.contrib.assign(N, -1);
            std::set<int> sa(ca.begin(), ca.end()), sb(cb.begin(), cb.end());
            for (int e : ca)
                if (!sb.count(e))
                {
                    const JointGeom g = joint_geom(t, e);
                    pl.contrib[t.control_id[e]] = contrib_pack(use_joint(e, s), g.translational, false);
                }
            for (int e : cb)
                if (!sa.count(e))
                {
                    const JointGeom g = joint_geom(t, e);
                    pl.contrib[t.control_id[e]] = contrib_pack(use_joint(e, s), g.translational, true);
                }
        }
        else if (em.kind == OP_EMIT_SPHERES)
        {
            pl.sphere_order.resize(m.n_frames);
            for (int i = 0; i < m.n_frames; ++i) pl.sphere_order[i] = i;
            std::stable_sort(pl.sphere_order.begin(), pl.sphere_order.end(), [&](int a, int b) { return m.group[a] < m.group[b]; });
            for (int i : pl.sphere_order)
            {
                const int a = anchor[m.elem_a[i]];
                pl.sphere_anchor_elem.push_back(a);
                pl.sphere_src.push_back(use_frame(a, s, true));
                if (a >= 0 && std::find(pl.acc_elems.begin(), pl.acc_elems.end(), a) == pl.acc_elems.end()) pl.acc_elems.push_back(a);
            }
            // Jacobian columns: every controlled ancestor of any anchored sphere
            pl.contrib.assign(N, -1);
            for (int a : pl.acc_elems)
            {
                std::vector<int> ch;
                controlled_chain(a, ch);
                for (int e : ch)
                {
                    const JointGeom g = joint_geom(t, e);
                    pl.contrib[t.control_id[e]] = contrib_pack(use_joint(e, s), g.translational, false);
                }
            }
        }
    }

    // ---- 6. slot allocation (linear scan over the sequence) -------------------------------------------------
    int n_frame_slots = 0, n_joint_slots = 0;
    {
        std::vector<int> free_f, free_j;
        std::vector<std::pair<int, int>> busy_f, busy_j;  // (last use, slot)
        for (int s = 0; s < (int)seq.size(); ++s)
        {
            // release slots whose last use is strictly before s
            for (auto it = busy_f.begin(); it != busy_f.end();)
                if (it->first < s)
                {
                    free_f.push_back(it->second);
                    it = busy_f.erase(it);
                }
                else
                    ++it;
            for (auto it = busy_j.begin(); it != busy_j.end();)
                if (it->first < s)
                {
                    free_j.push_back(it->second);
                    it = busy_j.erase(it);
                }
                else
                    ++it;
            if (seq[s].first) continue;
            Producer& p = prod[seq[s].second];
            if (p.needs_frame_slot)
            {
                if (free_f.empty())
                    p.frame_slot = n_frame_slots++;
                else
                {
                    std::sort(free_f.begin(), free_f.end());
                    p.frame_slot = free_f.front();
                    free_f.erase(free_f.begin());
                }
                busy_f.push_back({p.frame_last_use, p.frame_slot});
            }
            if (p.needs_joint_slot)
            {
                if (free_j.empty())
                    p.joint_slot = n_joint_slots++;
                else
                {
                    std::sort(free_j.begin(), free_j.end());
                    p.joint_slot = free_j.front();
                    free_j.erase(free_j.begin());
                }
                busy_j.push_back({p.joint_last_use, p.joint_slot});
            }
        }
    }
    // ---- 6b. two-pass mode ------------------------------------------------------------------------------------------
    // Eligible when every requested frame has a constant base (no frame-B chains, no Binv that depends on the state) and
    // there is no SphereCollision map (it reads joint slots).  Worth it when the joint slots dominate the scratch.
    bool two_pass = mode != COMPILE_ONE_PASS && n_joint_slots > 0;
    int n_pos_slots = 0;
    std::vector<int> pos_slot_of_emit(emits.size(), -1);
    for (size_t i = 0; i < emits.size(); ++i)
    {
        if (emits[i].kind == OP_EMIT_SPHERES) two_pass = false;
        if (emits[i].kind == OP_EMIT_FRAME)
        {
            if (plans[i].src_b != SRC_ROOT) two_pass = false;
            if (maps[emits[i].map].kind == EXO_MAP_DISTANCE) two_pass = false;  // its row needs Phi.p / |Phi.p| at column time
            pos_slot_of_emit[i] = n_pos_slots++;
        }
    }
    if (mode == COMPILE_AUTO && 6 * n_joint_slots <= 3 * n_pos_slots + 48) two_pass = false;  // < 400 bytes to gain
    // per joint op: the columns it owns, for every requested frame below it
    std::vector<std::vector<int32_t>> tp_entries(n_ops);
    std::vector<int> encoded_index(n_ops, -1);  // producer (joint / table op) -> index in out.ops
    auto patch_src = [&](int v) { return v >= 0 ? prod[v].frame_slot : v; };
    auto patch_contrib = [&](int v) { return v < 0 ? -1 : ((prod[v >> 2].joint_slot << 2) | (v & 3)); };

    // ---- 7. encode ---------------------------------------------------------------------------------------
    out = Program();
    int n_acc_max = 0;
    auto push_frame = [&](const Frame& f) {
        for (int k = 0; k < 9; ++k) out.dbls.push_back(f.M[k]);
        for (int k = 0; k < 3; ++k) out.dbls.push_back(f.p[k]);
    };
    for (int s = 0; s < (int)seq.size(); ++s)
    {
        Op op;
        memset(&op, 0, sizeof(op));
        op.dst = op.jslot = op.qidx = -1;
        op.src = op.src_b = SRC_ROOT;
        if (!seq[s].first)
        {
            const int k = seq[s].second;
            const int e = order[k];
            const Producer& p = prod[k];
            op.src = patch_src(op_src[k]);
            op.dst = p.frame_slot;
            op.cst = (int)out.dbls.size();
            const Frame& pr = pre[e];
            int flags = frame_is_identity(pr) ? F_PRE_NONE : (rot_is_identity(pr) ? F_PRE_TRANS : F_PRE_FULL);
            if (kind[e] == K_TABLE)
            {
                op.kind = OP_TABLE;
                op.qidx = t.generated[e];
                push_frame(pr);
            }
            else
            {
                const JointGeom g = joint_geom(t, e);
                op.kind = OP_JOINT;
                op.jslot = p.joint_slot;
                op.qidx = q_ctrl[e];
                const double* ue = &u[3 * e];
                int axis_flag = F_AXIS_GENERAL;
                for (int a = 0; a < 3; ++a)
                {
                    const int b = (a + 1) % 3, c = (a + 2) % 3;
                    if (ue[b] == 0.0 && ue[c] == 0.0 && (ue[a] == 1.0 || ue[a] == -1.0))
                        axis_flag = (a << 2) | (ue[a] < 0 ? F_AXIS_NEG : 0);
                }
                flags |= axis_flag;
                if (g.translational) flags |= F_PRISMATIC;
                if (op.qidx < 0) flags |= F_CONST_Q;
                push_frame(pr);
                for (int a = 0; a < 3; ++a) out.dbls.push_back(ue[a]);
                out.dbls.push_back(q_mult[e]);
                out.dbls.push_back(q_off[e]);
                out.n_joint_ops++;
            }
            op.flags = flags;
            encoded_index[k] = (int)out.ops.size();
            out.ops.push_back(op);
            continue;
        }
        const Emit& em = emits[seq[s].second];
        const EmitPlan& pl = plans[seq[s].second];
        const MapSpec& m = maps[em.map];
        op.kind = em.kind;
        op.map_kind = m.kind;
        op.rot_type = m.rot_type;
        op.cst = (int)out.dbls.size();
        op.tab = (int)out.ints.size();
        out.n_emit_ops++;
        if (em.kind == OP_EMIT_FRAME)
        {
            const int f = em.frame;
            int rows = 6, stride = 0;
            int row_mode = ROWS_VW;
            if (m.kind == EXO_MAP_EFF_FRAME) stride = 3 + rotation_length(m.rot_type);
            if (m.kind == EXO_MAP_EFF_POSITION) stride = 3, rows = 3, row_mode = ROWS_V;
            if (m.kind == EXO_MAP_FRAMES) stride = 12;
            if (m.kind == EXO_MAP_EFF_ORIENTATION) stride = rotation_length(m.rot_type), rows = 3, row_mode = ROWS_W;
            if (m.kind == EXO_MAP_EFF_POSITION_XY) stride = 2, rows = 2, row_mode = ROWS_VXY;
            if (m.kind == EXO_MAP_DISTANCE) stride = 1, rows = 1, row_mode = ROWS_DIST;
            if (m.kind == EXO_MAP_EFF_AXIS_ALIGNMENT) stride = 1, rows = 1, row_mode = ROWS_WDOT;
            if (m.kind == EXO_MAP_CENTER_OF_MASS) stride = 0, rows = 0, row_mode = m.enable_z ? ROWS_V_ACC : ROWS_VXY_ACC;  // all frames share the rows
            op.phi_off = m.start + f * stride;
            op.jrow_off = m.start_jacobian + f * rows;
            if (m.kind == EXO_MAP_CENTER_OF_MASS) rows = m.enable_z ? 3 : 2;
            op.n_rows = rows;
            op.src = patch_src(pl.src_a);
            op.src_b = patch_src(pl.src_b);
            int flags = frame_is_identity(pl.off_a) ? F_OFFA_NONE : (rot_is_identity(pl.off_a) ? F_OFFA_TRANS : F_OFFA_FULL);
            flags |= row_mode << F_ROWS_SHIFT;
            push_frame(pl.off_a);
            if (op.src_b == SRC_ROOT)
            {
                if (frame_is_identity(pl.off_b))
                {
                    flags |= F_B_IDENTITY;
                    push_frame(pl.off_b);
                }
                else
                {
                    flags |= F_B_CONST;
                    push_frame(frame_inverse(pl.off_b));
                }
            }
            else
            {
                flags |= F_B_SLOT;
                push_frame(pl.off_b);
            }
            op.flags = flags;
            if (m.kind == EXO_MAP_EFF_AXIS_ALIGNMENT)
            {
                // frames_[i + n] is frames_[i] with frame_A_offset.p = axis (eff_axis_alignment.cpp:53-57), so
                // p_{i+n} - p_i = Phi_i.M * a with a = M_off^T (axis - p_off); then the direction.  pl.off_a is the link's
                // constant tail C times the LinkOffset, so the same point difference in the anchor frame is C.M axis + C.p - off_a.p
                const Frame& oa = pl.off_a;
                const Frame& Cf = C[m.elem_a[f]];
                const double* ax = &m.axis[3 * f];
                double d[3];
                for (int k = 0; k < 3; ++k) d[k] = Cf.M[3 * k] * ax[0] + Cf.M[3 * k + 1] * ax[1] + Cf.M[3 * k + 2] * ax[2] + Cf.p[k] - oa.p[k];
                for (int k = 0; k < 3; ++k) out.dbls.push_back(oa.M[k] * d[0] + oa.M[3 + k] * d[1] + oa.M[6 + k] * d[2]);
                for (int k = 0; k < 3; ++k) out.dbls.push_back(m.direction[3 * f + k]);
            }
            if (m.kind == EXO_MAP_CENTER_OF_MASS)
            {
                double total = 0.0;
                for (double v : m.mass) total += v;
                out.dbls.push_back(total != 0.0 ? m.mass[f] / total : 0.0);  // M == 0: Phi and J stay zero (center_of_mass.cpp:41,66)
            }
            if (two_pass)
                for (int c = 0; c < N; ++c)
                    if (pl.contrib[c] >= 0)
                    {
                        const int k = pl.contrib[c] >> 2;
                        const int tpf = ((pl.contrib[c] & 2) ? TP_F_PRISMATIC : 0) | ((flags & F_B_MASK) == F_B_CONST ? TP_F_BCONST : 0);
                        const int32_t ent[TP_INTS] = {pos_slot_of_emit[seq[s].second], c, op.jrow_off, row_mode, tpf, op.cst + CE_B_M};
                        tp_entries[k].insert(tp_entries[k].end(), ent, ent + TP_INTS);
                    }
            op.n_items = pos_slot_of_emit[seq[s].second];  // two-pass mode: where the frame's position is parked
            for (int c = 0; c < N; ++c) out.ints.push_back(patch_contrib(pl.contrib[c]));
            // the same contributors keyed by joint-op ordinal (used by the chain kernel, whose joint data sits in
            // registers indexed by ordinal): (column << 2) | flags
            {
                std::vector<int32_t> by_ord(n_ops, -1);
                for (int c = 0; c < N; ++c)
                    if (pl.contrib[c] >= 0) by_ord[pl.contrib[c] >> 2] = (c << 2) | (pl.contrib[c] & 3);
                for (int k = 0; k < n_ops; ++k) out.ints.push_back(by_ord[k]);
            }
            // columns no joint below this frame owns (structural zeros of the Jacobian): [count, columns...]
            op.pad = (int)out.ints.size();
            {
                std::vector<int32_t> zeros;
                for (int c = 0; c < N; ++c)
                    if (pl.contrib[c] < 0 && row_mode < ROWS_V_ACC) zeros.push_back(c);  // accumulated rows were zeroed by their fill op
                out.ints.push_back((int32_t)zeros.size());
                out.ints.insert(out.ints.end(), zeros.begin(), zeros.end());
            }
        }
        else if (em.kind == OP_EMIT_JLIMIT)
        {
            op.phi_off = m.start;
            op.jrow_off = m.start_jacobian;
            op.n_rows = N;
            out.dbls.push_back(m.safe_percentage);
            for (int c = 0; c < N; ++c) out.dbls.push_back(t.lower[c]);
            for (int c = 0; c < N; ++c) out.dbls.push_back(t.upper[c]);
        }
        else if (em.kind == OP_EMIT_JPOSE)
        {
            // JointPose::Update (joint_pose.cpp:36-57): phi_i = q[map_i] - ref_i, J(i, map_i) = 1
            op.phi_off = m.start;
            op.jrow_off = m.start_jacobian;
            op.n_rows = (int)m.joint_map.size();
            op.n_items = (int)m.joint_map.size();
            for (size_t i = 0; i < m.joint_map.size(); ++i)
            {
                out.ints.push_back(m.joint_map[i]);
                out.dbls.push_back(i < m.joint_ref.size() ? m.joint_ref[i] : 0.0);
            }
        }
        else if (em.kind == OP_EMIT_FILL)
        {
            // Phi.SetZero(): zeros with identity rotation encodings at EffFrame's Lie-group entries
            op.phi_off = m.start;
            op.jrow_off = m.start_jacobian;
            op.n_rows = m.length_jacobian;
            op.n_items = m.length;
            std::vector<double> fill(m.length, 0.0);
            if (m.kind == EXO_MAP_EFF_FRAME || m.kind == EXO_MAP_EFF_ORIENTATION)
            {
                const int rl = rotation_length(m.rot_type);
                const int lead = m.kind == EXO_MAP_EFF_FRAME ? 3 : 0;
                for (int f = 0; f < m.n_frames; ++f)
                {
                    double* r = &fill[f * (lead + rl) + lead];
                    if (m.rot_type == EXO_ROT_QUATERNION) r[3] = 1.0;
                    if (m.rot_type == EXO_ROT_MATRIX) r[0] = r[4] = r[8] = 1.0;
                }
            }
            for (double v : fill) out.dbls.push_back(v);
        }
        else if (em.kind == OP_EMIT_SPHERES)
        {
            const int S = m.n_frames, G = m.n_groups;
            const int n_acc = (int)pl.acc_elems.size();
            if (n_acc > 64)
            {
                err = "SphereCollision: more than 64 distinct moving links carry spheres";
                return false;
            }
            n_acc_max = std::max(n_acc_max, n_acc);
            op.phi_off = m.start;
            op.jrow_off = m.start_jacobian;
            op.n_rows = G * (G - 1) / 2;
            op.n_items = S;
            {
                // bit 0: first SphereCollision op of the program (it initialises the per-evaluation minimum distance)
                bool first = true;
                for (const Op& prev : out.ops) first = first && prev.kind != OP_EMIT_SPHERES;
                op.flags = first ? 1 : 0;
            }
            // ints: G, n_acc, group_start[G+1], per sphere {src, acc, group, 0}, per column {contrib, mask_lo, mask_hi}
            out.ints.push_back(G);
            out.ints.push_back(n_acc);
            std::vector<int> gs(G + 1, 0);
            for (int i = 0; i < S; ++i) gs[m.group[i] + 1]++;
            for (int g = 0; g < G; ++g) gs[g + 1] += gs[g];
            for (int g = 0; g <= G; ++g) out.ints.push_back(gs[g]);
            out.dbls.push_back(5.0 * (1.0 / m.precision));  // 5 * eps_, sphere_collision.cpp:39,74
            for (int k = 0; k < S; ++k)
            {
                const int i = pl.sphere_order[k];
                const int a = pl.sphere_anchor_elem[k];
                int acc = -1;
                if (a >= 0) acc = (int)(std::find(pl.acc_elems.begin(), pl.acc_elems.end(), a) - pl.acc_elems.begin());
                out.ints.push_back(patch_src(pl.sphere_src[k]));
                out.ints.push_back(acc);
                out.ints.push_back(m.group[i]);
                out.ints.push_back(0);
                const Frame off = frame_mul(C[m.elem_a[i]], frame_from12(&m.off_a[12 * i]));
                out.dbls.push_back(off.p[0]);
                out.dbls.push_back(off.p[1]);
                out.dbls.push_back(off.p[2]);
                out.dbls.push_back(m.radius[i]);
            }
            for (int c = 0; c < N; ++c)
            {
                const int v = patch_contrib(pl.contrib[c]);
                uint64_t mask = 0;
                if (v >= 0)
                {
                    // accumulators whose anchor descends from (or is) the joint of column c
                    int elem_c = -1;
                    for (int e = 0; e < n; ++e)
                        if (t.control_id[e] == c && kind[e] == K_JOINT && needed[e]) elem_c = e;
                    for (int a = 0; a < n_acc; ++a)
                    {
                        int e = pl.acc_elems[a];
                        while (e >= 0 && e != elem_c) e = t.parent[e];
                        if (e == elem_c) mask |= (uint64_t)1 << a;
                    }
                }
                out.ints.push_back(v);
                out.ints.push_back((int32_t)(mask & 0xffffffffu));
                out.ints.push_back((int32_t)(mask >> 32));
            }
        }
        out.ops.push_back(op);
    }
    {
        Op end;
        memset(&end, 0, sizeof(end));
        end.kind = OP_END;
        out.ops.push_back(end);
    }

    ProgramHeader& h = out.hdr;
    memset(&h, 0, sizeof(h));
    h.n_ops = (int)out.ops.size();
    h.n_int = (int)out.ints.size();
    h.n_dbl = (int)out.dbls.size();
    h.n_frame_slots = n_frame_slots;
    h.n_joint_slots = n_joint_slots;
    h.n_ctrl = N;
    h.length_phi = length_phi;
    h.length_jac = length_jac;
    h.n_generated = t.n_generated;
    h.n_acc = n_acc_max;
    h.two_pass = two_pass ? 1 : 0;
    h.n_pos_slots = n_pos_slots;
    h.scratch_doubles = 12 * n_frame_slots + (two_pass ? 3 * n_pos_slots : 6 * n_joint_slots);
    if (two_pass)
        for (int k = 0; k < n_ops; ++k)
        {
            Op& jop = out.ops[encoded_index[k]];
            jop.tab = (int)out.ints.size();
            out.ints.push_back((int32_t)(tp_entries[k].size() / TP_INTS));
            out.ints.insert(out.ints.end(), tp_entries[k].begin(), tp_entries[k].end());
        }
    for (const Op& op : out.ops)
        if (op.kind == OP_EMIT_SPHERES) h.max_spheres = std::max(h.max_spheres, op.n_items);
    // chain eligibility: [JointLimit / fill emits]* [joint fed by its predecessor]+ [frame emits reading the final frame]*,
    // nothing deferred to frame slots, frame B constant
    {
        bool chain = n_ops > 0 && n_frame_slots == 0;
        int ordinal = 0;
        for (const Op& op : out.ops)
        {
            if (op.kind == OP_JOINT)
            {
                if (op.src != (ordinal == 0 ? SRC_ROOT : SRC_CUR)) chain = false;
                ++ordinal;
            }
            else if (op.kind == OP_TABLE || op.kind == OP_EMIT_SPHERES || op.kind == OP_EMIT_JPOSE)
                chain = false;
            else if (op.kind == OP_EMIT_FRAME)
            {
                if (op.src != SRC_CUR || op.src_b != SRC_ROOT || ordinal != n_ops) chain = false;
                if (((op.flags & F_ROWS_MASK) >> F_ROWS_SHIFT) >= ROWS_WDOT) chain = false;  // EffAxisAlignment, CenterOfMass: interpreter only
            }
            else if (op.kind == OP_EMIT_JLIMIT || op.kind == OP_EMIT_FILL)
            {
                if (ordinal != 0) chain = false;
            }
        }
        h.chain_joints = chain ? n_ops : 0;
        h.chain_tab = (int)out.ints.size();
        for (int i = 0; i < (int)out.ops.size(); ++i)
            if (out.ops[i].kind == OP_JOINT || out.ops[i].kind == OP_TABLE) out.ints.push_back(i);
        h.n_int = (int)out.ints.size();
    }

    // image
    size_t off = sizeof(ProgramHeader) + out.ops.size() * sizeof(Op) + out.ints.size() * sizeof(int32_t);
    off = (off + 7) & ~(size_t)7;
    out.dbl_offset = (int)off;
    out.image.assign(off + out.dbls.size() * sizeof(double), 0);
    unsigned char* w = out.image.data();
    memcpy(w, &h, sizeof(h));
    w += sizeof(h);
    memcpy(w, out.ops.data(), out.ops.size() * sizeof(Op));
    w += out.ops.size() * sizeof(Op);
    if (!out.ints.empty()) memcpy(w, out.ints.data(), out.ints.size() * sizeof(int32_t));
    if (!out.dbls.empty()) memcpy(out.image.data() + off, out.dbls.data(), out.dbls.size() * sizeof(double));
    return true;
}

}  // namespace exo
