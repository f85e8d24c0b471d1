// tree_codegen.cpp -- turns a compiled tree program into the CUDA source of a kernel specialised for that robot and problem.
//
// The interpreter (kernels.cu) decodes the op list per evaluation: 29 k warp-instructions per 32 evaluations on the 38-DoF
// humanoid of BASELINE.json config 4, two thirds of them decode, address arithmetic and per-lane stores of structural zeros.
// Here the same two-pass program (compiler.cpp) is unrolled on the host into straight-line CUDA C++ and compiled at run time
// (NVRTC, tree_jit.cu):
//   * every transform constant is a literal and exact 0 / +-1 entries of the constant frames are folded away;
//   * one warp = one tile of 32 evaluations; Phi and the non-zero Jacobian columns are staged [row][32] in the warp's own
//     shared-memory buffers and leave through the TMA unit (2-D boxes for Phi, 3-D boxes [rows][columns][32] of J[LJ][N][E]);
//     no CTA-wide barrier after start-up;
//   * structural zeros of J (63 % of the bytes of config 4) are TMA boxes read from one shared zero tile, issued by one lane
//     before the arithmetic starts: no per-lane instruction is spent on them.
// Reference path replaced: KinematicTree::UpdateTree / UpdateFK / ComputeJ (exotica_core/src/kinematic_tree.cpp:695-737,
// 829-867,919-945) + EffFrame / EffPosition / JointLimit::Update + Phi.SetZero / jacobian.setZero.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>

#include "../../include/exotica_b200.h"
#include "compiler.h"
#include "kernels.cuh"

namespace exo
{
namespace
{
struct Emitter
{
    std::string s;
    void line(const char* fmt, ...) __attribute__((format(printf, 2, 3)))
    {
        va_list ap, aq;
        va_start(ap, fmt);
        va_copy(aq, ap);
        const int n = vsnprintf(nullptr, 0, fmt, ap);
        va_end(ap);
        std::string buf((size_t)(n > 0 ? n : 0) + 1, '\0');
        vsnprintf(&buf[0], buf.size(), fmt, aq);
        va_end(aq);
        buf.resize((size_t)(n > 0 ? n : 0));
        s += buf;
        s += '\n';
    }
};

std::string lit(double v)
{
    if (v == 0.0) return "0.0";
    char buf[64];
    snprintf(buf, sizeof(buf), "%.17g", v);
    std::string r = buf;
    if (r.find_first_of(".eEn") == std::string::npos) r += ".0";  // keep it a double literal
    if (v < 0) r = "(" + r + ")";
    return r;
}

// sum_k a_k * K_k (+ tail) with exact zeros dropped, +-1 folded and the rest as an fma chain; "0.0" when nothing is left
std::string dot_const(const std::vector<std::string>& a, const std::vector<double>& K, const std::string& tail)
{
    std::string acc = tail;
    for (size_t k = 0; k < a.size(); ++k)
    {
        if (K[k] == 0.0) continue;
        if (K[k] == 1.0)
            acc = acc.empty() ? a[k] : "(" + acc + " + " + a[k] + ")";
        else if (K[k] == -1.0)
            acc = acc.empty() ? "(-" + a[k] + ")" : "(" + acc + " - " + a[k] + ")";
        else
            acc = acc.empty() ? "(" + a[k] + " * " + lit(K[k]) + ")" : "fma(" + a[k] + ", " + lit(K[k]) + ", " + acc + ")";
    }
    return acc.empty() ? std::string("0.0") : acc;
}

const char* kM[9] = {"m0", "m1", "m2", "m3", "m4", "m5", "m6", "m7", "m8"};
const char* kP[3] = {"x", "y", "z"};

// dst = a * C for a constant frame C (12 doubles: rotation row-major, translation); a and dst may be the same variable
void emit_mul_const(Emitter& o, const char* dst, const char* a, const double* C)
{
    o.line("    {");
    for (int r = 0; r < 3; ++r)
    {
        const std::vector<std::string> row = {std::string(a) + "." + kM[3 * r], std::string(a) + "." + kM[3 * r + 1], std::string(a) + "." + kM[3 * r + 2]};
        for (int c = 0; c < 3; ++c) o.line("        const double t%d = %s;", 3 * r + c, dot_const(row, {C[c], C[3 + c], C[6 + c]}, "").c_str());
        o.line("        const double p%d = %s;", r, dot_const(row, {C[9], C[10], C[11]}, std::string(a) + "." + kP[r]).c_str());
    }
    for (int k = 0; k < 9; ++k) o.line("        %s.%s = t%d;", dst, kM[k], k);
    for (int k = 0; k < 3; ++k) o.line("        %s.%s = p%d;", dst, kP[k], k);
    o.line("    }");
}

// dst = C * a for a constant frame C (the inverse of a constant temp_B)
void emit_const_mul(Emitter& o, const char* dst, const double* C, const char* a)
{
    o.line("    {");
    for (int r = 0; r < 3; ++r)
    {
        for (int c = 0; c < 3; ++c)
        {
            const std::vector<std::string> col = {std::string(a) + "." + kM[c], std::string(a) + "." + kM[3 + c], std::string(a) + "." + kM[6 + c]};
            o.line("        const double t%d = %s;", 3 * r + c, dot_const(col, {C[3 * r], C[3 * r + 1], C[3 * r + 2]}, "").c_str());
        }
        const std::vector<std::string> p = {std::string(a) + ".x", std::string(a) + ".y", std::string(a) + ".z"};
        o.line("        const double p%d = %s;", r, dot_const(p, {C[3 * r], C[3 * r + 1], C[3 * r + 2]}, lit(C[9 + r])).c_str());
    }
    for (int k = 0; k < 9; ++k) o.line("        %s.%s = t%d;", dst, kM[k], k);
    for (int k = 0; k < 3; ++k) o.line("        %s.%s = p%d;", dst, kP[k], k);
    o.line("    }");
}

void emit_set_const(Emitter& o, const char* dst, const double* C)
{
    for (int k = 0; k < 9; ++k) o.line("    %s.%s = %s;", dst, kM[k], lit(C[k]).c_str());
    for (int k = 0; k < 3; ++k) o.line("    %s.%s = %s;", dst, kP[k], lit(C[9 + k]).c_str());
}

bool is_identity12(const double* C)
{
    static const double I[12] = {1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0};
    for (int k = 0; k < 12; ++k)
        if (C[k] != I[k]) return false;
    return true;
}

// dst = the frame `src` names: a slot of the tile's shared memory, the running frame, or the root (identity: nothing is emitted)
void emit_frame_source(Emitter& o, int src, const char* dst, int TILE, bool& is_identity)
{
    is_identity = false;
    if (src == SRC_ROOT)
        is_identity = true;
    else if (src >= 0)
    {
        for (int k = 0; k < 9; ++k) o.line("    %s.%s = slot[%d];", dst, kM[k], (12 * src + k) * TILE);
        for (int k = 0; k < 3; ++k) o.line("    %s.%s = slot[%d];", dst, kP[k], (12 * src + 9 + k) * TILE);
    }
    else if (strcmp(dst, "cur") != 0)
        o.line("    %s = cur;", dst);
}

// a constant frame of the program with its kind flag applied (none / translation only / full)
void const_frame(const double* c, int kind, int k_none, int k_trans, double out[12])
{
    static const double I[12] = {1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0};
    memcpy(out, c, 12 * sizeof(double));
    if (kind == k_none)
        memcpy(out, I, 12 * sizeof(double));
    else if (kind == k_trans)
        memcpy(out, I, 9 * sizeof(double));
}

// OP_JOINT, first half: cur = parent frame * constant pre-multiplication
void emit_joint_pre(Emitter& o, const Op& op, const double* c, int TILE)
{
    bool ident = false;
    emit_frame_source(o, op.src, "cur", TILE, ident);
    double prec[12];
    const_frame(c, op.flags & F_PRE_MASK, F_PRE_NONE, F_PRE_TRANS, prec);
    if (ident)
        emit_set_const(o, "cur", prec);
    else if (!is_identity12(prec))
        emit_mul_const(o, "cur", "cur", prec);
}

// OP_JOINT, second half (inside the block the caller opened, after `q` is defined): the joint's world axis (if wanted), its
// recorded axis / origin (one-pass programs, record_op >= 0) and the motion about / along z
void emit_joint_motion(Emitter& o, const Op& op, bool want_w, int record_op)
{
    if ((op.flags & F_PRISMATIC) || want_w) o.line("        const double wx = cur.m2, wy = cur.m5, wz = cur.m8;");
    if (record_op >= 0)
    {
        // the z column and the origin of the frame in front of the joint: its world axis and a point on it
        const int i = record_op;
        o.line("        jw%d_x = cur.m2; jw%d_y = cur.m5; jw%d_z = cur.m8; jp%d_x = cur.x; jp%d_y = cur.y; jp%d_z = cur.z;", i, i, i, i, i, i);
    }
    if (op.flags & F_PRISMATIC)
    {
        o.line("        cur.x = fma(wx, q, cur.x); cur.y = fma(wy, q, cur.y); cur.z = fma(wz, q, cur.z);");
    }
    else
    {
        o.line("        double sn, cs;");
        o.line("        SINCOS(q, sn, cs);");
        o.line("        const double a0 = cur.m0, a3 = cur.m3, a6 = cur.m6;");
        o.line("        cur.m0 = fma(cs, a0, sn * cur.m1); cur.m1 = fma(cs, cur.m1, -(sn * a0));");
        o.line("        cur.m3 = fma(cs, a3, sn * cur.m4); cur.m4 = fma(cs, cur.m4, -(sn * a3));");
        o.line("        cur.m6 = fma(cs, a6, sn * cur.m7); cur.m7 = fma(cs, cur.m7, -(sn * a6));");
    }
}

void emit_slot_store(Emitter& o, int dst, int TILE)
{
    if (dst < 0) return;
    for (int k = 0; k < 9; ++k) o.line("    slot[%d] = cur.%s;", (12 * dst + k) * TILE, kM[k]);
    for (int k = 0; k < 3; ++k) o.line("    slot[%d] = cur.%s;", (12 * dst + 9 + k) * TILE, kP[k]);
}

// OP_EMIT_FRAME: A = frame * offset (its origin parked in pa<slot>), then A = Binv * A for a constant base (A is then Phi)
void emit_requested_frame(Emitter& o, const Op& op, const double* c, int TILE)
{
    o.line("    Fr A;");
    bool ident = false;
    emit_frame_source(o, op.src, "A", TILE, ident);
    double offc[12];
    const_frame(c + CE_OFFA_M, op.flags & F_OFFA_MASK, F_OFFA_NONE, F_OFFA_TRANS, offc);
    if (ident)
        emit_set_const(o, "A", offc);
    else if (!is_identity12(offc))
        emit_mul_const(o, "A", "A", offc);
    o.line("    pa%d_x = A.x; pa%d_y = A.y; pa%d_z = A.z;", op.n_items, op.n_items, op.n_items);
    if ((op.flags & F_B_MASK) == F_B_CONST) emit_const_mul(o, "A", c + CE_B_M, "A");
}

// what every generated kernel starts with
void emit_prelude(Emitter& o)
{
    o.line("#define EXO_STORE_CS 1");
    o.line("#define EXO_INLINE_ENCODERS 1");
    o.line("#include \"device_math.cuh\"");
    o.line("using namespace exo;");
    o.line("using namespace exo::dev;");
    o.line("static __device__ __noinline__ double2 slow_sincos(double q) { double s, c; sincos(q, &s, &c); return make_double2(s, c); }");
    o.line("#define SINCOS(q, s, c) do { if (fabs(q) <= 1.0e5) fast_sincos((q), (s), (c)); else { const double2 sc_ = slow_sincos(q); (s) = sc_.x; (c) = sc_.y; } } while (0)");
}

struct Item  // one block of values a lane stages: the Phi entries of a frame, or one Jacobian column of one frame
{
    int op;               // index of the op that produces it
    bool phi;
    int nrows;            // Phi entries / Jacobian rows
    int key, col;         // J: first Jacobian row of the frame, column
    int ps, flags, cst_b; // J: position slot, TP_F_* flags, constant offset of Binv
    int box = -1, col_in_box = 0;
    int batch = -1;       // index of its batch within its pass
    int jop = -1;         // J: the joint op whose axis / origin the column is built from
    int row_mode = ROWS_VW;  // J: which rows of the column the map keeps (program.h RowMode)
};
struct Box
{
    bool phi;
    int nrows, ncols;
    int off;        // first staging row
    int c_col, c_row;
    int map = -1;
};
struct Batch
{
    std::vector<int> items;
    std::vector<Box> boxes;
    int last_item = -1;  // flushed right after this item has been staged
};

int env_int(const char* name, int dflt, int lo, int hi)
{
    const char* v = getenv(name);
    int r = v ? atoi(v) : dflt;
    if (r < lo) r = lo;
    if (r > hi) r = hi;
    return r;
}
}  // namespace

bool tree_codegen(const Program& prog, TreeGenPlan& plan, std::string& why, const TreeCostSpec* cost)
{
    const ProgramHeader& H = prog.hdr;
    if (!H.two_pass)
    {
        why = "the program has no two-pass form (SphereCollision, Distance or state-dependent base frames)";
        return false;
    }
    if (H.n_generated > 0)
    {
        why = "trajectory-driven links";
        return false;
    }
    const int N = H.n_ctrl;
    const int n_ops = (int)prog.ops.size();
    // ---- eligibility + the items every op stages ------------------------------------------------------------------------------
    std::vector<Item> items;
    std::vector<std::vector<int>> op_items(n_ops);
    int n_pos = 0, max_op_values = 0, jl_ops = 0;
    for (int i = 0; i < n_ops; ++i)
    {
        const Op& op = prog.ops[i];
        if (op.kind == OP_END) break;
        if (op.kind == OP_JOINT)
        {
            const int32_t* tp = &prog.ints[op.tab];
            int vals = 0;
            for (int k = 0; k < tp[0]; ++k)
            {
                const int32_t* ent = tp + 1 + TP_INTS * k;
                if (ent[3] != ROWS_VW && ent[3] != ROWS_V && ent[3] != ROWS_W && ent[3] != ROWS_VXY)
                {
                    why = "a map keeps Jacobian rows other than (v, w) / (v) / (w) / (vx, vy)";
                    return false;
                }
                Item it;
                it.op = i; it.phi = false; it.nrows = ent[3] == ROWS_VW ? 6 : (ent[3] == ROWS_VXY ? 2 : 3);
                it.row_mode = ent[3];
                it.key = ent[2]; it.col = ent[1]; it.ps = ent[0]; it.flags = ent[4]; it.cst_b = ent[5];
                op_items[i].push_back((int)items.size());
                items.push_back(it);
                vals += it.nrows;
                max_op_values = std::max(max_op_values, it.nrows);
            }
        }
        else if (op.kind == OP_EMIT_FRAME)
        {
            const int row_mode = (op.flags & F_ROWS_MASK) >> F_ROWS_SHIFT;
            if ((row_mode != ROWS_VW && row_mode != ROWS_V && row_mode != ROWS_W && row_mode != ROWS_VXY) || (op.flags & F_B_MASK) == F_B_SLOT)
            {
                why = "a requested frame is outside the generated kernel's scope";
                return false;
            }
            int rows = 0;
            if (op.map_kind == EXO_MAP_EFF_POSITION) rows = 3;
            else if (op.map_kind == EXO_MAP_EFF_POSITION_XY) rows = 2;
            else if (op.map_kind == EXO_MAP_EFF_ORIENTATION && rotation_length(op.rot_type) > 0) rows = rotation_length(op.rot_type);
            else if (op.map_kind == EXO_MAP_FRAMES) rows = 12;
            else if (op.map_kind == EXO_MAP_EFF_FRAME && rotation_length(op.rot_type) > 0) rows = 3 + rotation_length(op.rot_type);
            else
            {
                why = "a task map is outside the generated kernel's scope";
                return false;
            }
            if (cost && ((op.map_kind != EXO_MAP_EFF_FRAME && op.map_kind != EXO_MAP_EFF_POSITION) || (row_mode != ROWS_VW && row_mode != ROWS_V) ||
                         op.n_items >= (int)cost->frame_map.size() || cost->frame_map[op.n_items] < 0))
            {
                why = "a requested frame is outside the cost epilogue's scope";
                return false;
            }
            Item it;
            it.op = i; it.phi = true; it.nrows = rows; it.key = op.phi_off; it.col = 0; it.ps = op.n_items; it.flags = 0; it.cst_b = 0;
            op_items[i].push_back((int)items.size());
            items.push_back(it);
            n_pos = std::max(n_pos, op.n_items + 1);
            max_op_values = std::max(max_op_values, rows);
        }
        else if (op.kind != OP_EMIT_JLIMIT && op.kind != OP_EMIT_FILL && op.kind != OP_EMIT_JPOSE)
        {
            why = "an op is outside the generated kernel's scope";
            return false;
        }
        else if (cost && op.kind == OP_EMIT_JLIMIT)
            jl_ops++;
    }
    if (cost && cost->jl_map >= 0 && jl_ops != 1)
    {
        why = "the cost names a JointLimit map the program does not evaluate exactly once";
        return false;
    }
    if (n_pos > 16)
    {
        why = "more than 16 requested frames";
        return false;
    }
    // One pass when the axes and origins of all joints fit in registers (serial arms): the columns of a frame are staged right
    // after its Phi, from the recorded (w, p) of the joints below it.  Larger trees recompose the tree for the Jacobian pass.
    int n_joint_ops = 0;
    for (int i = 0; i < n_ops && prog.ops[i].kind != OP_END; ++i) n_joint_ops += prog.ops[i].kind == OP_JOINT;
    const bool one_pass = n_joint_ops <= env_int("EXO_TREE_ONEPASS_MAX", 8, 0, 16);
    std::vector<char> joint_recorded(n_ops, 0);
    if (one_pass)
    {
        std::vector<int> frame_op_of_ps(n_pos, -1);
        for (int i = 0; i < n_ops && prog.ops[i].kind != OP_END; ++i)
            if (prog.ops[i].kind == OP_EMIT_FRAME) frame_op_of_ps[prog.ops[i].n_items] = i;
        for (int i = 0; i < n_ops && prog.ops[i].kind != OP_END; ++i)
        {
            if (prog.ops[i].kind != OP_JOINT) continue;
            for (int itx : op_items[i])
            {
                items[itx].jop = i;
                op_items[frame_op_of_ps[items[itx].ps]].push_back(itx);
                joint_recorded[i] = 1;
            }
            op_items[i].clear();
        }
    }
    else
        for (int i = 0; i < n_ops && prog.ops[i].kind != OP_END; ++i)
            for (int itx : op_items[i]) items[itx].jop = i;
    // ---- geometry ---------------------------------------------------------------------------------------------------------------
    const int TILE = env_int("EXO_TREE_TILE", 32, 32, 64) >= 64 ? 64 : 32;  // evaluations per tile: one warp, or a pair of warps (512-byte rows)
    int WPC = env_int("EXO_TREE_WPC", 4, 1, 8);                  // warps per CTA
    if (TILE == 64 && (WPC & 1)) ++WPC;
    const int TPC = WPC * 32 / TILE;                             // tiles in flight per CTA
    int CAP = env_int("EXO_TREE_CAP", 24, 3, 96);                // staging rows per buffer
    const int ZC = env_int("EXO_TREE_ZC", 4, 1, 8);              // columns of the zero tile
    if (CAP < max_op_values) CAP = max_op_values;
    if (CAP > 96)
    {
        why = "one frame has more entries than the staging buffer holds";
        return false;
    }
    // ---- batches: ops are taken in program order (pass 1: frame emits, pass 2: joints) until the buffer is full --------------------
    std::vector<Batch> batches[2];
    for (int pass = 0; pass < 2; ++pass)
    {
        Batch cur;
        int used = 0;
        for (int i = 0; i < n_ops; ++i)
        {
            const Op& op = prog.ops[i];
            if (op.kind == OP_END) break;
            if ((pass == 0) != (op.kind == OP_EMIT_FRAME)) continue;
            if (op.kind != OP_EMIT_FRAME && op.kind != OP_JOINT) continue;
            for (int it : op_items[i])
            {
                const bool kind_change = !cur.items.empty() && items[cur.items.back()].phi != items[it].phi;
                if ((used + items[it].nrows > CAP || kind_change) && !cur.items.empty())
                {
                    batches[pass].push_back(cur);
                    cur = Batch();
                    used = 0;
                }
                items[it].batch = (int)batches[pass].size();
                cur.items.push_back(it);
                cur.last_item = it;
                used += items[it].nrows;
            }
        }
        if (!cur.items.empty()) batches[pass].push_back(cur);
    }
    // boxes: Phi -> one 2-D box per frame; J -> per frame, runs of consecutive columns become one 3-D box [rows][columns][32]
    plan.maps.clear();
    auto map_id = [&](int kind, int ncols, int nrows) {  // kind: 0 J [rows][columns][32], 1 Phi [rows][32], 2 J flattened [LJ * N rows][32]
        for (size_t m = 0; m < plan.maps.size(); ++m)
            if (plan.maps[m].phi == kind && plan.maps[m].ncols == ncols && plan.maps[m].nrows == nrows) return (int)m;
        TreeGenPlan::MapShape sh;
        sh.phi = kind; sh.ncols = ncols; sh.nrows = nrows;
        plan.maps.push_back(sh);
        return (int)plan.maps.size() - 1;
    };
    for (int pass = 0; pass < 2; ++pass)
        for (Batch& b : batches[pass])
        {
            int off = 0;
            if (items[b.items[0]].phi)
            {
                for (int it : b.items)
                {
                    Box bx;
                    bx.phi = true; bx.nrows = items[it].nrows; bx.ncols = 1; bx.off = off; bx.c_col = 0; bx.c_row = items[it].key;
                    bx.map = map_id(1, 1, bx.nrows);
                    items[it].box = (int)b.boxes.size();
                    b.boxes.push_back(bx);
                    off += bx.nrows;
                }
                continue;
            }
            std::vector<int> order = b.items;
            std::sort(order.begin(), order.end(), [&](int a, int c) { return items[a].key != items[c].key ? items[a].key < items[c].key : items[a].col < items[c].col; });
            for (size_t k = 0; k < order.size();)
            {
                size_t k1 = k + 1;
                while (k1 < order.size() && items[order[k1]].key == items[order[k]].key && items[order[k1]].nrows == items[order[k]].nrows &&
                       items[order[k1]].col == items[order[k1 - 1]].col + 1)
                    ++k1;
                Box bx;
                bx.phi = false; bx.nrows = items[order[k]].nrows; bx.ncols = (int)(k1 - k); bx.off = off; bx.c_col = items[order[k]].col;
                bx.c_row = items[order[k]].key;
                bx.map = map_id(0, bx.ncols, bx.nrows);
                for (size_t q = k; q < k1; ++q)
                {
                    items[order[q]].box = (int)b.boxes.size();
                    items[order[q]].col_in_box = (int)(q - k);
                }
                b.boxes.push_back(bx);
                off += bx.nrows * bx.ncols;
                k = k1;
            }
        }
    // structural zeros: per frame emit, runs of columns no joint owns -> boxes [rows][<= ZC columns][32] from the zero tile
    struct ZBox
    {
        int map, c_col, c_row;  // c_col < 0: box of the flattened view J[LJ * N][E], c_row its first flattened row
    };
    std::vector<ZBox> zboxes;
    int z_rows_max = 0;
    const int ZFLAT = env_int("EXO_TREE_ZFLAT", 24, 1, 64);  // rows of the flattened zero boxes
    bool needs_flat = false;
    auto flat_zero_run = [&](int row0, int len) {  // flattened rows [row0, row0 + len) are zero
        if (len <= 0) return;
        const int h = std::min(len, ZFLAT);
        const int id = map_id(2, 1, h);
        for (int r = 0; r < len; r += h) zboxes.push_back({id, -1, row0 + std::min(r, len - h)});
        z_rows_max = std::max(z_rows_max, h);
        needs_flat = true;
    };
    for (int i = 0; i < n_ops; ++i)
    {
        const Op& op = prog.ops[i];
        if (op.kind == OP_END) break;
        if (op.kind == OP_EMIT_JLIMIT)
        {
            // rows jrow_off + i, i < N: the diagonal entry is a per-lane store, the N flattened rows between two diagonals are zero
            for (int r = 0; r + 1 < N; ++r) flat_zero_run((op.jrow_off + r) * N + r + 1, N);
            continue;
        }
        if (op.kind == OP_EMIT_FILL)
        {
            flat_zero_run(op.jrow_off * N, op.n_rows * N);
            continue;
        }
        if (op.kind == OP_EMIT_JPOSE)
        {
            // row i holds a single one in column map[i] (a per-lane store): zeros in front of it and behind it
            for (int r = 0; r < op.n_items; ++r)
            {
                const int c1 = prog.ints[op.tab + r];
                flat_zero_run((op.jrow_off + r) * N, c1);
                flat_zero_run((op.jrow_off + r) * N + c1 + 1, N - c1 - 1);
            }
            continue;
        }
        if (op.kind != OP_EMIT_FRAME) continue;
        const int nrows = op.n_rows;
        const int32_t* zl = &prog.ints[op.pad];
        for (int k = 1; k <= zl[0];)
        {
            int k1 = k + 1;
            while (k1 <= zl[0] && zl[k1] == zl[k1 - 1] + 1) ++k1;
            const int c0 = zl[k], len = k1 - k;
            const int w = std::min(len, ZC);
            const int id = map_id(0, w, nrows);
            for (int c = 0; c < len; c += w) zboxes.push_back({id, c0 + std::min(c, len - w), op.jrow_off});
            z_rows_max = std::max(z_rows_max, w * nrows);
            k = k1;
        }
    }
    plan.needs_flat_jacobian = needs_flat;
    if ((int)plan.maps.size() > TREE_MAXMAPS)
    {
        why = "more distinct TMA box shapes than the kernel parameter holds";
        return false;
    }
    const int ZROWS = std::max(z_rows_max, 1);
    const int n_slots = H.n_frame_slots;
    const int per_warp = 2 * CAP + 12 * n_slots;  // per tile: staging rows (two buffers) + frame slots, in rows of TILE doubles
    plan.warps_per_cta = WPC;
    plan.tile = TILE;
    plan.smem_bytes = (size_t)(ZROWS + TPC * per_warp) * TILE * sizeof(double);
    plan.n_ctrl = N;
    plan.length_phi = H.length_phi;
    plan.length_jac = H.length_jac;
    if (plan.smem_bytes > 200 * 1024)
    {
        why = "the program needs more shared memory per warp than a CTA can hold";
        return false;
    }

    // ---- source ---------------------------------------------------------------------------------------------------------------------
    Emitter o;
    o.line("// generated by exotica_b200 tree_codegen: %d ops, %d controlled joints, %d frame slots, %d staging rows x 2, %d zero rows, %d TMA shapes", n_ops, N,
           n_slots, CAP, ZROWS, (int)plan.maps.size());
    emit_prelude(o);
    o.line("struct alignas(64) TmaDesc { unsigned long long v[16]; };");
    o.line("struct TreeGenParams { const double* x; double* phi; double* jac; long long n_evals, x_sk, phi_sk, j_sr, j_sc; TmaDesc maps[%d];", TREE_MAXMAPS);
    o.line("    double* cost_eval; const double* rho; const double* goal; const double* goal_rot; const double* W; double ct; int T, t_first, T_goal, n_maps, length_phi, n_rot; };");
    plan.kernel_name = cost ? "exo_tree_update_cost_kernel" : "exo_tree_kernel";
    o.line("static __device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }");
    o.line("static __device__ __forceinline__ void tma2(const TmaDesc* m, const void* s, int c0, int c1, unsigned long long pol)");
    o.line("{ asm volatile(\"cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%%0, {%%2, %%3}], [%%1], %%4;\" :: \"l\"(m), \"r\"(smem_u32(s)), \"r\"(c0), \"r\"(c1), \"l\"(pol) : \"memory\"); }");
    o.line("static __device__ __forceinline__ void tma3(const TmaDesc* m, const void* s, int c0, int c1, int c2, unsigned long long pol)");
    o.line("{ asm volatile(\"cp.async.bulk.tensor.3d.global.shared::cta.bulk_group.L2::cache_hint [%%0, {%%2, %%3, %%4}], [%%1], %%5;\" :: \"l\"(m), \"r\"(smem_u32(s)), \"r\"(c0), \"r\"(c1), \"r\"(c2), \"l\"(pol) : \"memory\"); }");
    o.line("static __device__ __forceinline__ void tma_commit() { asm volatile(\"cp.async.bulk.commit_group;\" ::: \"memory\"); }");
    o.line("static __device__ __forceinline__ void tma_wait_read() { asm volatile(\"cp.async.bulk.wait_group.read 0;\" ::: \"memory\"); }");
    o.line("static __device__ __forceinline__ void fence_staged() { asm volatile(\"fence.proxy.async.shared::cta;\" ::: \"memory\"); }");
    const int min_blocks = env_int("EXO_TREE_MINB", 0, 0, 8);
    if (min_blocks)
        o.line("extern \"C\" __global__ void __launch_bounds__(%d, %d) %s(const __grid_constant__ TreeGenParams P)", WPC * 32, min_blocks, plan.kernel_name);
    else
        o.line("extern \"C\" __global__ void __launch_bounds__(%d) %s(const __grid_constant__ TreeGenParams P)", WPC * 32, plan.kernel_name);
    o.line("{");
    o.line("    extern __shared__ __align__(128) double smem[];");
    o.line("    const int lane = threadIdx.x & %d, warp = threadIdx.x / %d;  // position in the tile, tile of the CTA", TILE - 1, TILE);
    if (TILE == 32)
        o.line("#define TILE_SYNC() __syncwarp()");
    else
        o.line("#define TILE_SYNC() asm volatile(\"bar.sync %%0, 64;\" ::\"r\"(1 + warp) : \"memory\")  // the two warps of this tile");
    o.line("    for (int i = threadIdx.x; i < %d; i += %d) smem[i] = 0.0;  // the zero tile", ZROWS * TILE, WPC * 32);
    o.line("    fence_staged();");
    o.line("    __syncthreads();");
    o.line("    const long long e0 = ((long long)blockIdx.x * %d + warp) * %d;", TPC, TILE);
    o.line("    if (e0 >= P.n_evals) return;");
    o.line("    const bool valid = e0 + lane < P.n_evals;");
    o.line("    const long long e = valid ? e0 + lane : P.n_evals - 1;  // lanes past the batch redo the last evaluation; the TMA unit clips them");
    o.line("    const int e0i = (int)e0;");
    o.line("    double* const wsm = smem + %d + warp * %d;", ZROWS * TILE, per_warp * TILE);
    o.line("    double* const stg0 = wsm + lane;                 // staging buffer 0: row r at stg0[r * %d]", TILE);
    o.line("    double* const stg1 = wsm + %d + lane;", CAP * TILE);
    if (n_slots > 0) o.line("    double* const slot = wsm + %d + lane;            // frame slot s, word k at slot[(12 * s + k) * %d]", 2 * CAP * TILE, TILE);
    o.line("    const double* const xq = P.x + e;");
    o.line("    unsigned long long pol = 0;");
    o.line("    if (lane == 0) asm volatile(\"createpolicy.fractional.L2::evict_first.b64 %%0, 1.0;\" : \"=l\"(pol));");
    o.line("    Fr cur;");
    if (cost)
    {
        // cost epilogue: ct (ydiff^T S ydiff + xdiff^T W xdiff) of this evaluation, as exo_cost_kernel forms it from Phi in HBM
        // (TimeIndexedTask::Update, TaskSpaceVector::operator-, GetScalarTaskCost / GetScalarTransitionCost)
        o.line("    const int t = (int)(e %% P.T);");
        o.line("    const double* const rho = P.rho + (long long)t * P.n_maps;");
        o.line("    const int tg = P.T_goal > 1 ? t : 0;");
        o.line("    const double* const goal = P.goal + (long long)tg * P.length_phi;");
        o.line("    const double* const grot = P.goal_rot + (long long)tg * P.n_rot * 9;");
        o.line("    double c = 0.0;");
        if (cost->jl_map >= 0) o.line("    const double rho_jl = rho[%d];", cost->jl_map);
    }
    for (int k = 0; k < n_pos; ++k) o.line("    double pa%d_x = 0.0, pa%d_y = 0.0, pa%d_z = 0.0;", k, k, k);
    for (int i = 0; i < n_ops; ++i)
        if (joint_recorded[i]) o.line("    double jw%d_x, jw%d_y, jw%d_z, jp%d_x, jp%d_y, jp%d_z;  // axis and origin of the joint of op %d", i, i, i, i, i, i, i);

    // J batches in execution order: zero boxes ride along with their flushes in equal shares (one lane issues them: no per-lane
    // work at all)
    int n_jbatches = 0;
    for (int pass = 0; pass < 2; ++pass)
        for (const Batch& b : batches[pass]) n_jbatches += items[b.items[0]].phi ? 0 : 1;
    int jb_seen = 0;
    auto emit_zbox = [&](const ZBox& z, const char* ind) {
        if (z.c_col < 0)
            o.line("%stma2(&P.maps[%d], smem, e0i, %d, pol);", ind, z.map, z.c_row);
        else
            o.line("%stma3(&P.maps[%d], smem, e0i, %d, %d, pol);", ind, z.map, z.c_col, z.c_row);
    };
    int flip = 0;  // staging buffer of the first batch of the pass being emitted
    auto emit_flush = [&](const Batch& b, int buf) {
        const bool jb = !items[b.items[0]].phi;
        o.line("    if (lane == 0) tma_wait_read();  // the other staging buffer has been read out");
        o.line("    fence_staged();");
        o.line("    TILE_SYNC();");
        o.line("    if (lane == 0)");
        o.line("    {");
        for (const Box& bx : b.boxes)
        {
            if (bx.phi)
                o.line("        tma2(&P.maps[%d], wsm + %d, e0i, %d, pol);", bx.map, (buf * CAP + bx.off) * TILE, bx.c_row);
            else
                o.line("        tma3(&P.maps[%d], wsm + %d, e0i, %d, %d, pol);", bx.map, (buf * CAP + bx.off) * TILE, bx.c_col, bx.c_row);
        }
        if (jb)
        {
            const size_t z0 = zboxes.size() * jb_seen / n_jbatches, z1 = zboxes.size() * (jb_seen + 1) / n_jbatches;
            for (size_t z = z0; z < z1; ++z) emit_zbox(zboxes[z], "        ");
            ++jb_seen;
        }
        o.line("        tma_commit();");
        o.line("    }");
    };
    // Jacobian column of one (joint, frame) pair into its staging rows -- KinematicTree::ComputeJ (kinematic_tree.cpp:919-945);
    // w / p name the joint's axis and origin (registers of the current joint, or the recorded ones in one-pass mode)
    auto emit_column = [&](const Item& it, int pass, const char* wn, const char* pn) {
        const Batch& bt = batches[pass][it.batch];
        const Box& bx = bt.boxes[it.box];
        const int ibuf = (flip + it.batch) & 1;
        const char* stg = ibuf ? "stg1" : "stg0";
        o.line("        {");
        if (it.flags & TP_F_PRISMATIC)
            o.line("            double v0 = %sx, v1 = %sy, v2 = %sz, w0 = 0.0, w1 = 0.0, w2 = 0.0;", wn, wn, wn);
        else
        {
            o.line("            const double dx = pa%d_x - %sx, dy = pa%d_y - %sy, dz = pa%d_z - %sz;", it.ps, pn, it.ps, pn, it.ps, pn);
            o.line("            double v0 = %sy * dz - %sz * dy, v1 = %sz * dx - %sx * dz, v2 = %sx * dy - %sy * dx, w0 = %sx, w1 = %sy, w2 = %sz;", wn, wn, wn, wn, wn, wn, wn,
                   wn, wn);
        }
        if (it.flags & TP_F_BCONST)
        {
            const double* b = prog.dbls.data() + it.cst_b;
            for (int r = 0; r < 3; ++r) o.line("            const double tv%d = %s;", r, dot_const({"v0", "v1", "v2"}, {b[3 * r], b[3 * r + 1], b[3 * r + 2]}, "").c_str());
            o.line("            v0 = tv0; v1 = tv1; v2 = tv2;");
            if (it.row_mode == ROWS_VW || it.row_mode == ROWS_W)
            {
                for (int r = 0; r < 3; ++r)
                    o.line("            const double tw%d = %s;", r, dot_const({"w0", "w1", "w2"}, {b[3 * r], b[3 * r + 1], b[3 * r + 2]}, "").c_str());
                o.line("            w0 = tw0; w1 = tw1; w2 = tw2;");
            }
        }
        // EffFrame / raw frames: (v, w); EffPosition: v; EffOrientation: w; EffPositionXY: (vx, vy)
        const char* vals_vw[6] = {"v0", "v1", "v2", "w0", "w1", "w2"};
        const char* vals_w[3] = {"w0", "w1", "w2"};
        const char* const* vals = it.row_mode == ROWS_W ? vals_w : vals_vw;
        for (int r = 0; r < it.nrows; ++r) o.line("            %s[%d] = %s;", stg, (bx.off + r * bx.ncols + it.col_in_box) * TILE, vals[r]);
        o.line("            (void)v0; (void)v1; (void)v2; (void)w0; (void)w1; (void)w2;");
        o.line("        }");
        if (bt.last_item == (int)(&it - items.data())) emit_flush(bt, ibuf);
    };

    // joint values are loaded AHEAD joints before they are used (the asm statements of the TMA path pin loads in place, and a
    // load issued where it is needed costs its full latency: ~1 us from DRAM while the memory system is saturated with stores)
    const int AHEAD = env_int("EXO_TREE_AHEAD", 8, 0, 16);
    const int n_passes = one_pass ? 1 : 2;
    std::vector<std::pair<int, int>> qseq;  // (pass, op) of every joint that reads the state, in execution order
    for (int pass = 0; pass < n_passes; ++pass)
        for (int i = 0; i < n_ops && prog.ops[i].kind != OP_END; ++i)
            if (prog.ops[i].kind == OP_JOINT && prog.ops[i].qidx >= 0) qseq.push_back({pass, i});
    size_t q_issued = 0;
    auto issue_loads = [&](size_t upto, int cur_pass) {  // emit the loads of qseq[q_issued .. upto)
        for (; q_issued < upto && q_issued < qseq.size(); ++q_issued)
        {
            if (qseq[q_issued].first != cur_pass) break;  // pass-2 values are loaded inside the pass-2 scope
            o.line("    const double q_%d = xq[%d * P.x_sk];", (int)q_issued, prog.ops[qseq[q_issued].second].qidx);
        }
    };
    size_t q_used = 0;
    for (int pass = 0; pass < n_passes; ++pass)
    {
        if (pass == 1) o.line("    if (P.jac)\n    {");
        o.line("    // ---- pass %d: %s", pass,
               one_pass ? "compose the chain, record every joint's axis and origin, Phi and Jacobian columns of every requested frame"
                        : (pass == 0 ? "compose the tree, Phi of every requested frame, park the frame origins" : "compose it again, Jacobian columns of every joint"));
        issue_loads(q_used + AHEAD + 1, pass);
        for (int i = 0; i < n_ops; ++i)
        {
            const Op& op = prog.ops[i];
            if (op.kind == OP_END) break;
            const double* c = prog.dbls.data() + op.cst;
            if (op.kind == OP_JOINT)
            {
                o.line("    // op %d: joint, state %d", i, op.qidx);
                emit_joint_pre(o, op, c, TILE);
                o.line("    {");
                if (op.qidx >= 0)
                {
                    issue_loads(q_used + 1, pass);  // AHEAD == 0, or the first joints of a pass
                    if (c[CJ_QMULT] == 1.0 && c[CJ_QOFF] == 0.0)
                        o.line("        const double q = q_%d;", (int)q_used);
                    else
                        o.line("        const double q = fma(%s, q_%d, %s);", lit(c[CJ_QMULT]).c_str(), (int)q_used, lit(c[CJ_QOFF]).c_str());
                    ++q_used;
                }
                else
                    o.line("        const double q = %s;", lit(c[CJ_QOFF]).c_str());
                const bool columns_now = pass == 1 && !op_items[i].empty();
                emit_joint_motion(o, op, columns_now, joint_recorded[i] ? i : -1);
                if (columns_now)
                    for (int itx : op_items[i]) emit_column(items[itx], 1, "w", "cur.");
                o.line("    }");
                emit_slot_store(o, op.dst, TILE);
                issue_loads(q_used + AHEAD + 1, pass);
            }
            else if (op.kind == OP_EMIT_FRAME && pass == 0)
            {
                o.line("    // op %d: requested frame -> Phi[%d..], position slot %d", i, op.phi_off, op.n_items);
                o.line("    {");
                emit_requested_frame(o, op, c, TILE);
                const int itx = op_items[i][0];
                const Item& it = items[itx];
                const Batch& bt = batches[0][it.batch];
                const Box& bx = bt.boxes[it.box];
                const int ibuf = (flip + it.batch) & 1;
                o.line("    write_phi(%d, %d, A, %s + %d, %d);", op.map_kind, op.rot_type, ibuf ? "stg1" : "stg0", bx.off * TILE, TILE);
                if (cost)
                {
                    const int ps = op.n_items;
                    o.line("    {");
                    o.line("        const double y0 = A.x - goal[%d], y1 = A.y - goal[%d], y2 = A.z - goal[%d];", op.phi_off, op.phi_off + 1, op.phi_off + 2);
                    o.line("        double s = y0 * y0 + y1 * y1 + y2 * y2;");
                    if (cost->frame_rot[ps] >= 0)
                    {
                        // rotvec = M1 * (M2^-1 M1).GetRot()  (TaskSpaceVector::operator-, task_space_vector.cpp:64-91)
                        o.line("        const double* M2 = grot + %d;", 9 * cost->frame_rot[ps]);
                        o.line("        Fr M;");
                        for (int r = 0; r < 3; ++r)
                            for (int k = 0; k < 3; ++k) o.line("        M.m%d = M2[%d] * A.m%d + M2[%d] * A.m%d + M2[%d] * A.m%d;", 3 * r + k, r, k, 3 + r, 3 + k, 6 + r, 6 + k);
                        o.line("        M.x = M.y = M.z = 0.0;");
                        o.line("        double vx, vy, vz;");
                        o.line("        get_angle_axis(M, vx, vy, vz);");
                        o.line("        const double y3 = A.m0 * vx + A.m1 * vy + A.m2 * vz, y4 = A.m3 * vx + A.m4 * vy + A.m5 * vz, y5 = A.m6 * vx + A.m7 * vy + A.m8 * vz;");
                        o.line("        s += y3 * y3 + y4 * y4 + y5 * y5;");
                    }
                    o.line("        c += rho[%d] * s;", cost->frame_map[ps]);
                    o.line("    }");
                }
                o.line("    }");
                if (bt.last_item == itx) emit_flush(bt, ibuf);
                if (op_items[i].size() > 1)
                {
                    // one-pass mode: the frame's Jacobian columns from the recorded joint axes / origins
                    o.line("    if (P.jac)");
                    o.line("    {");
                    for (size_t k = 1; k < op_items[i].size(); ++k)
                    {
                        const Item& ji = items[op_items[i][k]];
                        char wn[32], pn[32];
                        snprintf(wn, sizeof(wn), "jw%d_", ji.jop);
                        snprintf(pn, sizeof(pn), "jp%d_", ji.jop);
                        emit_column(ji, 0, wn, pn);
                    }
                    o.line("    }");
                    o.line("    else");
                    o.line("    {");
                    o.line("        if (lane == 0) tma_wait_read();  // no Jacobian batches in between: the next Phi block reuses a buffer");
                    o.line("        TILE_SYNC();");
                    o.line("    }");
                }
            }
            else if (op.kind == OP_EMIT_JLIMIT && pass == 0)
            {
                // JointLimit::Update (joint_limit.cpp:63-99): per-lane streaming stores of phi and of the diagonal of its Jacobian block
                // (the zeros between the diagonal entries are TMA boxes of the zero tile)
                o.line("    // op %d: JointLimit", i);
                o.line("    if (valid)");
                o.line("    {");
                const double safe = c[0];
                for (int k = 0; k < N; ++k)
                {
                    const double lo = c[1 + k], hi = c[1 + N + k];
                    const double tau = (0.5 * safe) * (hi - lo);
                    const double lo_t = lo + tau, hi_t = hi - tau;
                    o.line("        {");
                    o.line("            const double q = xq[%d * P.x_sk];", k);
                    o.line("            double v = 0.0;");
                    o.line("            if (q < %s) v = q - %s - %s;", lit(lo_t).c_str(), lit(lo).c_str(), lit(tau).c_str());
                    o.line("            if (q > %s) v = q - %s + %s;", lit(hi_t).c_str(), lit(hi).c_str(), lit(tau).c_str());
                    if (cost && cost->jl_map >= 0) o.line("            { const double yd = v - goal[%d]; c += rho_jl * yd * yd; }", op.phi_off + k);
                    o.line("            __stcs(P.phi + %d * P.phi_sk + e, v);", op.phi_off + k);
                    o.line("            if (P.jac) __stcs(P.jac + %d * P.j_sr + %d * P.j_sc + e, (q >= %s && q <= %s) ? 0.0 : 1.0);", op.jrow_off + k, k, lit(lo_t).c_str(),
                           lit(hi_t).c_str());
                    o.line("        }");
                }
                o.line("    }");
            }
            else if (op.kind == OP_EMIT_JPOSE && pass == 0)
            {
                // JointPose::Update (joint_pose.cpp:36-57): phi_i = q[map_i] - ref_i, J(i, map_i) = 1 -- per-lane streaming stores; the zeros
                // around the ones are TMA boxes of the zero tile
                o.line("    // op %d: JointPose", i);
                o.line("    if (valid)");
                o.line("    {");
                for (int k = 0; k < op.n_items; ++k)
                {
                    const int c1 = prog.ints[op.tab + k];
                    o.line("        __stcs(P.phi + %d * P.phi_sk + e, xq[%d * P.x_sk] - %s);", op.phi_off + k, c1, lit(c[k]).c_str());
                    o.line("        if (P.jac) __stcs(P.jac + %d * P.j_sr + %d * P.j_sc + e, 1.0);", op.jrow_off + k, c1);
                }
                o.line("    }");
            }
            else if (op.kind == OP_EMIT_FILL && pass == 0)
            {
                // unused map: Phi.SetZero() values per lane; its zero Jacobian rows are TMA boxes of the zero tile
                o.line("    // op %d: unused map", i);
                o.line("    if (valid)");
                o.line("    {");
                for (int k = 0; k < op.n_items; ++k) o.line("        __stcs(P.phi + %d * P.phi_sk + e, %s);", op.phi_off + k, lit(c[k]).c_str());
                o.line("    }");
            }
        }
        flip = (int)((flip + batches[pass].size()) & 1);
        if (pass == 1) o.line("    }");
    }
    if (cost)
    {
        o.line("    // transition cost and the evaluation's total (the first timesteps, t < t_first, carry no cost: GetCost sums t = 1 .. T - 1)");
        o.line("    if (valid)");
        o.line("    {");
        o.line("        if (P.W && t > 0)");
        o.line("        {");
        for (int n = 0; n < N; ++n) o.line("            { const double d = xq[%d * P.x_sk] - xq[%d * P.x_sk - 1]; c += P.W[%d] * d * d; }", n, n, n);
        o.line("        }");
        o.line("        P.cost_eval[e] = t < P.t_first ? 0.0 : P.ct * c;");
        o.line("    }");
    }
    if (jb_seen == 0 && !zboxes.empty())
    {
        o.line("    if (lane == 0 && P.jac)");
        o.line("    {");
        for (const ZBox& z : zboxes) emit_zbox(z, "        ");
        o.line("        tma_commit();");
        o.line("    }");
    }
    o.line("    if (lane == 0) tma_wait_read();  // shared memory must outlive the TMA reads");
    o.line("}");
    plan.source.swap(o.s);
    return true;
}

// ---- fused forward kinematics + task cost + cost gradient for tree programs ---------------------------------------------------------
// The generated counterpart of exo_chain_costgrad_kernel (chain_kernel.cu): grad[e][n] = 2 ct (J^T S ydiff)[n] + the JointLimit and
// transition terms (AbstractTimeIndexedProblem::GetCostJacobian, abstract_time_indexed_problem.cpp:309-342; T = 1:
// UnconstrainedEndPoseProblem::GetScalarJacobian) without ever materialising J.  Pass 1 composes the tree, forms ydiff of every
// requested frame (SO(3) log for EffFrame: TaskSpaceVector::operator-, task_space_vector.cpp:64-91), adds its cost and keeps
// S ydiff (turned by Binv^T for a constant base) in registers; pass 2 composes the tree again and dots every joint's column with
// S ydiff of the frames below it -- the joint owns exactly one gradient entry, which is complete and stored on the spot.  HBM
// traffic per evaluation: N doubles in, N + 1 out.  The two passes are two copies of the straight-line code (9 k instructions for
// the 38-DoF humanoid; ncu shows `no_instruction` stalls of 3.3 cycles per issue): sharing one copy through a run-time pass loop
// with warp-uniform `if (pass == ...)` tests was measured slower (0.265 against 0.238 ms), so the copies stay.
bool tree_costgrad_codegen(const Program& prog, const TreeCostSpec& cs, TreeCostPlan& plan, std::string& why)
{
    const ProgramHeader& H = prog.hdr;
    if (!H.two_pass || H.n_generated > 0)
    {
        why = "the program has no two-pass form";
        return false;
    }
    const int N = H.n_ctrl;
    const int n_ops = (int)prog.ops.size();
    int n_pos = 0, jl_op = -1;
    for (int i = 0; i < n_ops && prog.ops[i].kind != OP_END; ++i)
    {
        const Op& op = prog.ops[i];
        if (op.kind == OP_EMIT_FRAME)
        {
            const int row_mode = (op.flags & F_ROWS_MASK) >> F_ROWS_SHIFT;
            if ((row_mode != ROWS_VW && row_mode != ROWS_V) || (op.flags & F_B_MASK) == F_B_SLOT ||
                (op.map_kind != EXO_MAP_EFF_FRAME && op.map_kind != EXO_MAP_EFF_POSITION))
            {
                why = "a requested frame is outside the generated cost kernel's scope";
                return false;
            }
            n_pos = std::max(n_pos, op.n_items + 1);
        }
        else if (op.kind == OP_EMIT_JLIMIT)
            jl_op = i;
        else if (op.kind != OP_JOINT && op.kind != OP_EMIT_FILL)
        {
            why = "an op is outside the generated cost kernel's scope";
            return false;
        }
    }
    if (n_pos > 16 || (int)cs.frame_map.size() < n_pos || (int)cs.frame_rot.size() < n_pos)
    {
        why = "more requested frames than the generated cost kernel keeps in registers";
        return false;
    }
    if (cs.jl_map >= 0 && jl_op < 0)
    {
        why = "the cost names a JointLimit map the program does not evaluate";
        return false;
    }
    const int WPC = env_int("EXO_TREE_WPC", 4, 1, 8);
    const int AHEAD = env_int("EXO_TREE_AHEAD", 8, 0, 16);
    const int TILE = 32;
    const int n_slots = H.n_frame_slots;
    plan.warps_per_cta = WPC;
    plan.smem_bytes = (size_t)WPC * 12 * n_slots * TILE * sizeof(double);

    Emitter o;
    o.line("// generated by exotica_b200 tree_costgrad_codegen: %d ops, %d controlled joints, %d frame slots", n_ops, N, n_slots);
    emit_prelude(o);
    o.line("struct TreeCostParams { const double* x; double* grad; double* cost_eval; const double* rho; const double* goal; const double* goal_rot; const double* W;");
    o.line("    long long n_evals, x_se, x_sk, g_se, g_sk; double ct; int T, t_first, T_goal, n_maps, length_phi, n_rot; };");
    o.line("extern \"C\" __global__ void __launch_bounds__(%d) exo_tree_costgrad_kernel(const __grid_constant__ TreeCostParams P)", WPC * 32);
    o.line("{");
    if (n_slots > 0) o.line("    extern __shared__ __align__(16) double smem[];");
    o.line("    const long long e = (long long)blockIdx.x * %d + threadIdx.x;", WPC * 32);
    o.line("    if (e >= P.n_evals) return;");
    if (n_slots > 0) o.line("    double* const slot = smem + (threadIdx.x >> 5) * %d + (threadIdx.x & 31);  // frame slot s, word k at slot[(12 * s + k) * 32]", 12 * n_slots * TILE);
    o.line("    const double* const xq = P.x + e * P.x_se;");
    o.line("    const int t = (int)(e %% P.T);");
    o.line("    const bool skip = t < P.t_first;  // GetCost sums t = 1 .. T - 1: the first timestep carries neither cost nor gradient");
    o.line("    const double* const rho = P.rho + (long long)t * P.n_maps;");
    o.line("    const int tg = P.T_goal > 1 ? t : 0;");
    o.line("    const double* const goal = P.goal + (long long)tg * P.length_phi;");
    o.line("    const double* const grot = P.goal_rot + (long long)tg * P.n_rot * 9;");
    o.line("    double* const ge = P.grad ? P.grad + e * P.g_se : nullptr;");
    o.line("    double c = 0.0;");
    o.line("    Fr cur;");
    for (int k = 0; k < n_pos; ++k)
    {
        o.line("    double pa%d_x = 0.0, pa%d_y = 0.0, pa%d_z = 0.0;", k, k, k);
        o.line("    double s%d_0 = 0.0, s%d_1 = 0.0, s%d_2 = 0.0, s%d_3 = 0.0, s%d_4 = 0.0, s%d_5 = 0.0;  // S ydiff of the frame (turned by Binv^T)", k, k, k, k, k, k);
    }
    if (cs.jl_map >= 0) o.line("    const double rho_jl = rho[%d];", cs.jl_map);

    // JointLimit and transition terms of column n (both diagonal in the joint index): their share of the cost goes into c, their
    // share of the gradient (before the factor 2 ct) into `s`
    const double* jlc = jl_op >= 0 ? prog.dbls.data() + prog.ops[jl_op].cst : nullptr;
    const int jl_phi_off = jl_op >= 0 ? prog.ops[jl_op].phi_off : 0;
    auto emit_extra = [&](int n, const char* ind) {
        o.line("%sconst double xn = xq[%d * P.x_sk];", ind, n);
        o.line("%sdouble s = 0.0;", ind);
        if (cs.jl_map >= 0)
        {
            const double safe = jlc[0], lo = jlc[1 + n], hi = jlc[1 + N + n];
            const double tau = (0.5 * safe) * (hi - lo);
            const double lo_t = lo + tau, hi_t = hi - tau;
            o.line("%s{", ind);
            o.line("%s    double v = 0.0;", ind);
            o.line("%s    if (xn < %s) v = xn - %s - %s;", ind, lit(lo_t).c_str(), lit(lo).c_str(), lit(tau).c_str());
            o.line("%s    if (xn > %s) v = xn - %s + %s;", ind, lit(hi_t).c_str(), lit(hi).c_str(), lit(tau).c_str());
            o.line("%s    const double yd = v - goal[%d];", ind, jl_phi_off + n);
            o.line("%s    c += rho_jl * yd * yd;", ind);
            o.line("%s    if (!(xn >= %s && xn <= %s)) s += rho_jl * yd;", ind, lit(lo_t).c_str(), lit(hi_t).c_str());
            o.line("%s}", ind);
        }
        o.line("%sif (P.W && t > 0)", ind);
        o.line("%s{", ind);
        o.line("%s    const double w = P.W[%d];", ind, n);
        o.line("%s    const double d0 = xn - xq[%d * P.x_sk - P.x_se];", ind, n);
        o.line("%s    c += w * d0 * d0;", ind);
        o.line("%s    s += w * d0;", ind);
        o.line("%s    if (t + 1 < P.T) s -= w * (xq[%d * P.x_sk + P.x_se] - xn);", ind, n);
        o.line("%s}", ind);
    };

    std::vector<std::pair<int, int>> qseq;
    for (int pass = 0; pass < 2; ++pass)
        for (int i = 0; i < n_ops && prog.ops[i].kind != OP_END; ++i)
            if (prog.ops[i].kind == OP_JOINT && prog.ops[i].qidx >= 0) qseq.push_back({pass, i});
    size_t q_issued = 0, q_used = 0;
    auto issue_loads = [&](size_t upto, int cur_pass) {
        for (; q_issued < upto && q_issued < qseq.size(); ++q_issued)
        {
            if (qseq[q_issued].first != cur_pass) break;
            o.line("    const double q_%d = xq[%d * P.x_sk];", (int)q_issued, prog.ops[qseq[q_issued].second].qidx);
        }
    };
    std::vector<char> owned(N, 0);  // columns some joint of the program owns
    for (int pass = 0; pass < 2; ++pass)
    {
        if (pass == 1) o.line("    if (ge)\n    {");
        o.line("    // ---- pass %d: %s", pass, pass == 0 ? "compose the tree; ydiff, cost and S ydiff of every requested frame" : "compose it again; every joint's gradient entry");
        issue_loads(q_used + AHEAD + 1, pass);
        for (int i = 0; i < n_ops; ++i)
        {
            const Op& op = prog.ops[i];
            if (op.kind == OP_END) break;
            const double* c = prog.dbls.data() + op.cst;
            if (op.kind == OP_JOINT)
            {
                const int32_t* tp = &prog.ints[op.tab];
                o.line("    // op %d: joint, state %d", i, op.qidx);
                emit_joint_pre(o, op, c, TILE);
                o.line("    {");
                if (op.qidx >= 0)
                {
                    issue_loads(q_used + 1, pass);
                    if (c[CJ_QMULT] == 1.0 && c[CJ_QOFF] == 0.0)
                        o.line("        const double q = q_%d;", (int)q_used);
                    else
                        o.line("        const double q = fma(%s, q_%d, %s);", lit(c[CJ_QMULT]).c_str(), (int)q_used, lit(c[CJ_QOFF]).c_str());
                    ++q_used;
                }
                else
                    o.line("        const double q = %s;", lit(c[CJ_QOFF]).c_str());
                const bool dots = pass == 1 && tp[0] > 0;
                emit_joint_motion(o, op, dots, -1);
                if (dots)
                {
                    // column . S ydiff for every frame below the joint: (w x d) . yv + w . yw (revolute), w . yv (prismatic)
                    const int col = tp[2];
                    owned[col] = 1;
                    o.line("        double g = 0.0;");
                    for (int k = 0; k < tp[0]; ++k)
                    {
                        const int32_t* ent = tp + 1 + TP_INTS * k;
                        const int ps = ent[0];
                        if (ent[1] != col)
                        {
                            why = "a joint feeds more than one Jacobian column";
                            return false;
                        }
                        if (ent[4] & TP_F_PRISMATIC)
                            o.line("        g += wx * s%d_0 + wy * s%d_1 + wz * s%d_2;", ps, ps, ps);
                        else
                        {
                            o.line("        {");
                            o.line("            const double dx = pa%d_x - cur.x, dy = pa%d_y - cur.y, dz = pa%d_z - cur.z;", ps, ps, ps);
                            o.line("            g += (wy * dz - wz * dy) * s%d_0 + (wz * dx - wx * dz) * s%d_1 + (wx * dy - wy * dx) * s%d_2;", ps, ps, ps);
                            if (ent[3] == ROWS_VW) o.line("            g += wx * s%d_3 + wy * s%d_4 + wz * s%d_5;", ps, ps, ps);
                            o.line("        }");
                        }
                    }
                    o.line("        {");
                    emit_extra(col, "            ");
                    o.line("            ge[%d * P.g_sk] = skip ? 0.0 : 2.0 * P.ct * (g + s);", col);
                    o.line("        }");
                }
                o.line("    }");
                emit_slot_store(o, op.dst, TILE);
                issue_loads(q_used + AHEAD + 1, pass);
            }
            else if (op.kind == OP_EMIT_FRAME && pass == 0)
            {
                const int ps = op.n_items;
                const bool six = cs.frame_rot[ps] >= 0;
                o.line("    // op %d: requested frame %d, cost map %d", i, ps, cs.frame_map[ps]);
                o.line("    {");
                emit_requested_frame(o, op, c, TILE);  // A = Phi
                o.line("    double y0 = A.x - goal[%d], y1 = A.y - goal[%d], y2 = A.z - goal[%d], y3 = 0.0, y4 = 0.0, y5 = 0.0;", op.phi_off, op.phi_off + 1, op.phi_off + 2);
                if (six)
                {
                    // rotvec = M1 * (M2^-1 M1).GetRot()  (TaskSpaceVector::operator-, task_space_vector.cpp:64-91)
                    o.line("    {");
                    o.line("        const double* M2 = grot + %d;", 9 * cs.frame_rot[ps]);
                    o.line("        Fr M;");
                    for (int r = 0; r < 3; ++r)
                        for (int k = 0; k < 3; ++k) o.line("        M.m%d = M2[%d] * A.m%d + M2[%d] * A.m%d + M2[%d] * A.m%d;", 3 * r + k, r, k, 3 + r, 3 + k, 6 + r, 6 + k);
                    o.line("        M.x = M.y = M.z = 0.0;");
                    o.line("        double vx, vy, vz;");
                    o.line("        get_angle_axis(M, vx, vy, vz);");
                    o.line("        y3 = A.m0 * vx + A.m1 * vy + A.m2 * vz;");
                    o.line("        y4 = A.m3 * vx + A.m4 * vy + A.m5 * vz;");
                    o.line("        y5 = A.m6 * vx + A.m7 * vy + A.m8 * vz;");
                    o.line("    }");
                }
                o.line("    const double r = rho[%d];", cs.frame_map[ps]);
                o.line("    c += r * (y0 * y0 + y1 * y1 + y2 * y2 + y3 * y3 + y4 * y4 + y5 * y5);");
                o.line("    y0 *= r; y1 *= r; y2 *= r; y3 *= r; y4 *= r; y5 *= r;  // S ydiff");
                if ((op.flags & F_B_MASK) == F_B_CONST)
                {
                    // column = Binv [v; w]  ->  dot against Binv^T (S ydiff)
                    const double* b = c + CE_B_M;
                    for (int k = 0; k < 3; ++k) o.line("    s%d_%d = %s;", ps, k, dot_const({"y0", "y1", "y2"}, {b[k], b[3 + k], b[6 + k]}, "").c_str());
                    for (int k = 0; k < 3; ++k) o.line("    s%d_%d = %s;", ps, 3 + k, dot_const({"y3", "y4", "y5"}, {b[k], b[3 + k], b[6 + k]}, "").c_str());
                }
                else
                    o.line("    s%d_0 = y0; s%d_1 = y1; s%d_2 = y2; s%d_3 = y3; s%d_4 = y4; s%d_5 = y5;", ps, ps, ps, ps, ps, ps);
                o.line("    }");
            }
        }
        if (pass == 1)
        {
            // columns no joint of the program owns: JointLimit and transition terms only
            for (int n = 0; n < N; ++n)
            {
                if (owned[n]) continue;
                o.line("    {");
                emit_extra(n, "        ");
                o.line("        ge[%d * P.g_sk] = skip ? 0.0 : 2.0 * P.ct * s;", n);
                o.line("    }");
            }
            o.line("    }");
            o.line("    else");
            o.line("    {");
            o.line("    // cost only: the JointLimit and transition shares of every column");
            for (int n = 0; n < N; ++n)
            {
                o.line("    {");
                emit_extra(n, "        ");
                o.line("        (void)s;");
                o.line("    }");
            }
            o.line("    }");
        }
    }
    o.line("    if (P.cost_eval) P.cost_eval[e] = skip ? 0.0 : P.ct * c;");
    o.line("}");
    plan.source.swap(o.s);
    return true;
}

// ---- front half of split SphereCollision programs ----------------------------------------------------------------------------------
// exo_spheres_kernel (kernels.cu) runs the sphere-pair loops one warp per evaluation from a row of parked frame / joint slots.  The
// interpreter used to produce those rows (0.20 ms of cfg5's 2.07 ms at 6 warps per SM: its 1 KB of per-lane scratch limits
// residency).  For programs made of joints, trajectory-driven links and SphereCollision maps only, this generator unrolls the same
// one-pass program into straight-line code that keeps nothing but the running frame in registers and writes every slot the moment
// it exists, into TILED rows ([sphere op][e / 32][word][e % 32]: every store of a warp is one full 256-byte segment;
// exo_spheres_kernel gathers its evaluation's words 8 bytes at a time, KernelParams::sph_tiled).  Evaluation-major rows written in
// pieces were measured no faster than the interpreter's transposed dump (180 vs 203 us of cfg5's 2.07 ms): partial lines.
bool tree_park_codegen(const Program& prog, TreeParkPlan& plan, std::string& why)
{
    const ProgramHeader& H = prog.hdr;
    if (H.two_pass)
    {
        why = "two-pass program";
        return false;
    }
    const int n_ops = (int)prog.ops.size();
    const int joint_base = 12 * H.n_frame_slots;
    // which producer is live in each slot at every SphereCollision op, and which slots later producers read back
    std::vector<int> owner_f(H.n_frame_slots, -1), owner_j(H.n_joint_slots, -1);
    std::vector<std::vector<int>> rows_f(n_ops), rows_j(n_ops);  // per producer op: the sphere-op ordinals whose row takes its frame / joint slot
    std::vector<char> read_back(H.n_frame_slots, 0);
    int n_sph = 0;
    for (int i = 0; i < n_ops && prog.ops[i].kind != OP_END; ++i)
    {
        const Op& op = prog.ops[i];
        if (op.kind == OP_JOINT || op.kind == OP_TABLE)
        {
            if (op.src >= 0) read_back[op.src] = 1;
            if (op.dst >= 0) owner_f[op.dst] = i;
            if (op.kind == OP_JOINT && op.jslot >= 0) owner_j[op.jslot] = i;
        }
        else if (op.kind == OP_EMIT_SPHERES)
        {
            for (int s = 0; s < H.n_frame_slots; ++s)
                if (owner_f[s] >= 0) rows_f[owner_f[s]].push_back(n_sph);
            for (int s = 0; s < H.n_joint_slots; ++s)
                if (owner_j[s] >= 0) rows_j[owner_j[s]].push_back(n_sph);
            ++n_sph;
        }
        else
        {
            why = "maps other than SphereCollision share the program";
            return false;
        }
    }
    if (n_sph == 0)
    {
        why = "no SphereCollision map";
        return false;
    }
    const int AHEAD = env_int("EXO_TREE_AHEAD", 8, 0, 16);
    plan.threads = 128;
    Emitter o;
    o.line("// generated by exotica_b200 tree_park_codegen: %d ops, %d frame slots, %d joint slots, %d SphereCollision maps", n_ops, H.n_frame_slots, H.n_joint_slots, n_sph);
    emit_prelude(o);
    o.line("struct TreeParkParams { const double* x; const double* generated; double* dump; long long n_evals, x_se, x_sk, e_base, sph_block; int T, n_generated, sph_row; };");
    o.line("extern \"C\" __global__ void __launch_bounds__(%d) exo_tree_park_kernel(const __grid_constant__ TreeParkParams P)", plan.threads);
    o.line("{");
    o.line("    const long long e = (long long)blockIdx.x * %d + threadIdx.x;", plan.threads);
    o.line("    if (e >= P.n_evals) return;");
    o.line("    const double* const xq = P.x + e * P.x_se;");
    o.line("    // tiled rows: word k of evaluation e at [e / 32][k][e %% 32] -- every store of a warp is one full 256-byte segment");
    o.line("    double* const row = P.dump + (e & ~31ll) * P.sph_row + (e & 31);  // sphere op k: + k * P.sph_block; word w: + 32 * w");
    o.line("    Fr cur;");
    for (int s = 0; s < H.n_frame_slots; ++s)
        if (read_back[s]) o.line("    Fr fs%d;  // frame slot %d: a later joint starts from it", s, s);
    std::vector<int> qseq;
    for (int i = 0; i < n_ops && prog.ops[i].kind != OP_END; ++i)
        if (prog.ops[i].kind == OP_JOINT && prog.ops[i].qidx >= 0) qseq.push_back(i);
    size_t q_issued = 0, q_used = 0;
    auto issue_loads = [&](size_t upto) {
        for (; q_issued < upto && q_issued < qseq.size(); ++q_issued) o.line("    const double q_%d = xq[%d * P.x_sk];", (int)q_issued, prog.ops[qseq[q_issued]].qidx);
    };
    issue_loads(AHEAD + 1);
    for (int i = 0; i < n_ops; ++i)
    {
        const Op& op = prog.ops[i];
        if (op.kind == OP_END) break;
        if (op.kind != OP_JOINT && op.kind != OP_TABLE) continue;
        const double* c = prog.dbls.data() + op.cst;
        o.line("    // op %d: %s", i, op.kind == OP_JOINT ? "joint" : "trajectory-driven link");
        // cur = parent * pre  (the parent is the running frame, the root, or a slot kept in registers)
        {
            double prec[12];
            const_frame(c, op.flags & F_PRE_MASK, F_PRE_NONE, F_PRE_TRANS, prec);
            if (op.src == SRC_ROOT)
                emit_set_const(o, "cur", prec);
            else
            {
                if (op.src >= 0) o.line("    cur = fs%d;", op.src);
                if (!is_identity12(prec)) emit_mul_const(o, "cur", "cur", prec);
            }
        }
        o.line("    {");
        if (op.kind == OP_TABLE)
        {
            o.line("        const double* g = P.generated + ((long long)((P.e_base + e) %% P.T) * P.n_generated + %d) * 12;", op.qidx);
            o.line("        Fr gf;");
            for (int k = 0; k < 9; ++k) o.line("        gf.%s = g[%d];", kM[k], k);
            for (int k = 0; k < 3; ++k) o.line("        gf.%s = g[%d];", kP[k], 9 + k);
            o.line("        cur = fr_mul(cur, gf);");
        }
        else
        {
            if (op.qidx >= 0)
            {
                issue_loads(q_used + 1);
                if (c[CJ_QMULT] == 1.0 && c[CJ_QOFF] == 0.0)
                    o.line("        const double q = q_%d;", (int)q_used);
                else
                    o.line("        const double q = fma(%s, q_%d, %s);", lit(c[CJ_QMULT]).c_str(), (int)q_used, lit(c[CJ_QOFF]).c_str());
                ++q_used;
            }
            else
                o.line("        const double q = %s;", lit(c[CJ_QOFF]).c_str());
            const bool park_joint = !rows_j[i].empty();
            emit_joint_motion(o, op, park_joint, -1);
            // joint slot: world axis (the z column in front of the motion) and origin (after a sliding joint's translation)
            for (int k : rows_j[i])
            {
                o.line("        {");
                o.line("            double* js = row + %d * P.sph_block + %d;", k, 32 * (joint_base + 6 * op.jslot));
                o.line("            js[0] = wx; js[32] = wy; js[64] = wz; js[96] = cur.x; js[128] = cur.y; js[160] = cur.z;");
                o.line("        }");
            }
        }
        o.line("    }");
        for (int k : rows_f[i])
        {
            o.line("    {");
            o.line("        double* fsl = row + %d * P.sph_block + %d;", k, 32 * 12 * op.dst);
            for (int w = 0; w < 9; ++w) o.line("        fsl[%d] = cur.%s;", 32 * w, kM[w]);
            for (int w = 0; w < 3; ++w) o.line("        fsl[%d] = cur.%s;", 32 * (9 + w), kP[w]);
            o.line("    }");
        }
        if (op.dst >= 0 && read_back[op.dst]) o.line("    fs%d = cur;", op.dst);
        issue_loads(q_used + AHEAD + 1);
    }
    o.line("}");
    plan.source.swap(o.s);
    return true;
}

}  // namespace exo
