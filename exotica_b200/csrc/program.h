// program.h -- the straight-line "tree program" the host compiler emits and the CUDA interpreter executes.
//
// The reference walks a pointer tree per configuration (KinematicTree::UpdateTree/ComputeJ,
// exotica_core/src/kinematic_tree.cpp:695-737,919-945).  Here the tree, the frame requests and the task
// maps of one problem are compiled ONCE on the host into a short sequence of ops over a small set of
// per-thread scratch slots; every lane of a warp executes the same op on its own configuration, so control
// flow is warp-uniform and all constants are shared-memory broadcasts.
//
// Canonical joint form (compiler.cpp): for a movable element m with nearest movable ancestor a,
//     G[m] = G[a] * C_pre(m) * J(u, q)        J = Rot(u,q) (revolute) | Trans(u*q) (prismatic)
// where C_pre folds every constant transform between a and m (fixed joints, uncontrolled joints at their
// model state, the joint origin and the tip rotation) and u is the joint axis in the joint's own frame.
// With this form the Jacobian column of a controlled joint m needs only  w = G[m].M * u  and  p = G[m].p.
#pragma once
#include <stdint.h>

namespace exo
{
enum OpKind : int32_t
{
    OP_END = 0,
    OP_JOINT = 1,        // G = src * pre * J(u, q)
    OP_TABLE = 2,        // G = src * pre * table[t][slot]      (trajectory-generated element)
    OP_EMIT_FRAME = 3,   // EffFrame / EffPosition / raw frame outputs of ONE requested frame
    OP_EMIT_JLIMIT = 4,  // JointLimit
    OP_EMIT_FILL = 5,    // unused map: Phi.SetZero() identity encodings + zero Jacobian rows
    OP_EMIT_SPHERES = 6, // SphereCollision
    OP_EMIT_JPOSE = 7    // JointPose
};

// frame sources
enum : int32_t
{
    SRC_ROOT = -2,  // identity
    SRC_CUR = -1    // the register-resident output of the most recent OP_JOINT / OP_TABLE
};

// OP_JOINT flags
enum : int32_t
{
    F_PRE_NONE = 0,
    F_PRE_TRANS = 1,  // pre = pure translation
    F_PRE_FULL = 2,   // pre = rotation + translation
    F_PRE_MASK = 3,
    F_AXIS_X = 0 << 2,
    F_AXIS_Y = 1 << 2,
    F_AXIS_Z = 2 << 2,
    F_AXIS_GENERAL = 3 << 2,
    F_AXIS_MASK = 3 << 2,
    F_AXIS_NEG = 1 << 4,    // axis-aligned, negative direction
    F_PRISMATIC = 1 << 5,
    F_CONST_Q = 1 << 6      // q does not depend on the state (should have been folded; kept for safety)
};

// OP_EMIT_FRAME flags
enum : int32_t
{
    F_OFFA_NONE = 0,
    F_OFFA_TRANS = 1,
    F_OFFA_FULL = 2,
    F_OFFA_MASK = 3,
    F_B_IDENTITY = 0 << 2,  // temp_B = identity
    F_B_CONST = 1 << 2,     // temp_B constant: consts hold its INVERSE
    F_B_SLOT = 2 << 2,      // temp_B = slot * offB
    F_B_MASK = 3 << 2,
    // which rows of the frame's 6 x N Jacobian the map keeps (bits 4-6)
    F_ROWS_SHIFT = 4,
    F_ROWS_MASK = 7 << 4
};
enum RowMode : int32_t
{
    ROWS_VW = 0,    // 6 rows: linear then angular velocity (EffFrame, raw frames)
    ROWS_V = 1,     // 3 rows: linear (EffPosition)
    ROWS_W = 2,     // 3 rows: angular (EffOrientation)
    ROWS_VXY = 3,   // 2 rows: x and y of the linear part (EffPositionXY)
    ROWS_DIST = 4,  // 1 row: (p . J_v) / |p| (Distance)
    ROWS_WDOT = 5,  // 1 row: J_w . k with k = (Phi.M a) x direction (EffAxisAlignment)
    ROWS_V_ACC = 6, // 3 rows: += weight * linear part (CenterOfMass; the rows are zeroed by a fill op first)
    ROWS_VXY_ACC = 7  // 2 rows: the same without z (CenterOfMass, EnableZ = false)
};

// One op = 16 int32 (64 B) + a slice of the double constant pool.
struct Op
{
    int32_t kind;
    int32_t flags;
    int32_t src;      // OP_JOINT/TABLE: parent frame source.  EMIT_FRAME: frame A source
    int32_t dst;      // OP_JOINT/TABLE: slot to store G (-1: keep in registers only)
    int32_t jslot;    // OP_JOINT: joint-data slot (w, p) or -1
    int32_t qidx;     // OP_JOINT: control id feeding q (-1: constant);  OP_TABLE: generated slot
    int32_t cst;      // offset into the double pool
    int32_t src_b;    // EMIT_FRAME: frame B source
    int32_t map_kind; // EMIT_*: exo_map_kind
    int32_t rot_type; // EMIT_FRAME / EMIT_FILL
    int32_t phi_off;  // first Phi entry written
    int32_t jrow_off; // first Jacobian row written
    int32_t n_rows;   // EMIT_FILL: Jacobian rows to zero;  EMIT_SPHERES: number of group pairs
    int32_t tab;      // offset into the int pool (contributor table, fill table, sphere tables)
    int32_t n_items;  // EMIT_FILL: number of Phi entries; EMIT_SPHERES: number of spheres; EMIT_FRAME: position slot (two-pass)
    int32_t pad;      // EMIT_FRAME: offset into the int pool of [count, columns without a contributor...]
};

// OP_JOINT double constants
enum
{
    CJ_PRE_M = 0,   // 9
    CJ_PRE_P = 9,   // 3
    CJ_AXIS = 12,   // 3
    CJ_QMULT = 15,  // q = mult * x[qidx] + off
    CJ_QOFF = 16,
    CJ_SIZE = 17
};
// OP_TABLE double constants: pre (12)
enum
{
    CT_SIZE = 12
};
// OP_EMIT_FRAME double constants
enum
{
    CE_OFFA_M = 0,
    CE_OFFA_P = 9,
    CE_B_M = 12,  // F_B_CONST: inverse of temp_B;  F_B_SLOT: offB
    CE_B_P = 21,
    CE_SIZE = 24
};
// EMIT_FRAME int table: [N] contributor entries by Jacobian column, then [n_joint_ops] entries by joint-op ordinal.
// column entry: -1 none, else (jslot << 2) | (prismatic << 1) | negative
// ordinal entry: -1 none, else (column << 2) | (prismatic << 1) | negative
#ifndef __CUDACC_RTC__  // host helper: hidden from the run-time compiled kernels, which include this header for the enums
static inline int32_t contrib_pack(int jslot, bool prismatic, bool negative) { return (jslot << 2) | (prismatic ? 2 : 0) | (negative ? 1 : 0); }
#endif

// OP_EMIT_SPHERES int table layout (per sphere, 4 ints): anchor frame source, anchor accumulator index (-1 if
// the sphere has no controlled ancestor), group, offset of its constants (3 offset doubles + radius) -- then
// per Jacobian column: n_anchor-bitmask words follow (see compiler.cpp).
enum
{
    SPH_INTS = 4
};

struct ProgramHeader
{
    int32_t n_ops;
    int32_t n_int;        // ints in the int pool
    int32_t n_dbl;        // doubles in the constant pool
    int32_t n_frame_slots;
    int32_t n_joint_slots;
    int32_t n_ctrl;       // N
    int32_t length_phi;
    int32_t length_jac;
    int32_t n_generated;
    int32_t n_acc;        // distinct moving links that carry spheres (SphereCollision Jacobian masks)
    int32_t scratch_doubles;  // per-thread scratch size in doubles (excluding the q staging)
    int32_t chain_joints;     // > 0: the program is a serial chain of that many joint ops and qualifies for the
                              // register-resident chain kernel (chain_kernel.cu); 0 otherwise
    int32_t chain_tab;        // offset into the int pool of the op index of every joint ordinal (chain programs)
    int32_t max_spheres;      // largest sphere count of any SphereCollision map (per-warp centre table in shared memory)
    int32_t two_pass;         // 1: the interpreter walks the ops twice; pass 2 recomputes each joint and emits its Jacobian
                              // columns on the fly (per-thread scratch holds 3 doubles per requested frame instead of 6 per joint)
    int32_t n_pos_slots;      // two-pass mode: requested frames (one position slot each)
};

// two-pass mode: OP_JOINT's `tab` points at [n, then n x {pos_slot, column, jrow_off, row mode, flags, cst_b}] in the int pool
enum
{
    TP_INTS = 6,
    TP_F_PRISMATIC = 1,
    TP_F_BCONST = 2
};

enum CompileMode : int32_t
{
    COMPILE_AUTO = 0,      // two-pass when eligible and the joint slots would dominate the scratch
    COMPILE_ONE_PASS = 1,
    COMPILE_TWO_PASS = 2   // two-pass whenever eligible (tests)
};

}  // namespace exo
