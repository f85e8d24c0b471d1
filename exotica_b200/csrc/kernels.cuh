// kernels.cuh -- launch interface of the CUDA interpreter (kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "program.h"

namespace exo
{
struct KernelParams
{
    const unsigned char* program;  // device copy of Program::image
    int32_t program_bytes;
    int32_t dbl_offset;            // byte offset of the double pool in the image
    int32_t scratch_offset;        // byte offset of the per-thread scratch inside dynamic shared memory
    int32_t T;                     // timesteps (evaluation e is timestep e % T)
    const double* x;
    double* phi;
    double* jac;                   // may be NULL
    const double* generated;       // [T][n_generated][12] or NULL
    double* min_dist;              // [n_evals] minimum of |p_i - p_j| - r_i - r_j over all sphere pairs, or NULL
    int64_t n_evals;
    // element strides (in doubles): addr = e * s_e + k * s_k  (Phi, X)   /  e * s_e + row * s_r + col * s_c  (J)
    int64_t x_se, x_sk;
    int64_t phi_se, phi_sk;
    int64_t j_se, j_sr, j_sc;
    // SphereCollision as its own launch: the interpreter parks every evaluation's frame / joint slots here ([sphere op][e][word],
    // one contiguous row per evaluation) instead of running the pair loops; exo_spheres_kernel picks them up one warp per row
    double* sph_dump;              // NULL: the interpreter evaluates SphereCollision itself
    int64_t sph_block;             // doubles per sphere op (rows of one launch)
    int32_t sph_row;               // doubles per row: the program's scratch words rounded up to an even count (16-byte rows)
    int32_t sph_n_ops;             // SphereCollision ops in the program (<= SPH_SPLIT_MAX_OPS)
    int64_t e_base;                // index of evaluation 0 of this launch in the caller's batch (timestep = (e_base + e) % T)
    int32_t sph_tiled;             // 1: the parked rows are tiled [e / 32][word][e % 32] (written by the generated park kernel)
    int32_t pad_tiled;
};

constexpr int SPH_SPLIT_MAX_OPS = 16;  // programs with more SphereCollision maps evaluate them inside the interpreter

struct LaunchConfig
{
    int threads = 128;
    int tile = 32;  // evaluations per warp
    int blocks = 0;
    size_t smem_bytes = 0;
    // second launch of split SphereCollision programs (exo_spheres_kernel)
    int sph_epc = 0;  // warps (= evaluations in flight) per CTA; 0: no second launch
    int sph_blocks = 0;
    size_t sph_smem_bytes = 0;
};

// Computes threads/blocks/shared memory for a program on the current device. split_spheres: plan the interpreter without the
// SphereCollision pair loops (they run in exo_spheres_kernel, planned here too).
cudaError_t plan_launch(const ProgramHeader& hdr, int program_bytes, int64_t n_evals, bool split_spheres, LaunchConfig& cfg, KernelParams& params);
cudaError_t launch_eval(const LaunchConfig& cfg, const KernelParams& params, cudaStream_t stream);
cudaError_t launch_spheres(const LaunchConfig& cfg, const KernelParams& params, cudaStream_t stream);

// ---- constant-bank kernel for serial-chain programs (chain_kernel.cu) ------------------------------------------------
constexpr int CH_MAXJ = 8;       // joints in the chain
constexpr int CH_MAXF = 4;       // requested frames hanging off the chain tip
constexpr int CH_MAXN = 16;      // controlled joints of the problem
constexpr int CH_MAXFILLS = 2;   // unused maps
constexpr int CH_MAXFILL = 16;   // Phi entries per unused map

struct ChainJointP
{
    double ca, sa, a;  // Rx(alpha) * Tx(a) in front of the joint (unused for joint 0: the base frame holds it)
    double th0;        // revolute: theta = th0 + qmult * x[qidx];  prismatic: constant rotation about z
    double c0, s0;     // cos / sin of th0 (used by prismatic joints)
    double d;          // translation along z after the joint (prismatic: d + qmult * x[qidx])
    double qmult;      // 0 when the joint value does not depend on the state
    int32_t qidx;      // state index feeding the joint (-1: none)
    int32_t col;       // Jacobian column or -1
    int32_t prismatic;
    int32_t pad;
    int64_t x_byte_off;  // 8 * qidx * x_sk            (filled per launch by chain_bind)
};
struct ChainFrameP
{
    double k[12];     // requested frame = D_{n-1} * k
    double binv[12];  // inverse of the constant temp_B
    int32_t k_kind;   // F_OFFA_NONE / F_OFFA_TRANS / F_OFFA_FULL
    int32_t b_const;  // 0: temp_B is the identity
    int32_t map_kind, rot_type, phi_off, jrow_off, n_rows, row_mode;  // row_mode: program.h RowMode
    int64_t phi_byte_off;  // 8 * phi_off * phi_sk     (filled per launch by chain_bind)
};
struct ChainFillP
{
    double value[CH_MAXFILL];
    int32_t phi_off, jrow_off, n_items, n_rows;
};
// Passed by value as the kernel parameter (constant bank 0): < 4 KB.
struct ChainParams
{
    const double* x;
    double* phi;
    double* jac;  // may be NULL
    int64_t n_evals;
    int64_t prefetch_ahead;                        // evaluations between a lane's work and the L2 prefetch it issues (full-grid launches)
    int64_t x_sk, phi_sk, j_sr, j_sc;              // element strides (doubles) of the entry / row / column index
    int64_t x_se_b, phi_se_b, j_se_b, j_sr_b;      // byte strides of the evaluation index and of a Jacobian row
    int64_t jf_byte_off[CH_MAXF][CH_MAXJ];         // byte offset of J(first row of frame f, column of joint j)
    int32_t n_joints, n_ctrl, n_frames, n_fill;
    int32_t has_prefix, jl_enabled, jl_phi_off, jl_jrow_off;
    uint32_t zero_cols;
    int32_t all_revolute;
    int32_t ortho, pad_ortho;  // every link has alpha = +-90 degrees and a = 0 (exact): dh_link's signed-swap form applies
    double base[12];  // frame in front of joint 0, z axis = joint 0's axis
    ChainJointP joint[CH_MAXJ];
    ChainFrameP frame[CH_MAXF];
    double jl_lo[CH_MAXN], jl_hi[CH_MAXN], jl_tau[CH_MAXN], jl_lo_tau[CH_MAXN], jl_hi_tau[CH_MAXN];
    ChainFillP fill[CH_MAXFILLS];
};
static_assert(sizeof(ChainParams) <= 4000, "ChainParams must fit the classic 4 KB kernel parameter space");

enum
{
    CHAIN_REGS = 0,      // one pass, every joint's axis and origin stay in registers until the Jacobian is emitted
    CHAIN_TWO_PASS = 1,  // the chain is recomposed for the Jacobian pass (nothing parked at all)
    CHAIN_DEFAULT = CHAIN_REGS
};
struct Program;
// Fills the problem constants of ChainParams from a chain-eligible program; false if the program does not qualify.
bool chain_lower(const Program& prog, ChainParams& P);
// Binds the buffers and strides of one launch (and the element offsets derived from them).
void chain_bind(ChainParams& P, const double* x, double* phi, double* jac, int64_t n_evals, int64_t x_se, int64_t x_sk, int64_t phi_se,
                int64_t phi_sk, int64_t j_se, int64_t j_sr, int64_t j_sc);
cudaError_t chain_launch(const ChainParams& P, int variant, cudaStream_t stream);

// ---- TMA tensor stores (tma.cu): SoA outputs leave the SM as 2-D tiles [rows][TILE evaluations] staged in shared memory -----------
// Opaque 128-byte tensor map (CUtensorMap) as a kernel parameter; encoded on the host by cuTensorMapEncodeTiled, which is
// resolved through cudaGetDriverEntryPoint (no link-time dependency on libcuda).
struct alignas(64) TmaDesc
{
    unsigned long long v[16];
};
// 2-D FP64 tensor [rows][inner] with row pitch pitch_bytes (a multiple of 16, base 16-byte aligned), box [box_rows][box_inner]
// (each <= 256).  False when the driver refuses the geometry (the caller takes its per-lane store path).
bool tma_encode_2d(TmaDesc& out, const void* base, uint64_t inner, uint64_t rows, uint64_t pitch_bytes, uint32_t box_inner, uint32_t box_rows);
// 3-D FP64 tensor [d2][d1][inner] with byte pitches pitch1 (between d1 entries) and pitch2 (between d2 entries), box [b2][b1][box_inner]
bool tma_encode_3d(TmaDesc& out, const void* base, uint64_t inner, uint64_t d1, uint64_t d2, uint64_t pitch1_bytes, uint64_t pitch2_bytes, uint32_t box_inner,
                   uint32_t b1, uint32_t b2);

// Serial-chain kernel with TMA output (chain_tma_kernel.cu): each CTA stages the Phi / J entries of its `tile` evaluations as
// [row][tile] in shared memory (immediate-offset st.shared, no 64-bit address arithmetic per entry) and one lane hands every
// frame's blocks to the TMA unit (cp.async.bulk.tensor.2d.global.shared::cta, L2 evict-first).  Eligible: SoA buffers with an
// even number of evaluations per row, every controlled joint in the chain in column order, EffFrame / EffPosition rows.
struct ChainTmaParams
{
    TmaDesc phi[CH_MAXF], jac[CH_MAXF];                      // per requested frame: its Phi rows / its n_rows x N Jacobian rows
    int32_t stage_phi_row[CH_MAXF], stage_j_row[CH_MAXF];    // first staging row of the frame's blocks
    int32_t stage_rows, tile;
};
// Plans tile size, staging layout and tensor maps for the buffers bound to P; false: not eligible (use chain_launch).
// tile: 0 = automatic, else 64 or 128.
bool chain_tma_plan(const ChainParams& P, int tile, ChainTmaParams& M);
cudaError_t chain_tma_launch(const ChainParams& P, const ChainTmaParams& M, cudaStream_t stream);

// ---- per-robot generated kernel for tree programs (tree_codegen.cpp, tree_jit.cu) --------------------------------------------------
// The two-pass tree program unrolled into straight-line CUDA C++ (constants as literals, one warp per tile of 32 evaluations,
// Phi / J leaving through TMA boxes, structural zeros from a shared zero tile), compiled at run time with NVRTC for sm_100a.
constexpr int TREE_MAXMAPS = 24;
struct TreeGenPlan
{
    struct MapShape
    {
        int phi;           // 1: 2-D box of Phi[LP][E]; 0: 3-D box of J[LJ][N][E]; 2: 2-D box of the flattened J[LJ * N][E]
        int ncols, nrows;  // box = [nrows][ncols][32]
    };
    std::string source;
    std::vector<MapShape> maps;
    int warps_per_cta = 4;
    int tile = 32;  // evaluations per tile (inner box dimension)
    size_t smem_bytes = 0;
    int n_ctrl = 0, length_phi = 0, length_jac = 0;
    bool needs_flat_jacobian = false;  // some boxes address J as [LJ * N][E]: the row pitch must be N column pitches
    const char* kernel_name = "exo_tree_kernel";  // "exo_tree_update_cost_kernel" when generated with a cost epilogue
};
struct TreeGenParams  // mirrored by the generated source
{
    const double* x;
    double* phi;
    double* jac;
    long long n_evals, x_sk, phi_sk, j_sr, j_sc;
    TmaDesc maps[TREE_MAXMAPS];
    // cost epilogue (exo_tree_update_cost_kernel only): the task + transition cost of every evaluation, formed from the Phi entries
    // while they are in registers
    double* cost_eval;  // [n_evals]
    const double *rho, *goal, *goal_rot, *W;
    double ct;
    int T, t_first, T_goal, n_maps, length_phi, n_rot;
};
struct TreeCostSpec;
// Generates the source; false + why when the program is outside the generated kernel's scope (the interpreter runs it).  With a
// cost structure the kernel also writes every evaluation's scalar cost (PlanningProblem::Update + GetScalarTaskCost /
// GetScalarTransitionCost in one pass over the state).
bool tree_codegen(const Program& prog, TreeGenPlan& plan, std::string& why, const TreeCostSpec* cost = nullptr);
// NVRTC only (no device needed): source -> sm_100a cubin
bool tree_jit_compile(const TreeGenPlan& plan, std::vector<char>& cubin, std::string& err);
// NVRTC + module cache (keyed by source text and device): an opaque kernel handle, nullptr + err on failure
void* tree_jit_get(const TreeGenPlan& plan, std::string& err);
// Tensor maps of one binding; false when the buffers do not qualify (layout, alignment, odd row pitch)
bool tree_jit_bind(const TreeGenPlan& plan, TreeGenParams& P, const double* x, double* phi, double* jac, int64_t n_evals, int64_t x_se, int64_t x_sk,
                   int64_t phi_se, int64_t phi_sk, int64_t j_se, int64_t j_sr, int64_t j_sc);
cudaError_t tree_jit_launch(void* kernel, const TreeGenPlan& plan, const TreeGenParams& P, cudaStream_t stream);

// Generated fused forward kinematics + task cost + cost gradient for tree programs (exo_tree_costgrad_kernel): what
// exo_chain_costgrad_kernel is for serial chains.  The cost STRUCTURE is compiled in (which cost map and goal rotation every
// requested frame belongs to, the JointLimit map); rho, goals and W stay run-time arrays.
struct TreeCostSpec
{
    std::vector<int> frame_map;  // per requested frame (position slot): index of its cost map (its rho)
    std::vector<int> frame_rot;  // per requested frame: index into goal_rot, or -1 for position-only frames (EffPosition)
    int jl_map = -1;             // cost map index of the JointLimit map or -1
};
struct TreeCostPlan
{
    std::string source;
    int warps_per_cta = 4;
    size_t smem_bytes = 0;
};
struct TreeCostParams  // mirrored by the generated source
{
    const double* x;
    double* grad;       // may be NULL: cost only
    double* cost_eval;  // [n_evals] or NULL
    const double *rho, *goal, *goal_rot, *W;
    long long n_evals, x_se, x_sk, g_se, g_sk;
    double ct;
    int T, t_first, T_goal, n_maps, length_phi, n_rot;
};
bool tree_costgrad_codegen(const Program& prog, const TreeCostSpec& cs, TreeCostPlan& plan, std::string& why);

// Generated front half of split SphereCollision programs (exo_tree_park_kernel): composes the tree and parks every evaluation's
// frame / joint slots as the row exo_spheres_kernel reads, in place of the interpreter.
struct TreeParkPlan
{
    std::string source;
    int threads = 128;
};
struct TreeParkParams  // mirrored by the generated source
{
    const double* x;
    const double* generated;
    double* dump;
    long long n_evals, x_se, x_sk, e_base, sph_block;
    int T, n_generated, sph_row;
};
bool tree_park_codegen(const Program& prog, TreeParkPlan& plan, std::string& why);
void* tree_park_jit_get(const TreeParkPlan& plan, std::string& err);
cudaError_t tree_park_jit_launch(void* kernel, const TreeParkPlan& plan, const TreeParkParams& P, cudaStream_t stream);
void* tree_costgrad_jit_get(const TreeCostPlan& plan, std::string& err);
bool tree_costgrad_jit_compile(const TreeCostPlan& plan, std::vector<char>& cubin, std::string& err);
cudaError_t tree_costgrad_jit_launch(void* kernel, const TreeCostPlan& plan, const TreeCostParams& P, cudaStream_t stream);

// Fused forward kinematics + task cost + cost gradient for chain programs (exo_chain_costgrad_kernel): Phi and J are formed in
// registers and contracted on the spot, only N doubles per evaluation (and its cost) reach HBM.
struct ChainGradParams
{
    const double* rho;       // device, [T][n_maps]
    const double* goal;      // device, [T_goal][length_phi]
    const double* goal_rot;  // device, [T_goal][n_rot][9]
    const double* W;         // device, [n_ctrl] or NULL
    double* grad;            // gradient of evaluation e, entry n at grad[e * g_se + n * g_sk]
    double* cost_eval;       // [n_evals] or NULL
    double ct;
    int64_t g_se, g_sk;
    int32_t T, t_first, T_goal, n_maps, length_phi, n_rot;
    int32_t jl_map, pad;               // map index of the JointLimit map or -1
    int32_t frame_map[CH_MAXF];        // map index of every requested frame (its rho)
    int32_t frame_rot[CH_MAXF];        // index into goal_rot, or -1 for position-only frames (EffPosition)
};
cudaError_t chain_costgrad_launch(const ChainParams& P, const ChainGradParams& G, cudaStream_t stream);
// cost only: G.cost_eval[e] for every evaluation (G.grad unused)
cudaError_t chain_cost_launch(const ChainParams& P, const ChainGradParams& G, cudaStream_t stream);

// Fused line search of the batched IK step for chain programs (exo_chain_ik_linesearch_kernel): the ten trial states
// q - alpha qd, alpha = 1 ... 0.1 (ik_solver.cpp:146-161), are evaluated in registers (forward kinematics + cost) until the cost
// drops; replaces ten trial / update / accept launch triples.  All arrays are [.][E] (SoA); P.x must be bound to q_trial.
struct ChainLineSearchParams
{
    double* q;               // [N][E] in / out
    double* q_trial;         // [N][E] scratch (= P.x)
    double* qd;              // [N][E] in / out
    const double* err_prev;  // [E]
    double* cost;            // [E] error at the last state evaluated
    const int32_t* active;   // [E]
    int32_t* ls_done;        // [E]
    double* steplen;         // [E] steplength_: the last alpha tried (read by the MaxIterations == 1 regularisation adaptation)
};
cudaError_t chain_ik_linesearch_launch(const ChainParams& P, const ChainGradParams& G, const ChainLineSearchParams& L, cudaStream_t stream);
struct IkParams;
struct CostParams;
// The whole IKSolver::Solve loop of every start state in ONE launch (exo_chain_ik_solve_kernel): each lane iterates update ->
// step -> line search -> termination tests on its own column of the SoA workspace until it stops; no host round trips.
// P must be bound to (q, workspace Phi, workspace J); K / C as for the staged kernels.
cudaError_t chain_ik_solve_launch(const ChainParams& P, const ChainGradParams& G, const ChainLineSearchParams& L, const IkParams& K,
                                  const CostParams& C, int max_iterations, cudaStream_t stream);

// ---- batched task cost and per-trajectory reductions (cost_kernel.cu) -------------------------------------------------
constexpr int COST_MAXMAPS = 16;
struct CostMapP
{
    int32_t kind, start, length, n_frames, rot_type, stride, start_jacobian, length_jacobian;  // stride: Phi entries per frame (EffFrame, EffOrientation)
};
struct CostParams
{
    const double* x;
    const double* phi;
    double* cost_traj;       // [n_traj] or NULL
    double* cost_eval;       // [n_traj * T] or NULL
    const double* rho;       // device, [T][n_maps]
    const double* goal;      // device, [T_goal][length_phi]
    const double* goal_rot;  // device, [T_goal][n_rot][9]: GetRotation of every EffFrame frame's goal, map order
    const double* W;         // device, [n_ctrl] diagonal of W or NULL
    double ct;
    int64_t n_traj;
    int64_t x_se, x_sk, phi_se, phi_sk;
    int32_t T, t_first, T_goal, n_maps, n_ctrl, length_phi, n_rot, pad;
    CostMapP map[COST_MAXMAPS];
};
cudaError_t launch_cost(const CostParams& P, cudaStream_t stream);
// gradient of the task cost with respect to every state: grad(e, n) at grad[e * g_se + n * g_sk]
struct CostGradParams
{
    CostParams c;
    const double* jac;
    double* grad;
    int64_t j_se, j_sr, j_sc, g_se, g_sk;
    int32_t length_jac, pad;
};
cudaError_t launch_cost_gradient(const CostGradParams& P, cudaStream_t stream);
// AICO task messages of every evaluation (aico_solver.cpp:364-386): R(e)(i, j) at R[e * R_se + i * R_si + j * R_sj], r(e)(i) at
// r[e * r_se + i * r_sk]
struct AicoParams
{
    CostParams c;          // c.x = qhat
    const double* jac;
    double *R, *r, *rhat, *cost;
    int64_t j_se, j_sr, j_sc, R_se, R_si, R_sj, r_se, r_sk;
    int32_t length_jac, pad;
};
cudaError_t launch_aico_messages(const AicoParams& P, cudaStream_t stream);
// op 0: min, 1: sum over t in [t_first, T) of in[b * T + t]
cudaError_t launch_reduce(const double* in, int64_t n_traj, int T, int t_first, int op, double* out, cudaStream_t stream);
cudaError_t launch_fill(double* out, int64_t n, double value, cudaStream_t stream);
void host_get_rotation(const double* data, int type, double* out9);

// ---- batched IK step (ik_kernel.cu) ----
constexpr int IK_MAXLJ = 16;  // rows of the task-space Jacobian the register / local-memory step kernel handles
constexpr int IK_BIG_MAXLJ = 64;  // ... and the warp-per-configuration step kernel (whole-body problems: several end-effector frames)
constexpr int IK_MAXN = 64;       // controlled joints
enum
{
    // exotica::TerminationCriterion (exotica_core/include/exotica_core/planning_problem.h:49-62)
    IK_TERM_ITERATION_LIMIT = 0,
    IK_TERM_STEP_TOLERANCE = 2,
    IK_TERM_FUNCTION_TOLERANCE = 3,
    IK_TERM_GRADIENT_TOLERANCE = 4,
    IK_TERM_DIVERGENCE = 5
};
struct IkParams
{
    int64_t n_evals;
    int32_t n_ctrl, length_jac;
    double tolerance, step_tolerance, gradient_tolerance, regmax;
    // MaxIterations == 1 (ik_solver.cpp:139-143,187-199): MaxStep != 0 replaces the line search by one clamped step; MaxStep == 0
    // adapts lambda from the accepted step length and stops with Divergence when it reaches MaximumRegularization
    double max_step, th_stepinc, th_stepdec, regmin;
    int32_t single_iteration, pad0;
    double* steplen;      // [E] steplength_
    const double* jac;    // [LJ][N][E] of the current states
    double* q;            // [N][E] current states (in / out)
    double* q_trial;      // [N][E]
    double* qd;           // [N][E]
    double* cost;         // [E] error_
    double* err_prev;     // [E]
    double* stop;         // [E] |cJ|_F
    double* lambda;       // [E]
    int32_t* active;      // [E]
    int32_t* ls_done;     // [E]
    int32_t* iterations;  // [E]
    int32_t* termination; // [E]
    int32_t* n_active;    // [1]
    int32_t warp_step;    // 1: task spaces beyond the register kernel's (IK_MAXLJ rows, CH_MAXN joints): exo_ik_step_warp_kernel, one warp per configuration
    int32_t pad_ws;
    int32_t* list;        // [E] the active configurations, compacted before every step of the warp-per-configuration kernel (order irrelevant)
    int32_t* n_list;      // [1]
    double row_rho[IK_BIG_MAXLJ];  // rho of the map that owns each Jacobian row (S = diag)
    double w_inv[IK_MAXN];         // 1 / W[c]
};
cudaError_t launch_ik_init(const IkParams& K, double lambda0, cudaStream_t s);
cudaError_t launch_ik_step(const IkParams& K, const CostParams& C, cudaStream_t s);  // K.warp_step: the warp-per-configuration kernel
cudaError_t launch_ik_trial(const IkParams& K, double alpha, bool first, cudaStream_t s);  // first: the first step length of a line search
cudaError_t launch_ik_accept(const IkParams& K, const CostParams& C, double alpha, cudaStream_t s);
cudaError_t launch_ik_finish(const IkParams& K, int last_iteration, cudaStream_t s);

// ---- task-map Hessians (cost_kernel.cu) ----
struct HessMapP
{
    int32_t kind, is_used, n_frames, frame_start, start_jacobian, length_jacobian, rows_per_frame, pad;
};
struct HessParams
{
    const double* frames_phi;  // raw frames of the shadow problem
    const double* frames_jac;
    const double* map_jac;     // the problem's own Jacobian
    double* hess;
    int64_t n_evals;
    int64_t fp_se, fp_sk, fj_se, fj_sr, fj_sc, mj_se, mj_sr, mj_sc, h_se, h_sr, h_si, h_sj;  // element strides
    int32_t n_maps, n_ctrl;
    HessMapP map[COST_MAXMAPS];
};
cudaError_t launch_hessian(const HessParams& P, cudaStream_t stream);

}  // namespace exo
