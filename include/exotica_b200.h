/*
 * exotica_b200.h -- C-ABI of the B200-native batched-kinematics backend for EXOTica.
 *
 * This is the drop-in boundary: plain C, opaque handles, caller-owned buffers,
 * int status codes + a thread-local message (exo_last_error).  No torch / Eigen / KDL types.
 * Handles are NOT thread-safe (same contract as the reference's stateful KinematicTree / Scene /
 * PlanningProblem); use one handle per host thread and one per device.
 *
 * Each entry point names the reference interface it replaces (paths relative to ipab-slmc/exotica
 * v6.2.0).  INTEGRATION.md shows the binding a maintainer adds on the reference side.
 *
 * Conventions
 *   - All arithmetic is IEEE FP64, indices are 32-bit, batch sizes 64-bit.
 *   - A frame is 12 doubles: KDL::Rotation::data[9] (row-major) then KDL::Vector p[3].
 *   - N  = number of controlled joints, E = number of evaluations in a batch.  For time-indexed
 *     problems E = B*T with the timestep innermost: evaluation e is timestep t = e % T of
 *     trajectory b = e / T.
 *   - Layouts (exo_layout): how the E evaluations are laid out in X / Phi / J.
 *       EXO_LAYOUT_SOA           X[N][E], Phi[length_Phi][E], J[length_jacobian][N][E]  (batch innermost;
 *                                the kernels' native, fully coalesced layout)
 *       EXO_LAYOUT_AOS_ROWMAJOR  X[E][N], Phi[E][length_Phi], J[E][length_jacobian][N]  (numpy-natural)
 *       EXO_LAYOUT_AOS_COLMAJOR  X[E][N], Phi[E][length_Phi], J[E][N][length_jacobian]  (one Eigen::MatrixXd
 *                                per evaluation, as problem->jacobian holds it)
 */
#ifndef EXOTICA_B200_H_
#define EXOTICA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EXO_OK 0
#define EXO_ERR_INVALID 1   /* bad argument / size mismatch (the reference throws exotica::Exception) */
#define EXO_ERR_CUDA 2      /* CUDA runtime failure, message in exo_last_error() */
#define EXO_ERR_STATE 3     /* call order violated (e.g. update before finalize) */
#define EXO_ERR_UNSUPPORTED 4

/* KDL::Joint::JointType (external, orocos_kdl joint.hpp) */
typedef enum
{
    EXO_JOINT_ROT_AXIS = 0,
    EXO_JOINT_ROT_X = 1,
    EXO_JOINT_ROT_Y = 2,
    EXO_JOINT_ROT_Z = 3,
    EXO_JOINT_TRANS_AXIS = 4,
    EXO_JOINT_TRANS_X = 5,
    EXO_JOINT_TRANS_Y = 6,
    EXO_JOINT_TRANS_Z = 7,
    EXO_JOINT_NONE = 8
} exo_joint_type;

/* exotica::RotationType, exotica_core/include/exotica_core/tools/conversions.h:80-88 */
typedef enum
{
    EXO_ROT_QUATERNION = 0,
    EXO_ROT_RPY = 1,
    EXO_ROT_ZYX = 2,
    EXO_ROT_ZYZ = 3,
    EXO_ROT_ANGLE_AXIS = 4,
    EXO_ROT_MATRIX = 5
} exo_rotation_type;

typedef enum
{
    EXO_LAYOUT_SOA = 0,
    EXO_LAYOUT_AOS_ROWMAJOR = 1,
    EXO_LAYOUT_AOS_COLMAJOR = 2
} exo_layout;

/* Task maps registered by the reference as "exotica/<Type>" (REGISTER_TASKMAP_TYPE,
 * exotations/exotica_core_task_maps/src/{eff_frame,eff_position,joint_limit,sphere_collision}.cpp)
 * plus EXO_MAP_FRAMES, the raw KinematicResponse (Phi frames + 6xN Jacobians) that
 * KinematicTree::Update fills (exotica_core/src/kinematic_tree.cpp:678-693). */
typedef enum
{
    EXO_MAP_EFF_FRAME = 0,
    EXO_MAP_EFF_POSITION = 1,
    EXO_MAP_JOINT_LIMIT = 2,
    EXO_MAP_SPHERE_COLLISION = 3,
    EXO_MAP_FRAMES = 4,
    /* sibling maps that reuse the same frames (exotations/exotica_core_task_maps/src/
     * {eff_orientation,eff_position_xy,distance,joint_pose}.cpp) */
    EXO_MAP_EFF_ORIENTATION = 5,
    EXO_MAP_EFF_POSITION_XY = 6,
    EXO_MAP_DISTANCE = 7,
    EXO_MAP_JOINT_POSE = 8,
    /* exotations/exotica_core_task_maps/src/eff_axis_alignment.cpp:116-145 (one row per frame:
     * phi = direction . (R_link axis) - 1) and center_of_mass.cpp:37-132 (mass-weighted mean of the frames' positions,
     * 3 rows, or 2 when EnableZ is false) */
    EXO_MAP_EFF_AXIS_ALIGNMENT = 9,
    EXO_MAP_CENTER_OF_MASS = 10
} exo_map_kind;

typedef struct exo_tree exo_tree;       /* backend of exotica::KinematicTree (scene.h:195 kinematica_) */
typedef struct exo_problem exo_problem; /* backend of a PlanningProblem's Update path */

/* Flattened KinematicTree::tree_ (kinematic_element.h:54-143), index 0 = root element.
 * Replaces the KDL::Segment each element holds: joint type/origin/axis and the tip frame f_tip
 * (Segment::pose(q) = Joint::pose(q) * f_tip). */
typedef struct
{
    int32_t n_elements;
    const int32_t* id;             /* KinematicElement::id (root -1) */
    const int32_t* parent;         /* parent element INDEX, -1 for the root */
    const int32_t* joint_type;     /* exo_joint_type */
    const double* joint_origin;    /* [n][3]  KDL::Joint origin */
    const double* joint_axis;      /* [n][3]  unit axis, parent-link coordinates */
    const double* f_tip;           /* [n][12] */
    const int32_t* control_id;     /* KinematicElement::control_id, -1 if not controlled */
    const int32_t* mimic_element;  /* element INDEX of the mimicked joint or -1 (may be NULL) */
    const double* mimic_multiplier;/* may be NULL */
    const double* mimic_offset;    /* may be NULL */
    const int32_t* generated_slot; /* >=0: trajectory-driven element (is_trajectory_generated), index into the
                                      per-timestep frame table; -1 otherwise (may be NULL) */
    int32_t n_controlled;          /* KinematicTree::GetNumControlledJoints() */
    const double* lower;           /* [N] KinematicTree::GetJointLimits() column 0 */
    const double* upper;           /* [N] column 1 */
} exo_tree_desc;

/* One task map.  Frames follow exotica_core/init/frame.in after name resolution
 * (KinematicTree::RequestFrames, kinematic_tree.cpp:628-676) and GetFrame() of the offsets. */
typedef struct
{
    int32_t kind;            /* exo_map_kind */
    int32_t n_frames;
    const int32_t* elem_a;   /* [n_frames] element index of frame A (Link) */
    const double* off_a;     /* [n_frames][12] LinkOffset */
    const int32_t* elem_b;   /* [n_frames] element index of frame B (Base; 0 = root) */
    const double* off_b;     /* [n_frames][12] BaseOffset */
    int32_t rotation_type;   /* EffFrame Type (init/eff_frame.in) */
    double safe_percentage;  /* JointLimit SafePercentage (init/joint_limit.in) */
    double precision;        /* SphereCollision Precision (init/sphere_collision.in) */
    const double* radius;    /* [n_frames] Sphere Radius (init/sphere.in) */
    const int32_t* group;    /* [n_frames] rank of the sphere's Group name in lexicographic order (std::map order) */
    int32_t is_used;         /* TaskMap::is_used (rho != 0 somewhere, tasks.cpp:246-259); unused maps only get the
                                SetZero fill (task_space_vector.cpp:54-62) */
    int32_t n_joint_map;     /* JointPose: entries of JointMap (0 = all controlled joints in order, init/joint_pose.in) */
    const int32_t* joint_map;/* [n_joint_map] controlled-joint indices */
    const double* joint_ref; /* [n_joint_map or N] JointRef, NULL = zeros */
    const double* axis;      /* EffAxisAlignment: [n_frames][3] Axis (init/frame_with_axis_and_direction.in; normalised here) */
    const double* direction; /* EffAxisAlignment: [n_frames][3] Direction (normalised here) */
    const double* mass;      /* CenterOfMass: [n_frames] mass of every frame (center_of_mass.cpp:150-205) */
    int32_t enable_z;        /* CenterOfMass EnableZ (init/center_of_mass.in): 0 keeps x and y only */
} exo_task_map_desc;

/* The library's run-time compiled kernels include this header for the enums above; the host entry points below are hidden
 * from NVRTC (which defines __CUDACC_RTC__ and rejects host function declarations). */
#ifndef __CUDACC_RTC__
/* ---- errors / device ------------------------------------------------------------------------- */
const char* exo_last_error(void);
int exo_device_count(int* count);
int exo_set_device(int device);
const char* exo_version(void);

/* ---- memory and streams (plumbing for callers that do not bring their own CUDA allocator) ------ */
int exo_device_alloc(void** ptr, size_t bytes);
int exo_device_free(void* ptr);
int exo_host_alloc_pinned(void** ptr, size_t bytes);
int exo_host_free_pinned(void* ptr);
int exo_memcpy_h2d(void* dst, const void* src, size_t bytes, void* stream);
int exo_memcpy_d2h(void* dst, const void* src, size_t bytes, void* stream);
int exo_memset_device(void* dst, int value, size_t bytes, void* stream);
int exo_stream_create(void** stream);
int exo_stream_destroy(void* stream);
int exo_stream_synchronize(void* stream); /* NULL = default stream */
int exo_device_synchronize(void);
/* CUDA events (device-side timing on the launching stream) and stream capture into a CUDA graph, for callers without their
 * own CUDA runtime binding.  A batched call and the collective behind it (exo_comm_*) capture into one graph once their
 * scratch buffers exist, i.e. after one ordinary call of the same shape. */
int exo_event_create(void** event);
int exo_event_destroy(void* event);
int exo_event_record(void* event, void* stream);
int exo_event_synchronize(void* event);
int exo_event_elapsed_ms(void* start, void* stop, float* ms);
typedef struct exo_graph exo_graph;
int exo_graph_begin_capture(void* stream); /* a stream from exo_stream_create */
int exo_graph_end_capture(void* stream, exo_graph** out);
int exo_graph_launch(exo_graph* graph, void* stream);
int exo_graph_destroy(exo_graph* graph);

/* ---- KinematicTree backend ----------------------------------------------------------------------
 * exo_tree_create        <- end of KinematicTree::BuildTree / UpdateModel (kinematic_tree.cpp:164-421,423-432)
 * exo_tree_set_model_state <- KinematicTree::SetModelState (tree_state_ of uncontrolled joints, :1035-1060)
 * exo_tree_set_joint_limits <- KinematicTree::SetJointLimitsLower/Upper, ResetJointLimits (:1035-1297)
 */
int exo_tree_create(const exo_tree_desc* desc, exo_tree** out);
int exo_tree_destroy(exo_tree* tree);
int exo_tree_set_model_state(exo_tree* tree, const double* tree_state /* [n_elements] by element index */);
int exo_tree_set_joint_limits(exo_tree* tree, const double* lower, const double* upper);
int exo_tree_num_controlled(const exo_tree* tree);

/* ---- Problem backend ------------------------------------------------------------------------------
 * exo_problem_create / add_task_map / finalize <- PlanningProblem::InstantiateBase (planning_problem.cpp:148-244):
 *     maps are appended in order; frames of all maps are concatenated into one KinematicsRequest; start /
 *     length / start_jacobian / length_jacobian are assigned in map order.
 * exo_problem_set_generated  <- Scene::UpdateTrajectoryGenerators (scene.cpp:329-335): generated_offset of every
 *     trajectory-driven element for each of the T timesteps.
 * exo_problem_update_batch   <- for each evaluation: AbstractTimeIndexedProblem::Update(x,t)
 *     (abstract_time_indexed_problem.cpp:220-292) / UnconstrainedEndPoseProblem::Update (:123-167) up to and
 *     including the task-map loop: Scene::Update, Phi.SetZero, jacobian.setZero, TaskMap::Update per used map.
 */
int exo_problem_create(exo_tree* tree, exo_problem** out);
int exo_problem_destroy(exo_problem* problem);
int exo_problem_add_task_map(exo_problem* problem, const exo_task_map_desc* map);
int exo_problem_finalize(exo_problem* problem);
int exo_problem_length_phi(const exo_problem* problem);
int exo_problem_length_jacobian(const exo_problem* problem);
int exo_problem_num_maps(const exo_problem* problem);
/* start, length, start_jacobian, length_jacobian of map i (TaskMap::start..., task_map.h:79-84) */
int exo_problem_map_indexing(const exo_problem* problem, int map, int32_t out[4]);
/* T timesteps, n_generated frames per timestep: frames[T][n_generated][12] (host pointer, copied). */
int exo_problem_set_generated(exo_problem* problem, int32_t T, int32_t n_generated, const double* frames);

/* Device buffers (already resident in HBM).  d_jac may be NULL (KIN_FK only).  Asynchronous on `stream`. */
int exo_problem_update_batch(exo_problem* problem, const double* d_x, int64_t n_evals, int32_t T, int32_t layout,
                             double* d_phi, double* d_jac, void* stream);
/* Host buffers: chunked H2D -> kernel -> D2H pipeline on internal streams; returns when the outputs are
 * complete in host memory.  Fastest with pinned buffers (exo_host_alloc_pinned). */
int exo_problem_update_batch_host(exo_problem* problem, const double* h_x, int64_t n_evals, int32_t T, int32_t layout,
                                  double* h_phi, double* h_jac);
/* Number of kernel launches issued so far by this handle (evidence for bench.py's gpu_launches). */
int64_t exo_problem_launch_count(const exo_problem* problem);
/* Device time of the most recent exo_problem_update_batch launch in milliseconds (CUDA events on the
 * launching stream; blocks until that launch finished). */
int exo_problem_last_kernel_ms(exo_problem* problem, float* ms);

/* Kernel selection: 0 = auto (the Denavit-Hartenberg chain kernel when the compiled program is a serial chain of at
 * most 8 joints whose frames hang off the last link, else the general interpreter), 1 = always the interpreter,
 * 2 / 3 = the chain kernel's variants (joint data kept in registers / chain composed twice; EXO_ERR_UNSUPPORTED at
 * launch if the program is not a chain), 4 / 5 = the interpreter with the one-pass program (joint axes and origins parked
 * in shared memory until a frame needs them) / the two-pass program (the tree is walked twice and Jacobian columns are
 * emitted on the fly; used automatically for large trees), 6 = the interpreter with SphereCollision evaluated inside it
 * (by default programs with a SphereCollision map run as two launches: the interpreter parks every evaluation's frames
 * in a scratch buffer -- at most EXO_SPH_DUMP_MIB MiB, default 1024, larger batches go in chunks -- and
 * exo_spheres_kernel runs the pair loops one warp per evaluation), 7 / 8 = the chain kernel with TMA output (each CTA stages
 * the Phi / J entries of its 64 / 128 evaluations in shared memory and the TMA unit stores them as 2-D tiles; needs
 * EXO_LAYOUT_SOA device buffers with an even, 16-byte aligned row pitch and a chain that owns every Jacobian column, else
 * EXO_ERR_UNSUPPORTED at launch; choice 0 takes this path automatically when those conditions hold, environment
 * EXO_CHAIN_TMA=0 keeps choice 0 on the per-lane stores), 9 = the per-robot generated kernel for tree programs (the
 * two-pass program unrolled into straight-line CUDA C++ with the robot's constants as literals, compiled at run time with
 * NVRTC for sm_100a and cached; one warp per 32 evaluations, Phi / J staged in shared memory and stored by the TMA unit,
 * structural zeros of J from a shared zero tile; EffFrame / EffPosition / EffOrientation / EffPositionXY / raw frames /
 * JointLimit / JointPose maps with constant base frames, EXO_LAYOUT_SOA buffers with an even row pitch; EXO_ERR_UNSUPPORTED otherwise; choice 0 takes it for batches of at
 * least EXO_TREE_MIN_EVALS = 4096 evaluations, and from the 16th call on for batches of at least 256 -- a solver loop earns the
 * 0.3-2 s of compilation, a one-off small call stays on the interpreter).  Tests compare all of them. */
int exo_problem_set_kernel(exo_problem* problem, int choice);
/* CUDA source of the per-robot generated kernel (choice 9 above) for the current model state: copies at most capacity - 1
 * characters into out (may be NULL) and reports the size needed; bit 0 of `compile` also runs NVRTC (no device needed) and returns
 * EXO_ERR_UNSUPPORTED with the compiler log if it fails; bit 1 selects the generated fused cost + gradient kernel of tree programs
 * (exo_tree_costgrad_kernel, after exo_problem_set_cost) instead of the update kernel.  Replaces nothing in the reference:
 * introspection for tests. */
int exo_problem_generated_kernel_source(exo_problem* problem, char* out, int64_t capacity, int64_t* needed, int32_t compile);
/* Name of the kernel the most recent launch used ("exo_chain_kernel" / "exo_chain_tma_kernel" / "exo_tree_kernel" / "exo_eval_kernel" / "exo_spheres_kernel" (the
 * second launch of a SphereCollision program) / "none"). */
const char* exo_problem_last_kernel(const exo_problem* problem);

/* Introspection of the compiled program (tests / DESIGN.md): number of joint ops, emit ops, frame slots,
 * joint slots, shared-memory bytes per thread. */
int exo_problem_program_info(const exo_problem* problem, int32_t out[8]);
/* Whether the program was lowered to the Denavit-Hartenberg chain kernel: out = {lowered (0/1), joints in the chain,
 * bit 0: all joints revolute, bit 1: every link has alpha = +-90 degrees and a = 0 (the signed-swap link form applies),
 * requested frames}.  Host-only (valid after exo_problem_finalize, no device needed). */
int exo_problem_chain_info(const exo_problem* problem, int32_t out[4]);

/* ---- task cost and per-trajectory reductions (what the multi-GPU path gathers; SURVEY.md section 8 rows e, f1) --------
 * exo_problem_set_cost   <- EndPoseTask / TimeIndexedTask::Initialize + SetRho + SetGoal (exotica_core/src/tasks.cpp:36-244),
 *                            ct = 1/tau/T and W (abstract_time_indexed_problem.cpp:105, time_indexed_problem.cpp:43-50)
 * exo_problem_cost_batch <- TimeIndexedTask::Update's ydiff = Phi - y with TaskSpaceVector::operator- (tasks.cpp:306-314,
 *                            task_space_vector.cpp:64-91) + GetScalarTaskCost + GetScalarTransitionCost + GetCost
 *                            (abstract_time_indexed_problem.cpp:299-342); T = 1: UnconstrainedEndPoseProblem::GetScalarCost
 *                            (unconstrained_end_pose_problem.cpp:79-82).
 * cost[b] = sum_{t = t_first}^{T-1} ct * ( sum_maps rho[t][map] |ydiff_map[t]|^2 + sum_i W[i] (x[t][i] - x[t-1][i])^2 ),
 * t_first = 1 for T > 1 (the reference's GetCost skips t = 0), 0 for end-pose problems. */
typedef struct
{
    int32_t T;           /* timesteps per trajectory (1: end-pose problem) */
    int32_t T_goal;      /* 1 (same goal at every timestep) or T */
    const double* rho;   /* [T][n_maps]: rho of every task map at every timestep (0: not a cost task) */
    const double* goal;  /* [T_goal][length_Phi] in the problem's Phi layout, or NULL = TaskSpaceVector::SetZero goal */
    double ct;           /* 1 / tau / T for time-indexed problems, 1 for end-pose problems */
    const double* W;     /* [N] diagonal of the transition-cost matrix W, or NULL (no transition cost) */
} exo_cost_desc;
int exo_problem_set_cost(exo_problem* problem, const exo_cost_desc* cost);
/* d_x / d_phi: the buffers of exo_problem_update_batch for n_traj * T evaluations in `layout`.  d_cost_traj[n_traj] and / or
 * d_cost_eval[n_traj * T] (per-timestep terms; 0 for t < t_first).  Asynchronous on `stream`. */
int exo_problem_cost_batch(exo_problem* problem, const double* d_x, const double* d_phi, int64_t n_traj, int32_t layout,
                           double* d_cost_traj, double* d_cost_eval, void* stream);
/* Host buffers: X[n_traj][T][N] (evaluation-major) -> cost[n_traj].  Chunked H2D -> forward kinematics + task maps (no
 * Jacobian is computed or moved) -> cost kernel -> D2H of 8 bytes per trajectory; returns when h_cost_traj is complete. */
int exo_problem_cost_batch_host(exo_problem* problem, const double* h_x, int64_t n_traj, int32_t layout, double* h_cost_traj);
/* Gradient of the task cost with respect to every state, from Phi and J already on the device (same layout; the gradient has
 * the shape and layout of d_x): AbstractTimeIndexedProblem::GetCostJacobian with GetScalarTaskJacobian = 2 ct J^T S ydiff
 * and GetScalarTransitionJacobian = 2 ct W xdiff (abstract_time_indexed_problem.cpp:309-342) --
 *   grad(b, t) = task(t) + transition(t) - transition(t + 1)   for t >= 1, and grad(b, 0) = 0 (the reference's row vector
 * starts at t = 1) -- or, for end-pose problems (T = 1), UnconstrainedEndPoseProblem::GetScalarJacobian = 2 J^T S ydiff
 * (unconstrained_end_pose_problem.cpp:84-87).  N doubles per evaluation leave the device instead of length_jacobian x N. */
int exo_problem_cost_jacobian_batch(exo_problem* problem, const double* d_x, const double* d_phi, const double* d_jac, int64_t n_traj,
                                    int32_t layout, double* d_grad, void* stream);
/* AICO task messages of every (trajectory, timestep) from the Phi / J of an update at the states d_qhat (same layout):
 * AICOSolver::GetTaskCosts (exotations/solvers/exotica_aico_solver/src/aico_solver.cpp:364-386) -- for every task with
 * precision rho > 0:  C += rho |ydiff|^2,  R += rho J^T J,  v = -ydiff + J qhat,  r += rho J^T v,  rhat += rho |v|^2;
 * d_cost[e] = ct * C.  EXO_LAYOUT_SOA: R[N][N][E], r[N][E]; evaluation-major layouts: R[E][N][N], r[E][N]; rhat[E], cost[E].
 * Together with exo_problem_cost_batch (GetScalarTaskCost + GetScalarTransitionCost per timestep = the solver's roll-out
 * cost, aico_solver.cpp:440-470) this is what AICO needs from the problem per sweep. */
int exo_problem_aico_task_messages_batch(exo_problem* problem, const double* d_qhat, const double* d_phi, const double* d_jac, int64_t n_traj,
                                         int32_t layout, double* d_R, double* d_r, double* d_rhat, double* d_cost, void* stream);
/* States in, cost and gradient out: d_cost_traj[n_traj] (may be NULL) and d_grad (shape and layout of d_x) as the two calls
 * above define them, without caller-side Phi / J buffers.  Serial-chain programs whose cost is made of EffFrame / EffPosition /
 * JointLimit maps run ONE fused kernel (exo_chain_costgrad_kernel): forward kinematics, ydiff with its SO(3) log, and every
 * Jacobian column dotted with S ydiff the moment it exists -- J is never materialised, N + 1 doubles per evaluation reach
 * HBM instead of (1 + N) x length_jacobian.  Other programs go through an internal scratch buffer and the three kernels
 * (exo_problem_set_kernel(1) forces that path; tests compare the two).  d_grad may be NULL: cost only (one kernel for chain
 * programs, exo_chain_cost_kernel; forward kinematics without Jacobians + the cost kernel otherwise).  Calls on one problem
 * must be serialised. */
int exo_problem_cost_gradient_batch(exo_problem* problem, const double* d_x, int64_t n_traj, int32_t layout, double* d_cost_traj,
                                    double* d_grad, void* stream);
/* exo_problem_update_batch and exo_problem_cost_batch in one call -- what a time-indexed solver does per roll-out: Update(x, t)
 * for every timestep (tasks.cpp:306-314 fills Phi, J and ydiff), then GetCost() (abstract_time_indexed_problem.cpp:299-307).
 * d_phi / d_jac (d_jac may be NULL) receive exactly what exo_problem_update_batch writes for the n_traj * T evaluations,
 * d_cost_traj[n_traj] and / or d_cost_eval[n_traj * T] what exo_problem_cost_batch writes.  EXO_LAYOUT_SOA buffers of problems
 * whose cost is made of EffFrame / EffPosition / JointLimit maps run ONE generated kernel (exo_tree_update_cost_kernel, compiled at
 * run time like exo_tree_kernel: every evaluation's cost is formed while its Phi entries are still in registers, so Phi is not
 * read back from HBM and the cost kernel's launch disappears) plus the per-trajectory sum; everything else is the two calls
 * back to back.  Asynchronous on `stream`; calls on one problem must be serialised. */
int exo_problem_update_cost_batch(exo_problem* problem, const double* d_x, int64_t n_traj, int32_t layout, double* d_phi, double* d_jac,
                                  double* d_cost_traj, double* d_cost_eval, void* stream);
/* Host buffers, states X[B][T][N] evaluation-major: H2D of the states -> forward kinematics + Jacobians (they stay on the
 * device) -> cost and cost-gradient kernels -> D2H of h_grad[B][T][N] and, if h_cost_traj is not NULL, one cost per
 * trajectory.  What a gradient-based solver on the host needs per iteration, at N + 1/T doubles per evaluation over PCIe. */
int exo_problem_cost_jacobian_batch_host(exo_problem* problem, const double* h_x, int64_t n_traj, int32_t layout, double* h_cost_traj,
                                         double* h_grad);
/* New output of SphereCollision (not in the reference): min over all sphere pairs of |p_i - p_j| - r_i - r_j, written per
 * evaluation to d_min_distance[n_evals] by every following exo_problem_update_batch (NULL switches it off). */
int exo_problem_set_min_distance_output(exo_problem* problem, double* d_min_distance);
#define EXO_REDUCE_MIN 0
#define EXO_REDUCE_SUM 1
#define EXO_REDUCE_MAX 2 /* exo_comm_allreduce only */
/* d_out[b] = min / sum over t in [t_first, T) of d_in[b * T + t] */
int exo_trajectory_reduce(const double* d_in, int64_t n_traj, int32_t T, int32_t t_first, int32_t op, double* d_out, void* stream);

/* ---- multi-GPU collectives (SURVEY.md section 8 row e) ----------------------------------------------------------------
 * The reference is single-process; its per-trajectory scalars are AbstractTimeIndexedProblem::GetCost()
 * (exotica_core/src/problems/abstract_time_indexed_problem.cpp:299-307) and, new here, the minimum SphereCollision surface
 * distance.  One process per GPU evaluates a contiguous block of whole trajectories (no data-path exchange); these calls are
 * the only exchange step: NCCL over NVLink / NVSwitch on the CALLER'S stream, right behind the kernel that produced their
 * input, no host synchronisation.  NCCL (libnccl.so.2, or EXO_NCCL_LIB) is loaded on first use; single-GPU callers never
 * need it (world == 1 communicators work without NCCL and degrade to device copies).
 *   exo_comm_get_unique_id  rank 0 creates the 128-byte id; the caller ships it to the other ranks (MPI_Bcast, a file, a
 *                           TCP store: the rendezvous is the launcher's business, as with ncclGetUniqueId)
 *   exo_comm_create         ncclCommInitRank on the current device (exo_set_device first); collective over all ranks
 *   exo_comm_allgather      d_recv[world][count_per_rank] <- every rank's d_send[count_per_rank]; in place when
 *                           d_send == d_recv + rank * count_per_rank
 *   exo_comm_allreduce      element-wise EXO_REDUCE_MIN / SUM / MAX over ranks (in place when d_send == d_recv)
 *   exo_comm_barrier        all ranks' streams meet (1-element all-reduce), then the stream is synchronised */
typedef struct exo_comm exo_comm;
#define EXO_COMM_ID_BYTES 128
int exo_comm_nccl_version(int* version);
int exo_comm_get_unique_id(unsigned char* id /* [EXO_COMM_ID_BYTES] */);
int exo_comm_create(int rank, int world, const unsigned char* id /* [EXO_COMM_ID_BYTES], may be NULL when world == 1 */, exo_comm** out);
int exo_comm_destroy(exo_comm* comm);
int exo_comm_rank(const exo_comm* comm);
int exo_comm_world_size(const exo_comm* comm);
int64_t exo_comm_collective_count(const exo_comm* comm); /* NCCL calls issued so far by this communicator */
/* Messages of at most 8192 doubles per rank do not go through NCCL: every rank owns a window in its HBM that all peers map
 * (cudaIpc, NVLink / NVSwitch peer access), and ONE kernel stores the rank's message into every peer's window -- as 16-byte cells
 * that carry their own sequence tags, so no fence or flag store follows (EXO_COMM_LL=0: plain doubles, a system-scope fence and one
 * flag per peer) --, polls its own window until the peers' cells are complete and copies / reduces the result into d_recv: 8.7 us
 * (measured back to back, 10 us inside a step's graph) instead of NCCL's 22-26 us for the 512-cost messages of an eight-GPU
 * multi-start step, and capturable in a CUDA graph.  Number of collectives that took
 * this path so far; -1 when the windows could not be mapped on every rank (the communicator then uses NCCL throughout;
 * EXO_COMM_PEER=0 forces that).  As with NCCL, the collectives of one communicator must reach every rank in the same order:
 * issue them on one stream, or order the streams yourself. */
int64_t exo_comm_peer_collective_count(const exo_comm* comm);
int exo_comm_allgather(exo_comm* comm, const double* d_send, double* d_recv, int64_t count_per_rank, void* stream);
int exo_comm_allreduce(exo_comm* comm, const double* d_send, double* d_recv, int64_t count, int32_t op, void* stream);
int exo_comm_barrier(exo_comm* comm, void* stream);
/* One batched call per step of a sharded multi-start optimisation: the costs of this rank's n_traj_local trajectories
 * (exo_problem_cost_gradient_batch with a NULL gradient: one fused kernel for chain programs) are written straight into this
 * rank's slot of d_cost_all[world][n_traj_per_rank] and all-gathered in place on `stream`; slots of ranks holding fewer than
 * n_traj_per_rank trajectories are zero-padded.  Every rank ends up with every trajectory's GetCost(). */
int exo_problem_cost_allgather(exo_problem* problem, exo_comm* comm, const double* d_x, int64_t n_traj_local, int64_t n_traj_per_rank,
                               int32_t layout, double* d_cost_all, void* stream);
/* SphereCollision: d_min_eval[n_traj_local * T] is the per-evaluation minimum surface distance an update wrote
 * (exo_problem_set_min_distance_output).  Per-trajectory minima go into this rank's slot of d_min_traj_all[world][n_traj_per_rank]
 * (padded with +inf) and are all-gathered in place; if d_min_global is not NULL it receives the minimum over all trajectories
 * of all ranks (all-reduce(min) of one double).  d_min_traj_all may be NULL when only the global minimum is wanted. */
int exo_min_distance_allgather(exo_comm* comm, const double* d_min_eval, int64_t n_traj_local, int32_t T, int64_t n_traj_per_rank,
                               double* d_min_traj_all, double* d_min_global, void* stream);

/* ---- task-map Hessians ---------------------------------------------------------------------------
 * exo_problem_hessian_batch <- KinematicTree::ComputeH (exotica_core/src/kinematic_tree.cpp:947-978) + the Hessian overloads of
 *     EffFrame / EffPosition / EffOrientation / EffPositionXY / Distance (exotations/exotica_core_task_maps/src/), the
 *     zero Hessian of JointPose and the base-class fallback J_row^T J_row of maps without an overload
 *     (exotica_core/src/task_map.cpp:99-108: JointLimit, SphereCollision).
 * d_x / d_jac: the state and Jacobian buffers of exo_problem_update_batch for the same n_evals, T and layout.
 * d_hess: EXO_LAYOUT_SOA  H[length_jacobian][N][N][E];  AoS layouts  H[E][length_jacobian][N][N]  (one N x N matrix per row of
 * the task-space Jacobian, as exotica::Hessian holds it).  Asynchronous on `stream`. */
int exo_problem_hessian_batch(exo_problem* problem, const double* d_x, const double* d_jac, int64_t n_evals, int32_t T, int32_t layout,
                              double* d_hess, void* stream);

/* ---- batched IK --------------------------------------------------------------------------------
 * exo_problem_ik_solve_batch <- IKSolver::Solve (exotations/solvers/exotica_ik_solver/src/ik_solver.cpp:73-240) on an
 *     UnconstrainedEndPoseProblem, for n_evals independent start states at once (multi-start IK): damped weighted
 *     pseudo-inverse step (Cholesky of S J W^-1 J^T S + lambda I, lambda x10 on failure), backtracking line search over
 *     alpha = linspace(1, 0.1, 10), the reference's termination tests.  Parameters: init/ik_solver.in + MaxIterations of
 *     exotica_core/init/motion_solver.in.  MaxIterations == 1 (the reference's interactive mode, exotica_examples/resources/configs/
 *     example_ik.xml:5) follows ik_solver.cpp:139-143,187-199: with MaxStep != 0 the line search is replaced by one step clamped to
 *     MaxStep per joint; with MaxStep == 0 lambda is adapted from the accepted step length and reaching MaximumRegularization ends
 *     the solve with Divergence.
 * Requires exo_problem_set_cost (T = 1: rho, goal, W).  d_q: SoA [N][n_evals] start states in, solutions out.  d_cost[n_evals]:
 * cost at the returned state; d_iterations[n_evals]; d_termination[n_evals]: exotica::TerminationCriterion values
 * (planning_problem.h:49-62: 0 IterationLimit, 2 StepTolerance, 3 FunctionTolerance, 4 GradientTolerance, 5 Divergence). */
typedef struct
{
    int32_t max_iterations;          /* MaxIterations = 100 */
    double tolerance;                /* Tolerance = 1e-5 */
    double regularization_rate;      /* RegularizationRate = 1e-9 */
    double maximum_regularization;   /* MaximumRegularization = 1e9 */
    double gradient_tolerance;       /* GradientToleranceConvergenceThreshold = 1e-9 */
    double step_tolerance;           /* StepToleranceConvergenceThreshold = 1e-5 */
    double max_step;                         /* MaxStep = 0 (only read when max_iterations == 1) */
    double threshold_regularization_increase;/* ThresholdRegularizationIncrease = 0.1 */
    double threshold_regularization_decrease;/* ThresholdRegularizationDecrease = 0.5 */
    double minimum_regularization;           /* MinimumRegularization = 1e-12 */
} exo_ik_params;
int exo_problem_ik_solve_batch(exo_problem* problem, double* d_q, int64_t n_evals, const exo_ik_params* params, double* d_cost,
                               int32_t* d_iterations, int32_t* d_termination, void* stream);
#endif /* __CUDACC_RTC__ */

#ifdef __cplusplus
}
#endif
#endif /* EXOTICA_B200_H_ */
