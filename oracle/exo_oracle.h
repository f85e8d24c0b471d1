/*
 * exo_oracle.h -- CPU oracle for the EXOTica batched-kinematics hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C, double-precision, one-configuration-at-a-time
 * restatement of the reference algorithm (ipab-slmc/exotica v6.2.0):
 *
 *   KinematicTree::Update/UpdateTree/FK/ComputeJ   exotica_core/src/kinematic_tree.cpp:678-737,829-867,919-945
 *   KinematicElement::GetPose                      exotica_core/include/exotica_core/kinematic_element.h:93-107
 *   SetRotation / GetFrame                         exotica_core/src/tools/conversions.cpp:52-69,150-179
 *   EffFrame / EffPosition / JointLimit / SphereCollision::Update
 *                                                  exotations/exotica_core_task_maps/src/{eff_frame.cpp:63-73,
 *                                                  eff_position.cpp:46-55,joint_limit.cpp:63-99,sphere_collision.cpp:72-162}
 *   AbstractTimeIndexedProblem::Update(x,t)        exotica_core/src/problems/abstract_time_indexed_problem.cpp:220-292
 *   TaskSpaceVector::SetZero                       exotica_core/src/task_space_vector.cpp:54-62
 *
 * The frame/twist arithmetic of orocos_kdl (liborocos-kdl 1.4.0 on the reference's CI, NOT vendored
 * in the reference, unpinned in exotica_core/package.xml:22-23) and Eigen's Quaternion(Matrix3)
 * conversion are restated from their published algorithms.
 *
 * Parity status: pinned against every known-answer vector the reference's tests hold for maps on this path
 * (tests/golden/reference_goldens.json, extracted from exotations/exotica_core_task_maps/test/test_maps.cpp: EffPosition
 * :340-347, Distance :493-500, SphereCollision :651-673, JointPose :700-780, CenterOfMass :789-892) and the ten Valkyrie
 * centre-of-mass goldens of exotica_examples/test/test_valkyrie_com:13-120, in tests/test_oracle_golden.py.  EffFrame /
 * EffOrientation rotation encodings, EffAxisAlignment, EffPositionXY and JointLimit have no golden values in the reference:
 * "parity unpinned" beyond the reference's own finite-difference tests (test_maps.cpp:115-200), which
 * tests/test_oracle_fd.py reproduces on this oracle.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use
 * this library.  The product (exotica_b200/) never links, loads or calls it.
 */
#ifndef EXO_ORACLE_H_
#define EXO_ORACLE_H_

#ifdef __cplusplus
extern "C" {
#endif

/* KDL::Joint::JointType order */
enum
{
    EXO_ORC_ROT_AXIS = 0,
    EXO_ORC_ROT_X = 1,
    EXO_ORC_ROT_Y = 2,
    EXO_ORC_ROT_Z = 3,
    EXO_ORC_TRANS_AXIS = 4,
    EXO_ORC_TRANS_X = 5,
    EXO_ORC_TRANS_Y = 6,
    EXO_ORC_TRANS_Z = 7,
    EXO_ORC_NONE = 8
};

/* exotica::RotationType order (tools/conversions.h:80-88) */
enum
{
    EXO_ORC_QUATERNION = 0,
    EXO_ORC_RPY = 1,
    EXO_ORC_ZYX = 2,
    EXO_ORC_ZYZ = 3,
    EXO_ORC_ANGLE_AXIS = 4,
    EXO_ORC_MATRIX = 5
};

enum
{
    EXO_ORC_MAP_EFF_FRAME = 0,
    EXO_ORC_MAP_EFF_POSITION = 1,
    EXO_ORC_MAP_JOINT_LIMIT = 2,
    EXO_ORC_MAP_SPHERE_COLLISION = 3,
    /* 4 is the C-ABI's raw-frames pseudo map; numbering follows exo_map_kind */
    EXO_ORC_MAP_EFF_ORIENTATION = 5,
    EXO_ORC_MAP_EFF_POSITION_XY = 6,
    EXO_ORC_MAP_DISTANCE = 7,
    EXO_ORC_MAP_JOINT_POSE = 8,
    EXO_ORC_MAP_EFF_AXIS_ALIGNMENT = 9, /* eff_axis_alignment.cpp */
    EXO_ORC_MAP_CENTER_OF_MASS = 10     /* center_of_mass.cpp */
};

typedef struct exo_oracle exo_oracle;

/* Tree in model_tree_ order (index 0 = world root element, id -1).  Frames are 12 doubles:
 * KDL::Rotation::data[9] (row-major) followed by p[3]. */
exo_oracle* exo_oracle_create(int n_elem, const int* id, const int* parent, const int* joint_type,
                              const double* joint_origin /*[n][3]*/, const double* joint_axis /*[n][3]*/,
                              const double* f_tip /*[n][12]*/, const int* control_id,
                              const int* mimic_elem /* element index or -1 */, const double* mimic_mult,
                              const double* mimic_off, const int* generated_index /* -1 or slot in the per-timestep table */,
                              int n_controlled);
exo_oracle* exo_oracle_clone(const exo_oracle* o);
void exo_oracle_destroy(exo_oracle* o);

/* tree_state_ for every element by element INDEX (uncontrolled joints keep these values). */
void exo_oracle_set_tree_state(exo_oracle* o, const double* state);
void exo_oracle_set_limits(exo_oracle* o, const double* lower, const double* upper);
/* generated_offset of a trajectory-driven element (scene.cpp:329-335). */
void exo_oracle_set_generated(exo_oracle* o, int slot, const double* frame12);

/* KinematicTree::RequestFrames (kinematic_tree.cpp:628-676); elements by index. */
int exo_oracle_request_frames(exo_oracle* o, int n_frames, const int* elem_a, const double* off_a /*[F][12]*/,
                              const int* elem_b, const double* off_b /*[F][12]*/);
/* KinematicTree::Update (kinematic_tree.cpp:678-693) with KIN_FK | KIN_J. */
int exo_oracle_update(exo_oracle* o, const double* x);
void exo_oracle_get_phi(const exo_oracle* o, double* out /*[F][12]*/);
void exo_oracle_get_jacobian(const exo_oracle* o, double* out /*[F][6*N] column-major 6xN*/);
void exo_oracle_get_element_frame(const exo_oracle* o, int elem, double* out12);

/* exotica::SetRotation (conversions.cpp:150-179); returns the encoding length. */
int exo_oracle_set_rotation(const double* rot9, int type, double* out);
/* exotica::GetFrame(vector) (conversions.cpp:52-69), n in {3,6,7}; returns 0 on success. */
int exo_oracle_get_frame(const double* v, int n, double* out12);

/* Problem = ordered list of task maps over consecutive slices of the requested frames
 * (PlanningProblem::InstantiateBase, planning_problem.cpp:206-238). */
void exo_oracle_problem_clear(exo_oracle* o);
int exo_oracle_problem_add_map(exo_oracle* o, int kind, int frame_start, int n_frames, int rot_type, double param,
                               const int* sphere_group /*[n_frames] lexicographic group rank*/,
                               const double* sphere_radius /*[n_frames]*/, int is_used);
int exo_oracle_problem_length_phi(const exo_oracle* o);
int exo_oracle_problem_length_jacobian(const exo_oracle* o);
/* One evaluation: Scene::Update + Phi.SetZero + jacobian.setZero + used task maps.
 * Phi[length_Phi], J column-major [length_jacobian x N]. */
int exo_oracle_problem_update(exo_oracle* o, const double* x, double* phi, double* jac);
/* E evaluations, one at a time.  X[E][N]; evaluation e uses timestep t = e % T of the generated
 * table gen[T][n_gen][12] (may be NULL).  Phi[E][length_Phi]; J[E][length_jacobian*N] column-major
 * per evaluation (NULL to discard outputs: timing mode).  n_threads > 1 uses one cloned tree per
 * thread over contiguous chunks. */
int exo_oracle_problem_update_batch(exo_oracle* o, const double* X, long E, int T, const double* gen, int n_gen,
                                    double* phi, double* jac, int n_threads);

/* ---- task cost ----
 * exotica::GetRotation (tools/conversions.cpp:114-148); TaskSpaceVector::operator- (task_space_vector.cpp:64-91);
 * AbstractTimeIndexedProblem::GetCost (abstract_time_indexed_problem.cpp:299-342). */
int exo_oracle_get_rotation(const double* data, int type, double* rot9);
int exo_oracle_problem_ydiff(const exo_oracle* o, const double* phi, const double* goal, double* ydiff);
double exo_oracle_trajectory_cost(exo_oracle* o, const double* X, int T, int t_first, const double* rho, const double* goal, int T_goal,
                                  double ct, const double* W, const double* gen, int n_gen, double* cost_t);
void exo_oracle_trajectory_cost_jacobian(exo_oracle* o, const double* X, int T, int t_first, const double* rho, const double* goal, int T_goal,
                                         double ct, const double* W, const double* gen, int n_gen, double* grad);
void exo_oracle_aico_task_messages(exo_oracle* o, const double* Q, int T, const double* rho, const double* goal, int T_goal, double ct,
                                   const double* gen, int n_gen, double* R, double* r, double* rhat, double* cost);

/* ---- task-map Hessians: KinematicTree::ComputeH (kinematic_tree.cpp:947-978) + the maps' Hessian
 * overloads + the base-class fallback (task_map.cpp:99-108).  hess[length_jacobian][N][N]. */
int exo_oracle_problem_hessian(exo_oracle* o, const double* x, double* hess);

/* ---- IKSolver::Solve (exotations/solvers/exotica_ik_solver/src/ik_solver.cpp:73-240) ---- */
int exo_oracle_ik_solve(exo_oracle* o, double* q, const double* rho, const double* goal, const double* W, int max_iterations, double tolerance,
                        double lambda, double regmax, double gradient_tolerance, double step_tolerance, double* cost_out, int* iterations_out);
int exo_oracle_ik_solve_ex(exo_oracle* o, double* q, const double* rho, const double* goal, const double* W, int max_iterations, double tolerance,
                           double lambda, double regmax, double gradient_tolerance, double step_tolerance, double max_step, double th_stepinc,
                           double th_stepdec, double regmin, double* cost_out, int* iterations_out, double* lambda_out);

#ifdef __cplusplus
}
#endif
#endif
