/*
 * exotica_b200_host.h -- C-ABI over the C++ host mirror (exotica_b200/csrc/host/exotica_host.h) for language
 * bindings.  The reference binds its C++ classes with pybind11 (exotica_python/src/pyexotica.cpp); here the same
 * subset is exposed as plain C so that the Python package only forwards (ctypes), with no logic of its own.
 *
 * Every function returns EXO_OK or an error code with the message in exo_last_error() (the host classes throw
 * exotica::Exception; the wrappers translate).  String results are '\n'-joined lists owned by the handle and valid
 * until the next call on it.
 */
#ifndef EXOTICA_B200_HOST_H_
#define EXOTICA_B200_HOST_H_

#include "exotica_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct exo_host_scene exo_host_scene;     /* exotica::Scene (+ its KinematicTree) */
typedef struct exo_host_problem exo_host_problem; /* exotica::PlanningProblem */

/* Setup: {package} resolution without ROS (tools.cpp:150-190) and the registered plugin names
 * (pyexotica.cpp:683-706 Setup.get_maps / get_problems). */
int exo_host_set_package_path(const char* package, const char* directory);
const char* exo_host_registered_maps(void);
const char* exo_host_registered_problems(void);
/* exotica::GetFrame (tools/conversions.cpp:52-69): n in {3,6,7}. */
int exo_host_get_frame(const double* v, int n, double out12[12]);

/* ---- Scene / KinematicTree (pyexotica.cpp:1221-1300,1376-1455) ---- */
int exo_host_scene_create(const char* urdf_xml, const char* srdf_xml, const char* joint_group, exo_host_scene** out);
int exo_host_scene_destroy(exo_host_scene* scene);
/* which: 0 controlled joint names, 1 controlled link names, 2 model joint names, 3 model link names (all tree
 * elements, scene objects included), 4 root frame name, 5 root joint name */
const char* exo_host_scene_names(exo_host_scene* scene, int which);
int exo_host_scene_num_controlled(exo_host_scene* scene);
int exo_host_scene_num_elements(exo_host_scene* scene);
int exo_host_scene_base_types(exo_host_scene* scene, int32_t out[2]); /* model, controlled: 0 fixed, 10 floating, 20 planar */
/* flattened tree for cross-checks: id, parent, joint_type, control_id [n]; origin, axis [n][3]; f_tip [n][12] */
int exo_host_scene_tree(exo_host_scene* scene, int32_t* id, int32_t* parent, int32_t* joint_type, int32_t* control_id,
                        double* origin, double* axis, double* f_tip);
int exo_host_scene_element_index(exo_host_scene* scene, const char* link_name); /* -1 if unknown */
int exo_host_scene_joint_limits(exo_host_scene* scene, double* lower, double* upper);
int exo_host_scene_set_joint_limits(exo_host_scene* scene, const double* lower, const double* upper);
int exo_host_scene_set_seed(exo_host_scene* scene, unsigned int seed);
int exo_host_scene_random_state(exo_host_scene* scene, double* x);
int exo_host_scene_get_model_state(exo_host_scene* scene, double* state /* [num model joints] */);
int exo_host_scene_set_model_state(exo_host_scene* scene, const double* state, int n);
int exo_host_scene_num_model_joints(exo_host_scene* scene);
int exo_host_scene_add_object(exo_host_scene* scene, const char* name, const double frame12[12], const char* parent);
/* Scene::LoadScene / LoadSceneFile / CleanScene (scene.cpp:610-660): MoveIt ".scene" geometry text; every object becomes a frame
 * (its first shape's pose) with one "<name>_collision_<i>" child per shape.  offset12 NULL = identity. */
int exo_host_scene_load_scene(exo_host_scene* scene, const char* scene_text, const double offset12[12]);
int exo_host_scene_load_scene_file(exo_host_scene* scene, const char* file_name, const double offset12[12]);
int exo_host_scene_clean(exo_host_scene* scene);
int exo_host_scene_add_trajectory(exo_host_scene* scene, const char* link, int T, const double* frames /*[T][12]*/);
/* Scene::AddTrajectory(link, traj) (scene.cpp:922-935) with the reference's trajectory text format
 * ("radius\n rows cols\n time x y z [rotation]..." -- exotica_core/src/trajectory.cpp:51-68); time-indexed problems sample it
 * at t * tau for every timestep */
int exo_host_scene_add_trajectory_string(exo_host_scene* scene, const char* link, const char* trajectory);
/* Scene::GetTrajectory(link)->ToString() into out[capacity] */
int exo_host_scene_get_trajectory(exo_host_scene* scene, const char* link, char* out, int capacity);
/* exotica::Trajectory::GetPosition at n times -> frames[n][12]; duration may be NULL */
int exo_host_trajectory_sample(const char* trajectory, int n, const double* times, double* frames, double* duration);
/* KinematicTree::RequestFrames / Update / UpdateBatch: link names '\n'-joined, offsets [F][12] (NULL = identity) */
int exo_host_scene_request_frames(exo_host_scene* scene, int n_frames, const char* links_a, const double* off_a,
                                  const char* links_b, const double* off_b, int with_jacobian);
int exo_host_scene_update(exo_host_scene* scene, const double* x, int n, double t, double* phi /*[F][12]*/, double* jac /*[F][6*N] col-major*/);
int exo_host_scene_update_batch(exo_host_scene* scene, const double* x, int64_t n_evals, int layout, double* phi, double* jac);
/* Scene.fk / Scene.jacobian at the current model state */
int exo_host_scene_fk(exo_host_scene* scene, const char* link_a, const double* off_a, const char* link_b, const double* off_b, double out12[12]);
int exo_host_scene_jacobian(exo_host_scene* scene, const char* link_a, const double* off_a, const char* link_b, const double* off_b, double* out /*6xN col-major*/);

/* ---- Problems (pyexotica.cpp:935-1060) ---- */
int exo_host_load_problem_xml(const char* xml_text, exo_host_problem** out); /* Setup::LoadProblem on XML text */
int exo_host_problem_destroy(exo_host_problem* problem);
int exo_host_problem_dims(exo_host_problem* problem, int32_t out[4]); /* N, length_Phi, length_jacobian, T (1 for end-pose) */
const char* exo_host_problem_type(exo_host_problem* problem);
int exo_host_problem_num_tasks(exo_host_problem* problem);
const char* exo_host_problem_task_name(exo_host_problem* problem, int task);
int exo_host_problem_task_indexing(exo_host_problem* problem, int task, int32_t out[6]); /* id,start,length,start_jacobian,length_jacobian,is_used */
exo_host_scene* exo_host_problem_scene(exo_host_problem* problem); /* borrowed, owned by the problem */
int exo_host_problem_start_state(exo_host_problem* problem, double* x);
/* One evaluation: t < 0 -> UnconstrainedEndPoseProblem::Update(x); t >= 0 -> TimeIndexedProblem::Update(x, t).
 * phi[length_Phi], jac[length_jacobian x N] column-major (the Eigen::MatrixXd the reference exposes). */
int exo_host_problem_update(exo_host_problem* problem, const double* x, int n, int t, double* phi, double* jac);
/* AbstractTimeIndexedProblem::Update(x_trajectory) (abstract_time_indexed_problem.cpp:209-218): n = (T-1)*N values, timesteps
 * 1..T-1 are evaluated (one launch); t = 0 keeps what the last Update(x, 0) left.  phi[T][length_Phi], jac[T][length_jacobian x N]
 * column-major: the problem's Phi / jacobian members after the call. */
int exo_host_problem_update_trajectory(exo_host_problem* problem, const double* x, int n, double* phi, double* jac);
/* NEW batched PlanningProblem::Update: E evaluations (E = B*T for time-indexed problems), host buffers. */
int exo_host_problem_update_batch(exo_host_problem* problem, const double* x, int64_t n_evals, int layout, double* phi, double* jac);
int exo_host_problem_update_batch_device(exo_host_problem* problem, const double* d_x, int64_t n_evals, int layout, double* d_phi,
                                         double* d_jac, void* stream);
/* NEW scalar cost of every trajectory: X[n_traj][T][N] (T = 1: X[n_traj][N]) -> cost[n_traj]
 * (GetCost, abstract_time_indexed_problem.cpp:299-307; GetScalarCost, unconstrained_end_pose_problem.cpp:79-82). */
int exo_host_problem_cost_batch(exo_host_problem* problem, const double* x, int64_t n_traj, double* cost);
/* NEW gradient of that cost with respect to every state: jacobian[n_traj][T][N], row t = 0 zero (GetCostJacobian,
 * abstract_time_indexed_problem.cpp:309-318; T = 1: GetScalarJacobian, unconstrained_end_pose_problem.cpp:84-87);
 * cost[n_traj] may be NULL.  Phi and J stay on the device. */
int exo_host_problem_cost_jacobian_batch(exo_host_problem* problem, const double* x, int64_t n_traj, double* cost, double* jacobian);
int exo_host_problem_cost_batch_device(exo_host_problem* problem, const double* d_x, const double* d_phi, int64_t n_traj, int layout,
                                       double* d_cost, double* d_cost_eval, void* stream);
/* NEW Update(x, t) of every timestep + GetCost() of every trajectory on device buffers (exo_problem_update_cost_batch): d_phi /
 * d_jac as exo_host_problem_update_batch_device fills them, d_cost[n_traj] / d_cost_eval[n_traj * T] (either may be NULL) as
 * exo_host_problem_cost_batch_device does. */
int exo_host_problem_update_cost_batch_device(exo_host_problem* problem, const double* d_x, int64_t n_traj, int layout, double* d_phi, double* d_jac,
                                              double* d_cost, double* d_cost_eval, void* stream);
/* NEW multi-start IK: IKSolver::Solve (exotations/solvers/exotica_ik_solver/src/ik_solver.cpp:73-240) from n start states
 * q0[n][N] -> solutions[n][N], cost[n], iterations[n], termination[n] (exotica::TerminationCriterion values). */
int exo_host_problem_ik_solve_batch(exo_host_problem* problem, const double* q0, int64_t n, const exo_ik_params* params, double* solutions,
                                    double* cost, int32_t* iterations, int32_t* termination);
/* NEW exotica::Hessian of every evaluation: x[n][N] -> hess[n][length_jacobian][N][N]. */
int exo_host_problem_hessian_batch(exo_host_problem* problem, const double* x, int64_t n_evals, double* hess);
int exo_host_problem_last_kernel_ms(exo_host_problem* problem, float* ms);
int64_t exo_host_problem_launch_count(exo_host_problem* problem);
int64_t exo_host_problem_update_count(exo_host_problem* problem); /* number_of_problem_updates_ */
exo_problem* exo_host_problem_backend(exo_host_problem* problem);

#ifdef __cplusplus
}
#endif
#endif
