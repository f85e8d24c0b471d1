// host_capi.cpp -- C wrappers over the C++ host mirror (include/exotica_b200_host.h).
#include <string.h>

#include <sstream>

#include "../../../include/exotica_b200_host.h"
#include "exotica_host.h"

using namespace exotica;

// defined in capi.cu: sets the thread-local message returned by exo_last_error()
extern "C" void exo_internal_set_error(const char* msg);

struct exo_host_scene
{
    ScenePtr scene;
    std::string scratch;
    bool with_jacobian = true;
    int n_frames = 0;
};

struct exo_host_problem
{
    PlanningProblemPtr problem;
    exo_host_scene scene_view;
    std::string scratch;
    std::string type;
};

namespace
{
int Fail(const std::exception& e)
{
    exo_internal_set_error(e.what());
    return EXO_ERR_INVALID;
}

#define HOST_TRY(...)                    \
    try                                  \
    {                                    \
        __VA_ARGS__;                     \
        return EXO_OK;                   \
    }                                    \
    catch (const std::exception& e)      \
    {                                    \
        return Fail(e);                  \
    }

std::string Join(const std::vector<std::string>& v)
{
    std::string out;
    for (size_t i = 0; i < v.size(); ++i)
    {
        if (i) out += "\n";
        out += v[i];
    }
    return out;
}

std::vector<std::string> Split(const char* s, int expected)
{
    std::vector<std::string> out;
    if (!s)
    {
        out.assign(expected, "");
        return out;
    }
    std::string cur;
    for (const char* p = s;; ++p)
    {
        if (*p == '\n' || *p == 0)
        {
            out.push_back(cur);
            cur.clear();
            if (*p == 0) break;
        }
        else
            cur += *p;
    }
    if ((int)out.size() != expected) ThrowPretty("expected " + std::to_string(expected) + " link names, got " + std::to_string(out.size()));
    return out;
}

Frame FrameOr(const double* f)
{
    Frame out = IdentityFrame();
    if (f) memcpy(out.data(), f, 12 * sizeof(double));
    return out;
}
}  // namespace

extern "C" {

int exo_host_set_package_path(const char* package, const char* directory)
{
    HOST_TRY(Setup::SetPackagePath(package ? package : "", directory ? directory : ""))
}

const char* exo_host_registered_maps(void)
{
    static thread_local std::string s;
    s = Join(Setup::GetMaps());
    return s.c_str();
}

const char* exo_host_registered_problems(void)
{
    static thread_local std::string s;
    s = Join(Setup::GetProblems());
    return s.c_str();
}

int exo_host_get_frame(const double* v, int n, double out12[12])
{
    HOST_TRY({
        if (!v || !out12 || n < 0) ThrowPretty("NULL argument");
        const Frame f = GetFrame(std::vector<double>(v, v + n));
        memcpy(out12, f.data(), 12 * sizeof(double));
    })
}

int exo_host_scene_create(const char* urdf_xml, const char* srdf_xml, const char* joint_group, exo_host_scene** out)
{
    HOST_TRY({
        if (!urdf_xml || !out) ThrowPretty("NULL argument");
        exo_host_scene* h = new exo_host_scene();
        try
        {
            h->scene.reset(new Scene());
            h->scene->InstantiateFromStrings(urdf_xml, srdf_xml ? srdf_xml : "", joint_group ? joint_group : "");
        }
        catch (...)
        {
            delete h;
            throw;
        }
        *out = h;
    })
}

int exo_host_scene_destroy(exo_host_scene* scene)
{
    delete scene;
    return EXO_OK;
}

const char* exo_host_scene_names(exo_host_scene* s, int which)
{
    if (!s) return "";
    KinematicTree& kt = s->scene->GetKinematicTree();
    switch (which)
    {
        case 0: s->scratch = Join(kt.GetControlledJointNames()); break;
        case 1: s->scratch = Join(kt.GetControlledLinkNames()); break;
        case 2: s->scratch = Join(kt.GetModelJointNames()); break;
        case 3:
        {
            std::vector<std::string> names;
            for (const KinematicElement& e : kt.GetTree()) names.push_back(e.link_name);
            s->scratch = Join(names);
            break;
        }
        case 4: s->scratch = kt.GetRootFrameName(); break;
        case 5: s->scratch = kt.GetRootJointName(); break;
        case 6: s->scratch = Join(kt.GetModelLinkNames()); break;
        case 7:
        {
            std::vector<std::string> names;
            for (const KinematicElement& e : kt.GetTree()) names.push_back(e.joint_name);
            s->scratch = Join(names);
            break;
        }
        default: s->scratch.clear();
    }
    return s->scratch.c_str();
}

int exo_host_scene_num_controlled(exo_host_scene* s)
{
    return s ? s->scene->GetKinematicTree().GetNumControlledJoints() : -1;
}

int exo_host_scene_num_elements(exo_host_scene* s)
{
    return s ? (int)s->scene->GetKinematicTree().GetTree().size() : -1;
}

int exo_host_scene_num_model_joints(exo_host_scene* s)
{
    return s ? s->scene->GetKinematicTree().GetNumModelJoints() : -1;
}

int exo_host_scene_base_types(exo_host_scene* s, int32_t out[2])
{
    HOST_TRY({
        out[0] = (int)s->scene->GetKinematicTree().GetModelBaseType();
        out[1] = (int)s->scene->GetKinematicTree().GetControlledBaseType();
    })
}

int exo_host_scene_tree(exo_host_scene* s, int32_t* id, int32_t* parent, int32_t* joint_type, int32_t* control_id, double* origin, double* axis, double* f_tip)
{
    HOST_TRY({
        const std::vector<KinematicElement>& t = s->scene->GetKinematicTree().GetTree();
        for (size_t i = 0; i < t.size(); ++i)
        {
            if (id) id[i] = t[i].id;
            if (parent) parent[i] = t[i].parent;
            if (joint_type) joint_type[i] = t[i].joint_type;
            if (control_id) control_id[i] = t[i].control_id;
            if (origin) memcpy(origin + 3 * i, t[i].origin, 3 * sizeof(double));
            if (axis) memcpy(axis + 3 * i, t[i].axis, 3 * sizeof(double));
            if (f_tip) memcpy(f_tip + 12 * i, t[i].f_tip.data(), 12 * sizeof(double));
        }
    })
}

int exo_host_scene_element_index(exo_host_scene* s, const char* link_name)
{
    if (!s || !link_name) return -1;
    return s->scene->GetKinematicTree().GetElementIndex(link_name);
}

int exo_host_scene_joint_limits(exo_host_scene* s, double* lower, double* upper)
{
    HOST_TRY({
        KinematicTree& kt = s->scene->GetKinematicTree();
        memcpy(lower, kt.GetLowerLimits().data(), kt.GetNumControlledJoints() * sizeof(double));
        memcpy(upper, kt.GetUpperLimits().data(), kt.GetNumControlledJoints() * sizeof(double));
    })
}

int exo_host_scene_set_joint_limits(exo_host_scene* s, const double* lower, const double* upper)
{
    HOST_TRY({
        KinematicTree& kt = s->scene->GetKinematicTree();
        const int n = kt.GetNumControlledJoints();
        if (lower) kt.SetJointLimitsLower(std::vector<double>(lower, lower + n));
        if (upper) kt.SetJointLimitsUpper(std::vector<double>(upper, upper + n));
    })
}

int exo_host_scene_set_seed(exo_host_scene* s, unsigned int seed)
{
    HOST_TRY(s->scene->GetKinematicTree().SetSeed(seed))
}

int exo_host_scene_random_state(exo_host_scene* s, double* x)
{
    HOST_TRY({
        const std::vector<double> q = s->scene->GetKinematicTree().GetRandomControlledState();
        memcpy(x, q.data(), q.size() * sizeof(double));
    })
}

int exo_host_scene_get_model_state(exo_host_scene* s, double* state)
{
    HOST_TRY({
        const std::vector<double> q = s->scene->GetModelState();
        memcpy(state, q.data(), q.size() * sizeof(double));
    })
}

int exo_host_scene_set_model_state(exo_host_scene* s, const double* state, int n)
{
    HOST_TRY(s->scene->SetModelState(std::vector<double>(state, state + n)))
}

int exo_host_scene_add_object(exo_host_scene* s, const char* name, const double frame12[12], const char* parent)
{
    HOST_TRY(s->scene->AddObject(name, FrameOr(frame12), parent ? parent : ""))
}

int exo_host_scene_load_scene(exo_host_scene* s, const char* scene_text, const double offset12[12])
{
    HOST_TRY(s->scene->LoadScene(scene_text ? scene_text : "", FrameOr(offset12)))
}

int exo_host_scene_load_scene_file(exo_host_scene* s, const char* file_name, const double offset12[12])
{
    HOST_TRY(s->scene->LoadSceneFile(file_name ? file_name : "", FrameOr(offset12)))
}

int exo_host_scene_clean(exo_host_scene* s)
{
    HOST_TRY(s->scene->CleanScene())
}

int exo_host_scene_add_trajectory(exo_host_scene* s, const char* link, int T, const double* frames)
{
    HOST_TRY({
        if (T < 1 || !frames) ThrowPretty("trajectory needs T >= 1 frames");
        std::vector<Frame> f(T);
        for (int t = 0; t < T; ++t) memcpy(f[t].data(), frames + 12 * t, 12 * sizeof(double));
        s->scene->AddTrajectoryFrames(link, f);
    })
}

int exo_host_scene_add_trajectory_string(exo_host_scene* s, const char* link, const char* trajectory)
{
    HOST_TRY(s->scene->AddTrajectory(link, std::string(trajectory ? trajectory : "")))
}

int exo_host_scene_get_trajectory(exo_host_scene* s, const char* link, char* out, int capacity)
{
    HOST_TRY({
        const std::string text = s->scene->GetTrajectory(link)->ToString();
        if ((int)text.size() + 1 > capacity) ThrowPretty("trajectory text does not fit the buffer");
        memcpy(out, text.c_str(), text.size() + 1);
    })
}

int exo_host_trajectory_sample(const char* trajectory, int n, const double* times, double* frames, double* duration)
{
    HOST_TRY({
        exotica::Trajectory traj(std::string(trajectory ? trajectory : ""));
        for (int i = 0; i < n; ++i)
        {
            const Frame f = traj.GetPosition(times[i]);
            memcpy(frames + 12 * i, f.data(), 12 * sizeof(double));
        }
        if (duration) *duration = traj.GetDuration();
    })
}

int exo_host_scene_request_frames(exo_host_scene* s, int n_frames, const char* links_a, const double* off_a, const char* links_b, const double* off_b, int with_jacobian)
{
    HOST_TRY({
        const std::vector<std::string> a = Split(links_a, n_frames), b = Split(links_b, n_frames);
        KinematicsRequest req;
        req.flags = with_jacobian ? (KIN_FK | KIN_J) : KIN_FK;
        for (int i = 0; i < n_frames; ++i) req.frames.push_back(KinematicFrameRequest(a[i], FrameOr(off_a ? off_a + 12 * i : nullptr), b[i], FrameOr(off_b ? off_b + 12 * i : nullptr)));
        s->scene->RequestKinematics(req);
        s->with_jacobian = with_jacobian != 0;
        s->n_frames = n_frames;
    })
}

int exo_host_scene_update(exo_host_scene* s, const double* x, int n, double t, double* phi, double* jac)
{
    HOST_TRY({
        s->scene->Update(std::vector<double>(x, x + n), t);
        std::shared_ptr<KinematicResponse> r = s->scene->GetKinematicTree().GetKinematicResponse();
        if (!r) return EXO_OK;  // no frames requested: the state and the trajectory-driven links were updated, nothing to return
        const int N = s->scene->GetKinematicTree().GetNumControlledJoints();
        for (size_t f = 0; f < r->Phi.size(); ++f)
        {
            if (phi) memcpy(phi + 12 * f, r->Phi[f].data(), 12 * sizeof(double));
            if (jac && !r->jacobian.empty()) memcpy(jac + 6 * (size_t)N * f, r->jacobian[f].data(), 6 * (size_t)N * sizeof(double));
        }
    })
}

int exo_host_scene_update_batch(exo_host_scene* s, const double* x, int64_t n_evals, int layout, double* phi, double* jac)
{
    HOST_TRY(s->scene->GetKinematicTree().UpdateBatch(x, n_evals, layout, phi, s->with_jacobian ? jac : nullptr))
}

int exo_host_scene_fk(exo_host_scene* s, const char* link_a, const double* off_a, const char* link_b, const double* off_b, double out12[12])
{
    HOST_TRY({
        const Frame f = s->scene->GetKinematicTree().FK(link_a ? link_a : "", FrameOr(off_a), link_b ? link_b : "", FrameOr(off_b));
        memcpy(out12, f.data(), 12 * sizeof(double));
    })
}

int exo_host_scene_jacobian(exo_host_scene* s, const char* link_a, const double* off_a, const char* link_b, const double* off_b, double* out)
{
    HOST_TRY({
        const std::vector<double> j = s->scene->GetKinematicTree().Jacobian(link_a ? link_a : "", FrameOr(off_a), link_b ? link_b : "", FrameOr(off_b));
        memcpy(out, j.data(), j.size() * sizeof(double));
    })
}

// ---- problems ------------------------------------------------------------------------------------------------------
int exo_host_load_problem_xml(const char* xml_text, exo_host_problem** out)
{
    HOST_TRY({
        if (!xml_text || !out) ThrowPretty("NULL argument");
        const Initializer init = Setup::LoadProblemInitializer(xml_text);
        exo_host_problem* h = new exo_host_problem();
        try
        {
            h->problem = Setup::CreateProblem(init);
            h->type = init.name;
            h->scene_view.scene = h->problem->GetScene();
        }
        catch (...)
        {
            delete h;
            throw;
        }
        *out = h;
    })
}

int exo_host_problem_destroy(exo_host_problem* p)
{
    delete p;
    return EXO_OK;
}

int exo_host_problem_dims(exo_host_problem* p, int32_t out[4])
{
    HOST_TRY({
        out[0] = p->problem->N;
        out[1] = p->problem->length_Phi;
        out[2] = p->problem->length_jacobian;
        out[3] = p->problem->GetT();
    })
}

const char* exo_host_problem_type(exo_host_problem* p)
{
    return p ? p->type.c_str() : "";
}

int exo_host_problem_num_tasks(exo_host_problem* p)
{
    return p ? (int)p->problem->GetTasks().size() : -1;
}

const char* exo_host_problem_task_name(exo_host_problem* p, int task)
{
    if (!p || task < 0 || task >= (int)p->problem->GetTasks().size()) return "";
    p->scratch = p->problem->GetTasks()[task]->GetObjectName();
    return p->scratch.c_str();
}

int exo_host_problem_task_indexing(exo_host_problem* p, int task, int32_t out[6])
{
    HOST_TRY({
        if (task < 0 || task >= (int)p->problem->GetTasks().size()) ThrowPretty("task index out of range");
        const TaskMapPtr& m = p->problem->GetTasks()[task];
        out[0] = m->id;
        out[1] = m->start;
        out[2] = m->length;
        out[3] = m->start_jacobian;
        out[4] = m->length_jacobian;
        out[5] = m->is_used ? 1 : 0;
    })
}

exo_host_scene* exo_host_problem_scene(exo_host_problem* p)
{
    return p ? &p->scene_view : nullptr;
}

int exo_host_problem_start_state(exo_host_problem* p, double* x)
{
    HOST_TRY({
        const std::vector<double> s = p->problem->GetStartState();
        memcpy(x, s.data(), s.size() * sizeof(double));
    })
}

int exo_host_problem_update(exo_host_problem* p, const double* x, int n, int t, double* phi, double* jac)
{
    HOST_TRY({
        const std::vector<double> xv(x, x + n);
        const size_t lp = p->problem->length_Phi, lj = (size_t)p->problem->length_jacobian * p->problem->N;
        if (UnconstrainedEndPoseProblem* ep = dynamic_cast<UnconstrainedEndPoseProblem*>(p->problem.get()))
        {
            ep->Update(xv);
            if (phi) memcpy(phi, ep->Phi.data(), lp * sizeof(double));
            if (jac) memcpy(jac, ep->jacobian.data(), lj * sizeof(double));
        }
        else if (TimeIndexedProblem* tp = dynamic_cast<TimeIndexedProblem*>(p->problem.get()))
        {
            if (t < 0) ThrowPretty("TimeIndexedProblem::Update needs a time index");
            tp->Update(xv, t);
            if (phi) memcpy(phi, tp->Phi[t].data(), lp * sizeof(double));
            if (jac) memcpy(jac, tp->jacobian[t].data(), lj * sizeof(double));
        }
        else
            ThrowPretty("unsupported problem type");
    })
}

int exo_host_problem_update_trajectory(exo_host_problem* p, const double* x, int n, double* phi, double* jac)
{
    HOST_TRY({
        TimeIndexedProblem* tp = dynamic_cast<TimeIndexedProblem*>(p->problem.get());
        if (!tp) ThrowPretty("Update(x_trajectory) needs a time-indexed problem");
        tp->Update(std::vector<double>(x, x + n));
        const size_t lp = tp->length_Phi, lj = (size_t)tp->length_jacobian * tp->N;
        for (int t = 0; t < tp->GetT(); ++t)
        {
            if (phi) memcpy(phi + (size_t)t * lp, tp->Phi[t].data(), lp * sizeof(double));
            if (jac) memcpy(jac + (size_t)t * lj, tp->jacobian[t].data(), lj * sizeof(double));
        }
    })
}

int exo_host_problem_update_batch(exo_host_problem* p, const double* x, int64_t n_evals, int layout, double* phi, double* jac)
{
    HOST_TRY({
        if (TimeIndexedProblem* tp = dynamic_cast<TimeIndexedProblem*>(p->problem.get()))
        {
            if (n_evals % tp->GetT() != 0) ThrowPretty("batch is not a whole number of trajectories");
            tp->UpdateTrajectories(x, n_evals / tp->GetT(), layout, phi, jac);
        }
        else
            p->problem->UpdateBatch(x, n_evals, layout, phi, jac);
    })
}

int exo_host_problem_update_batch_device(exo_host_problem* p, const double* d_x, int64_t n_evals, int layout, double* d_phi, double* d_jac, void* stream)
{
    HOST_TRY(p->problem->UpdateBatchDevice(d_x, n_evals, layout, d_phi, d_jac, stream))
}

int exo_host_problem_cost_batch(exo_host_problem* p, const double* x, int64_t n_traj, double* cost)
{
    HOST_TRY(p->problem->CostBatch(x, n_traj, cost))
}

int exo_host_problem_cost_jacobian_batch(exo_host_problem* p, const double* x, int64_t n_traj, double* cost, double* jacobian)
{
    HOST_TRY(p->problem->CostJacobianBatch(x, n_traj, cost, jacobian))
}

int exo_host_problem_cost_batch_device(exo_host_problem* p, const double* d_x, const double* d_phi, int64_t n_traj, int layout, double* d_cost,
                                       double* d_cost_eval, void* stream)
{
    HOST_TRY(p->problem->CostBatchDevice(d_x, d_phi, n_traj, layout, d_cost, d_cost_eval, stream))
}

int exo_host_problem_update_cost_batch_device(exo_host_problem* p, const double* d_x, int64_t n_traj, int layout, double* d_phi, double* d_jac, double* d_cost,
                                              double* d_cost_eval, void* stream)
{
    HOST_TRY(p->problem->UpdateCostBatchDevice(d_x, n_traj, layout, d_phi, d_jac, d_cost, d_cost_eval, stream))
}

int exo_host_problem_ik_solve_batch(exo_host_problem* p, const double* q0, int64_t n, const exo_ik_params* params, double* solutions, double* cost,
                                    int32_t* iterations, int32_t* termination)
{
    HOST_TRY(p->problem->IkSolveBatch(q0, n, *params, solutions, cost, iterations, termination))
}

int exo_host_problem_hessian_batch(exo_host_problem* p, const double* x, int64_t n_evals, double* hess)
{
    HOST_TRY(p->problem->HessianBatch(x, n_evals, hess))
}

int exo_host_problem_last_kernel_ms(exo_host_problem* p, float* ms)
{
    HOST_TRY(*ms = p->problem->LastKernelMs())
}

int64_t exo_host_problem_launch_count(exo_host_problem* p)
{
    return p ? p->problem->LaunchCount() : -1;
}

int64_t exo_host_problem_update_count(exo_host_problem* p)
{
    return p ? (int64_t)p->problem->GetNumberOfProblemUpdates() : -1;
}

exo_problem* exo_host_problem_backend(exo_host_problem* p)
{
    try
    {
        return p ? p->problem->GetBackend() : nullptr;
    }
    catch (const std::exception& e)
    {
        Fail(e);
        return nullptr;
    }
}

}  // extern "C"
