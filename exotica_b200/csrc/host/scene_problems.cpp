// scene_problems.cpp -- host mirror of Scene, the four task maps, the factories and the problem Update paths.
// Reference: exotica_core/src/{scene.cpp:320-348, task_map.cpp:42-60, planning_problem.cpp:148-266, setup.cpp},
// exotica_core/src/problems/{abstract_time_indexed_problem.cpp:209-292, unconstrained_end_pose_problem.cpp:123-167},
// exotations/exotica_core_task_maps/src/{eff_frame,eff_position,joint_limit,sphere_collision}.cpp.
#include <string.h>

#include <algorithm>
#include <fstream>
#include <set>
#include <sstream>

#include "exotica_host.h"

namespace exotica
{
static void CheckRc(int rc)
{
    if (rc != EXO_OK) throw Exception(exo_last_error());
}

// ---- Initializer ------------------------------------------------------------------------------------------
std::string Initializer::Require(const std::string& key) const
{
    auto it = properties.find(key);
    if (it == properties.end()) ThrowPretty("Initializer " + name + " requires property " + key + " to be set!");
    return it->second;
}

const std::vector<Initializer>& Initializer::List(const std::string& key) const
{
    static const std::vector<Initializer> empty;
    auto it = children.find(key);
    return it == children.end() ? empty : it->second;
}

Initializer InitializerFromXml(const xml::Element& e, const std::string& prefix)
{
    Initializer init;
    init.name = prefix + e.name;
    for (const auto& a : e.attributes) init.properties[a.first] = a.second;
    for (const auto& c : e.children)
    {
        if (c->children.empty())
        {
            // leaf element: <T>100</T>  or an attribute-only initializer inside a list (handled by the parent below)
            if (!c->attributes.empty() && xml::Trim(c->text).empty())
                init.children[c->name].push_back(InitializerFromXml(*c, prefix));
            else
                init.properties[c->name] = xml::Trim(c->text);
        }
        else
        {
            // container element: its children are initializers (<Maps><EffFrame .../></Maps>, <EndEffector><Frame/></EndEffector>)
            std::vector<Initializer>& list = init.children[c->name];
            for (const auto& g : c->children) list.push_back(InitializerFromXml(*g, prefix));
        }
    }
    return init;
}

// ---- Scene ----------------------------------------------------------------------------------------------------
Scene::Scene() {}

void Scene::InstantiateFromStrings(const std::string& urdf_xml, const std::string& srdf_xml, const std::string& joint_group)
{
    RobotModelPtr model(new RobotModel(urdf_xml, srdf_xml));
    kinematica_.Instantiate(joint_group, model, name_);
    num_positions_ = kinematica_.GetNumControlledJoints();
}

void Scene::Instantiate(const Initializer& init)
{
    name_ = init.Get("Name", "");
    std::string urdf = init.Get("URDF", "");
    std::string srdf = init.Get("SRDF", "");
    if (urdf.empty()) ThrowPretty("Scene requires a URDF (file path or XML text); there is no ROS parameter server here");
    auto load = [](const std::string& v) {
        const std::string t = xml::Trim(v);
        if (!t.empty() && t[0] == '<') return t;
        return Setup::LoadFile(Setup::ParsePath(t));
    };
    InstantiateFromStrings(load(urdf), srdf.empty() ? std::string() : load(srdf), init.Get("JointGroup", ""));
    // scene.cpp:292-303
    if (init.Has("JointPositionLimitsLower")) kinematica_.SetJointLimitsLower(ParseVector(init.Get("JointPositionLimitsLower", "")));
    if (init.Has("JointPositionLimitsUpper")) kinematica_.SetJointLimitsUpper(ParseVector(init.Get("JointPositionLimitsUpper", "")));
    // scene.cpp:107-112: LoadScene = ';'-separated list of .scene files, identity offset
    if (init.Has("LoadScene"))
    {
        std::string list = init.Get("LoadScene", "");
        size_t pos = 0;
        while (pos <= list.size())
        {
            const size_t end = std::min(list.find(';', pos), list.size());
            const std::string file = xml::Trim(list.substr(pos, end - pos));
            if (!file.empty()) LoadSceneFile(file);
            pos = end + 1;
        }
    }
    // custom links (scene.cpp:114-170; init/link.in: Name, Transform, Parent -- shapes belong to the collision scene, out of scope)
    for (const Initializer& link : init.List("Links"))
    {
        Frame tf = IdentityFrame();
        if (link.Has("Transform")) tf = GetFrame(ParseVector(link.Get("Transform", "")));
        AddObject(link.Require("Name"), tf, link.Get("Parent", ""));
    }
    // trajectory generators (scene.cpp:247-259; init/trajectory.in: Link, File | Trajectory)
    for (const Initializer& tr : init.List("Trajectories"))
    {
        const std::string file = tr.Get("File", "");
        AddTrajectory(tr.Require("Link"), file.empty() ? tr.Get("Trajectory", "") : Setup::LoadFile(Setup::ParsePath(file)));
    }
}

void Scene::UpdateTrajectoryGenerators(double t)
{
    // scene.cpp:329-335: generated_offset = trajectory->GetPosition(t) for every trajectory-driven link.  Links driven by an
    // explicit per-timestep table (AddTrajectoryFrames) have no time base of their own: they show their first frame here; the
    // time-indexed problems index the table by timestep.
    for (const auto& kv : trajectories_)
        if (!kv.second.empty()) kinematica_.SetGeneratedOffset(kv.first, kv.second.front());
    for (const auto& kv : trajectory_generators_) kinematica_.SetGeneratedOffset(kv.second.first, kv.second.second->GetPosition(t));
}

void Scene::Update(const std::vector<double>& x, double t)
{
    UpdateTrajectoryGenerators(t);  // scene.cpp:337-348
    kinematica_.Update(x);
}

void Scene::AddObject(const std::string& name, const Frame& transform, const std::string& parent)
{
    kinematica_.AddEnvironmentElement(name, transform, parent, false);
    custom_links_.push_back({name, parent, transform});
}

// ---- MoveIt ".scene" geometry text (planning_scene::PlanningScene::loadGeometryFromStream, moveit_core -- external to the reference,
// not vendored: restated from the published file format; the reference's fixtures are exotica_examples/resources/scenes/*.scene) -----
//   (SceneName)+
//   * ObjectName
//   [x y z qx qy qz qw]          object pose, only in the newer format
//   n_shapes
//   shape (box: x y z | sphere: r | cylinder: r l | cone: r l | plane: a b c d | mesh: nv nt, nv vertices, nt triangles)
//   x y z
//   qx qy qz qw
//   r g b a
//   ...
//   .
namespace
{
Frame PoseFromQuaternion(double x, double y, double z, double qx, double qy, double qz, double qw)
{
    // Eigen::Translation3d(x, y, z) * Eigen::Quaterniond(qw, qx, qy, qz): Eigen's toRotationMatrix, no normalisation
    const double tx = 2 * qx, ty = 2 * qy, tz = 2 * qz;
    const double twx = tx * qw, twy = ty * qw, twz = tz * qw, txx = tx * qx, txy = ty * qx, txz = tz * qx, tyy = ty * qy, tyz = tz * qy, tzz = tz * qz;
    Frame f;
    double* d = f.data();
    d[0] = 1 - (tyy + tzz); d[1] = txy - twz; d[2] = txz + twy;
    d[3] = txy + twz; d[4] = 1 - (txx + tzz); d[5] = tyz - twx;
    d[6] = txz - twy; d[7] = tyz + twx; d[8] = 1 - (txx + tyy);
    d[9] = x; d[10] = y; d[11] = z;
    return f;
}

Frame Compose(const Frame& a, const Frame& b)
{
    Frame out;
    const double *A = a.data(), *B = b.data();
    double* o = out.data();
    for (int r = 0; r < 3; ++r)
    {
        for (int c = 0; c < 3; ++c) o[3 * r + c] = A[3 * r] * B[c] + A[3 * r + 1] * B[3 + c] + A[3 * r + 2] * B[6 + c];
        o[9 + r] = A[3 * r] * B[9] + A[3 * r + 1] * B[10] + A[3 * r + 2] * B[11] + A[9 + r];
    }
    return out;
}

Frame Inverse(const Frame& a)
{
    Frame out;
    const double* A = a.data();
    double* o = out.data();
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) o[3 * r + c] = A[3 * c + r];
    for (int r = 0; r < 3; ++r) o[9 + r] = -(o[3 * r] * A[9] + o[3 * r + 1] * A[10] + o[3 * r + 2] * A[11]);
    return out;
}

void SkipNumbers(std::istream& in, int n, const std::string& what)
{
    double v;
    for (int i = 0; i < n; ++i)
        if (!(in >> v)) ThrowPretty("Malformed scene geometry: expected " + std::to_string(n) + " numbers for " + what);
}
}  // namespace

void Scene::LoadScene(const std::string& scene_text, const Frame& offset)
{
    std::istringstream in(scene_text);
    std::string line;
    if (!std::getline(in, line)) ThrowPretty("Bad input stream when loading scene geometry");  // the scene name line
    std::map<std::string, WorldObject> loaded;
    while (true)
    {
        std::string marker;
        if (!(in >> marker)) ThrowPretty("Bad input stream when loading marker in scene geometry");
        if (marker == ".") break;
        if (marker != "*") ThrowPretty("Unknown marker in scene geometry file: " + marker);
        std::string name;
        std::getline(in, name);
        name = xml::Trim(name);
        if (name.empty()) ThrowPretty("Bad input stream when loading object name in scene geometry");
        // newer MoveIt files carry the object pose before the shape count: a line of seven numbers instead of one
        std::string next;
        do
        {
            if (!std::getline(in, next)) ThrowPretty("Bad input stream when loading shape count of '" + name + "'");
        } while (xml::Trim(next).empty());
        Frame object_pose = IdentityFrame();
        std::vector<double> nums = ParseVector(next);
        if (nums.size() == 7)
        {
            object_pose = PoseFromQuaternion(nums[0], nums[1], nums[2], nums[3], nums[4], nums[5], nums[6]);
            int count = 0;
            if (!(in >> count)) ThrowPretty("Bad input stream when loading shape count of '" + name + "'");
            nums.assign(1, (double)count);
        }
        if (nums.size() != 1 || nums[0] < 0) ThrowPretty("Bad shape count for object '" + name + "' in scene geometry");
        WorldObject obj;
        for (int i = 0; i < (int)nums[0]; ++i)
        {
            std::string type;
            if (!(in >> type)) ThrowPretty("Bad input stream when loading shape type of '" + name + "'");
            if (type == "box") SkipNumbers(in, 3, type);
            else if (type == "sphere") SkipNumbers(in, 1, type);
            else if (type == "cylinder" || type == "cone") SkipNumbers(in, 2, type);
            else if (type == "plane") SkipNumbers(in, 4, type);
            else if (type == "mesh")
            {
                long nv = 0, nt = 0;
                if (!(in >> nv >> nt) || nv < 0 || nt < 0) ThrowPretty("Malformed mesh in scene geometry");
                SkipNumbers(in, (int)(3 * nv + 3 * nt), type);
            }
            else
                ThrowPretty("Unknown shape type: '" + type + "'");
            double p[7];
            for (double& v : p)
                if (!(in >> v)) ThrowPretty("Failed to read pose of a shape of '" + name + "' from scene geometry");
            SkipNumbers(in, 4, "the shape colour");
            obj.shape_poses.push_back(Compose(offset, Compose(object_pose, PoseFromQuaternion(p[0], p[1], p[2], p[3], p[4], p[5], p[6]))));
        }
        loaded[name] = obj;
    }
    for (auto& kv : loaded) world_objects_[kv.first] = kv.second;
    RebuildEnvironment();
}

void Scene::LoadSceneFile(const std::string& file_name, const Frame& offset)
{
    LoadScene(Setup::LoadFile(Setup::ParsePath(file_name)), offset);
}

void Scene::CleanScene()
{
    world_objects_.clear();
    RebuildEnvironment();
}

std::vector<std::string> Scene::GetWorldObjectNames() const
{
    std::vector<std::string> names;
    for (const auto& kv : world_objects_) names.push_back(kv.first);
    return names;
}

void Scene::RebuildEnvironment()
{
    // UpdateSceneFrames (scene.cpp:702-758): reset, then every world object in name order -- its frame is the first shape's pose,
    // every shape hangs below it as "<name>_collision_<i>" -- and UpdateInternalFrames (scene.cpp:662-700): the custom links, then
    // the trajectory generators again
    kinematica_.ResetModel();
    for (const auto& kv : world_objects_)
    {
        if (kv.second.shape_poses.empty()) continue;  // "Object with no shapes"
        const Frame& origin = kv.second.shape_poses[0];
        kinematica_.AddEnvironmentElement(kv.first, origin, "", false);
        const Frame inv = Inverse(origin);
        for (size_t i = 0; i < kv.second.shape_poses.size(); ++i)
            kinematica_.AddEnvironmentElement(kv.first + "_collision_" + std::to_string(i), Compose(inv, kv.second.shape_poses[i]), kv.first, false);
    }
    for (const CustomLink& l : custom_links_) kinematica_.AddEnvironmentElement(l.name, l.transform, l.parent, false);
    trajectories_.clear();
    for (auto& kv : trajectory_tables_)
    {
        int idx = kinematica_.GetElementIndex(kv.first);
        if (idx < 0) idx = kinematica_.AddEnvironmentElement(kv.first, IdentityFrame(), "", false);
        trajectories_[kinematica_.MakeTrajectoryGenerated(idx)] = kv.second;
    }
    for (auto& kv : trajectory_generators_)
    {
        const int idx = kinematica_.GetElementIndex(kv.first);
        if (idx < 0) ThrowPretty("Can't find link '" + kv.first + "'!");
        kv.second.first = kinematica_.MakeTrajectoryGenerated(idx);
    }
    ++trajectory_version_;
    UpdateTrajectoryGenerators(0.0);
}

void Scene::AddTrajectoryFrames(const std::string& link, const std::vector<Frame>& frames_per_timestep)
{
    if (frames_per_timestep.empty()) ThrowPretty("Trajectory for '" + link + "' is empty");
    int idx = kinematica_.GetElementIndex(link);
    if (idx < 0)
    {
        idx = kinematica_.AddEnvironmentElement(link, IdentityFrame(), "", false);
        custom_links_.push_back({link, "", IdentityFrame()});
    }
    trajectory_generators_.erase(link);
    trajectory_tables_[link] = frames_per_timestep;
    trajectories_[kinematica_.MakeTrajectoryGenerated(idx)] = frames_per_timestep;
    ++trajectory_version_;
    UpdateTrajectoryGenerators(0.0);
}

void Scene::AddTrajectory(const std::string& link, const std::string& traj)
{
    AddTrajectory(link, std::shared_ptr<Trajectory>(new Trajectory(traj)));
}

void Scene::AddTrajectory(const std::string& link, std::shared_ptr<Trajectory> traj)
{
    if (!traj || traj->GetDuration() == 0.0) ThrowPretty("The trajectory is empty!");
    const int idx = kinematica_.GetElementIndex(link);
    if (idx < 0) ThrowPretty("Can't find link '" + link + "'!");
    trajectory_tables_.erase(link);
    trajectory_generators_[link] = std::make_pair(kinematica_.MakeTrajectoryGenerated(idx), traj);
    ++trajectory_version_;
    UpdateTrajectoryGenerators(0.0);
}

std::shared_ptr<Trajectory> Scene::GetTrajectory(const std::string& link) const
{
    const auto it = trajectory_generators_.find(link);
    if (it == trajectory_generators_.end()) ThrowPretty("No trajectory generator defined for link '" + link + "'!");
    return it->second.second;
}

void Scene::SampleTrajectories(int T, double tau)
{
    // Scene::UpdateTrajectoryGenerators(t * tau) for every timestep at once
    for (const auto& kv : trajectory_generators_)
    {
        std::vector<Frame> frames(T);
        for (int t = 0; t < T; ++t) frames[t] = kv.second.second->GetPosition((double)t * tau);
        trajectories_[kv.second.first] = frames;
    }
}

const std::vector<double>& Scene::GetGeneratedTable(int& T, int& n_generated) const
{
    n_generated = kinematica_.GetNumGenerated();
    T = 0;
    generated_table_.clear();
    if (n_generated == 0) return generated_table_;
    for (const auto& kv : trajectories_) T = std::max(T, (int)kv.second.size());
    if ((int)trajectories_.size() != n_generated) ThrowPretty("Every trajectory-driven link needs AddTrajectoryFrames");
    generated_table_.assign((size_t)T * n_generated * 12, 0.0);
    for (const auto& kv : trajectories_)
        for (int t = 0; t < T; ++t)
        {
            const Frame& f = kv.second[std::min<size_t>(t, kv.second.size() - 1)];
            memcpy(&generated_table_[((size_t)t * n_generated + kv.first) * 12], f.data(), 12 * sizeof(double));
        }
    return generated_table_;
}

// ---- TaskMap base ------------------------------------------------------------------------------------------------
void TaskMap::InstantiateBase(const Initializer& init)
{
    object_name_ = init.Get("Name", "");
    is_used = true;
    frames_.clear();
    for (const Initializer& eff : init.List("EndEffector"))
    {
        // init/frame.in: Link (required), LinkOffset, Base, BaseOffset
        const std::string link = eff.Require("Link");
        Frame oa = IdentityFrame(), ob = IdentityFrame();
        if (eff.Has("LinkOffset")) oa = GetFrame(ParseVector(eff.Get("LinkOffset", "")));
        if (eff.Has("BaseOffset")) ob = GetFrame(ParseVector(eff.Get("BaseOffset", "")));
        frames_.push_back(KinematicFrameRequest(link, oa, eff.Get("Base", ""), ob));
    }
}

void TaskMap::ResolveFrames()
{
    if (!scene_) ThrowPretty("Scene is not initialised!");
    const int F = (int)frames_.size();
    elem_a_.resize(F);
    elem_b_.resize(F);
    off_a_.resize(12 * (size_t)F);
    off_b_.resize(12 * (size_t)F);
    for (int i = 0; i < F; ++i)
    {
        elem_a_[i] = scene_->GetKinematicTree().ResolveFrame(frames_[i].frame_A_link_name, "frame_A");
        elem_b_[i] = scene_->GetKinematicTree().ResolveFrame(frames_[i].frame_B_link_name, "frame_B");
        memcpy(&off_a_[12 * i], frames_[i].frame_A_offset.data(), 12 * sizeof(double));
        memcpy(&off_b_[12 * i], frames_[i].frame_B_offset.data(), 12 * sizeof(double));
    }
}

static void AddMapToBackend(exo_problem* p, TaskMap& m, const std::vector<int32_t>& ea, const std::vector<double>& oa, const std::vector<int32_t>& eb,
                            const std::vector<double>& ob)
{
    exo_task_map_desc d;
    memset(&d, 0, sizeof(d));
    d.kind = m.Kind();
    d.n_frames = (int)ea.size();
    d.elem_a = ea.data();
    d.off_a = oa.data();
    d.elem_b = eb.data();
    d.off_b = ob.data();
    d.rotation_type = EXO_ROT_RPY;
    d.precision = 1.0;
    d.is_used = m.is_used ? 1 : 0;
    m.FillDescriptor(d);
    CheckRc(exo_problem_add_task_map(p, &d));
}

void TaskMap::ReleaseOwnProblem()
{
    if (own_problem_) exo_problem_destroy(own_problem_);
    own_problem_ = nullptr;
}

void TaskMap::EnsureOwnProblem()
{
    if (!scene_) ThrowPretty("Scene is not initialised!");
    KinematicTree& kt = scene_->GetKinematicTree();
    // the C-ABI tree is rebuilt when scene objects / trajectories are added: a problem built on the old one is stale
    if (own_problem_ && own_problem_generation_ != kt.GetBackendGeneration()) ReleaseOwnProblem();
    if (own_problem_) return;
    ResolveFrames();
    exo_tree* tree = kt.GetBackend();
    own_problem_generation_ = kt.GetBackendGeneration();
    CheckRc(exo_problem_create(tree, &own_problem_));
    const bool used = is_used;
    is_used = true;
    try
    {
        AddMapToBackend(own_problem_, *this, elem_a_, off_a_, elem_b_, off_b_);
    }
    catch (...)
    {
        is_used = used;
        ReleaseOwnProblem();
        throw;
    }
    is_used = used;
    CheckRc(exo_problem_finalize(own_problem_));
}

void TaskMap::UpdateBatch(const double* x, int64_t n_evals, int layout, double* phi, double* jacobian)
{
    EnsureOwnProblem();
    CheckRc(exo_problem_update_batch_host(own_problem_, x, n_evals, 1, layout, phi, jacobian));
}

void TaskMap::Update(const std::vector<double>& x, std::vector<double>& phi)
{
    const int N = scene_ ? scene_->GetKinematicTree().GetNumControlledJoints() : 0;
    if ((int)x.size() != N) ThrowPretty("Wrong state vector size! Got " + std::to_string(x.size()) + " expected " + std::to_string(N));
    if ((int)phi.size() != TaskSpaceDim()) ThrowPretty("Wrong size of Phi!");
    UpdateBatch(x.data(), 1, EXO_LAYOUT_AOS_COLMAJOR, phi.data(), nullptr);
}

void TaskMap::Update(const std::vector<double>& x, std::vector<double>& phi, std::vector<double>& jacobian)
{
    const int N = scene_ ? scene_->GetKinematicTree().GetNumControlledJoints() : 0;
    if ((int)x.size() != N) ThrowPretty("Wrong state vector size! Got " + std::to_string(x.size()) + " expected " + std::to_string(N));
    if ((int)phi.size() != TaskSpaceDim()) ThrowPretty("Wrong size of Phi!");
    if ((int)jacobian.size() != TaskSpaceJacobianDim() * N) ThrowPretty("Wrong size of jacobian! " + std::to_string(N));
    UpdateBatch(x.data(), 1, EXO_LAYOUT_AOS_COLMAJOR, phi.data(), jacobian.data());
}

// ---- the four maps ----------------------------------------------------------------------------------------------------
void EffFrame::Instantiate(const Initializer& init)
{
    rotation_type_ = GetRotationTypeFromString(init.Get("Type", "RPY"));  // init/eff_frame.in
    small_stride_ = GetRotationTypeLength(rotation_type_);
    big_stride_ = small_stride_ + 3;
}

void EffOrientation::Instantiate(const Initializer& init)
{
    rotation_type_ = GetRotationTypeFromString(init.Get("Type", "RPY"));  // init/eff_orientation.in
    stride_ = GetRotationTypeLength(rotation_type_);
}

void JointPose::Instantiate(const Initializer& init)
{
    // init/joint_pose.in: JointRef, JointMap (both optional)
    init_map_.clear();
    for (double v : ParseVector(init.Get("JointMap", ""))) init_map_.push_back((int32_t)v);
    init_ref_ = ParseVector(init.Get("JointRef", ""));
}

void JointPose::AssignScene(ScenePtr scene)
{
    // JointPose::Initialize (joint_pose.cpp:66-97)
    scene_ = scene;
    const int n = scene_->GetKinematicTree().GetNumControlledJoints();
    joint_map_.clear();
    if (!init_map_.empty())
    {
        for (int32_t v : init_map_)
        {
            if (v >= n || v < 0)
                ThrowPretty("JointMap index (" + std::to_string(v) + ") is greater than number of controlled joints (" + std::to_string(n) + ")!");
            joint_map_.push_back(v);
        }
    }
    else
        for (int i = 0; i < n; ++i) joint_map_.push_back(i);
    if (!init_ref_.empty())
    {
        if (init_ref_.size() != joint_map_.size())
            ThrowPretty("Invalid joint reference size! Expecting " + std::to_string(joint_map_.size()) + " but received " + std::to_string(init_ref_.size()));
        joint_ref_ = init_ref_;
    }
    else
        joint_ref_.assign(joint_map_.size(), 0.0);
}

void JointLimit::Instantiate(const Initializer& init)
{
    const std::vector<double> v = ParseVector(init.Get("SafePercentage", "0.0"));  // init/joint_limit.in
    if (v.size() != 1) ThrowPretty("SafePercentage must be a single number");
    safe_percentage_ = v[0];
}

void JointLimit::AssignScene(ScenePtr scene)
{
    scene_ = scene;
    N_ = scene_->GetKinematicTree().GetNumControlledJoints();  // joint_limit.cpp:42-56 (kinematic case)
}

// ---- EffAxisAlignment (eff_axis_alignment.cpp:39-58; init/frame_with_axis_and_direction.in) ---------------------------------------
static void Normalise3(std::vector<double>& v, const char* what)
{
    if (v.size() != 3) ThrowPretty(std::string(what) + " must have 3 entries");
    const double n = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    for (double& x : v) x /= n;  // Eigen normalized()
}

void EffAxisAlignment::InstantiateBase(const Initializer& init)
{
    TaskMap::InstantiateBase(init);
    axis_.clear();
    dir_.clear();
    for (const Initializer& eff : init.List("EndEffector"))
    {
        std::vector<double> a = eff.Has("Axis") ? ParseVector(eff.Get("Axis", "")) : std::vector<double>{1.0, 0.0, 0.0};
        std::vector<double> d = eff.Has("Direction") ? ParseVector(eff.Get("Direction", "")) : std::vector<double>{1.0, 0.0, 0.0};
        Normalise3(a, "Axis");
        Normalise3(d, "Direction");
        axis_.insert(axis_.end(), a.begin(), a.end());
        dir_.insert(dir_.end(), d.begin(), d.end());
    }
}

std::vector<double> EffAxisAlignment::GetDirection(const std::string& frame_name) const
{
    for (size_t i = 0; i < frames_.size(); ++i)
        if (frames_[i].frame_A_link_name == frame_name) return std::vector<double>(dir_.begin() + 3 * i, dir_.begin() + 3 * i + 3);
    ThrowPretty("Direction for frame with name " + frame_name + " could not be found.");
}

void EffAxisAlignment::SetDirection(const std::string& frame_name, const std::vector<double>& dir_in)
{
    for (size_t i = 0; i < frames_.size(); ++i)
        if (frames_[i].frame_A_link_name == frame_name)
        {
            std::vector<double> d = dir_in;
            Normalise3(d, "Direction");
            std::copy(d.begin(), d.end(), dir_.begin() + 3 * i);
            ReleaseOwnProblem();
            return;
        }
    ThrowPretty("Could not find frame with name " + frame_name + ".");
}

// ---- CenterOfMass (center_of_mass.cpp:145-170, 88-117; init/center_of_mass.in) -----------------------------------------------------
void CenterOfMass::Instantiate(const Initializer& init)
{
    const std::string z = init.Get("EnableZ", "1");
    enable_z_ = !(z == "0" || z == "false" || z == "False");
    dim_ = enable_z_ ? 3 : 2;
}

void CenterOfMass::AssignScene(ScenePtr scene)
{
    scene_ = scene;
    const KinematicTree& kt = scene_->GetKinematicTree();
    mass_.clear();
    if (!frames_.empty())
    {
        // explicit frames: the offset's position becomes the link's centre of gravity, the weight its mass (:160-167)
        for (KinematicFrameRequest& f : frames_)
        {
            if (!f.frame_B_link_name.empty()) ThrowPretty("Requesting CenterOfMass frame with base other than root! '" + f.frame_A_link_name + "'");
            const KinematicElement& e = kt.GetTree()[kt.ResolveFrame(f.frame_A_link_name, "frame_A")];
            for (int k = 0; k < 3; ++k) f.frame_A_offset[9 + k] = e.cog[k];
            mass_.push_back(e.mass);
        }
    }
    else
    {
        // no frames: every robot link (and attached object) with mass, at its centre of gravity (:88-117)
        for (const KinematicElement& e : kt.GetTree())
            if (e.is_robot_link && e.mass > 0)
            {
                Frame off = IdentityFrame();
                for (int k = 0; k < 3; ++k) off[9 + k] = e.cog[k];
                frames_.push_back(KinematicFrameRequest(e.link_name, off));
                mass_.push_back(e.mass);
            }
    }
}

void SphereCollision::InstantiateBase(const Initializer& init)
{
    TaskMap::InstantiateBase(init);
    radiuses_.clear();
    group_names_.clear();
    for (const Initializer& eff : init.List("EndEffector"))
    {
        // init/sphere.in: Radius (required), Group = "default"
        const std::vector<double> r = ParseVector(eff.Require("Radius"));
        if (r.size() != 1) ThrowPretty("Sphere Radius must be a single number");
        radiuses_.push_back(r[0]);
        group_names_.push_back(eff.Get("Group", "default"));
    }
}

void SphereCollision::Instantiate(const Initializer& init)
{
    const std::vector<double> p = ParseVector(init.Require("Precision"));  // init/sphere_collision.in
    if (p.size() != 1) ThrowPretty("Precision must be a single number");
    precision_ = p[0];
    // groups_ is a std::map<string, vector<int>>: lexicographic group order (sphere_collision.cpp:42-47)
    std::set<std::string> names(group_names_.begin(), group_names_.end());
    std::vector<std::string> sorted(names.begin(), names.end());
    group_rank_.clear();
    for (const std::string& g : group_names_) group_rank_.push_back((int32_t)(std::find(sorted.begin(), sorted.end(), g) - sorted.begin()));
    dim_ = (int)(sorted.size() * (sorted.size() - 1) / 2);
}

void SphereCollision::FillDescriptor(exo_task_map_desc& desc)
{
    desc.precision = precision_;
    desc.radius = radiuses_.data();
    desc.group = group_rank_.data();
}

// ---- factories ----------------------------------------------------------------------------------------------------------------
template <class T>
static TaskMapPtr MakeMap()
{
    return TaskMapPtr(new T());
}
template <class T>
static PlanningProblemPtr MakeProblem()
{
    return PlanningProblemPtr(new T());
}

Setup& Setup::Instance()
{
    static Setup* s = nullptr;
    if (!s)
    {
        s = new Setup();
        s->RegisterMap("exotica/EffFrame", &MakeMap<EffFrame>);
        s->RegisterMap("exotica/EffPosition", &MakeMap<EffPosition>);
        s->RegisterMap("exotica/JointLimit", &MakeMap<JointLimit>);
        s->RegisterMap("exotica/SphereCollision", &MakeMap<SphereCollision>);
        s->RegisterMap("exotica/EffOrientation", &MakeMap<EffOrientation>);
        s->RegisterMap("exotica/EffPositionXY", &MakeMap<EffPositionXY>);
        s->RegisterMap("exotica/Distance", &MakeMap<Distance>);
        s->RegisterMap("exotica/EffAxisAlignment", &MakeMap<EffAxisAlignment>);
        s->RegisterMap("exotica/CenterOfMass", &MakeMap<CenterOfMass>);
        s->RegisterMap("exotica/JointPose", &MakeMap<JointPose>);
        s->RegisterProblem("exotica/UnconstrainedEndPoseProblem", &MakeProblem<UnconstrainedEndPoseProblem>);
        s->RegisterProblem("exotica/TimeIndexedProblem", &MakeProblem<TimeIndexedProblem>);
        s->RegisterProblem("exotica/UnconstrainedTimeIndexedProblem", &MakeProblem<TimeIndexedProblem>);
    }
    return *s;
}

void Setup::RegisterMap(const std::string& type, MapCreator creator)
{
    if (maps_.count(type)) ThrowPretty("Trying to register already existing type '" + type + "'!");  // factory.h:76-83
    maps_[type] = creator;
}

void Setup::RegisterProblem(const std::string& type, ProblemCreator creator)
{
    if (problems_.count(type)) ThrowPretty("Trying to register already existing type '" + type + "'!");
    problems_[type] = creator;
}

std::vector<std::string> Setup::GetMaps()
{
    std::vector<std::string> out;
    for (const auto& kv : Instance().maps_) out.push_back(kv.first);
    return out;
}

std::vector<std::string> Setup::GetProblems()
{
    std::vector<std::string> out;
    for (const auto& kv : Instance().problems_) out.push_back(kv.first);
    return out;
}

TaskMapPtr Setup::CreateMap(const Initializer& init)
{
    auto it = Instance().maps_.find(init.name);
    if (it == Instance().maps_.end()) ThrowPretty("Unable to load task map of type '" + init.name + "' (registered on the GPU path: EffFrame, EffPosition, JointLimit, SphereCollision, EffOrientation, EffPositionXY, Distance, JointPose, EffAxisAlignment, CenterOfMass)");
    TaskMapPtr m = it->second();
    m->InstantiateBase(init);  // InstantiateInternal = InstantiateBase + Instantiate (property.h:113-119)
    m->Instantiate(init);
    return m;
}

PlanningProblemPtr Setup::CreateProblem(const std::string& type)
{
    auto it = Instance().problems_.find(type);
    if (it == Instance().problems_.end())
    {
        std::string all;
        for (const auto& kv : Instance().problems_) all += " '" + kv.first + "'";
        ThrowPretty("This factory does not recognize type '" + type + "' (registered:" + all + ")");  // factory.h:100-106
    }
    return it->second();
}

PlanningProblemPtr Setup::CreateProblem(const Initializer& init)
{
    PlanningProblemPtr p = CreateProblem(init.name);
    p->InstantiateBase(init);
    p->Instantiate(init);
    return p;
}

Initializer Setup::LoadProblemInitializer(const std::string& xml_text)
{
    std::shared_ptr<xml::Element> doc;
    try
    {
        doc = xml::Parse(xml_text);
    }
    catch (const std::exception& e)
    {
        ThrowPretty(std::string("Can't load XML: ") + e.what());
    }
    auto is_problem = [](const std::string& n) { return Instance().problems_.count("exotica/" + n) > 0; };
    if (is_problem(doc->name)) return InitializerFromXml(*doc);
    for (const auto& c : doc->children)
        if (is_problem(c->name)) return InitializerFromXml(*c);
    ThrowPretty("No problem definition found in the XML (root <" + doc->name + ">)");
}

PlanningProblemPtr Setup::LoadProblem(const std::string& xml_text)
{
    return CreateProblem(LoadProblemInitializer(xml_text));
}

void Setup::SetPackagePath(const std::string& package, const std::string& directory)
{
    Instance().packages_[package] = directory;
}

std::string Setup::ParsePath(const std::string& path)
{
    // tools.cpp:150-190: "{package}/rest" and "package://package/rest"
    std::string ret = path;
    const size_t open = ret.find('{');
    const size_t close = ret.find('}');
    if (open != std::string::npos && close != std::string::npos && close > open)
    {
        const std::string pkg = ret.substr(open + 1, close - open - 1);
        auto it = Instance().packages_.find(pkg);
        if (it == Instance().packages_.end()) ThrowPretty("Unknown package '" + pkg + "' (use Setup::SetPackagePath)");
        ret = ret.substr(0, open) + it->second + ret.substr(close + 1);
    }
    else if (ret.compare(0, 10, "package://") == 0)
    {
        const size_t slash = ret.find('/', 10);
        const std::string pkg = ret.substr(10, slash == std::string::npos ? std::string::npos : slash - 10);
        auto it = Instance().packages_.find(pkg);
        if (it == Instance().packages_.end()) ThrowPretty("Unknown package '" + pkg + "' (use Setup::SetPackagePath)");
        ret = it->second + (slash == std::string::npos ? std::string() : ret.substr(slash));
    }
    return ret;
}

std::string Setup::LoadFile(const std::string& path)
{
    std::ifstream f(path.c_str(), std::ios::in | std::ios::binary);
    if (!f) ThrowPretty("File does not exist: '" + path + "'");  // tools.cpp:98-112
    std::stringstream ss;
    ss << f.rdbuf();
    return ss.str();
}

// ---- PlanningProblem --------------------------------------------------------------------------------------------------------------
PlanningProblem::~PlanningProblem()
{
    InvalidateBackend();
}

void PlanningProblem::InvalidateBackend()
{
    if (backend_) exo_problem_destroy(backend_);
    backend_ = nullptr;
    cost_pushed_ = false;
}

void PlanningProblem::InstantiateBase(const Initializer& init)
{
    object_name_ = init.Get("Name", "");
    task_maps_.clear();
    tasks_.clear();
    InvalidateBackend();

    // Create the scene
    const std::vector<Initializer>& ps = init.List("PlanningScene");
    if (ps.empty()) ThrowPretty("Initializer " + init.name + " requires property PlanningScene to be set!");
    scene_.reset(new Scene());
    scene_->Instantiate(ps[0]);
    N = scene_->GetKinematicTree().GetNumControlledJoints();

    start_state_.assign(N, 0.0);
    if (init.Has("StartState")) SetStartState(ParseVector(init.Get("StartState", "")));
    const std::vector<double> st = ParseVector(init.Get("StartTime", "0"));
    if (st.size() != 1 || st[0] < 0) ThrowPretty("Invalid start time " + init.Get("StartTime", "0"));
    t_start = st[0];

    const int order = init.Has("DerivativeOrder") ? (int)ParseVector(init.Get("DerivativeOrder", "-1")).at(0) : -1;
    switch (order)
    {
        case -1:
            break;
        case 0:
            flags_ = KIN_FK;
            break;
        case 1:
            flags_ = KIN_FK | KIN_J;
            break;
        case 2:
            ThrowPretty("DerivativeOrder 2 (kinematic Hessian) is outside the GPU hot path");
        default:
            ThrowPretty("Unsupported DerivativeOrder: " + std::to_string(order));
    }

    // Create the maps (planning_problem.cpp:204-223)
    for (const Initializer& map_init : init.List("Maps"))
    {
        TaskMapPtr new_map = Setup::CreateMap(map_init);
        new_map->AssignScene(scene_);
        new_map->ns_ = object_name_ + "/" + new_map->GetObjectName();
        if (task_maps_.find(new_map->GetObjectName()) != task_maps_.end()) ThrowPretty("Map '" + new_map->GetObjectName() + "' already exists!");
        task_maps_[new_map->GetObjectName()] = new_map;
        tasks_.push_back(new_map);
    }
    // planning_problem.cpp:227-238
    int id = 0, idJ = 0;
    for (int i = 0; i < (int)tasks_.size(); ++i)
    {
        tasks_[i]->ResolveFrames();  // RequestKinematics: unknown links throw here (kinematic_tree.cpp:647,658)
        tasks_[i]->id = i;
        tasks_[i]->start = id;
        tasks_[i]->length = tasks_[i]->TaskSpaceDim();
        tasks_[i]->start_jacobian = idJ;
        tasks_[i]->length_jacobian = tasks_[i]->TaskSpaceJacobianDim();
        id += tasks_[i]->length;
        idJ += tasks_[i]->length_jacobian;
    }
    length_Phi = id;
    length_jacobian = idJ;
}

void PlanningProblem::SetStartState(const std::vector<double>& x)
{
    // planning_problem.cpp:93-136: N values (controlled) or the full model state
    if ((int)x.size() == N)
        start_state_ = x;
    else if ((int)x.size() == scene_->GetKinematicTree().GetNumModelJoints())
    {
        scene_->SetModelState(x);
        start_state_ = scene_->GetControlledState();
    }
    else
        ThrowPretty("Wrong start state vector size, expected " + std::to_string(N) + ", got " + std::to_string(x.size()));
}

void PlanningProblem::ParseCost(const Initializer& init, int T)
{
    // Task::Initialize (tasks.cpp:36-66) + Rho handling (tasks.cpp:68-118,166-244): a map is used iff some rho != 0
    cost_tasks_.clear();
    cost_rho_.clear();
    cost_goal_.clear();
    for (TaskMapPtr& m : tasks_) m->is_used = false;
    const char* blocks[3] = {"Cost", "Inequality", "Equality"};
    for (const char* block : blocks)
        for (const Initializer& task : init.List(block))
        {
            const std::string name = task.Require("Task");
            if (task_maps_.find(name) == task_maps_.end()) ThrowPretty("Task map '" + name + "' has not been defined!");
            std::vector<double> rho = {1.0};
            if (task.Has("Rho"))
            {
                rho = ParseVector(task.Get("Rho", ""));
                if (rho.size() != 1 && (int)rho.size() != T) ThrowPretty("Invalid task rho size! Expecting 1 or " + std::to_string(T) + " got " + std::to_string(rho.size()));
            }
            std::vector<double> goal;
            if (task.Has("Goal"))
            {
                goal = ParseVector(task.Get("Goal", ""));
                const int len = task_maps_[name]->TaskSpaceDim();
                if (!goal.empty() && (int)goal.size() != len)
                    ThrowPretty("Invalid task goal size! Expecting " + std::to_string(len) + " got " + std::to_string(goal.size()));  // tasks.cpp:96
            }
            for (double r : rho)
                if (r != 0.0) task_maps_[name]->is_used = true;
            if (std::string(block) != "Cost") continue;  // only Cost tasks enter the scalar cost; all blocks drive is_used
            cost_tasks_.push_back(name);
            cost_rho_.push_back(rho);
            cost_goal_.push_back(goal);
        }
}

void PlanningProblem::PushCost()
{
    // EndPoseTask / TimeIndexedTask::Initialize (tasks.cpp:68-118,166-244): rho and goal of every Cost task, per map
    exo_problem* p = GetBackend();
    if (cost_pushed_) return;
    const int T = GetT(), n_maps = (int)tasks_.size();
    std::vector<double> rho((size_t)T * n_maps, 0.0), goal(length_Phi, 0.0);
    for (int i = 0; i < n_maps; ++i) tasks_[i]->FillZeroGoal(goal.data() + tasks_[i]->start);  // TaskSpaceVector::SetZero
    for (size_t c = 0; c < cost_tasks_.size(); ++c)
    {
        const TaskMapPtr& m = task_maps_[cost_tasks_[c]];
        for (int t = 0; t < T; ++t) rho[(size_t)t * n_maps + m->id] = cost_rho_[c].size() == 1 ? cost_rho_[c][0] : cost_rho_[c][t];
        if (!cost_goal_[c].empty()) std::copy(cost_goal_[c].begin(), cost_goal_[c].end(), goal.begin() + m->start);
    }
    exo_cost_desc d;
    d.T = T;
    d.T_goal = 1;
    d.rho = rho.data();
    d.goal = goal.data();
    d.ct = cost_ct_;
    d.W = cost_W_.empty() ? nullptr : cost_W_.data();
    CheckRc(exo_problem_set_cost(p, &d));
    cost_pushed_ = true;
}

void PlanningProblem::CostBatch(const double* x, int64_t n_trajectories, double* cost)
{
    PushCost();
    PrepareBackend();
    CheckRc(exo_problem_cost_batch_host(GetBackend(), x, n_trajectories, EXO_LAYOUT_AOS_ROWMAJOR, cost));
    number_of_problem_updates_ += (unsigned int)(n_trajectories * GetT());
}

void PlanningProblem::CostJacobianBatch(const double* x, int64_t n_trajectories, double* cost, double* jacobian)
{
    PushCost();
    PrepareBackend();
    CheckRc(exo_problem_cost_jacobian_batch_host(GetBackend(), x, n_trajectories, EXO_LAYOUT_AOS_ROWMAJOR, cost, jacobian));
    number_of_problem_updates_ += (unsigned int)(n_trajectories * GetT());
}

namespace
{
struct DeviceMem  // RAII over the C-ABI allocator
{
    void* p = nullptr;
    explicit DeviceMem(size_t bytes) { CheckRc(exo_device_alloc(&p, bytes)); }
    ~DeviceMem() { exo_device_free(p); }
    DeviceMem(const DeviceMem&) = delete;
    DeviceMem& operator=(const DeviceMem&) = delete;
};
// q[n][N] (evaluation-major) <-> q[N][n] (what the batched solver works on), in blocks of 64 evaluations so that both sides stay
// in cache (the plain double loop scatters every element to its own cache line; 65 536 x 38: 35 -> 32.7 ms for the whole call)
void TransposeToSoA(const double* aos, int64_t n, int N, double* soa)
{
    for (int64_t e0 = 0; e0 < n; e0 += 64)
    {
        const int64_t e1 = std::min<int64_t>(n, e0 + 64);
        for (int c = 0; c < N; ++c)
            for (int64_t e = e0; e < e1; ++e) soa[(size_t)c * n + e] = aos[(size_t)e * N + c];
    }
}
void TransposeFromSoA(const double* soa, int64_t n, int N, double* aos)
{
    for (int64_t e0 = 0; e0 < n; e0 += 64)
    {
        const int64_t e1 = std::min<int64_t>(n, e0 + 64);
        for (int c = 0; c < N; ++c)
            for (int64_t e = e0; e < e1; ++e) aos[(size_t)e * N + c] = soa[(size_t)c * n + e];
    }
}
}  // namespace

void PlanningProblem::IkSolveBatch(const double* q0, int64_t n, const exo_ik_params& params, double* solutions, double* cost, int32_t* iterations,
                                   int32_t* termination)
{
    if (GetT() != 1) ThrowPretty("This IKSolver can't solve problem of type 'exotica::TimeIndexedProblem'!");  // ik_solver.cpp:38-41
    if (n <= 0) return;
    PushCost();
    PrepareBackend();
    // SoA on the device: transpose on the way in and out
    std::vector<double> soa((size_t)n * N);
    TransposeToSoA(q0, n, N, soa.data());
    DeviceMem dq((size_t)n * N * sizeof(double)), dc((size_t)n * sizeof(double)), di((size_t)n * sizeof(int32_t)), dt((size_t)n * sizeof(int32_t));
    CheckRc(exo_memcpy_h2d(dq.p, soa.data(), soa.size() * sizeof(double), nullptr));
    CheckRc(exo_problem_ik_solve_batch(GetBackend(), (double*)dq.p, n, &params, (double*)dc.p, (int32_t*)di.p, (int32_t*)dt.p, nullptr));
    CheckRc(exo_memcpy_d2h(soa.data(), dq.p, soa.size() * sizeof(double), nullptr));
    CheckRc(exo_memcpy_d2h(cost, dc.p, (size_t)n * sizeof(double), nullptr));
    CheckRc(exo_memcpy_d2h(iterations, di.p, (size_t)n * sizeof(int32_t), nullptr));
    CheckRc(exo_memcpy_d2h(termination, dt.p, (size_t)n * sizeof(int32_t), nullptr));
    CheckRc(exo_stream_synchronize(nullptr));
    TransposeFromSoA(soa.data(), n, N, solutions);
}

void PlanningProblem::HessianBatch(const double* x, int64_t n_evals, double* hess)
{
    if (n_evals <= 0) return;
    const int T = GetT();
    if (T > 1 && n_evals % T != 0) ThrowPretty("Batch of " + std::to_string(n_evals) + " evaluations is not a whole number of trajectories (T=" + std::to_string(T) + ")");
    exo_problem* p = GetBackend();
    PrepareBackend();
    const size_t n = (size_t)n_evals;
    DeviceMem dx(n * N * sizeof(double)), dphi(n * std::max(length_Phi, 1) * sizeof(double)), djac(n * std::max(length_jacobian, 1) * N * sizeof(double)),
        dh(n * std::max(length_jacobian, 1) * N * N * sizeof(double));
    CheckRc(exo_memcpy_h2d(dx.p, x, n * N * sizeof(double), nullptr));
    CheckRc(exo_problem_update_batch(p, (const double*)dx.p, n_evals, T, EXO_LAYOUT_AOS_ROWMAJOR, (double*)dphi.p, (double*)djac.p, nullptr));
    CheckRc(exo_problem_hessian_batch(p, (const double*)dx.p, (const double*)djac.p, n_evals, T, EXO_LAYOUT_AOS_ROWMAJOR, (double*)dh.p, nullptr));
    CheckRc(exo_memcpy_d2h(hess, dh.p, n * length_jacobian * N * N * sizeof(double), nullptr));
    CheckRc(exo_stream_synchronize(nullptr));
}

void PlanningProblem::CostBatchDevice(const double* d_x, const double* d_phi, int64_t n_trajectories, int layout, double* d_cost, double* d_cost_eval,
                                      void* stream)
{
    PushCost();
    PrepareBackend();
    CheckRc(exo_problem_cost_batch(GetBackend(), d_x, d_phi, n_trajectories, layout, d_cost, d_cost_eval, stream));
}

void PlanningProblem::UpdateCostBatchDevice(const double* d_x, int64_t n_trajectories, int layout, double* d_phi, double* d_jacobian, double* d_cost,
                                            double* d_cost_eval, void* stream)
{
    PushCost();
    PrepareBackend();
    CheckRc(exo_problem_update_cost_batch(GetBackend(), d_x, n_trajectories, layout, d_phi, (flags_ & KIN_J) ? d_jacobian : nullptr, d_cost, d_cost_eval, stream));
    number_of_problem_updates_ += (unsigned int)(n_trajectories * GetT());
}

exo_problem* PlanningProblem::GetBackend()
{
    KinematicTree& kt = scene_->GetKinematicTree();
    // Scene::AddObject / AddTrajectory after a first Update destroy the C-ABI tree (KinematicTree::InvalidateBackend): the problem
    // handle built on it is stale -- rebuild it (and push the cost again) instead of dereferencing a freed tree
    if (backend_ && backend_generation_ != kt.GetBackendGeneration()) InvalidateBackend();
    if (backend_) return backend_;
    exo_tree* tree = kt.GetBackend();
    backend_generation_ = kt.GetBackendGeneration();
    CheckRc(exo_problem_create(tree, &backend_));
    ++backend_serial_;
    try
    {
        for (TaskMapPtr& m : tasks_)
        {
            m->ResolveFrames();
            AddMapToBackend(backend_, *m, m->elem_a_, m->off_a_, m->elem_b_, m->off_b_);
        }
        CheckRc(exo_problem_finalize(backend_));
        if (exo_problem_length_phi(backend_) != length_Phi || exo_problem_length_jacobian(backend_) != length_jacobian)
            ThrowPretty("backend indexing disagrees with the host task maps");
    }
    catch (...)
    {
        InvalidateBackend();
        throw;
    }
    return backend_;
}

void PlanningProblem::UpdateBatch(const double* x, int64_t n_evals, int layout, double* phi, double* jacobian)
{
    exo_problem* p = GetBackend();
    PrepareBackend();
    const int T = GetT();
    if (T > 1 && n_evals % T != 0) ThrowPretty("Batch of " + std::to_string(n_evals) + " evaluations is not a whole number of trajectories (T=" + std::to_string(T) + ")");
    CheckRc(exo_problem_update_batch_host(p, x, n_evals, T, layout, phi, (flags_ & KIN_J) ? jacobian : nullptr));
    number_of_problem_updates_ += (unsigned int)n_evals;
}

void PlanningProblem::UpdateBatchDevice(const double* d_x, int64_t n_evals, int layout, double* d_phi, double* d_jacobian, void* stream)
{
    exo_problem* p = GetBackend();
    PrepareBackend();
    const int T = GetT();
    if (T > 1 && n_evals % T != 0) ThrowPretty("Batch of " + std::to_string(n_evals) + " evaluations is not a whole number of trajectories (T=" + std::to_string(T) + ")");
    CheckRc(exo_problem_update_batch(p, d_x, n_evals, T, layout, d_phi, (flags_ & KIN_J) ? d_jacobian : nullptr, stream));
    number_of_problem_updates_ += (unsigned int)n_evals;
}

float PlanningProblem::LastKernelMs()
{
    float ms = 0;
    CheckRc(exo_problem_last_kernel_ms(GetBackend(), &ms));
    return ms;
}

int64_t PlanningProblem::LaunchCount()
{
    return backend_ ? exo_problem_launch_count(backend_) : 0;
}

// ---- UnconstrainedEndPoseProblem ------------------------------------------------------------------------------------------------------
void UnconstrainedEndPoseProblem::Instantiate(const Initializer& init)
{
    ParseCost(init, 1);
    if (init.Has("W"))
    {
        // unconstrained_end_pose_problem.cpp:56-66: W = I, or diag(W); IKSolver uses W^-1 (ik_solver.cpp:45)
        const std::vector<double> w = ParseVector(init.Get("W", ""));
        if ((int)w.size() == N)
            cost_W_ = w;
        else if (!w.empty())
            ThrowPretty("W dimension mismatch! Expected " + std::to_string(N) + ", got " + std::to_string(w.size()));
    }
    Phi.assign(length_Phi, 0.0);
    jacobian.assign((size_t)length_jacobian * N, 0.0);
}

void UnconstrainedEndPoseProblem::Update(const std::vector<double>& x)
{
    if ((int)x.size() != N) ThrowPretty("Wrong state vector size! Got " + std::to_string(x.size()) + " expected " + std::to_string(N));
    UpdateBatch(x.data(), 1, EXO_LAYOUT_AOS_COLMAJOR, Phi.data(), jacobian.data());
}

// ---- TimeIndexedProblem ------------------------------------------------------------------------------------------------------------------
void TimeIndexedProblem::Instantiate(const Initializer& init)
{
    const std::vector<double> t = ParseVector(init.Require("T"));
    if (t.size() != 1) ThrowPretty("T must be a single number");
    T_ = (int)t[0];
    if (T_ <= 2) ThrowPretty("Invalid number of timesteps: " + std::to_string(T_));  // abstract_time_indexed_problem.cpp:118-125
    const std::vector<double> tau = ParseVector(init.Require("tau"));
    if (tau.size() != 1 || tau[0] <= 0.) ThrowPretty("tau_ is expected to be greater than 0. (tau_=" + init.Get("tau", "") + ")");
    tau_ = tau[0];
    ParseCost(init, T_);
    // ct = 1 / tau / T (abstract_time_indexed_problem.cpp:105); W = I * Wrate, or diag(W) * Wrate (time_indexed_problem.cpp:43-50)
    cost_ct_ = 1.0 / tau_ / T_;
    {
        double wrate = 1.0;
        if (init.Has("Wrate"))
        {
            const std::vector<double> wr = ParseVector(init.Get("Wrate", ""));
            if (wr.size() == 1) wrate = wr[0];
        }
        cost_W_.assign(N, wrate);
        if (init.Has("W"))
        {
            const std::vector<double> w = ParseVector(init.Get("W", ""));
            if ((int)w.size() == N)
                for (int i = 0; i < N; ++i) cost_W_[i] = w[i] * wrate;
            else if (!w.empty())
                ThrowPretty("W dimension mismatch! Expected " + std::to_string(N) + ", got " + std::to_string(w.size()));
        }
    }
    // time_indexed_problem.cpp:57-72
    if (init.Has("LowerBound"))
    {
        const std::vector<double> lb = ParseVector(init.Get("LowerBound", ""));
        if ((int)lb.size() == N) scene_->GetKinematicTree().SetJointLimitsLower(lb);
        else if (!lb.empty()) ThrowPretty("Lower bound size incorrect! Expected " + std::to_string(N) + " got " + std::to_string(lb.size()));
    }
    if (init.Has("UpperBound"))
    {
        const std::vector<double> ub = ParseVector(init.Get("UpperBound", ""));
        if ((int)ub.size() == N) scene_->GetKinematicTree().SetJointLimitsUpper(ub);
        else if (!ub.empty()) ThrowPretty("Upper bound size incorrect! Expected " + std::to_string(N) + " got " + std::to_string(ub.size()));
    }
    x.assign(T_, std::vector<double>(N, 0.0));
    Phi.assign(T_, std::vector<double>(length_Phi, 0.0));
    jacobian.assign(T_, std::vector<double>((size_t)length_jacobian * N, 0.0));
}

void TimeIndexedProblem::PushGenerated()
{
    int T = 0, n_gen = 0;
    exo_problem* p = GetBackend();
    // nothing changed since the last push to this backend (Scene::AddTrajectory* bump the version; a Trajectory object edited in
    // place after get_trajectory() is the caller's to re-add, as Scene::AddTrajectory is the reference's only way to install one)
    if (pushed_serial_ == backend_serial_ && pushed_version_ == scene_->GetTrajectoryVersion()) return;
    scene_->SampleTrajectories(T_, tau_);
    const std::vector<double>& table = scene_->GetGeneratedTable(T, n_gen);
    if (n_gen > 0)
    {
        if (T != T_) ThrowPretty("Trajectory tables have " + std::to_string(T) + " timesteps, the problem has T=" + std::to_string(T_));
        CheckRc(exo_problem_set_generated(p, T_, n_gen, table.data()));
    }
    pushed_serial_ = backend_serial_;
    pushed_version_ = scene_->GetTrajectoryVersion();
}

void TimeIndexedProblem::Update(const std::vector<double>& x_in, int t)
{
    if (t >= T_ || t < -1) ThrowPretty("Requested t=" + std::to_string(t) + " out of range, needs to be 0 =< t < " + std::to_string(T_));
    if (t == -1) t = T_ - 1;
    if ((int)x_in.size() != N) ThrowPretty("Wrong state vector size! Got " + std::to_string(x_in.size()) + " expected " + std::to_string(N));
    x[t] = x_in;
    exo_problem* p = GetBackend();
    PushGenerated();
    // one evaluation at timestep t: a length-T batch would waste work, so shift the table lookup by evaluating a
    // padded mini-batch only when trajectory-driven links exist
    if (scene_->GetKinematicTree().GetNumGenerated() == 0)
    {
        CheckRc(exo_problem_update_batch_host(p, x_in.data(), 1, T_, EXO_LAYOUT_AOS_COLMAJOR, Phi[t].data(), (flags_ & KIN_J) ? jacobian[t].data() : nullptr));
    }
    else
    {
        std::vector<double> X((size_t)(t + 1) * N, 0.0), P((size_t)(t + 1) * length_Phi), J((size_t)(t + 1) * length_jacobian * N);
        for (int k = 0; k <= t; ++k) memcpy(&X[(size_t)k * N], x_in.data(), N * sizeof(double));
        CheckRc(exo_problem_update_batch_host(p, X.data(), t + 1, T_, EXO_LAYOUT_AOS_COLMAJOR, P.data(), (flags_ & KIN_J) ? J.data() : nullptr));
        memcpy(Phi[t].data(), &P[(size_t)t * length_Phi], length_Phi * sizeof(double));
        if (flags_ & KIN_J) memcpy(jacobian[t].data(), &J[(size_t)t * length_jacobian * N], (size_t)length_jacobian * N * sizeof(double));
    }
    ++number_of_problem_updates_;
}

void TimeIndexedProblem::Update(const std::vector<double>& x_trajectory_in)
{
    if ((int)x_trajectory_in.size() != (T_ - 1) * N)
        ThrowPretty("To update using the trajectory Update method, please use a trajectory of size N x (T-1) (" + std::to_string(N * (T_ - 1)) + "), given: " + std::to_string(x_trajectory_in.size()));
    // t = 0 holds the start state (App. C #9); evaluate all T timesteps of this one trajectory in one launch
    std::vector<double> X((size_t)T_ * N);
    memcpy(X.data(), x[0].data(), N * sizeof(double));
    memcpy(X.data() + N, x_trajectory_in.data(), (size_t)(T_ - 1) * N * sizeof(double));
    std::vector<double> P((size_t)T_ * length_Phi), J((size_t)T_ * length_jacobian * N);
    UpdateTrajectories(X.data(), 1, EXO_LAYOUT_AOS_COLMAJOR, P.data(), J.data());
    for (int t = 1; t < T_; ++t)
    {
        x[t].assign(X.begin() + (size_t)t * N, X.begin() + (size_t)(t + 1) * N);
        memcpy(Phi[t].data(), &P[(size_t)t * length_Phi], length_Phi * sizeof(double));
        if (flags_ & KIN_J) memcpy(jacobian[t].data(), &J[(size_t)t * length_jacobian * N], (size_t)length_jacobian * N * sizeof(double));
    }
}

void TimeIndexedProblem::UpdateTrajectories(const double* xin, int64_t n_trajectories, int layout, double* phi, double* jac)
{
    UpdateBatch(xin, n_trajectories * T_, layout, phi, jac);
}

}  // namespace exotica
