// exotica_host.h -- host-side mirror of the reference's interface for the batched-kinematics hot path.
//
// Same class / method names, argument meaning and error behaviour as the reference (ipab-slmc/exotica v6.2.0),
// minus KDL / Eigen / ROS types, with every evaluation forwarded to the CUDA backend through the C-ABI
// (include/exotica_b200.h).  There is no CPU evaluation path in these classes.
//
//   exotica::KinematicTree   <- exotica_core/include/exotica_core/kinematic_tree.h:145-250
//   exotica::Scene           <- exotica_core/include/exotica_core/scene.h (Update, RequestKinematics, name getters)
//   exotica::TaskMap + EffFrame/EffPosition/JointLimit/SphereCollision
//                            <- exotica_core/include/exotica_core/task_map.h:52-87, exotations/exotica_core_task_maps
//   exotica::PlanningProblem, UnconstrainedEndPoseProblem, TimeIndexedProblem (UnconstrainedTimeIndexedProblem)
//                            <- exotica_core/src/planning_problem.cpp:148-266, problems/*.cpp
//   exotica::Setup / factories / XML loading
//                            <- exotica_core/include/exotica_core/{factory.h:47-140,setup.h:64-137}, loaders/xml_loader.h
#pragma once
#include <stdint.h>

#include <array>
#include <map>
#include <memory>
#include <random>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../include/exotica_b200.h"
#include "xml.h"

namespace exotica
{
// exotica::Exception (tools/exception.h:36-47)
class Exception : public std::runtime_error
{
public:
    explicit Exception(const std::string& msg) : std::runtime_error(msg) {}
};
#define ThrowPretty(m)                                                                            \
    {                                                                                             \
        std::string _msg = std::string("In ") + __FILE__ + ":" + std::to_string(__LINE__) + ": "; \
        _msg += (m);                                                                              \
        throw ::exotica::Exception(_msg);                                                         \
    }

constexpr double inf = 1e20;  // kinematic_tree.h:70
constexpr double pi = 3.14159265358979323846;

// stand-in for KDL::Frame: rotation data[9] row-major, then p[3]
using Frame = std::array<double, 12>;
Frame IdentityFrame();
Frame FrameMul(const Frame& a, const Frame& b);
Frame FrameInverse(const Frame& a);
Frame RotationQuaternion(double x, double y, double z, double w);  // KDL::Rotation::Quaternion
Frame RotationRPY(double roll, double pitch, double yaw);          // KDL::Rotation::RPY
// exotica::GetFrame (tools/conversions.cpp:52-69): 3 (xyz), 6 (xyz rpy), 7 (xyz + qx qy qz qw, normalised)
Frame GetFrame(const std::vector<double>& val);
// ParseVector (tools/conversions.h:260-290): whitespace-separated doubles
std::vector<double> ParseVector(const std::string& value);
// kdl_parser toKdl(urdf::Pose): urdf rpy -> quaternion (setFromRPY, normalised) -> KDL::Rotation::Quaternion
Frame UrdfOriginToFrame(const double* xyz, const double* rpy);

enum class RotationType
{
    QUATERNION,
    RPY,
    ZYX,
    ZYZ,
    ANGLE_AXIS,
    MATRIX
};
int GetRotationTypeLength(const RotationType& type);
RotationType GetRotationTypeFromString(const std::string& rotation_type);

enum class BaseType
{
    FIXED = 0,
    FLOATING = 10,
    PLANAR = 20
};

enum KinematicRequestFlags
{
    KIN_FK = 0,
    KIN_J = 2,
    KIN_FK_VEL = 4,
    KIN_H = 8
};

// ---- robot description (urdfdom + srdfdom + MoveIt RobotModel behaviour the reference relies on) -----------
struct UrdfJoint
{
    std::string name, type, parent_link, child_link;
    double xyz[3] = {0, 0, 0}, rpy[3] = {0, 0, 0}, axis[3] = {1, 0, 0};
    bool has_limits = false, has_safety = false, has_mimic = false;
    double lower = 0, upper = 0, velocity = 0, soft_lower = 0, soft_upper = 0;
    std::string mimic_joint;
    double mimic_multiplier = 1.0, mimic_offset = 0.0;
};

struct UrdfLink
{
    std::string name;
    double mass = 0.0;
    double cog[3] = {0, 0, 0};
    std::vector<Frame> collision_origins;
    std::string parent_joint;               // empty for the root
    std::vector<std::string> child_links;   // ordered by ascending joint name (urdfdom initTree)
};

struct SrdfGroup
{
    std::string name;
    std::vector<std::pair<std::string, std::string>> chains;
    std::vector<std::string> joints, links, subgroups;
};

struct ModelJoint  // one moveit::core::JointModel in DFS pre-order
{
    std::string name, type;  // fixed | revolute | continuous | prismatic | floating | planar
    std::vector<std::string> variables;
    std::string parent_link, child_link;
    bool mimic = false;
};

class RobotModel
{
public:
    RobotModel(const std::string& urdf_xml, const std::string& srdf_xml);
    const UrdfLink& GetLink(const std::string& name) const;
    const UrdfJoint* GetJoint(const std::string& name) const;
    bool HasGroup(const std::string& name) const { return groups_.count(name) > 0; }
    std::vector<std::string> GetGroupNames() const;
    // JointModelGroup::getVariableNames()
    std::vector<std::string> GetGroupVariableNames(const std::string& group) const;
    // RobotModel::getVariableNames()
    std::vector<std::string> GetVariableNames() const;
    // names of active joints (>0 variables, not mimic)
    std::vector<std::string> GetActiveJointNames() const;
    struct Bounds
    {
        bool position_bounded = false;
        double min_position = 0, max_position = 0;
        bool velocity_bounded = false;
        double max_velocity = 0;
    };
    Bounds GetVariableBounds(const std::string& variable) const;  // jointBoundsFromURDF

    std::string name;
    std::map<std::string, UrdfLink> links;
    std::map<std::string, UrdfJoint> joints;
    std::string root_link;
    std::string root_joint_name, root_joint_type;  // virtual joint (SRDF) or ASSUMED_FIXED_ROOT_JOINT / fixed
    std::string world_frame_name;
    std::vector<ModelJoint> model_joints;

private:
    std::map<std::string, SrdfGroup> groups_;
    std::vector<std::string> group_order_;
    std::map<std::string, int> joint_index_;
    std::map<std::string, std::string> link_parent_joint_;
    void BuildRecursive(const std::string& link);
    void CollectGroupJoints(const std::string& group, std::vector<std::string>& seen, std::vector<std::string>& out) const;
};
typedef std::shared_ptr<RobotModel> RobotModelPtr;

// ---- KinematicTree -------------------------------------------------------------------------------------------
struct KinematicFrameRequest  // kinematic_tree.h:83-92
{
    KinematicFrameRequest() : frame_A_offset(IdentityFrame()), frame_B_offset(IdentityFrame()) {}
    KinematicFrameRequest(std::string a, Frame a_off = IdentityFrame(), std::string b = "", Frame b_off = IdentityFrame())
        : frame_A_link_name(std::move(a)), frame_A_offset(a_off), frame_B_link_name(std::move(b)), frame_B_offset(b_off) {}
    std::string frame_A_link_name;
    Frame frame_A_offset;
    std::string frame_B_link_name;
    Frame frame_B_offset;
};

struct KinematicsRequest  // kinematic_tree.h:94-100
{
    int flags = KIN_FK;
    std::vector<KinematicFrameRequest> frames;
};

struct KinematicElement  // kinematic_element.h:54-143 (flattened: parent is an index into the tree)
{
    int id = -1;
    int parent = -1;
    std::string link_name, joint_name, parent_name;
    int joint_type = EXO_JOINT_NONE;
    double origin[3] = {0, 0, 0}, axis[3] = {0, 0, 0};
    Frame f_tip = IdentityFrame();
    int control_id = -1;
    bool is_controlled = false;
    bool is_mimic_joint = false;
    int mimic_element = -1;
    double mimic_multiplier = 1.0, mimic_offset = 0.0;
    std::vector<double> joint_limits;
    double velocity_limit = 0.0;
    bool is_robot_link = false;
    bool is_trajectory_generated = false;
    int generated_slot = -1;
    double mass = 0.0;
    double cog[3] = {0, 0, 0};
};

// What KinematicTree::Update fills (kinematic_tree.h:112-126), for ONE configuration
struct KinematicResponse
{
    int flags = KIN_FK;
    std::vector<double> x;
    std::vector<Frame> Phi;                     // [F]
    std::vector<std::vector<double>> jacobian;  // [F] 6 x N column-major (KDL::Jacobian::data)
};

class KinematicTree
{
public:
    KinematicTree();
    ~KinematicTree();
    KinematicTree(const KinematicTree&) = delete;
    KinematicTree& operator=(const KinematicTree&) = delete;

    void Instantiate(const std::string& joint_group, RobotModelPtr model, const std::string& name);
    const std::string& GetRootFrameName() const;
    const std::string& GetRootJointName() const;
    BaseType GetModelBaseType() const { return model_base_type_; }
    BaseType GetControlledBaseType() const { return controlled_base_type_; }

    std::shared_ptr<KinematicResponse> RequestFrames(const KinematicsRequest& request);
    void Update(const std::vector<double>& x);  // one configuration, evaluated on the GPU
    // NEW: batched KinematicTree::Update.  Host buffers; Phi[F][12] and J[F][6][N] per evaluation in `layout`
    // (as an EXO_MAP_FRAMES problem).  J may be NULL.
    void UpdateBatch(const double* x, int64_t n_evals, int layout, double* phi, double* jacobian);
    std::shared_ptr<KinematicResponse> GetKinematicResponse() { return solution_; }

    Frame FK(const std::string& element_A, const Frame& offset_a, const std::string& element_B, const Frame& offset_b);
    // 6 x N, column-major
    std::vector<double> Jacobian(const std::string& element_A, const Frame& offset_a, const std::string& element_B, const Frame& offset_b);

    int GetNumControlledJoints() const { return num_controlled_joints_; }
    int GetNumModelJoints() const { return num_joints_; }
    const std::vector<std::string>& GetControlledJointNames() const { return controlled_joints_names_; }
    const std::vector<std::string>& GetControlledLinkNames() const { return controlled_link_names_; }
    const std::vector<std::string>& GetModelJointNames() const { return model_joints_names_; }
    const std::vector<std::string>& GetModelLinkNames() const { return model_link_names_; }
    bool HasModelLink(const std::string& link) const;
    int IsControlledLink(const std::string& link_name) const;

    // N x 2, row-major [lower, upper]
    std::vector<double> GetJointLimits() const;
    const std::vector<double>& GetLowerLimits() const { return lower_; }
    const std::vector<double>& GetUpperLimits() const { return upper_; }
    const std::vector<double>& GetVelocityLimits() const { return velocity_limits_; }
    void SetJointLimitsLower(const std::vector<double>& lower_in);
    void SetJointLimitsUpper(const std::vector<double>& upper_in);
    void ResetJointLimits();
    std::vector<double> GetRandomControlledState();
    void SetSeed(unsigned int seed) { generator_.seed(seed); }

    std::vector<double> GetModelState() const;
    std::map<std::string, double> GetModelStateMap() const;
    void SetModelState(const std::vector<double>& x);
    void SetModelState(const std::map<std::string, double>& x);
    std::vector<double> GetControlledState() const;

    // Scene objects (KinematicTree::AddEnvironmentElement / AddElement, kinematic_tree.cpp:434-578, reduced to frames)
    int AddEnvironmentElement(const std::string& name, const Frame& transform, const std::string& parent = "", bool trajectory_generated = false);
    int AddElement(const std::string& name, const Frame& transform, const std::string& parent = "");
    int MakeTrajectoryGenerated(int element_index);  // returns the element's slot in the per-timestep frame table
    // KinematicTree::ResetModel (kinematic_tree.cpp:436-452): drops every environment element, the robot model stays
    void ResetModel();
    int GetNumGenerated() const { return num_generated_; }
    // KinematicElement::generated_offset of a trajectory-driven element (Scene::UpdateTrajectoryGenerators, scene.cpp:329-335): the
    // pose the one-configuration entry points (Update, FK, Jacobian) use; time-indexed problems push whole per-timestep tables instead
    void SetGeneratedOffset(int slot, const Frame& frame);

    const std::vector<KinematicElement>& GetTree() const { return tree_; }
    int GetElementIndex(const std::string& link_name) const;  // tree_map_ lookup, -1 if missing
    int ResolveFrame(const std::string& link_name, const char* what) const;
    exo_tree* GetBackend();  // builds / rebuilds the C-ABI tree when the structure changed
    // Bumped whenever the C-ABI tree is destroyed (scene objects / trajectories added, tree rebuilt): every exo_problem built on
    // the old tree is stale and its owner (PlanningProblem, TaskMap, the frames problem here) rebuilds it on the next call.
    uint64_t GetBackendGeneration() const { return backend_generation_; }

private:
    void BuildTree();
    int IsControlled(const KinematicElement& joint) const;
    void UpdateJointLimits();
    void InvalidateBackend();
    void PushModelStateToBackend();

    RobotModelPtr model_;
    std::string name_;
    std::string root_joint_name_;
    BaseType model_base_type_ = BaseType::FIXED, controlled_base_type_ = BaseType::FIXED;
    std::vector<KinematicElement> tree_;  // model_tree_ followed by scene elements
    int num_model_elements_ = 0;
    std::map<std::string, int> tree_map_;
    std::map<std::string, int> model_joints_map_, controlled_joints_map_;
    std::vector<int> controlled_joints_;  // control id -> element index
    std::vector<std::string> model_joints_names_, controlled_joints_names_, model_link_names_, controlled_link_names_;
    std::vector<double> tree_state_;  // by element index
    std::vector<double> lower_, upper_, velocity_limits_;
    int num_joints_ = 0, num_controlled_joints_ = 0, num_generated_ = 0;
    std::mt19937 generator_;
    std::vector<std::uniform_real_distribution<double>> random_state_distributions_;

    exo_tree* backend_ = nullptr;
    uint64_t backend_generation_ = 0;
    exo_problem* frames_problem_ = nullptr;  // EXO_MAP_FRAMES problem behind RequestFrames/Update
    bool frames_requested_ = false;
    void EnsureFramesProblem();              // (re)builds frames_problem_ from request_ after the tree changed
    std::vector<Frame> generated_offsets_;   // by generated slot
public:
    void PushCurrentGenerated(exo_problem* p);  // the current generated offsets as a one-timestep table
private:
    KinematicsRequest request_;
    std::shared_ptr<KinematicResponse> solution_;
};

// ---- Initializer (property.h:44-86, reduced to strings + nested initializers) ----------------------------------------
struct Initializer
{
    std::string name;                                               // e.g. "exotica/EffFrame"
    std::map<std::string, std::string> properties;                  // scalar / vector properties as text
    std::map<std::string, std::vector<Initializer>> children;       // list-of-initializer properties
    bool Has(const std::string& key) const { return properties.count(key) > 0; }
    std::string Get(const std::string& key, const std::string& fallback) const
    {
        auto it = properties.find(key);
        return it == properties.end() ? fallback : it->second;
    }
    std::string Require(const std::string& key) const;
    const std::vector<Initializer>& List(const std::string& key) const;
};
// XMLLoader semantics (loaders/xml_loader.cpp): attributes and leaf child elements become properties, child elements
// with element children become lists of initializers.
Initializer InitializerFromXml(const xml::Element& e, const std::string& prefix = "exotica/");

// exotica::Trajectory (exotica_core/include/exotica_core/trajectory.h, src/trajectory.cpp): piecewise Cartesian trajectory with
// KDL's line path / single-axis rotation / cubic velocity profile semantics (restated in host/trajectory.cpp)
class Trajectory
{
public:
    Trajectory() {}
    explicit Trajectory(const std::string& data);  // "radius\n rows cols\n time x y z [rotation]..."
    Trajectory(const std::vector<std::vector<double>>& data, double radius = 1.0);
    Frame GetPosition(double t) const;
    double GetDuration() const;
    const std::vector<std::vector<double>>& GetData() const { return data_; }
    double GetRadius() const { return radius_; }
    std::string ToString() const;

private:
    struct Segment
    {
        Frame start, end;
        double duration = 0.0;
        bool stationary = true;
        double dir[3] = {0, 0, 0}, axis[3] = {0, 0, 1};
        double path_length = 0.0, scale_rot = 1.0, scale_lin = 1.0;
    };
    void ConstructFromData(const std::vector<std::vector<double>>& data, double radius);
    Frame SegmentPos(const Segment& s, double time) const;
    std::vector<Segment> segments_;
    std::vector<std::vector<double>> data_;
    double radius_ = 1.0;
};

// ---- Scene -------------------------------------------------------------------------------------------------------
class Scene
{
public:
    Scene();
    // init/scene.in: JointGroup, URDF, SRDF (file contents or paths), optional JointPositionLimitsLower/Upper
    void Instantiate(const Initializer& init);
    void InstantiateFromStrings(const std::string& urdf_xml, const std::string& srdf_xml, const std::string& joint_group);
    KinematicTree& GetKinematicTree() { return kinematica_; }
    std::shared_ptr<KinematicResponse> RequestKinematics(const KinematicsRequest& request) { return kinematica_.RequestFrames(request); }
    void Update(const std::vector<double>& x, double t = 0);  // scene.cpp:337-348
    void UpdateTrajectoryGenerators(double t = 0);            // scene.cpp:329-335
    std::string GetName() const { return name_; }
    std::string GetRootFrameName() { return kinematica_.GetRootFrameName(); }
    std::string GetRootJointName() { return kinematica_.GetRootJointName(); }
    std::vector<std::string> GetControlledJointNames() { return kinematica_.GetControlledJointNames(); }
    std::vector<std::string> GetControlledLinkNames() { return kinematica_.GetControlledLinkNames(); }
    std::vector<std::string> GetModelJointNames() { return kinematica_.GetModelJointNames(); }
    std::vector<std::string> GetModelLinkNames() { return kinematica_.GetModelLinkNames(); }
    std::vector<double> GetControlledState() { return kinematica_.GetControlledState(); }
    std::vector<double> GetModelState() { return kinematica_.GetModelState(); }
    void SetModelState(const std::vector<double>& x) { kinematica_.SetModelState(x); }
    void SetModelState(const std::map<std::string, double>& x) { kinematica_.SetModelState(x); }
    int get_num_positions() const { return num_positions_; }
    int get_num_velocities() const { return 0; }
    // Scene::AddObject (scene.cpp:829-...): a named frame attached to `parent` (default: world)
    void AddObject(const std::string& name, const Frame& transform = IdentityFrame(), const std::string& parent = "");
    // Scene::LoadScene / LoadSceneFile (scene.cpp:610-645): world objects of a MoveIt ".scene" geometry text.  Every object becomes
    // a frame named after it, posed like its first shape (UpdateSceneFrames, scene.cpp:702-758), with one child frame
    // "<name>_collision_<i>" per shape; the shapes themselves belong to the collision scene (out of scope) and are skipped.
    void LoadScene(const std::string& scene_text, const Frame& offset = IdentityFrame());
    void LoadSceneFile(const std::string& file_name, const Frame& offset = IdentityFrame());
    void CleanScene();                           // scene.cpp:655-660
    std::vector<std::string> GetWorldObjectNames() const;
    // Scene::AddTrajectory (scene.cpp:917-935) replaced by an explicit per-timestep frame table [T][12]
    void AddTrajectoryFrames(const std::string& link, const std::vector<Frame>& frames_per_timestep);
    // Scene::AddTrajectory(link, traj) (scene.cpp:922-935): the link follows the trajectory; time-indexed problems sample it at
    // t * tau for every timestep (AbstractTimeIndexedProblem::Update -> Scene::Update(x, t * tau), scene.cpp:337-348)
    void AddTrajectory(const std::string& link, const std::string& traj);
    void AddTrajectory(const std::string& link, std::shared_ptr<Trajectory> traj);
    std::shared_ptr<Trajectory> GetTrajectory(const std::string& link) const;
    void SampleTrajectories(int T, double tau);
    // generated frames [T][n_generated][12] (empty if no link is trajectory driven)
    const std::vector<double>& GetGeneratedTable(int& T, int& n_generated) const;

private:
    std::string name_;
    KinematicTree kinematica_;
    int num_positions_ = 0;
    std::map<int, std::vector<Frame>> trajectories_;  // generated slot -> frames per timestep
    std::map<std::string, std::pair<int, std::shared_ptr<Trajectory>>> trajectory_generators_;  // link -> (slot, trajectory)
    mutable std::vector<double> generated_table_;
    uint64_t trajectory_version_ = 0;  // bumped by AddTrajectory / AddTrajectoryFrames (problems re-push their frame tables)
    struct WorldObject
    {
        std::vector<Frame> shape_poses;  // world pose of every shape
    };
    struct CustomLink
    {
        std::string name, parent;
        Frame transform;
    };
    std::map<std::string, WorldObject> world_objects_;  // planning_scene::World iterates its objects by name (std::map)
    std::vector<CustomLink> custom_links_;              // Links initializers / AddObject, in insertion order
    std::map<std::string, std::vector<Frame>> trajectory_tables_;  // link -> explicit per-timestep frames (AddTrajectoryFrames)
    void RebuildEnvironment();  // UpdateSceneFrames + UpdateInternalFrames: world objects (by name), custom links, trajectories

public:
    uint64_t GetTrajectoryVersion() const { return trajectory_version_; }
};
typedef std::shared_ptr<Scene> ScenePtr;

// ---- Task maps --------------------------------------------------------------------------------------------------------
class TaskMap
{
public:
    virtual ~TaskMap() {}
    virtual void InstantiateBase(const Initializer& init);  // task_map.cpp:42-55
    virtual void Instantiate(const Initializer& init) {}
    virtual void AssignScene(ScenePtr scene) { scene_ = scene; }
    virtual int TaskSpaceDim() = 0;
    virtual int TaskSpaceJacobianDim() { return TaskSpaceDim(); }
    virtual int Kind() const = 0;  // exo_map_kind
    virtual void FillDescriptor(exo_task_map_desc& desc) {}
    // TaskSpaceVector::SetZero for this map's entries (task_space_vector.cpp:54-62 over GetLieGroupIndices): y is pre-zeroed
    virtual void FillZeroGoal(double* y) {}
    std::vector<KinematicFrameRequest> GetFrames() const { return frames_; }
    const std::string& GetObjectName() const { return object_name_; }
    // One configuration through the GPU (a private one-map problem): phi[length], jacobian[length_jacobian x N] col-major
    void Update(const std::vector<double>& x, std::vector<double>& phi);
    void Update(const std::vector<double>& x, std::vector<double>& phi, std::vector<double>& jacobian);
    // NEW: batched TaskMap::Update (host buffers)
    void UpdateBatch(const double* x, int64_t n_evals, int layout, double* phi, double* jacobian);

    int id = -1;
    int start = -1, length = -1, start_jacobian = -1, length_jacobian = -1;
    bool is_used = true;
    std::string ns_;

protected:
    friend class PlanningProblem;
    std::vector<KinematicFrameRequest> frames_;
    ScenePtr scene_;
    std::string object_name_;
    // resolved frame arrays kept alive for descriptors
    std::vector<int32_t> elem_a_, elem_b_;
    std::vector<double> off_a_, off_b_;
    void ResolveFrames();
    exo_problem* own_problem_ = nullptr;
    uint64_t own_problem_generation_ = 0;  // KinematicTree::GetBackendGeneration() own_problem_ was built against
    void EnsureOwnProblem();

public:
    void ReleaseOwnProblem();
};
typedef std::shared_ptr<TaskMap> TaskMapPtr;

class EffFrame : public TaskMap  // eff_frame.cpp
{
public:
    void Instantiate(const Initializer& init) override;
    int TaskSpaceDim() override { return (int)frames_.size() * big_stride_; }
    int TaskSpaceJacobianDim() override { return (int)frames_.size() * 6; }
    int Kind() const override { return EXO_MAP_EFF_FRAME; }
    void FillDescriptor(exo_task_map_desc& desc) override { desc.rotation_type = (int)rotation_type_; }
    void FillZeroGoal(double* y) override
    {
        // identity rotation encodings at the Lie-group entries (eff_frame.cpp:43-51): quaternion 0 0 0 1, matrix I, else zeros
        for (int f = 0; f < (int)frames_.size(); ++f)
        {
            double* r = y + f * big_stride_ + 3;
            if (rotation_type_ == RotationType::QUATERNION) r[3] = 1.0;
            if (rotation_type_ == RotationType::MATRIX) r[0] = r[4] = r[8] = 1.0;
        }
    }
    const RotationType& get_rotation_type() const { return rotation_type_; }

private:
    RotationType rotation_type_ = RotationType::RPY;
    int small_stride_ = 3, big_stride_ = 6;
};

class EffPosition : public TaskMap  // eff_position.cpp
{
public:
    int TaskSpaceDim() override { return (int)frames_.size() * 3; }
    int Kind() const override { return EXO_MAP_EFF_POSITION; }
};

// ---- sibling maps on the same frames ----
class EffOrientation : public TaskMap  // eff_orientation.cpp
{
public:
    void Instantiate(const Initializer& init) override;
    int TaskSpaceDim() override { return (int)frames_.size() * stride_; }
    int TaskSpaceJacobianDim() override { return (int)frames_.size() * 3; }
    int Kind() const override { return EXO_MAP_EFF_ORIENTATION; }
    void FillDescriptor(exo_task_map_desc& desc) override { desc.rotation_type = (int)rotation_type_; }
    void FillZeroGoal(double* y) override
    {
        for (int f = 0; f < (int)frames_.size(); ++f)
        {
            double* r = y + f * stride_;
            if (rotation_type_ == RotationType::QUATERNION) r[3] = 1.0;
            if (rotation_type_ == RotationType::MATRIX) r[0] = r[4] = r[8] = 1.0;
        }
    }

private:
    RotationType rotation_type_ = RotationType::RPY;
    int stride_ = 3;
};

class EffPositionXY : public TaskMap  // eff_position_xy.cpp
{
public:
    int TaskSpaceDim() override { return (int)frames_.size() * 2; }
    int Kind() const override { return EXO_MAP_EFF_POSITION_XY; }
};

class Distance : public TaskMap  // distance.cpp
{
public:
    int TaskSpaceDim() override { return (int)frames_.size(); }
    int Kind() const override { return EXO_MAP_DISTANCE; }
};

class EffAxisAlignment : public TaskMap  // eff_axis_alignment.cpp
{
public:
    void InstantiateBase(const Initializer& init) override;  // EndEffector entries are FrameWithAxisAndDirection initializers
    int TaskSpaceDim() override { return (int)frames_.size(); }
    int Kind() const override { return EXO_MAP_EFF_AXIS_ALIGNMENT; }
    void FillDescriptor(exo_task_map_desc& desc) override
    {
        desc.axis = axis_.data();
        desc.direction = dir_.data();
    }
    // eff_axis_alignment.cpp:176-200 (normalised, like the reference)
    std::vector<double> GetDirection(const std::string& frame_name) const;
    void SetDirection(const std::string& frame_name, const std::vector<double>& dir);

private:
    std::vector<double> axis_, dir_;  // [n][3]
};

class CenterOfMass : public TaskMap  // center_of_mass.cpp
{
public:
    void Instantiate(const Initializer& init) override;
    void AssignScene(ScenePtr scene) override;  // Initialize(): offsets = centres of gravity, weights = link masses
    int TaskSpaceDim() override { return dim_; }
    int Kind() const override { return EXO_MAP_CENTER_OF_MASS; }
    void FillDescriptor(exo_task_map_desc& desc) override
    {
        desc.mass = mass_.data();
        desc.enable_z = enable_z_ ? 1 : 0;
    }
    const std::vector<double>& get_mass() const { return mass_; }

private:
    bool enable_z_ = true;
    int dim_ = 3;
    std::vector<double> mass_;
};

class JointPose : public TaskMap  // joint_pose.cpp
{
public:
    void Instantiate(const Initializer& init) override;
    void AssignScene(ScenePtr scene) override;
    int TaskSpaceDim() override { return (int)joint_map_.size(); }
    int Kind() const override { return EXO_MAP_JOINT_POSE; }
    void FillDescriptor(exo_task_map_desc& desc) override
    {
        desc.n_joint_map = (int)joint_map_.size();
        desc.joint_map = joint_map_.data();
        desc.joint_ref = joint_ref_.data();
    }
    const std::vector<int32_t>& get_joint_map() const { return joint_map_; }
    const std::vector<double>& get_joint_ref() const { return joint_ref_; }

private:
    std::vector<int32_t> joint_map_, init_map_;
    std::vector<double> joint_ref_, init_ref_;
};

class JointLimit : public TaskMap  // joint_limit.cpp
{
public:
    void Instantiate(const Initializer& init) override;
    void AssignScene(ScenePtr scene) override;
    int TaskSpaceDim() override { return N_; }
    int Kind() const override { return EXO_MAP_JOINT_LIMIT; }
    void FillDescriptor(exo_task_map_desc& desc) override { desc.safe_percentage = safe_percentage_; }

private:
    double safe_percentage_ = 0.0;
    int N_ = 0;
};

class SphereCollision : public TaskMap  // sphere_collision.cpp
{
public:
    void InstantiateBase(const Initializer& init) override;
    void Instantiate(const Initializer& init) override;
    int TaskSpaceDim() override { return dim_; }
    int Kind() const override { return EXO_MAP_SPHERE_COLLISION; }
    void FillDescriptor(exo_task_map_desc& desc) override;

private:
    double precision_ = 1.0;
    std::vector<double> radiuses_;
    std::vector<std::string> group_names_;  // per sphere
    std::vector<int32_t> group_rank_;       // lexicographic rank (std::map iteration order)
    int dim_ = 0;
};

// ---- Factories (factory.h:47-140, setup.h:64-137) --------------------------------------------------------------------------
class PlanningProblem;
typedef std::shared_ptr<PlanningProblem> PlanningProblemPtr;

class Setup
{
public:
    typedef TaskMapPtr (*MapCreator)();
    typedef PlanningProblemPtr (*ProblemCreator)();
    static Setup& Instance();
    void RegisterMap(const std::string& type, MapCreator creator);          // REGISTER_TASKMAP_TYPE
    void RegisterProblem(const std::string& type, ProblemCreator creator);  // REGISTER_PROBLEM_TYPE
    static std::vector<std::string> GetMaps();
    static std::vector<std::string> GetProblems();
    static TaskMapPtr CreateMap(const Initializer& init);
    static PlanningProblemPtr CreateProblem(const Initializer& init);
    static PlanningProblemPtr CreateProblem(const std::string& type);
    // XMLLoader::LoadProblem: first child of the document root that names a registered problem
    static PlanningProblemPtr LoadProblem(const std::string& xml_text);
    static Initializer LoadProblemInitializer(const std::string& xml_text);
    // {package} / package:// resolution (tools.cpp:150-190) without ROS: explicit map package -> directory
    static void SetPackagePath(const std::string& package, const std::string& directory);
    static std::string ParsePath(const std::string& path);
    static std::string LoadFile(const std::string& path);

private:
    std::map<std::string, MapCreator> maps_;
    std::map<std::string, ProblemCreator> problems_;
    std::map<std::string, std::string> packages_;
};

// ---- Problems --------------------------------------------------------------------------------------------------------------
class PlanningProblem
{
public:
    virtual ~PlanningProblem();
    virtual void InstantiateBase(const Initializer& init);  // planning_problem.cpp:148-244
    virtual void Instantiate(const Initializer& init) = 0;
    ScenePtr GetScene() const { return scene_; }
    std::vector<TaskMapPtr>& GetTasks() { return tasks_; }
    std::map<std::string, TaskMapPtr>& GetTaskMaps() { return task_maps_; }
    const std::string& GetObjectName() const { return object_name_; }
    int N = 0;
    int length_Phi = 0, length_jacobian = 0;
    std::vector<double> GetStartState() const { return start_state_; }
    void SetStartState(const std::vector<double>& x);
    int GetFlags() const { return flags_; }
    virtual int GetT() const { return 1; }
    unsigned int GetNumberOfProblemUpdates() const { return number_of_problem_updates_; }
    void ResetNumberOfProblemUpdates() { number_of_problem_updates_ = 0; }

    // NEW batched entry point (fills Phi and J for all evaluations in one call).
    //  host buffers: X/Phi/J in `layout`; device variant takes device pointers + stream.
    void UpdateBatch(const double* x, int64_t n_evals, int layout, double* phi, double* jacobian);
    void UpdateBatchDevice(const double* d_x, int64_t n_evals, int layout, double* d_phi, double* d_jacobian, void* stream);
    // NEW: scalar cost of every trajectory (time-indexed: GetCost, abstract_time_indexed_problem.cpp:299-307; end-pose:
    // GetScalarCost, unconstrained_end_pose_problem.cpp:79-82) for X[n_trajectories][T][N]; forward kinematics only.
    void CostBatch(const double* x, int64_t n_trajectories, double* cost);
    // NEW: gradient of that cost with respect to every state, jacobian[n_trajectories][T][N] (time-indexed: GetCostJacobian,
    // abstract_time_indexed_problem.cpp:309-318, with a zero row for t = 0; end-pose: GetScalarJacobian,
    // unconstrained_end_pose_problem.cpp:84-87); cost may be NULL.  Phi and J stay on the device.
    void CostJacobianBatch(const double* x, int64_t n_trajectories, double* cost, double* jacobian);
    // device variant on the buffers of UpdateBatchDevice; cost_eval (per timestep) may be NULL
    void CostBatchDevice(const double* d_x, const double* d_phi, int64_t n_trajectories, int layout, double* d_cost, double* d_cost_eval, void* stream);
    // UpdateBatchDevice + CostBatchDevice in one call: Update(x, t) of every timestep and GetCost() of every trajectory
    void UpdateCostBatchDevice(const double* d_x, int64_t n_trajectories, int layout, double* d_phi, double* d_jacobian, double* d_cost, double* d_cost_eval,
                               void* stream);
    // NEW: multi-start IK (IKSolver::Solve, exotations/solvers/exotica_ik_solver/src/ik_solver.cpp:73-240) from n start states
    // q0[n][N]; solutions[n][N], cost[n], iterations[n], termination[n] (exotica::TerminationCriterion values)
    void IkSolveBatch(const double* q0, int64_t n, const exo_ik_params& params, double* solutions, double* cost, int32_t* iterations,
                      int32_t* termination);
    // NEW: exotica::Hessian of every evaluation, hess[n][length_jacobian][N][N] (KinematicTree::ComputeH + the maps' overloads)
    void HessianBatch(const double* x, int64_t n_evals, double* hess);
    exo_problem* GetBackend();
    float LastKernelMs();
    int64_t LaunchCount();

protected:
    void UpdateUsage(const std::vector<std::vector<double>>& rho_per_task);  // TaskMap::is_used from rho (tasks.cpp:246-259)
    void InvalidateBackend();
    // Called by every batched entry point after GetBackend(): state the backend needs beyond its maps (time-indexed problems
    // push the per-timestep frames of trajectory-driven links here).
    virtual void PrepareBackend() {}
    uint64_t backend_generation_ = 0;  // KinematicTree::GetBackendGeneration() backend_ was built against
    uint64_t backend_serial_ = 0;      // counts the exo_problem handles this problem has built (a new one needs every push again)
    std::string object_name_;
    ScenePtr scene_;
    std::vector<TaskMapPtr> tasks_;
    std::map<std::string, TaskMapPtr> task_maps_;
    std::vector<double> start_state_;
    int flags_ = KIN_FK | KIN_J;
    double t_start = 0.0;
    unsigned int number_of_problem_updates_ = 0;
    exo_problem* backend_ = nullptr;
    // cost task indexing (Task::Initialize, tasks.cpp:36-66): names of the maps used by Cost
    std::vector<std::string> cost_tasks_;
    std::vector<std::vector<double>> cost_rho_;  // per cost task: rho (1 value, or T values)
    std::vector<std::vector<double>> cost_goal_; // per cost task: Goal (empty = the SetZero goal)
    double cost_ct_ = 1.0;                       // 1/tau/T for time-indexed problems
    std::vector<double> cost_W_;                 // diagonal of W (time-indexed problems), empty = no transition cost
    bool cost_pushed_ = false;
    void ParseCost(const Initializer& init, int T);
    void PushCost();
};

class UnconstrainedEndPoseProblem : public PlanningProblem  // unconstrained_end_pose_problem.cpp
{
public:
    void Instantiate(const Initializer& init) override;
    void Update(const std::vector<double>& x);  // :123-167 up to the task-map loop
    std::vector<double> Phi;                    // length_Phi
    std::vector<double> jacobian;               // length_jacobian x N, column-major (Eigen::MatrixXd)
};

class TimeIndexedProblem : public PlanningProblem  // abstract_time_indexed_problem.cpp, time_indexed_problem.cpp
{
public:
    void Instantiate(const Initializer& init) override;
    int GetT() const override { return T_; }
    double GetTau() const { return tau_; }
    void Update(const std::vector<double>& x_in, int t);     // :220-292 up to the task-map loop
    void Update(const std::vector<double>& x_trajectory_in); // :209-218, (T-1)*N values, t = 1..T-1
    // NEW: B trajectories x T timesteps in one call; X[B][T][N] (AoS) or X[N][B*T] (SoA)
    void UpdateTrajectories(const double* x, int64_t n_trajectories, int layout, double* phi, double* jacobian);
    std::vector<std::vector<double>> x;         // [T][N]
    std::vector<std::vector<double>> Phi;       // [T][length_Phi]
    std::vector<std::vector<double>> jacobian;  // [T] length_jacobian x N column-major

private:
    int T_ = 0;
    double tau_ = 0.0;
    void PushGenerated();
    void PrepareBackend() override { PushGenerated(); }
    uint64_t pushed_serial_ = 0;            // backend serial / trajectory version the frame table was last pushed for
    uint64_t pushed_version_ = ~0ull;
};

}  // namespace exotica
