// kinematic_tree.cpp -- host mirror of exotica::KinematicTree (exotica_core/src/kinematic_tree.cpp), with every
// evaluation forwarded to the CUDA backend through the C-ABI.
#include <math.h>
#include <string.h>

#include <algorithm>

#include "exotica_host.h"

namespace exotica
{
Frame UrdfOriginToFrame(const double* xyz, const double* rpy);

static void CheckRc(int rc)
{
    if (rc != EXO_OK) throw Exception(exo_last_error());
}

static double KdlNorm(const double* d)
{
    // KDL::Vector::Norm
    double tmp1 = fabs(d[0]), tmp2 = fabs(d[1]);
    if (tmp1 >= tmp2)
    {
        tmp2 = fabs(d[2]);
        if (tmp1 >= tmp2)
        {
            if (tmp1 == 0) return 0;
            return tmp1 * sqrt(1 + (d[1] / d[0]) * (d[1] / d[0]) + (d[2] / d[0]) * (d[2] / d[0]));
        }
        return tmp2 * sqrt(1 + (d[0] / d[2]) * (d[0] / d[2]) + (d[1] / d[2]) * (d[1] / d[2]));
    }
    tmp1 = fabs(d[2]);
    if (tmp2 > tmp1) return tmp2 * sqrt(1 + (d[0] / d[1]) * (d[0] / d[1]) + (d[2] / d[1]) * (d[2] / d[1]));
    return tmp1 * sqrt(1 + (d[0] / d[2]) * (d[0] / d[2]) + (d[1] / d[2]) * (d[1] / d[2]));
}

KinematicTree::KinematicTree() : generator_(std::random_device{}()) {}

KinematicTree::~KinematicTree()
{
    InvalidateBackend();
}

void KinematicTree::InvalidateBackend()
{
    if (frames_problem_) exo_problem_destroy(frames_problem_);
    frames_problem_ = nullptr;
    if (backend_) exo_tree_destroy(backend_);
    backend_ = nullptr;
    ++backend_generation_;  // dependants (PlanningProblem::GetBackend, TaskMap::EnsureOwnProblem) rebuild their exo_problem
}

// kinematic_tree.cpp:101-162
void KinematicTree::Instantiate(const std::string& joint_group, RobotModelPtr model, const std::string& name)
{
    if (!model) ThrowPretty("No robot model provided!");
    model_joints_names_ = model->GetVariableNames();
    name_ = name;
    controlled_joints_names_.clear();
    if (!model->HasGroup(joint_group))
    {
        if (!joint_group.empty())
        {
            std::string msg = "Joint group '" + joint_group + "' not defined in the robot model. " + std::to_string(model->GetGroupNames().size()) + " joint groups available";
            const auto names = model->GetGroupNames();
            if (!names.empty())
            {
                msg += ": ";
                for (const auto& n : names) msg += n + ", ";
            }
            ThrowPretty(msg);
        }
        // default: all active joints (no fixed, no mimic)
        controlled_joints_names_ = model->GetActiveJointNames();
    }
    else
    {
        controlled_joints_names_ = model->GetGroupVariableNames(joint_group);
    }
    model_ = model;
    BuildTree();
}

// kinematic_tree.cpp:164-421 (kinematic part)
void KinematicTree::BuildTree()
{
    InvalidateBackend();
    tree_.clear();
    tree_map_.clear();
    model_joints_map_.clear();
    controlled_joints_map_.clear();
    model_link_names_.clear();
    controlled_link_names_.clear();
    num_generated_ = 0;

    root_joint_name_ = model_->root_joint_name;
    const std::string world_frame_name = model_->world_frame_name;

    KinematicElement world;
    world.id = -1;
    world.parent = -1;
    world.link_name = world_frame_name;
    world.joint_name = root_joint_name_;
    tree_.push_back(world);
    if (model_->root_joint_type == "fixed")
    {
        model_base_type_ = BaseType::FIXED;
    }
    else if (model_->root_joint_type == "floating")
    {
        model_base_type_ = BaseType::FLOATING;
        const int types[6] = {EXO_JOINT_TRANS_X, EXO_JOINT_TRANS_Y, EXO_JOINT_TRANS_Z, EXO_JOINT_ROT_Z, EXO_JOINT_ROT_Y, EXO_JOINT_ROT_X};
        const char* suffix[6] = {"/trans_x", "/trans_y", "/trans_z", "/rot_z", "/rot_y", "/rot_x"};
        for (int i = 0; i < 6; ++i)
        {
            KinematicElement e;
            e.id = i;
            e.parent = i;  // model_tree_[i]
            e.parent_name = tree_[i].link_name;
            e.link_name = world_frame_name + suffix[i];
            e.joint_name = root_joint_name_ + suffix[i];
            e.joint_type = types[i];
            tree_.push_back(e);
        }
        // exotica follows an RPY virtual-joint convention: drop rot_w from the variable lists
        const std::string rot_w = root_joint_name_ + "/rot_w";
        auto it = std::find(controlled_joints_names_.begin(), controlled_joints_names_.end(), rot_w);
        if (it != controlled_joints_names_.end()) controlled_joints_names_.erase(it);
        it = std::find(model_joints_names_.begin(), model_joints_names_.end(), rot_w);
        if (it != model_joints_names_.end()) model_joints_names_.erase(it);
    }
    else if (model_->root_joint_type == "planar")
    {
        model_base_type_ = BaseType::PLANAR;
        const int types[3] = {EXO_JOINT_TRANS_X, EXO_JOINT_TRANS_Y, EXO_JOINT_ROT_Z};
        const char* suffix[3] = {"/x", "/y", "/theta"};
        for (int i = 0; i < 3; ++i)
        {
            KinematicElement e;
            e.id = i;
            e.parent = i;
            e.parent_name = tree_[i].link_name;
            e.link_name = root_joint_name_ + suffix[i];
            e.joint_name = root_joint_name_ + suffix[i];
            e.joint_type = types[i];
            tree_.push_back(e);
        }
    }
    else
    {
        ThrowPretty("Unsupported root joint type: " + model_->root_joint_type);
    }

    // AddElementFromSegmentMapIterator (:580-589) over the kdl_parser tree: recursive pre-order, children in urdfdom order
    struct Rec
    {
        static void Add(KinematicTree& kt, const RobotModel& m, const std::string& link_name, int parent)
        {
            const UrdfLink& link = m.GetLink(link_name);
            KinematicElement e;
            e.id = (int)kt.tree_.size();
            e.parent = parent;
            e.parent_name = kt.tree_[parent].link_name;
            e.link_name = link_name;
            e.mass = link.mass;
            memcpy(e.cog, link.cog, sizeof(e.cog));
            if (link.parent_joint.empty())
            {
                e.joint_name = "NoName";  // KDL root segment: default Joint(None)
            }
            else
            {
                const UrdfJoint& j = *m.GetJoint(link.parent_joint);
                e.joint_name = j.name;
                const Frame F = UrdfOriginToFrame(j.xyz, j.rpy);
                if (j.type == "revolute" || j.type == "continuous" || j.type == "prismatic")
                {
                    // kdl_parser: Joint(name, F.p, F.M * axis, RotAxis | TransAxis); KDL::Joint normalises the axis
                    double a[3];
                    for (int r = 0; r < 3; ++r) a[r] = F[3 * r] * j.axis[0] + F[3 * r + 1] * j.axis[1] + F[3 * r + 2] * j.axis[2];
                    const double n = KdlNorm(a);
                    for (int r = 0; r < 3; ++r) e.axis[r] = a[r] / n;
                    for (int r = 0; r < 3; ++r) e.origin[r] = F[9 + r];
                    e.joint_type = j.type == "prismatic" ? EXO_JOINT_TRANS_AXIS : EXO_JOINT_ROT_AXIS;
                    // Segment: f_tip = joint.pose(0)^-1 * F = (F.M, F.p - origin)
                    e.f_tip = F;
                    for (int r = 0; r < 3; ++r) e.f_tip[9 + r] = F[9 + r] - F[9 + r];
                }
                else
                {
                    e.joint_type = EXO_JOINT_NONE;
                    e.f_tip = F;
                }
            }
            kt.tree_.push_back(e);
            const int idx = (int)kt.tree_.size() - 1;
            for (const std::string& c : link.child_links) Add(kt, m, c, idx);
        }
    };
    Rec::Add(*this, *model_, model_->root_link, (int)tree_.size() - 1);
    num_model_elements_ = (int)tree_.size();
    // "Set root inertial" (:264-276): the URDF root link's inertial goes to model_tree_[2] (fixed base), [7] (floating) or [4]
    // (planar).  [7] and [4] are the root link itself; with a FIXED base the root link is model_tree_[1], so the first child
    // of the root carries the root's inertial instead of its own -- the CenterOfMass goldens of test_maps.cpp:789-827 depend on it.
    {
        const size_t k = model_->root_joint_type == "floating" ? 7 : (model_->root_joint_type == "planar" ? 4 : 2);
        if (k < tree_.size())
        {
            const UrdfLink& root = model_->GetLink(model_->root_link);
            tree_[k].mass = root.mass;
            memcpy(tree_[k].cog, root.cog, sizeof(root.cog));
        }
    }

    // UpdateModel (:423-432)
    for (int i = 0; i < (int)tree_.size(); ++i) tree_map_[tree_[i].link_name] = i;
    tree_state_.assign(tree_.size(), 0.0);

    num_joints_ = (int)model_joints_names_.size();
    num_controlled_joints_ = (int)controlled_joints_names_.size();
    if (num_controlled_joints_ < 1) ThrowPretty("No update joints specified!");
    controlled_joints_.assign(num_controlled_joints_, -1);
    controlled_base_type_ = BaseType::FIXED;
    for (int i = 0; i < (int)tree_.size(); ++i)
    {
        KinematicElement& joint = tree_[i];
        joint.is_robot_link = true;
        joint.control_id = IsControlled(joint);
        joint.is_controlled = joint.control_id >= 0;
        model_joints_map_[joint.joint_name] = i;
        model_link_names_.push_back(joint.link_name);
        if (joint.is_controlled)
        {
            controlled_joints_[joint.control_id] = i;
            controlled_joints_map_[joint.joint_name] = i;
            controlled_link_names_.push_back(joint.link_name);
            if (joint.joint_name.find(root_joint_name_) != std::string::npos) controlled_base_type_ = model_base_type_;
        }
    }
    tree_[0].is_robot_link = false;
    for (int c = 0; c < num_controlled_joints_; ++c)
        if (controlled_joints_[c] < 0) ThrowPretty("Controlled joint '" + controlled_joints_names_[c] + "' is not part of the kinematic tree");

    lower_.assign(num_controlled_joints_, 0.0);
    upper_.assign(num_controlled_joints_, 0.0);
    velocity_limits_.assign(num_controlled_joints_, 0.0);
    ResetJointLimits();

    // mimic joints (:342-356)
    for (const auto& kv : model_->links)
    {
        const UrdfLink& l = kv.second;
        if (l.parent_joint.empty()) continue;
        const UrdfJoint& j = *model_->GetJoint(l.parent_joint);
        if (!j.has_mimic) continue;
        auto target = model_joints_map_.find(j.mimic_joint);
        if (target == model_joints_map_.end()) ThrowPretty("Mimic joint '" + j.name + "' refers to unknown joint '" + j.mimic_joint + "'");
        KinematicElement& e = tree_[tree_map_.at(l.name)];
        e.is_mimic_joint = true;
        e.mimic_multiplier = j.mimic_multiplier;
        e.mimic_offset = j.mimic_offset;
        e.mimic_element = target->second;
    }
}

int KinematicTree::IsControlled(const KinematicElement& joint) const
{
    for (int i = 0; i < (int)controlled_joints_names_.size(); ++i)
        if (controlled_joints_names_[i] == joint.joint_name) return i;
    return -1;
}

int KinematicTree::IsControlledLink(const std::string& link_name) const
{
    auto it = tree_map_.find(link_name);
    if (it == tree_map_.end()) return -1;
    int e = it->second;
    while (e >= 0)
    {
        if (tree_[e].is_controlled) return tree_[e].control_id;
        e = tree_[e].parent;
    }
    return -1;
}

// kinematic_tree.cpp:1217-1297
void KinematicTree::ResetJointLimits()
{
    for (const std::string& var : model_->GetVariableNames())
    {
        auto it = controlled_joints_map_.find(var);
        if (it == controlled_joints_map_.end()) continue;
        KinematicElement& e = tree_[it->second];
        const RobotModel::Bounds b = model_->GetVariableBounds(var);
        if (b.position_bounded)
            e.joint_limits = {b.min_position, b.max_position};
        else
            e.joint_limits = {-inf, inf};
        e.velocity_limit = b.velocity_bounded ? b.max_velocity : inf;
    }
    if (controlled_base_type_ == BaseType::FLOATING)
    {
        for (int i = 0; i < 3; ++i) tree_[controlled_joints_[i]].joint_limits = {-inf, inf};
        for (int i = 3; i < 6; ++i) tree_[controlled_joints_[i]].joint_limits = {-pi, pi};
    }
    else if (controlled_base_type_ == BaseType::PLANAR)
    {
        for (int i = 0; i < 2; ++i) tree_[controlled_joints_[i]].joint_limits = {-inf, inf};
        tree_[controlled_joints_[2]].joint_limits = {-pi, pi};
    }
    UpdateJointLimits();
}

void KinematicTree::UpdateJointLimits()
{
    random_state_distributions_.clear();
    for (int i = 0; i < num_controlled_joints_; ++i)
    {
        KinematicElement& e = tree_[controlled_joints_[i]];
        if (e.joint_limits.size() != 2) e.joint_limits = {-inf, inf};
        lower_[i] = e.joint_limits[0];
        upper_[i] = e.joint_limits[1];
        velocity_limits_[i] = e.velocity_limit;
        random_state_distributions_.push_back(std::uniform_real_distribution<double>(lower_[i], upper_[i]));
    }
    if (backend_) CheckRc(exo_tree_set_joint_limits(backend_, lower_.data(), upper_.data()));
}

void KinematicTree::SetJointLimitsLower(const std::vector<double>& lower_in)
{
    if ((int)lower_in.size() != num_controlled_joints_) ThrowPretty("Got " + std::to_string(lower_in.size()) + " but " + std::to_string(num_controlled_joints_) + " expected.");
    for (int i = 0; i < num_controlled_joints_; ++i) tree_[controlled_joints_[i]].joint_limits[0] = lower_in[i];
    UpdateJointLimits();
}

void KinematicTree::SetJointLimitsUpper(const std::vector<double>& upper_in)
{
    if ((int)upper_in.size() != num_controlled_joints_) ThrowPretty("Got " + std::to_string(upper_in.size()) + " but " + std::to_string(num_controlled_joints_) + " expected.");
    for (int i = 0; i < num_controlled_joints_; ++i) tree_[controlled_joints_[i]].joint_limits[1] = upper_in[i];
    UpdateJointLimits();
}

std::vector<double> KinematicTree::GetJointLimits() const
{
    std::vector<double> out(2 * (size_t)num_controlled_joints_);
    for (int i = 0; i < num_controlled_joints_; ++i)
    {
        out[2 * i] = lower_[i];
        out[2 * i + 1] = upper_[i];
    }
    return out;
}

std::vector<double> KinematicTree::GetRandomControlledState()
{
    std::vector<double> q(num_controlled_joints_);
    for (int i = 0; i < num_controlled_joints_; ++i) q[i] = random_state_distributions_[i](generator_);
    return q;
}

const std::string& KinematicTree::GetRootFrameName() const
{
    return tree_[0].link_name;
}

const std::string& KinematicTree::GetRootJointName() const
{
    return root_joint_name_;
}

bool KinematicTree::HasModelLink(const std::string& link) const
{
    return std::find(model_link_names_.begin(), model_link_names_.end(), link) != model_link_names_.end();
}

int KinematicTree::GetElementIndex(const std::string& link_name) const
{
    auto it = tree_map_.find(link_name);
    return it == tree_map_.end() ? -1 : it->second;
}

int KinematicTree::ResolveFrame(const std::string& link_name, const char* what) const
{
    if (link_name.empty()) return 0;  // root
    auto it = tree_map_.find(link_name);
    if (it == tree_map_.end()) ThrowPretty(std::string("No ") + what + " link exists named '" + link_name + "'");
    return it->second;
}

// Model state: one value per model joint name (GetModelState / SetModelState, kinematic_tree.cpp:1090-1176)
std::vector<double> KinematicTree::GetModelState() const
{
    std::vector<double> out(model_joints_names_.size(), 0.0);
    for (size_t i = 0; i < model_joints_names_.size(); ++i)
    {
        auto it = model_joints_map_.find(model_joints_names_[i]);
        if (it != model_joints_map_.end()) out[i] = tree_state_[it->second];
    }
    return out;
}

std::map<std::string, double> KinematicTree::GetModelStateMap() const
{
    std::map<std::string, double> out;
    const std::vector<double> s = GetModelState();
    for (size_t i = 0; i < model_joints_names_.size(); ++i) out[model_joints_names_[i]] = s[i];
    return out;
}

void KinematicTree::SetModelState(const std::vector<double>& x)
{
    if (x.size() != model_joints_names_.size()) ThrowPretty("Model state vector has wrong size, expected " + std::to_string(model_joints_names_.size()) + " got " + std::to_string(x.size()));
    for (size_t i = 0; i < model_joints_names_.size(); ++i)
    {
        auto it = model_joints_map_.find(model_joints_names_[i]);
        if (it != model_joints_map_.end()) tree_state_[it->second] = x[i];
    }
    PushModelStateToBackend();
}

void KinematicTree::SetModelState(const std::map<std::string, double>& x)
{
    for (const auto& kv : x)
    {
        auto it = model_joints_map_.find(kv.first);
        if (it == model_joints_map_.end()) ThrowPretty("Robot model does not contain joint '" + kv.first + "'");
        tree_state_[it->second] = kv.second;
    }
    PushModelStateToBackend();
}

std::vector<double> KinematicTree::GetControlledState() const
{
    std::vector<double> x(num_controlled_joints_);
    for (int i = 0; i < num_controlled_joints_; ++i) x[i] = tree_state_[controlled_joints_[i]];
    return x;
}

void KinematicTree::PushModelStateToBackend()
{
    if (backend_) CheckRc(exo_tree_set_model_state(backend_, tree_state_.data()));
}

int KinematicTree::AddEnvironmentElement(const std::string& name, const Frame& transform, const std::string& parent, bool trajectory_generated)
{
    int parent_index = 0;
    if (!parent.empty())
    {
        auto it = tree_map_.find(parent);
        if (it == tree_map_.end()) ThrowPretty("Can't find parent '" + parent + "'!");
        parent_index = it->second;
    }
    KinematicElement e;
    e.id = (int)tree_.size();
    e.parent = parent_index;
    e.parent_name = tree_[parent_index].link_name;
    e.link_name = name;
    e.joint_name = "NoName";
    e.joint_type = EXO_JOINT_NONE;
    e.f_tip = transform;
    e.is_trajectory_generated = trajectory_generated;
    if (trajectory_generated) e.generated_slot = num_generated_++;
    tree_.push_back(e);
    tree_map_[name] = (int)tree_.size() - 1;
    tree_state_.push_back(0.0);
    InvalidateBackend();
    return (int)tree_.size() - 1;
}

void KinematicTree::ResetModel()
{
    tree_.resize(num_model_elements_);
    tree_state_.resize(num_model_elements_);
    tree_map_.clear();
    for (int i = 0; i < (int)tree_.size(); ++i) tree_map_[tree_[i].link_name] = i;
    num_generated_ = 0;
    generated_offsets_.clear();
    for (KinematicElement& e : tree_)
    {
        e.is_trajectory_generated = false;
        e.generated_slot = -1;
    }
    InvalidateBackend();
}

int KinematicTree::MakeTrajectoryGenerated(int index)
{
    // Scene::AddTrajectory (scene.cpp:927-935): is_trajectory_generated = true; UpdateTree then takes generated_offset for the
    // element's pose (kinematic_tree.cpp:711-714)
    if (index < 0 || index >= (int)tree_.size()) ThrowPretty("No such element");
    KinematicElement& e = tree_[index];
    if (!e.is_trajectory_generated)
    {
        e.is_trajectory_generated = true;
        e.generated_slot = num_generated_++;
        InvalidateBackend();
    }
    return e.generated_slot;
}

int KinematicTree::AddElement(const std::string& name, const Frame& transform, const std::string& parent)
{
    return AddEnvironmentElement(name, transform, parent, false);
}

void KinematicTree::SetGeneratedOffset(int slot, const Frame& frame)
{
    if (slot < 0 || slot >= num_generated_) ThrowPretty("No such trajectory-driven element");
    if ((int)generated_offsets_.size() < num_generated_) generated_offsets_.resize(num_generated_, IdentityFrame());
    generated_offsets_[slot] = frame;
}

void KinematicTree::PushCurrentGenerated(exo_problem* p)
{
    if (num_generated_ == 0) return;
    if ((int)generated_offsets_.size() < num_generated_) generated_offsets_.resize(num_generated_, IdentityFrame());
    std::vector<double> table(12 * (size_t)num_generated_);
    for (int g = 0; g < num_generated_; ++g) memcpy(&table[12 * (size_t)g], generated_offsets_[g].data(), 12 * sizeof(double));
    CheckRc(exo_problem_set_generated(p, 1, num_generated_, table.data()));
}

exo_tree* KinematicTree::GetBackend()
{
    if (backend_) return backend_;
    const int n = (int)tree_.size();
    std::vector<int32_t> id(n), parent(n), jt(n), cid(n), mim(n), gen(n);
    std::vector<double> origin(3 * (size_t)n), axis(3 * (size_t)n), tip(12 * (size_t)n), mm(n), mo(n);
    for (int i = 0; i < n; ++i)
    {
        const KinematicElement& e = tree_[i];
        id[i] = e.id;
        parent[i] = e.parent;
        jt[i] = e.joint_type;
        cid[i] = e.control_id;
        mim[i] = e.is_mimic_joint ? e.mimic_element : -1;
        gen[i] = e.generated_slot;
        mm[i] = e.mimic_multiplier;
        mo[i] = e.mimic_offset;
        memcpy(&origin[3 * i], e.origin, 3 * sizeof(double));
        memcpy(&axis[3 * i], e.axis, 3 * sizeof(double));
        memcpy(&tip[12 * i], e.f_tip.data(), 12 * sizeof(double));
    }
    exo_tree_desc d;
    memset(&d, 0, sizeof(d));
    d.n_elements = n;
    d.id = id.data();
    d.parent = parent.data();
    d.joint_type = jt.data();
    d.joint_origin = origin.data();
    d.joint_axis = axis.data();
    d.f_tip = tip.data();
    d.control_id = cid.data();
    d.mimic_element = mim.data();
    d.mimic_multiplier = mm.data();
    d.mimic_offset = mo.data();
    d.generated_slot = gen.data();
    d.n_controlled = num_controlled_joints_;
    d.lower = lower_.data();
    d.upper = upper_.data();
    CheckRc(exo_tree_create(&d, &backend_));
    CheckRc(exo_tree_set_model_state(backend_, tree_state_.data()));
    return backend_;
}

void KinematicTree::EnsureFramesProblem()
{
    if (frames_problem_) return;
    if (!frames_requested_) ThrowPretty("RequestFrames must be called before Update");
    const int F = (int)request_.frames.size();
    std::vector<int32_t> ea(F), eb(F);
    std::vector<double> oa(12 * (size_t)F), ob(12 * (size_t)F);
    for (int i = 0; i < F; ++i)
    {
        ea[i] = ResolveFrame(request_.frames[i].frame_A_link_name, "frame_A");
        eb[i] = ResolveFrame(request_.frames[i].frame_B_link_name, "frame_B");
        memcpy(&oa[12 * i], request_.frames[i].frame_A_offset.data(), 12 * sizeof(double));
        memcpy(&ob[12 * i], request_.frames[i].frame_B_offset.data(), 12 * sizeof(double));
    }
    exo_tree* tree = GetBackend();
    CheckRc(exo_problem_create(tree, &frames_problem_));
    exo_task_map_desc d;
    memset(&d, 0, sizeof(d));
    d.kind = EXO_MAP_FRAMES;
    d.n_frames = F;
    d.elem_a = ea.data();
    d.off_a = oa.data();
    d.elem_b = eb.data();
    d.off_b = ob.data();
    d.is_used = 1;
    CheckRc(exo_problem_add_task_map(frames_problem_, &d));
    CheckRc(exo_problem_finalize(frames_problem_));
}

// kinematic_tree.cpp:628-676
std::shared_ptr<KinematicResponse> KinematicTree::RequestFrames(const KinematicsRequest& request)
{
    request_ = request;
    if (request_.flags & KIN_H) ThrowPretty("KIN_H (kinematic Hessian) is outside the GPU hot path");
    const int F = (int)request.frames.size();
    frames_requested_ = true;
    if (frames_problem_) exo_problem_destroy(frames_problem_);
    frames_problem_ = nullptr;
    EnsureFramesProblem();
    solution_.reset(new KinematicResponse());
    solution_->flags = request.flags;
    solution_->x.assign(num_controlled_joints_, 0.0);
    solution_->Phi.assign(F, IdentityFrame());
    if (request.flags & KIN_J) solution_->jacobian.assign(F, std::vector<double>(6 * (size_t)num_controlled_joints_, 0.0));
    return solution_;
}

void KinematicTree::UpdateBatch(const double* x, int64_t n_evals, int layout, double* phi, double* jacobian)
{
    EnsureFramesProblem();
    PushCurrentGenerated(frames_problem_);  // trajectory-driven links sit where the last Scene::Update(x, t) put them
    CheckRc(exo_problem_update_batch_host(frames_problem_, x, n_evals, 1, layout, phi, jacobian));
}

// kinematic_tree.cpp:678-693
void KinematicTree::Update(const std::vector<double>& x)
{
    if ((int)x.size() != num_controlled_joints_) ThrowPretty("Wrong state vector size! Got " + std::to_string(x.size()) + " expected " + std::to_string(num_controlled_joints_));
    for (int i = 0; i < num_controlled_joints_; ++i) tree_state_[controlled_joints_[i]] = x[i];
    if (!frames_requested_) return;  // the reference updates the tree whether or not frames were requested (kinematic_tree.cpp:678-693)
    EnsureFramesProblem();
    const int F = (int)solution_->Phi.size();
    std::vector<double> phi(12 * (size_t)F), jac;
    const bool want_j = (solution_->flags & KIN_J) != 0;
    if (want_j) jac.resize(6 * (size_t)F * num_controlled_joints_);
    // column-major per evaluation == J[N][6F]: rows of frame f are 6f..6f+5
    PushCurrentGenerated(frames_problem_);
    CheckRc(exo_problem_update_batch_host(frames_problem_, x.data(), 1, 1, EXO_LAYOUT_AOS_COLMAJOR, phi.data(), want_j ? jac.data() : nullptr));
    solution_->x = x;
    for (int f = 0; f < F; ++f)
    {
        memcpy(solution_->Phi[f].data(), &phi[12 * f], 12 * sizeof(double));
        if (want_j)
            for (int c = 0; c < num_controlled_joints_; ++c)
                for (int r = 0; r < 6; ++r) solution_->jacobian[f][6 * c + r] = jac[(size_t)c * 6 * F + 6 * f + r];
    }
}

static void OneFrame(KinematicTree& kt, exo_tree* backend, int n, int ea, const Frame& oa, int eb, const Frame& ob, const std::vector<double>& x, Frame* phi,
                     std::vector<double>* jac)
{
    exo_problem* p = nullptr;
    CheckRc(exo_problem_create(backend, &p));
    exo_task_map_desc d;
    memset(&d, 0, sizeof(d));
    d.kind = EXO_MAP_FRAMES;
    d.n_frames = 1;
    const int32_t a = ea, b = eb;
    d.elem_a = &a;
    d.elem_b = &b;
    d.off_a = oa.data();
    d.off_b = ob.data();
    d.is_used = 1;
    int rc = exo_problem_add_task_map(p, &d);
    if (rc == EXO_OK) rc = exo_problem_finalize(p);
    std::vector<double> j(6 * (size_t)n);
    Frame f;
    if (rc == EXO_OK && kt.GetNumGenerated() > 0)
    {
        try
        {
            kt.PushCurrentGenerated(p);
        }
        catch (...)
        {
            exo_problem_destroy(p);
            throw;
        }
    }
    if (rc == EXO_OK) rc = exo_problem_update_batch_host(p, x.data(), 1, 1, EXO_LAYOUT_AOS_COLMAJOR, f.data(), jac ? j.data() : nullptr);
    exo_problem_destroy(p);
    CheckRc(rc);
    if (phi) *phi = f;
    if (jac) *jac = j;
}

Frame KinematicTree::FK(const std::string& element_A, const Frame& offset_a, const std::string& element_B, const Frame& offset_b)
{
    auto A = tree_map_.find(element_A.empty() ? tree_[0].link_name : element_A);
    if (A == tree_map_.end()) ThrowPretty("Can't find link '" + element_A + "'!");
    auto B = tree_map_.find(element_B.empty() ? tree_[0].link_name : element_B);
    if (B == tree_map_.end()) ThrowPretty("Can't find link '" + element_B + "'!");
    Frame out;
    OneFrame(*this, GetBackend(), num_controlled_joints_, A->second, offset_a, B->second, offset_b, GetControlledState(), &out, nullptr);
    return out;
}

std::vector<double> KinematicTree::Jacobian(const std::string& element_A, const Frame& offset_a, const std::string& element_B, const Frame& offset_b)
{
    auto A = tree_map_.find(element_A.empty() ? tree_[0].link_name : element_A);
    if (A == tree_map_.end()) ThrowPretty("Can't find link '" + element_A + "'!");
    auto B = tree_map_.find(element_B.empty() ? tree_[0].link_name : element_B);
    if (B == tree_map_.end()) ThrowPretty("Can't find link '" + element_B + "'!");
    std::vector<double> out;
    OneFrame(*this, GetBackend(), num_controlled_joints_, A->second, offset_a, B->second, offset_b, GetControlledState(), nullptr, &out);
    return out;
}

}  // namespace exotica
