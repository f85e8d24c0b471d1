// robot_model.cpp -- URDF / SRDF ingestion with the ordering semantics the reference inherits from urdfdom,
// kdl_parser and MoveIt's RobotModel (none of which are vendored in the reference):
//   * urdfdom keeps joints in a std::map, so a link's children are ordered by ascending JOINT name;
//   * MoveIt orders joints (and their variables) in DFS pre-order over those children;
//   * a JointModelGroup lists its variables in model joint order, whatever the SRDF listing order;
//   * variable bounds prefer the safety_controller soft limits clipped by the hard limits.
// Consumed by KinematicTree::Instantiate/BuildTree (exotica_core/src/kinematic_tree.cpp:101-421).
#include <math.h>

#include <algorithm>
#include <fstream>
#include <limits>
#include <sstream>

#include "exotica_host.h"

namespace exotica
{
// ---- frames / parsing helpers ---------------------------------------------------------------------------
Frame IdentityFrame()
{
    return Frame{{1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0}};
}

Frame FrameMul(const Frame& a, const Frame& b)
{
    Frame r;
    for (int i = 0; i < 3; ++i)
    {
        for (int j = 0; j < 3; ++j) r[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
        r[9 + i] = (a[3 * i] * b[9] + a[3 * i + 1] * b[10] + a[3 * i + 2] * b[11]) + a[9 + i];
    }
    return r;
}

Frame FrameInverse(const Frame& a)
{
    Frame r;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) r[3 * i + j] = a[3 * j + i];
    for (int i = 0; i < 3; ++i) r[9 + i] = -(a[i] * a[9] + a[3 + i] * a[10] + a[6 + i] * a[11]);
    return r;
}

Frame RotationQuaternion(double x, double y, double z, double w)
{
    Frame f = IdentityFrame();
    const double x2 = x * x, y2 = y * y, z2 = z * z, w2 = w * w;
    f[0] = w2 + x2 - y2 - z2;
    f[1] = 2 * x * y - 2 * w * z;
    f[2] = 2 * x * z + 2 * w * y;
    f[3] = 2 * x * y + 2 * w * z;
    f[4] = w2 - x2 + y2 - z2;
    f[5] = 2 * y * z - 2 * w * x;
    f[6] = 2 * x * z - 2 * w * y;
    f[7] = 2 * y * z + 2 * w * x;
    f[8] = w2 - x2 - y2 + z2;
    return f;
}

Frame RotationRPY(double roll, double pitch, double yaw)
{
    Frame f = IdentityFrame();
    const double ca1 = cos(yaw), sa1 = sin(yaw), cb1 = cos(pitch), sb1 = sin(pitch), cc1 = cos(roll), sc1 = sin(roll);
    f[0] = ca1 * cb1;
    f[1] = ca1 * sb1 * sc1 - sa1 * cc1;
    f[2] = ca1 * sb1 * cc1 + sa1 * sc1;
    f[3] = sa1 * cb1;
    f[4] = sa1 * sb1 * sc1 + ca1 * cc1;
    f[5] = sa1 * sb1 * cc1 - ca1 * sc1;
    f[6] = -sb1;
    f[7] = cb1 * sc1;
    f[8] = cb1 * cc1;
    return f;
}

Frame GetFrame(const std::vector<double>& val)
{
    Frame f = IdentityFrame();
    switch (val.size())
    {
        case 7:
        {
            const double norm = sqrt(val[3] * val[3] + val[4] * val[4] + val[5] * val[5] + val[6] * val[6]);
            if (norm <= 0.0) ThrowPretty("Invalid quaternion!");
            f = RotationQuaternion(val[3] / norm, val[4] / norm, val[5] / norm, val[6] / norm);
            break;
        }
        case 6:
            f = RotationRPY(val[3], val[4], val[5]);
            break;
        case 3:
            break;
        default:
            ThrowPretty("Eigen vector has incorrect length! (" + std::to_string(val.size()) + ")");
    }
    f[9] = val[0];
    f[10] = val[1];
    f[11] = val[2];
    return f;
}

std::vector<double> ParseVector(const std::string& value)
{
    std::vector<double> out;
    std::istringstream ss(value);
    std::string tok;
    while (ss >> tok)
    {
        try
        {
            size_t used = 0;
            const double v = std::stod(tok, &used);
            if (used != tok.size()) throw std::invalid_argument(tok);
            out.push_back(v);
        }
        catch (const std::exception&)
        {
            ThrowPretty("Can't parse value '" + tok + "' in vector '" + value + "'");
        }
    }
    return out;
}

int GetRotationTypeLength(const RotationType& type)
{
    static const int types[] = {4, 3, 3, 3, 3, 9};
    return types[static_cast<int>(type)];
}

RotationType GetRotationTypeFromString(const std::string& rotation_type)
{
    if (rotation_type == "Quaternion") return RotationType::QUATERNION;
    if (rotation_type == "RPY") return RotationType::RPY;
    if (rotation_type == "ZYX") return RotationType::ZYX;
    if (rotation_type == "ZYZ") return RotationType::ZYZ;
    if (rotation_type == "AngleAxis") return RotationType::ANGLE_AXIS;
    if (rotation_type == "Matrix") return RotationType::MATRIX;
    ThrowPretty("Unsupported rotation type '" + rotation_type + "'");
}

namespace
{
void Parse3(const xml::Element* e, const char* attr, double* out)
{
    if (!e) return;
    const std::string* v = e->Attribute(attr);
    if (!v) return;
    const std::vector<double> d = ParseVector(*v);
    if (d.size() != 3) ThrowPretty(std::string("URDF: attribute '") + attr + "' needs 3 numbers, got '" + *v + "'");
    out[0] = d[0];
    out[1] = d[1];
    out[2] = d[2];
}

double ParseDouble(const xml::Element* e, const char* attr, double fallback)
{
    if (!e) return fallback;
    const std::string* v = e->Attribute(attr);
    if (!v) return fallback;
    const std::vector<double> d = ParseVector(*v);
    if (d.size() != 1) ThrowPretty(std::string("URDF: attribute '") + attr + "' needs one number");
    return d[0];
}

// urdf::Rotation::setFromRPY + normalize, then KDL::Rotation::Quaternion (kdl_parser toKdl(urdf::Pose))
Frame FrameFromOrigin(const double* xyz, const double* rpy)
{
    const double phi = rpy[0] / 2.0, the = rpy[1] / 2.0, psi = rpy[2] / 2.0;
    double x = sin(phi) * cos(the) * cos(psi) - cos(phi) * sin(the) * sin(psi);
    double y = cos(phi) * sin(the) * cos(psi) + sin(phi) * cos(the) * sin(psi);
    double z = cos(phi) * cos(the) * sin(psi) - sin(phi) * sin(the) * cos(psi);
    double w = cos(phi) * cos(the) * cos(psi) + sin(phi) * sin(the) * sin(psi);
    const double s = sqrt(x * x + y * y + z * z + w * w);
    if (s == 0.0)
    {
        x = y = z = 0.0;
        w = 1.0;
    }
    else
    {
        x /= s;
        y /= s;
        z /= s;
        w /= s;
    }
    Frame f = RotationQuaternion(x, y, z, w);
    f[9] = xyz[0];
    f[10] = xyz[1];
    f[11] = xyz[2];
    return f;
}
}  // namespace

Frame UrdfOriginToFrame(const double* xyz, const double* rpy)
{
    return FrameFromOrigin(xyz, rpy);
}

RobotModel::RobotModel(const std::string& urdf_xml, const std::string& srdf_xml)
{
    std::shared_ptr<xml::Element> robot;
    try
    {
        robot = xml::Parse(urdf_xml);
    }
    catch (const std::exception& e)
    {
        ThrowPretty(std::string("Can't load URDF model! ") + e.what());
    }
    if (robot->name != "robot") ThrowPretty("Can't load URDF model! Root element is <" + robot->name + ">");
    name = robot->AttributeOr("name", "");
    for (const auto& le : robot->Children("link"))
    {
        UrdfLink l;
        const std::string* n = le->Attribute("name");
        if (!n) ThrowPretty("URDF: link without a name");
        l.name = *n;
        if (auto in = le->FirstChild("inertial"))
        {
            l.mass = ParseDouble(in->FirstChild("mass").get(), "value", 0.0);
            Parse3(in->FirstChild("origin").get(), "xyz", l.cog);
        }
        for (const auto& ce : le->Children("collision"))
        {
            double xyz[3] = {0, 0, 0}, rpy[3] = {0, 0, 0};
            Parse3(ce->FirstChild("origin").get(), "xyz", xyz);
            Parse3(ce->FirstChild("origin").get(), "rpy", rpy);
            l.collision_origins.push_back(FrameFromOrigin(xyz, rpy));
        }
        if (links.count(l.name)) ThrowPretty("URDF: link '" + l.name + "' is not unique");
        links[l.name] = l;
    }
    for (const auto& je : robot->Children("joint"))
    {
        UrdfJoint j;
        const std::string* n = je->Attribute("name");
        const std::string* t = je->Attribute("type");
        if (!n || !t) ThrowPretty("URDF: joint without name or type");
        j.name = *n;
        j.type = *t;
        Parse3(je->FirstChild("origin").get(), "xyz", j.xyz);
        Parse3(je->FirstChild("origin").get(), "rpy", j.rpy);
        Parse3(je->FirstChild("axis").get(), "xyz", j.axis);
        auto pe = je->FirstChild("parent");
        auto ce = je->FirstChild("child");
        if (!pe || !ce || !pe->Attribute("link") || !ce->Attribute("link")) ThrowPretty("URDF: joint '" + j.name + "' needs parent and child links");
        j.parent_link = *pe->Attribute("link");
        j.child_link = *ce->Attribute("link");
        if (auto lim = je->FirstChild("limit"))
        {
            j.has_limits = true;
            j.lower = ParseDouble(lim.get(), "lower", 0.0);
            j.upper = ParseDouble(lim.get(), "upper", 0.0);
            j.velocity = ParseDouble(lim.get(), "velocity", 0.0);
        }
        if (auto saf = je->FirstChild("safety_controller"))
        {
            j.has_safety = true;
            j.soft_lower = ParseDouble(saf.get(), "soft_lower_limit", 0.0);
            j.soft_upper = ParseDouble(saf.get(), "soft_upper_limit", 0.0);
        }
        if (auto mim = je->FirstChild("mimic"))
        {
            j.has_mimic = true;
            j.mimic_joint = mim->AttributeOr("joint", "");
            j.mimic_multiplier = ParseDouble(mim.get(), "multiplier", 1.0);
            j.mimic_offset = ParseDouble(mim.get(), "offset", 0.0);
        }
        if (joints.count(j.name)) ThrowPretty("URDF: joint '" + j.name + "' is not unique");
        joints[j.name] = j;
    }
    // urdfdom initTree: iterate the joint map (ascending name)
    for (const auto& kv : joints)
    {
        const UrdfJoint& j = kv.second;
        if (!links.count(j.child_link)) ThrowPretty("URDF: child link '" + j.child_link + "' of joint '" + j.name + "' not found");
        if (!links.count(j.parent_link)) ThrowPretty("URDF: parent link '" + j.parent_link + "' of joint '" + j.name + "' not found");
        if (!links[j.child_link].parent_joint.empty()) ThrowPretty("URDF: link '" + j.child_link + "' has two parent joints");
        links[j.child_link].parent_joint = j.name;
        links[j.parent_link].child_links.push_back(j.child_link);
    }
    int roots = 0;
    for (const auto& kv : links)
        if (kv.second.parent_joint.empty())
        {
            root_link = kv.first;
            ++roots;
        }
    if (roots != 1) ThrowPretty("Can't load URDF model! Expected exactly one root link, found " + std::to_string(roots));

    // SRDF
    std::vector<std::array<std::string, 4>> virtual_joints;
    if (!srdf_xml.empty())
    {
        std::shared_ptr<xml::Element> srdf;
        try
        {
            srdf = xml::Parse(srdf_xml);
        }
        catch (const std::exception& e)
        {
            ThrowPretty(std::string("Can't load SRDF model! ") + e.what());
        }
        for (const auto& ve : srdf->Children("virtual_joint"))
            virtual_joints.push_back({ve->AttributeOr("name", ""), ve->AttributeOr("type", "fixed"), ve->AttributeOr("parent_frame", ""), ve->AttributeOr("child_link", "")});
        for (const auto& ge : srdf->Children("group"))
        {
            SrdfGroup g;
            g.name = ge->AttributeOr("name", "");
            for (const auto& c : ge->children)
            {
                if (c->name == "chain") g.chains.emplace_back(c->AttributeOr("base_link", ""), c->AttributeOr("tip_link", ""));
                else if (c->name == "joint") g.joints.push_back(c->AttributeOr("name", ""));
                else if (c->name == "link") g.links.push_back(c->AttributeOr("name", ""));
                else if (c->name == "group") g.subgroups.push_back(c->AttributeOr("name", ""));
            }
            groups_[g.name] = g;
            group_order_.push_back(g.name);
        }
    }
    root_joint_name = "ASSUMED_FIXED_ROOT_JOINT";
    root_joint_type = "fixed";
    for (const auto& vj : virtual_joints)
        if (vj[3] == root_link)
        {
            root_joint_name = vj[0];
            root_joint_type = vj[1];
        }
    // kinematic_tree.cpp:178-191: world frame = parent_frame of the virtual joint named like the root joint
    for (const auto& vj : virtual_joints)
        if (vj[0] == root_joint_name) world_frame_name = vj[2];
    if (world_frame_name.empty()) world_frame_name = "world_frame";

    BuildRecursive(root_link);
    for (size_t i = 0; i < model_joints.size(); ++i)
    {
        joint_index_[model_joints[i].name] = (int)i;
        link_parent_joint_[model_joints[i].child_link] = model_joints[i].name;
    }
}

void RobotModel::BuildRecursive(const std::string& link_name)
{
    const UrdfLink& link = links.at(link_name);
    ModelJoint mj;
    mj.child_link = link_name;
    if (link.parent_joint.empty())
    {
        mj.name = root_joint_name;
        mj.type = root_joint_type;
        if (mj.type == "floating")
            for (const char* s : {"/trans_x", "/trans_y", "/trans_z", "/rot_x", "/rot_y", "/rot_z", "/rot_w"}) mj.variables.push_back(mj.name + s);
        else if (mj.type == "planar")
            for (const char* s : {"/x", "/y", "/theta"}) mj.variables.push_back(mj.name + s);
    }
    else
    {
        const UrdfJoint& j = joints.at(link.parent_joint);
        mj.name = j.name;
        mj.parent_link = j.parent_link;
        mj.mimic = j.has_mimic;
        if (j.type == "revolute" || j.type == "continuous" || j.type == "prismatic")
        {
            mj.type = j.type;
            mj.variables.push_back(j.name);
        }
        else if (j.type == "floating" || j.type == "planar")
            mj.type = j.type;  // kdl_parser treats these as fixed; MoveIt would add variables (not used by the reference's fixtures)
        else
            mj.type = "fixed";
    }
    model_joints.push_back(mj);
    for (const std::string& c : link.child_links) BuildRecursive(c);
}

const UrdfLink& RobotModel::GetLink(const std::string& n) const
{
    auto it = links.find(n);
    if (it == links.end()) ThrowPretty("Link '" + n + "' not found");
    return it->second;
}

const UrdfJoint* RobotModel::GetJoint(const std::string& n) const
{
    auto it = joints.find(n);
    return it == joints.end() ? nullptr : &it->second;
}

std::vector<std::string> RobotModel::GetGroupNames() const
{
    return group_order_;
}

std::vector<std::string> RobotModel::GetVariableNames() const
{
    std::vector<std::string> out;
    for (const ModelJoint& j : model_joints) out.insert(out.end(), j.variables.begin(), j.variables.end());
    return out;
}

std::vector<std::string> RobotModel::GetActiveJointNames() const
{
    std::vector<std::string> out;
    for (const ModelJoint& j : model_joints)
        if (!j.variables.empty() && !j.mimic) out.push_back(j.name);
    return out;
}

void RobotModel::CollectGroupJoints(const std::string& group, std::vector<std::string>& seen, std::vector<std::string>& out) const
{
    if (std::find(seen.begin(), seen.end(), group) != seen.end()) return;
    seen.push_back(group);
    auto git = groups_.find(group);
    if (git == groups_.end()) return;
    const SrdfGroup& g = git->second;
    auto parent_link_of = [&](const std::string& link) -> std::string {
        auto it = link_parent_joint_.find(link);
        if (it == link_parent_joint_.end()) return std::string();
        return model_joints[joint_index_.at(it->second)].parent_link;
    };
    for (const auto& ch : g.chains)
    {
        const std::string &base = ch.first, &tip = ch.second;
        if (!links.count(base) || !links.count(tip)) continue;
        // RobotModel::addJointModelGroup: from the tip up the tree until the base link (whose parent joint is NOT part of the chain)
        std::vector<std::string> cj, tip_path;
        bool found = false;
        std::string lm = tip;
        while (true)
        {
            tip_path.push_back(lm);
            if (lm == base)
            {
                found = true;
                break;
            }
            cj.push_back(link_parent_joint_.at(lm));
            if (lm == root_link) break;
            lm = parent_link_of(lm);
        }
        if (!found)
        {
            // base is not an ancestor of tip: go up from base until a link on the tip path is met
            std::vector<std::string> cj2;
            lm = base;
            std::string meet;
            while (true)
            {
                if (std::find(tip_path.begin(), tip_path.end(), lm) != tip_path.end())
                {
                    meet = lm;
                    break;
                }
                cj2.push_back(link_parent_joint_.at(lm));
                if (lm == root_link) break;
                lm = parent_link_of(lm);
            }
            if (!meet.empty())
            {
                cj.clear();
                for (const std::string& l : tip_path)
                {
                    if (l == meet) break;
                    cj.push_back(link_parent_joint_.at(l));
                }
                cj.insert(cj.end(), cj2.begin(), cj2.end());
                found = true;
            }
        }
        if (found) out.insert(out.end(), cj.begin(), cj.end());
    }
    for (const std::string& j : g.joints)
        if (joint_index_.count(j)) out.push_back(j);
    for (const std::string& l : g.links)
        if (link_parent_joint_.count(l)) out.push_back(link_parent_joint_.at(l));
    for (const std::string& sg : g.subgroups) CollectGroupJoints(sg, seen, out);
}

std::vector<std::string> RobotModel::GetGroupVariableNames(const std::string& group) const
{
    std::vector<std::string> seen, names;
    CollectGroupJoints(group, seen, names);
    std::vector<int> idx;
    for (const std::string& n : names) idx.push_back(joint_index_.at(n));
    std::sort(idx.begin(), idx.end());
    idx.erase(std::unique(idx.begin(), idx.end()), idx.end());
    std::vector<std::string> out;
    for (int i : idx) out.insert(out.end(), model_joints[i].variables.begin(), model_joints[i].variables.end());
    return out;
}

RobotModel::Bounds RobotModel::GetVariableBounds(const std::string& variable) const
{
    Bounds b;
    const UrdfJoint* j = GetJoint(variable);
    if (!j) return b;
    if (j->type == "continuous")
    {
        // RevoluteJointModel::setContinuous: [-pi, pi], position_bounded_ = false
        b.min_position = -pi;
        b.max_position = pi;
    }
    else if (j->has_safety)
    {
        b.position_bounded = true;
        b.min_position = j->soft_lower;
        b.max_position = j->soft_upper;
        if (j->has_limits)
        {
            if (j->lower > b.min_position) b.min_position = j->lower;
            if (j->upper < b.max_position) b.max_position = j->upper;
        }
    }
    else if (j->has_limits)
    {
        b.position_bounded = true;
        b.min_position = j->lower;
        b.max_position = j->upper;
    }
    if (j->has_limits)
    {
        b.max_velocity = fabs(j->velocity);
        b.velocity_bounded = b.max_velocity > std::numeric_limits<double>::epsilon();
    }
    return b;
}

}  // namespace exotica
