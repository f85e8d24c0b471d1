// compiler.h -- host-side compiler: flattened KinematicTree + frame requests + task maps -> tree program.
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

#include "program.h"

namespace exo
{
struct Frame
{
    double M[9];
    double p[3];
};

Frame frame_identity();
Frame frame_mul(const Frame& a, const Frame& b);
Frame frame_inverse(const Frame& a);
Frame frame_from12(const double* f);
bool frame_is_identity(const Frame& f);
bool rot_is_identity(const Frame& f);

// Flattened KinematicTree::tree_ (see exo_tree_desc in include/exotica_b200.h)
struct TreeModel
{
    int n = 0;
    std::vector<int32_t> id, parent, joint_type, control_id, mimic_elem, generated;
    std::vector<double> origin, axis, tip, mimic_mult, mimic_off;
    int n_ctrl = 0;
    std::vector<double> lower, upper;
    std::vector<double> tree_state;  // by element index
    int n_generated = 0;
};

struct MapSpec
{
    int kind = 0;
    int n_frames = 0;
    std::vector<int32_t> elem_a, elem_b;
    std::vector<double> off_a, off_b;  // [n_frames][12]
    int rot_type = 1;
    double safe_percentage = 0.0;
    double precision = 1.0;
    std::vector<double> radius;
    std::vector<int32_t> group;
    int n_groups = 0;
    std::vector<int32_t> joint_map;  // JointPose
    std::vector<double> joint_ref;
    std::vector<double> axis, direction;  // EffAxisAlignment: [n_frames][3], unit vectors
    std::vector<double> mass;             // CenterOfMass: [n_frames]
    bool enable_z = true;
    bool is_used = true;
    // assigned by the compiler (TaskMap::start / length / start_jacobian / length_jacobian)
    int start = 0, length = 0, start_jacobian = 0, length_jacobian = 0;
};

struct Program
{
    ProgramHeader hdr;
    std::vector<Op> ops;
    std::vector<int32_t> ints;
    std::vector<double> dbls;
    // serialised image staged into shared memory by the kernel:
    // [ProgramHeader][Op x n_ops][int32 x n_int (padded to 8 B)][double x n_dbl]
    std::vector<unsigned char> image;
    int dbl_offset = 0;  // byte offset of the double pool inside the image
    int n_joint_ops = 0, n_emit_ops = 0;
};

int rotation_length(int rot_type);
// Assigns start/length/... to the maps (planning_problem.cpp:227-238).  Returns false + err on bad sizes.
bool assign_indexing(const TreeModel& tree, std::vector<MapSpec>& maps, int& length_phi, int& length_jac, std::string& err);
bool compile_program(const TreeModel& tree, std::vector<MapSpec>& maps, Program& out, std::string& err, int mode = COMPILE_AUTO);

}  // namespace exo
