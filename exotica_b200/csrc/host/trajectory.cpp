// trajectory.cpp -- exotica::Trajectory (exotica_core/src/trajectory.cpp:47-141): piecewise Cartesian trajectories that drive
// scene objects over time (Scene::AddTrajectory, scene.cpp:917-935; Scene::UpdateTrajectoryGenerators, scene.cpp:329-335).
//
// The reference builds a KDL::Trajectory_Composite of, per pair of consecutive rows [time, frame]:
//   * KDL::Trajectory_Stationary(dt, f1)                                      if KDL::Equal(f1, f2, 1e-6)
//   * KDL::Trajectory_Segment(Path_Line(f1, f2, RotationalInterpolation_SingleAxis, radius), VelocityProfile_Spline, dt)   otherwise
// orocos_kdl (1.4, not vendored in the reference) is restated here from its published algorithm:
//   Path_Line: translation along the unit vector start -> end, rotation about the single axis of start^-1 end; the path length is
//     max(distance, radius * angle), the shorter of the two motions is scaled so that both end together;
//   VelocityProfile_Spline::SetProfileDuration(0, L, dt): the cubic s(t) = L (3 (t/dt)^2 - 2 (t/dt)^3) (zero end velocities);
//   Trajectory_Composite::Pos(t): the segment whose cumulative end time exceeds t (t < 0: the start; beyond the end: the end).
// Host-side only: the GPU path consumes the sampled frames (one table row per timestep, exo_problem_set_generated).
#include <cmath>
#include <sstream>

#include "exotica_host.h"

namespace exotica
{
namespace
{
const double kPi = 3.14159265358979323846;

// KDL::Rotation::GetRotAngle(axis, eps = 1e-6); R row-major
double RotAngleAxis(const double* d, double axis[3])
{
    const double epsilon = 1e-6, epsilon2 = 1e-5;
    if ((std::fabs(d[1] - d[3]) < epsilon) && (std::fabs(d[2] - d[6]) < epsilon) && (std::fabs(d[5] - d[7]) < epsilon))
    {
        if ((std::fabs(d[1] + d[3]) < epsilon2) && (std::fabs(d[2] + d[6]) < epsilon2) && (std::fabs(d[5] + d[7]) < epsilon2) &&
            (std::fabs(d[0] + d[4] + d[8] - 3) < epsilon2))
        {
            axis[0] = 0.0; axis[1] = 0.0; axis[2] = 1.0;
            return 0.0;
        }
        const double xx = (d[0] + 1) / 2, yy = (d[4] + 1) / 2, zz = (d[8] + 1) / 2;
        const double xy = (d[1] + d[3]) / 4, xz = (d[2] + d[6]) / 4, yz = (d[5] + d[7]) / 4;
        double x, y, z;
        if ((xx > yy) && (xx > zz))
        {
            x = std::sqrt(xx); y = xy / x; z = xz / x;
        }
        else if (yy > zz)
        {
            y = std::sqrt(yy); x = xy / y; z = yz / y;
        }
        else
        {
            z = std::sqrt(zz); x = xz / z; y = yz / z;
        }
        axis[0] = x; axis[1] = y; axis[2] = z;
        return kPi;
    }
    double f = (d[0] + d[4] + d[8] - 1) / 2;
    f = std::min(1.0, std::max(-1.0, f));
    double x = d[7] - d[5], y = d[2] - d[6], z = d[3] - d[1];
    const double n = std::sqrt(x * x + y * y + z * z);
    axis[0] = x / n; axis[1] = y / n; axis[2] = z / n;
    return std::acos(f);
}

// KDL::Rotation::Rot2(unit axis, angle)
void Rot2(const double v[3], double angle, double R[9])
{
    const double ct = std::cos(angle), st = std::sin(angle), vt = 1 - ct;
    R[0] = ct + vt * v[0] * v[0];         R[1] = -v[2] * st + vt * v[0] * v[1]; R[2] = v[1] * st + vt * v[0] * v[2];
    R[3] = v[2] * st + vt * v[1] * v[0];  R[4] = ct + vt * v[1] * v[1];         R[5] = -v[0] * st + vt * v[1] * v[2];
    R[6] = -v[1] * st + vt * v[2] * v[0]; R[7] = v[0] * st + vt * v[2] * v[1];  R[8] = ct + vt * v[2] * v[2];
}

bool FramesEqual(const Frame& a, const Frame& b, double eps)
{
    // KDL::Equal(Frame, Frame, eps): every entry of p and M within eps
    for (int i = 0; i < 12; ++i)
        if (std::fabs(a[i] - b[i]) > eps) return false;
    return true;
}
}  // namespace

Trajectory::Trajectory(const std::string& data)
{
    // "radius\n n m\n rows..." (trajectory.cpp:51-68)
    std::istringstream ss(data);
    int n = 0, m = 0;
    if (!(ss >> radius_ >> n >> m)) ThrowPretty("Invalid trajectory data!");
    std::vector<std::vector<double>> rows(n, std::vector<double>(m, 0.0));
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < m; ++j)
            if (!(ss >> rows[i][j])) ThrowPretty("Invalid trajectory data!");
    ConstructFromData(rows, radius_);
}

Trajectory::Trajectory(const std::vector<std::vector<double>>& data, double radius)
{
    ConstructFromData(data, radius);
}

void Trajectory::ConstructFromData(const std::vector<std::vector<double>>& data, double radius)
{
    const size_t cols = data.empty() ? 0 : data[0].size();
    if (!(cols == 4 || cols == 7 || cols == 8) || data.size() < 2)
        ThrowPretty("Invalid trajectory data size!\nNeeds to contain 4, 7, or 8 columns and at least 2 rows.");
    segments_.clear();
    for (size_t i = 0; i + 1 < data.size(); ++i)
    {
        Segment s;
        s.start = GetFrame(std::vector<double>(data[i].begin() + 1, data[i].end()));
        s.end = GetFrame(std::vector<double>(data[i + 1].begin() + 1, data[i + 1].end()));
        s.duration = data[i + 1][0] - data[i][0];
        if (s.duration <= 0) ThrowPretty("Time indices must be monotonically increasing! " + std::to_string(i) + " (" + std::to_string(s.duration) + ")");
        s.stationary = FramesEqual(s.start, s.end, 1e-6);
        if (!s.stationary)
        {
            // Path_Line: unit direction and distance; single-axis rotation start^-1 * end
            double dist = 0.0;
            for (int k = 0; k < 3; ++k)
            {
                s.dir[k] = s.end[9 + k] - s.start[9 + k];
                dist += s.dir[k] * s.dir[k];
            }
            dist = std::sqrt(dist);
            if (dist > 0.0)
                for (int k = 0; k < 3; ++k) s.dir[k] /= dist;
            double R[9];
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c) R[3 * r + c] = s.start[r] * s.end[c] + s.start[3 + r] * s.end[3 + c] + s.start[6 + r] * s.end[6 + c];
            const double alpha = RotAngleAxis(R, s.axis);
            if (alpha != 0 && alpha * radius > dist)
            {
                s.path_length = alpha * radius;  // rotation is the limitation
                s.scale_rot = 1 / radius;
                s.scale_lin = dist / s.path_length;
            }
            else if (dist != 0)
            {
                s.path_length = dist;  // translation is the limitation
                s.scale_rot = alpha / s.path_length;
                s.scale_lin = 1;
            }
            else
            {
                s.path_length = 0;
                s.scale_rot = 1;
                s.scale_lin = 1;
            }
        }
        segments_.push_back(s);
    }
    data_ = data;
    radius_ = radius;
}

double Trajectory::GetDuration() const
{
    double d = 0.0;
    for (const Segment& s : segments_) d += s.duration;
    return d;
}

Frame Trajectory::SegmentPos(const Segment& s, double time) const
{
    if (s.stationary) return s.start;
    // VelocityProfile_Spline::SetProfileDuration(0, L, dt) -> cubic with zero end velocities; Pos(time) is the polynomial as is
    const double T = s.duration, L = s.path_length;
    const double c2 = 3.0 * L / (T * T), c3 = -2.0 * L / (T * T * T);
    const double sp = c2 * time * time + c3 * time * time * time;
    // Path_Line::Pos(sp)
    double Rrel[9];
    Rot2(s.axis, sp * s.scale_rot, Rrel);
    Frame f;
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) f[3 * r + c] = s.start[3 * r] * Rrel[c] + s.start[3 * r + 1] * Rrel[3 + c] + s.start[3 * r + 2] * Rrel[6 + c];
    for (int k = 0; k < 3; ++k) f[9 + k] = s.start[9 + k] + s.dir[k] * sp * s.scale_lin;
    return f;
}

Frame Trajectory::GetPosition(double t) const
{
    if (segments_.empty()) ThrowPretty("The trajectory is empty!");
    // Trajectory_Composite::Pos
    if (t < 0) return SegmentPos(segments_.front(), 0.0);
    double previous = 0.0;
    for (const Segment& s : segments_)
    {
        if (t < previous + s.duration) return SegmentPos(s, t - previous);
        previous += s.duration;
    }
    return SegmentPos(segments_.back(), segments_.back().duration);
}

std::string Trajectory::ToString() const
{
    std::ostringstream ss;
    ss.precision(17);
    ss << radius_ << "\n" << data_.size() << " " << (data_.empty() ? 0 : data_[0].size()) << "\n";
    for (const auto& row : data_)
    {
        for (size_t j = 0; j < row.size(); ++j) ss << (j ? " " : "") << row[j];
        ss << "\n";
    }
    return ss.str();
}
}  // namespace exotica
