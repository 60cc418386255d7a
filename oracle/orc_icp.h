// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_math.h header). PARITY UNPINNED.
//
// orc_icp.h — CT_ICP_Registration: Gauss-Newton solver (DoRegisterGaussNewton) and dispatch.
//   reference: src/ct_icp/ct_icp.cpp:709-996 (GN), :998-1037 (Register / SELECT_SOLVER)
#pragma once
#include <sstream>
#include <string>

#include "orc_core.h"

namespace orc {

struct ICPSummary {   // include/ct_icp/ct_icp.h:155-169
    bool success = false;
    int num_residuals_used = 0;
    int num_iters = 0;
    std::string error_log;
    double duration_total = 0, duration_init = 0, avg_duration_iter = 0, avg_duration_neighborhood = 0,
           avg_duration_solve = 0;
    // oracle-only counters for the bench's algorithmic-bytes figure
    size_t keypoint_iterations = 0, stencil_points = 0;
};

// PreviousFrameMotionModel state + options (include/ct_icp/motion_model.h:33-92)
struct MotionModel {
    bool present = false;
    cticp_motion_model_options options;
    TrajectoryFrame previous_frame;
};

struct GNLinearSystem {
    double A[12][12];
    std::array<double, 12> b;
    int num_used = 0;
};

// One linearisation of src/ct_icp/ct_icp.cpp:746-910 at `frame` (world points of the keypoints as given).
inline void GNBuildSystem(const VoxelMap &map, const cticp_icp_options &options, const std::vector<WPoint3D> &kpts,
                          const TrajectoryFrame &frame, const MotionModel *motion_model, GNLinearSystem &sys,
                          ICPSummary *counters = nullptr) {
    const Pose &pose_begin = frame.begin_pose;
    const Pose &pose_end = frame.end_pose;
    for (int i = 0; i < 12; ++i) {
        sys.b[i] = 0;
        for (int j = 0; j < 12; ++j) sys.A[i][j] = 0;
    }
    int number_keypoints_used = 0;
    const int kMinNumNeighbors = options.min_number_neighbors;
    for (size_t pid = 0; pid < kpts.size(); ++pid) {
        const Vec3 pt_keypoint = kpts[pid].world;
        const Vec3 raw_pt_keypoint = kpts[pid].raw;
        const double timestamp = kpts[pid].timestamp;

        Neighborhood neighborhood;   // :762
        size_t stencil = 0;
        map.ComputeNeighborhoodInPlace(pt_keypoint, options.max_number_neighbors, neighborhood, &stencil);
        if (counters) {
            counters->keypoint_iterations++;
            counters->stencil_points += stencil;
        }
        if ((int) neighborhood.points.size() < kMinNumNeighbors) continue;   // :769

        neighborhood.ComputeNeighborhood();   // :778
        double planarity_weight = neighborhood.description.a2D;
        Vec3 normal = neighborhood.description.normal;
        if (normal.dot(frame.BeginTr() - pt_keypoint) < 0) normal = -1.0 * normal;   // :782-784

        double alpha_timestamp = pose_begin.GetAlphaTimestamp(timestamp, pose_end);   // :786
        double weight = planarity_weight * planarity_weight;                          // :787-788
        Vec3 closest_pt_normal = weight * normal;
        Vec3 closest_point = neighborhood.points[0];                                  // :791 (farthest of the k kept)

        double dist_to_plane = normal[0] * (pt_keypoint[0] - closest_point[0]) +
                               normal[1] * (pt_keypoint[1] - closest_point[1]) +
                               normal[2] * (pt_keypoint[2] - closest_point[2]);
        if (std::fabs(dist_to_plane) < options.max_dist_to_plane_ct_icp) {            // :803
            double scalar = closest_pt_normal[0] * (pt_keypoint[0] - closest_point[0]) +
                            closest_pt_normal[1] * (pt_keypoint[1] - closest_point[1]) +
                            closest_pt_normal[2] * (pt_keypoint[2] - closest_point[2]);
            number_keypoints_used++;

            Vec3 ob = frame.BeginQuat() * raw_pt_keypoint;   // :813-816
            Vec3 oe = frame.EndQuat() * raw_pt_keypoint;
            const double am = 1 - alpha_timestamp, a = alpha_timestamp;
            const Vec3 &n = closest_pt_normal;
            double u[12] = {am * (ob[1] * n[2] - ob[2] * n[1]), am * (ob[2] * n[0] - ob[0] * n[2]),
                            am * (ob[0] * n[1] - ob[1] * n[0]), am * n[0], am * n[1], am * n[2],
                            a * (oe[1] * n[2] - oe[2] * n[1]),  a * (oe[2] * n[0] - oe[0] * n[2]),
                            a * (oe[0] * n[1] - oe[1] * n[0]),  a * n[0],  a * n[1],  a * n[2]};
            for (int i = 0; i < 12; i++) {   // :845-850
                for (int j = 0; j < 12; j++) sys.A[i][j] = sys.A[i][j] + u[i] * u[j];
                sys.b[i] = sys.b[i] - u[i] * scalar;
            }
        }
    }
    sys.num_used = number_keypoints_used;
    if (number_keypoints_used < 100) return;   // :860 (caller reports the failure)

    for (int i = 0; i < 12; i++) {   // :877-882
        for (int j = 0; j < 12; j++) sys.A[i][j] = sys.A[i][j] / number_keypoints_used;
        sys.b[i] = sys.b[i] / number_keypoints_used;
    }
    if (motion_model && motion_model->present) {   // :885-910
        const double ALPHA_C = motion_model->options.beta_location_consistency;
        const double ALPHA_E = motion_model->options.beta_constant_velocity;
        Vec3 diff_traj = frame.BeginTr() - frame.EndTr();   // :892 (the current frame's own begin-end, sic)
        for (int d = 0; d < 3; ++d) {
            sys.A[3 + d][3 + d] += ALPHA_C;
            sys.b[3 + d] -= ALPHA_C * diff_traj[d];
        }
        Vec3 diff_ego = frame.EndTr() - frame.BeginTr() - motion_model->previous_frame.EndTr() +
                        motion_model->previous_frame.BeginTr();
        for (int d = 0; d < 3; ++d) {
            sys.A[9 + d][9 + d] += ALPHA_E;
            sys.b[9 + d] -= ALPHA_E * diff_ego[d];
        }
    }
}

inline Mat3 EulerZYX(double alpha, double beta, double gamma) {   // :916-932
    Mat3 R;
    R(0, 0) = cos(gamma) * cos(beta);
    R(0, 1) = -sin(gamma) * cos(alpha) + cos(gamma) * sin(beta) * sin(alpha);
    R(0, 2) = sin(gamma) * sin(alpha) + cos(gamma) * sin(beta) * cos(alpha);
    R(1, 0) = sin(gamma) * cos(beta);
    R(1, 1) = cos(gamma) * cos(alpha) + sin(gamma) * sin(beta) * sin(alpha);
    R(1, 2) = -cos(gamma) * sin(alpha) + sin(gamma) * sin(beta) * cos(alpha);
    R(2, 0) = -sin(beta);
    R(2, 1) = cos(beta) * sin(alpha);
    R(2, 2) = cos(beta) * cos(alpha);
    return R;
}

// DoRegisterGaussNewton, src/ct_icp/ct_icp.cpp:709-996
inline ICPSummary DoRegisterGaussNewton(const VoxelMap &map, const cticp_icp_options &options,
                                        std::vector<WPoint3D> &kpts, TrajectoryFrame &frame,
                                        const MotionModel *motion_model) {
    frame.begin_pose.pose.quat.normalize();
    frame.end_pose.pose.quat.normalize();
    ICPSummary summary;
    GNLinearSystem sys;
    int iter = 0;
    for (; iter < options.num_iters_icp; iter++) {
        GNBuildSystem(map, options, kpts, frame, motion_model, sys, &summary);
        if (sys.num_used < 100) {
            std::stringstream ss;
            ss << "[CT_ICP]Error : not enough keypoints selected in ct-icp !" << std::endl;
            ss << "[CT_ICP]Number_of_residuals : " << sys.num_used << std::endl;
            summary.error_log = ss.str();
            summary.success = false;
            return summary;
        }
        std::array<double, 12> x = LDLTSolve<12>(sys.A, sys.b);   // :914

        Mat3 rotation_begin = EulerZYX(x[0], x[1], x[2]);
        Mat3 rotation_end = EulerZYX(x[6], x[7], x[8]);
        frame.begin_pose.pose.quat = Quat::fromRotationMatrix(rotation_begin * frame.BeginQuat().toRotationMatrix());
        frame.begin_pose.pose.tr += Vec3(x[3], x[4], x[5]);
        frame.end_pose.pose.quat = Quat::fromRotationMatrix(rotation_end * frame.EndQuat().toRotationMatrix());
        frame.end_pose.pose.tr += Vec3(x[9], x[10], x[11]);
        frame.begin_pose.pose.quat.normalize();
        frame.end_pose.pose.quat.normalize();

        for (auto &kp : kpts)   // :964-966
            kp.world = frame.begin_pose.InterpolatePose(frame.end_pose, kp.timestamp) * kp.raw;

        double norm = 0;
        for (double v : x) norm += v * v;
        summary.num_iters = iter + 1;   // oracle-only bookkeeping (the reference leaves num_iters at 0 for GN)
        if (std::sqrt(norm) < options.threshold_orientation_norm) break;   // :978
    }
    summary.success = true;
    summary.num_residuals_used = sys.num_used;
    return summary;
}

ICPSummary DoRegisterCeres(const VoxelMap &map, const cticp_icp_options &options,
                           const cticp_strategy_options &strategy, std::vector<WPoint3D> &kpts,
                           TrajectoryFrame &frame, const MotionModel *motion_model);   // orc_ceres.cpp
ICPSummary DoRegisterRobust(const VoxelMap &map, const cticp_icp_options &options, std::vector<WPoint3D> &kpts,
                            TrajectoryFrame &frame, const MotionModel *motion_model);   // orc_ceres.cpp

// CT_ICP_Registration::Register + SELECT_SOLVER, src/ct_icp/ct_icp.cpp:998-1037
inline ICPSummary Register(const VoxelMap &map, const cticp_icp_options &options,
                           const cticp_strategy_options &strategy, std::vector<WPoint3D> &kpts,
                           TrajectoryFrame &frame, const MotionModel *motion_model) {
    switch (options.solver) {
        case CTICP_SOLVER_GN:
            return DoRegisterGaussNewton(map, options, kpts, frame, motion_model);
        case CTICP_SOLVER_CERES:
            return DoRegisterCeres(map, options, strategy, kpts, frame, motion_model);
        case CTICP_SOLVER_ROBUST:
            return DoRegisterRobust(map, options, kpts, frame, motion_model);
        default:
            throw std::runtime_error("Unsupported Solver Type");
    }
}

}  // namespace orc
