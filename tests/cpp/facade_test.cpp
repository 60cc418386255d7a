// C++ drop-in check of the facade (ct_icp_b200/include/ct_icp_b200/odometry.hpp): the reference's integration test
// test/integration/testint_odometry.cpp:57-114 restated — a 6-plane box scene (testint_utils.h:39-96), a sensor
// moving through it, per-point timestamps; asserts `result.success` for every frame (and, beyond the reference,
// that the trajectory tracks the ground truth).
//   build: g++ -std=c++17 -I include -I ct_icp_b200/include tests/cpp/facade_test.cpp -L ct_icp_b200 -lcticp_b200
#include <cstdio>
#include <cmath>
#include <random>

#include "ct_icp_b200/odometry.hpp"
#include "slam_pointcloud_stub.h"

namespace {
const double kScale = 30., kPlaneLoc = 4. + kScale;

struct GtPose {
    double yaw, tx, ty, tz;
};
GtPose gt_pose(double t) { return {0.01 * t, 0.35 * t, 0.05 * std::sin(0.5 * t), 0.02 * t}; }

// world → sensor frame with the ground-truth pose at time t
void to_sensor(const GtPose &p, const double w[3], double out[3]) {
    const double dx = w[0] - p.tx, dy = w[1] - p.ty, dz = w[2] - p.tz;
    const double c = std::cos(-p.yaw), s = std::sin(-p.yaw);
    out[0] = c * dx - s * dy;
    out[1] = s * dx + c * dy;
    out[2] = dz;
}

std::vector<slam::WPoint3D> generate_frame(int frame_id, int num_points, std::mt19937_64 &rng) {
    std::uniform_real_distribution<double> u(-1.0, 1.0);
    std::vector<slam::WPoint3D> pts;
    pts.reserve(num_points * 6);
    for (int i = 0; i < num_points; ++i) {
        for (int plane = 0; plane < 6; ++plane) {
            double w[3] = {u(rng) * kScale, u(rng) * kScale, u(rng) * kScale};
            w[plane / 2] = (plane % 2 ? -kPlaneLoc : kPlaneLoc);
            const double alpha = (double) i / (num_points - 1);
            slam::WPoint3D p;
            p.raw_point.timestamp = frame_id + alpha;
            double raw[3];
            to_sensor(gt_pose(p.raw_point.timestamp), w, raw);
            for (int d = 0; d < 3; ++d) {
                p.raw_point.point[d] = (double) (float) raw[d];
                p.world_point[d] = w[d];
            }
            p.index_frame = frame_id;
            pts.push_back(p);
        }
    }
    return pts;
}
}  // namespace

int main(int argc, char **argv) {
    const char *solver = argc > 1 ? argv[1] : "CERES";
    ct_icp::OdometryOptions options;                 // reference defaults (3-resolution map, CERES)
    options.ct_icp_options.solver = std::string(solver) == "GN" ? ct_icp::GN : ct_icp::CERES;
    options.initialization = ct_icp::INIT_NONE;      // like testint_odometry.cpp:73
    options.debug_print = 0;
    options.ct_icp_options.num_iters_icp = 30;
    options.ct_icp_options.ls_max_num_iters = 5;
    options.ct_icp_options.min_number_neighbors = 10;
    options.map_options.capacity_voxels = 1 << 18;
    options.init_num_frames = 3;
    options.voxel_size = 1.0;
    options.init_voxel_size = 1.0;
    options.sample_voxel_size = 1.5;
    options.init_sample_voxel_size = 1.5;
    options.map_options.num_resolutions = 1;
    options.map_options.resolutions[0].resolution = 2.0;
    options.map_options.resolutions[0].max_num_points = 30;
    options.map_options.resolutions[0].min_distance_between_points = 0.3;
    options.map_options.default_radius = 2.0;
    options.max_distance = 200.0;
    try {
        ct_icp::Odometry odometry(options);
        std::mt19937_64 rng(42);
        const int kFrames = 10;
        double err = 0;
        for (int i = 0; i < kFrames; ++i) {
            auto frame = generate_frame(i, 5000, rng);
            auto result = odometry.RegisterFrame(frame);
            if (!result.success) {
                std::printf("Odometry failed at frame %d: %s\n", i, result.error_message.c_str());
                return 1;
            }
            if (result.all_corrected_points.size() != frame.size() || result.corrected_points.empty()) {
                std::printf("point vectors not filled\n");
                return 1;
            }
            // frames 0 is registered at identity (gt_pose(0) is the identity too)
            const GtPose g = gt_pose(i + 1.0);
            const auto &t = result.frame.end_pose.pose.tr;
            err = std::sqrt((t[0] - g.tx) * (t[0] - g.tx) + (t[1] - g.ty) * (t[1] - g.ty) + (t[2] - g.tz) * (t[2] - g.tz));
            std::printf("frame %d: keypoints %zu residuals %d end_tr (%.3f %.3f %.3f) gt (%.3f %.3f %.3f) err %.4f\n", i,
                        result.keypoints.size(), result.number_of_residuals, t[0], t[1], t[2], g.tx, g.ty, g.tz, err);
        }
        // ---- the rest of the boundary (include/ct_icp/odometry.h:226-272): a second odometry on the same frames through the
        // slam::PointCloud overload (what odometry_runner.cpp:194 and the ROS node call), with a caller-owned motion model,
        // callbacks at the three events and Reset(options) — it must reproduce the first trajectory
        {
            struct Counter : ct_icp::Odometry::OdometryCallback {
                int runs = 0;
                size_t frame_points = 0, keypoints = 0;
                bool Run(const ct_icp::Odometry &, const std::vector<slam::WPoint3D> &current_frame,
                         const std::vector<slam::WPoint3D> *kp, const ct_icp::Odometry::RegistrationSummary *) override {
                    ++runs;
                    frame_points = current_frame.size();
                    if (kp) keypoints = kp->size();
                    return true;
                }
            } before, completed, finished;
            auto options2 = options;
            options2.voxel_size = 2.0;                      // wrong on purpose: Reset(options) below restores the real ones
            ct_icp::Odometry second(&options2);             // the pointer constructor (:228)
            second.Reset(options);
            second.RegisterCallback(ct_icp::Odometry::OdometryCallback::BEFORE_ITERATION, before);
            second.RegisterCallback(ct_icp::Odometry::OdometryCallback::ITERATION_COMPLETED, completed);
            second.RegisterCallback(ct_icp::Odometry::OdometryCallback::FINISHED_REGISTRATION, finished);
            ct_icp::PreviousFrameMotionModel model;         // same options as the default model: same result expected
            model.GetOptions() = ct_icp::PreviousFrameMotionModel::Options();
            static_cast<cticp_motion_model_options &>(model.GetOptions()) = options.default_motion_model;
            std::mt19937_64 rng2(42);
            const auto first_trajectory = odometry.Trajectory();
            for (int i = 0; i < kFrames; ++i) {
                const auto frame = generate_frame(i, 5000, rng2);
                slam::PointCloud cloud;
                for (const auto &p : frame) cloud.push_back(p.raw_point.point[0], p.raw_point.point[1], p.raw_point.point[2], p.raw_point.timestamp);
                const auto result = second.RegisterFrame(cloud, (slam::frame_id_t) i, &model);
                if (!result.success) {
                    std::printf("PointCloud overload failed at frame %d: %s\n", i, result.error_message.c_str());
                    return 1;
                }
                model.UpdateState(result.frame, i);
                const auto &a = result.frame.end_pose.pose.tr, &b = first_trajectory[i].end_pose.pose.tr;
                const double d = (a - b).norm();
                if (d > 1e-9) {
                    std::printf("PointCloud overload / caller motion model: frame %d differs by %.3e m\n", i, d);
                    return 1;
                }
                if (result.logged_values.count("odometry_total_duration(ms)") == 0 || result.logged_values.count("icp_total_duration") == 0 ||
                    (i > 0 && !(result.icp_summary.duration_total > 0.0))) {
                    std::printf("logged_values / icp_summary durations missing\n");
                    return 1;
                }
            }
            // frame 0 has no ICP: BEFORE / COMPLETED run on kFrames - 1 frames, FINISHED on all
            if (before.runs != kFrames - 1 || completed.runs != kFrames - 1 || finished.runs != kFrames || before.keypoints == 0 ||
                finished.frame_points == 0) {
                std::printf("callbacks: before %d completed %d finished %d\n", before.runs, completed.runs, finished.runs);
                return 1;
            }
            if (second.Map().NumPoints() != odometry.MapConst().NumPoints()) return 1;
            struct Veto : ct_icp::Odometry::OdometryCallback {
                bool Run(const ct_icp::Odometry &, const std::vector<slam::WPoint3D> &, const std::vector<slam::WPoint3D> *,
                         const ct_icp::Odometry::RegistrationSummary *) override { return false; }
            } veto;
            second.RegisterCallback(ct_icp::Odometry::OdometryCallback::BEFORE_ITERATION, veto);
            bool threw = false;
            try {
                second.RegisterFrame(generate_frame(kFrames, 5000, rng2));
            } catch (const ct_icp::CticpFailure &e) {
                threw = e.code == CTICP_ERR_CALLBACK;   // the reference CHECK-aborts here (odometry.cpp:748)
            }
            if (!threw) {
                std::printf("a callback returning false must abort the registration\n");
                return 1;
            }
        }
        auto trajectory = odometry.Trajectory();
        if ((int) trajectory.size() != kFrames || odometry.MapSize() == 0 || odometry.GetMapPointer()->NumPoints() != odometry.MapSize())
            return 1;
        if (err > 0.25) {
            std::printf("trajectory drifted: %.3f m\n", err);
            return 1;
        }
        std::printf("FACADE OK (%s) final error %.4f m, map %zu points\n", solver, err, odometry.MapSize());
    } catch (const ct_icp::CticpFailure &e) {
        std::printf("CTICP failure code %d: %s\n", e.code, e.what());
        return e.code == CTICP_ERR_NO_DEVICE ? 42 : 2;
    }
    return 0;
}
