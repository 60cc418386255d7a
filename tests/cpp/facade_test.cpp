// C++ drop-in check of the facade (ct_icp_b200/include/ct_icp_b200/odometry.hpp): the reference's integration test
// test/integration/testint_odometry.cpp:57-114 restated — a 6-plane box scene (testint_utils.h:39-96), a sensor
// moving through it, per-point timestamps; asserts `result.success` for every frame (and, beyond the reference,
// that the trajectory tracks the ground truth).
//   build: g++ -std=c++17 -I include -I ct_icp_b200/include tests/cpp/facade_test.cpp -L ct_icp_b200 -lcticp_b200
#include <cstdio>
#include <cmath>
#include <random>

#include "ct_icp_b200/odometry.hpp"

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
