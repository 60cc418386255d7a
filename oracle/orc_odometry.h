// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_math.h header). PARITY UNPINNED.
//
// orc_odometry.h — ct_icp::Odometry orchestration.
//   reference: src/ct_icp/odometry.cpp (199-214 RegisterFrame, 276-330 InitializeMotion, 333-382 InitializeFrame,
//              386-501 DoRegister, 525-601 TryRegister, 604-684 AssessRegistration, 780-852 RobustRegistration,
//              855-953 UpdateMap, 978-988 ComputeSummaryMetrics, 996-1018 IncreaseRobustnessLevel)
#pragma once
#include <chrono>
#include <map>
#include <memory>

#include "orc_icp.h"

namespace orc {

struct RegistrationSummary {   // include/ct_icp/odometry.h:163-199
    TrajectoryFrame frame, initial_frame;
    int sample_size = 0, number_of_residuals = 0, robust_level = 0;
    double distance_correction = 0, relative_distance = 0, relative_orientation = 0, ego_orientation = 0;
    bool success = true, points_added = false;
    int number_of_attempts = 0;
    std::string error_message;
    std::vector<WPoint3D> corrected_points, all_corrected_points, keypoints;
    ICPSummary icp_summary;
    std::map<std::string, double> logged_values;
};

class Odometry {
public:
    struct FrameInfo {
        int registered_fid = -1;
        uint32_t frame_id = uint32_t(-1);
        double begin_timestamp = -1, end_timestamp = -1;
    };

    // Odometry::Odometry, odometry.cpp:697-734
    explicit Odometry(const cticp_odometry_options &options) : options_(options) {
        map_ = std::make_shared<VoxelMap>(options.map_options);
        switch (options_.motion_compensation) {
            case CTICP_MC_NONE:
            case CTICP_MC_CONSTANT_VELOCITY:
                options_.ct_icp_options.point_to_plane_with_distortion = false;
                options_.ct_icp_options.distance = CTICP_DIST_POINT_TO_PLANE;
                options_.ct_icp_options.parametrization = CTICP_PARAM_SIMPLE;
                break;
            case CTICP_MC_ITERATIVE:
                options_.ct_icp_options.point_to_plane_with_distortion = true;
                options_.ct_icp_options.distance = CTICP_DIST_POINT_TO_PLANE;
                options_.ct_icp_options.parametrization = CTICP_PARAM_SIMPLE;
                break;
            case CTICP_MC_CONTINUOUS:
                options_.ct_icp_options.point_to_plane_with_distortion = true;
                options_.ct_icp_options.parametrization = CTICP_PARAM_CONTINUOUS_TIME;
                options_.ct_icp_options.distance = CTICP_DIST_POINT_TO_PLANE;
                break;
        }
        next_robust_level_ = options.robust_minimal_level;
    }

    // RegisterFrame / RegisterFrameWithEstimate, odometry.cpp:199-236
    RegistrationSummary RegisterFrame(const std::vector<Vec3> &xyz, const std::vector<double> &timestamps,
                                      uint32_t frame_id, const TrajectoryFrame *initial_estimate = nullptr) {
        auto start = clock::now();
        if (timestamps.empty()) throw std::runtime_error("The registered frame cannot be empty");
        FrameInfo info;
        info.registered_fid = registered_frames_++;
        auto mm = std::minmax_element(timestamps.begin(), timestamps.end());
        info.begin_timestamp = *mm.first;
        info.end_timestamp = *mm.second;
        info.frame_id = frame_id;
        InitializeMotion(info, initial_estimate);
        auto end_init = clock::now();
        auto summary = DoRegister(xyz, timestamps, info);
        summary.logged_values["odometry_total"] = ms(clock::now() - start);
        summary.logged_values["odometry_initialization"] += ms(end_init - start);
        return summary;
    }

    const std::vector<TrajectoryFrame> &Trajectory() const { return trajectory_; }
    size_t MapSize() const { return map_->NumPoints(); }
    VoxelMap &Map() { return *map_; }
    std::shared_ptr<VoxelMap> GetMapPointer() { return map_; }
    const cticp_odometry_options &Options() const { return options_; }

    void Reset() {   // odometry.cpp:956-965
        trajectory_.clear();
        map_->Clear();
        registered_frames_ = 0;
        robust_num_consecutive_failures_ = 0;
        suspect_registration_error_ = false;
        next_robust_level_ = 0;
        tracker_ = FrameInsertionTracker();
        default_motion_model_ = MotionModel();
    }

private:
    using clock = std::chrono::steady_clock;
    static double ms(clock::duration d) { return std::chrono::duration<double, std::milli>(d).count(); }

    // InitializeMotion, odometry.cpp:276-330
    void InitializeMotion(const FrameInfo &info, const TrajectoryFrame *initial_estimate) {
        if (initial_estimate) {
            trajectory_.push_back(*initial_estimate);
            return;
        }
        const int k = info.registered_fid;
        trajectory_.emplace_back();
        trajectory_[k].begin_pose = Pose(SE3(), info.begin_timestamp, info.frame_id);
        trajectory_[k].end_pose = Pose(SE3(), info.end_timestamp, info.frame_id);
        if (k <= 1) {
        } else if (k == 2) {
            if (options_.initialization == CTICP_INIT_CONSTANT_VELOCITY) {
                trajectory_[k].begin_pose.pose = trajectory_[k - 1].end_pose.pose;
                trajectory_[k].end_pose.pose = trajectory_[k - 1].end_pose.pose *
                                               trajectory_[k - 2].end_pose.pose.Inverse() *
                                               trajectory_[k - 1].end_pose.pose;
            } else {
                trajectory_[k].begin_pose.pose = trajectory_[k - 1].begin_pose.pose;
                trajectory_[k].end_pose.pose = trajectory_[k].begin_pose.pose;
            }
        } else {
            const auto &m1 = trajectory_[k - 1];
            const auto &m2 = trajectory_[k - 2];
            if (options_.initialization == CTICP_INIT_CONSTANT_VELOCITY) {
                if (options_.motion_compensation == CTICP_MC_CONTINUOUS) {
                    trajectory_[k].begin_pose.pose = m1.begin_pose.pose * m2.begin_pose.pose.Inverse() *
                                                     m1.begin_pose.pose;
                } else {
                    trajectory_[k].begin_pose.pose = m1.end_pose.pose;
                }
                trajectory_[k].end_pose.pose = trajectory_[k - 1].end_pose.pose *
                                               trajectory_[k - 2].end_pose.pose.Inverse() *
                                               trajectory_[k - 1].end_pose.pose;
            } else {
                trajectory_[k].begin_pose.pose = m1.end_pose.pose;
                trajectory_[k].end_pose.pose = m1.end_pose.pose;
            }
        }
    }

    // TransformPoint, odometry.cpp:171-184
    void TransformPoint(WPoint3D &point, const Pose &begin_pose, const Pose &end_pose) const {
        SE3 pose = end_pose.pose;
        switch (options_.motion_compensation) {
            case CTICP_MC_NONE:
            case CTICP_MC_CONSTANT_VELOCITY:
                break;
            case CTICP_MC_CONTINUOUS:
            case CTICP_MC_ITERATIVE:
                pose = begin_pose.InterpolatePose(end_pose, point.timestamp).pose;
                break;
        }
        point.world = pose * point.raw;
    }

    // InitializeFrame, odometry.cpp:333-382 (std::shuffle → order-contract permutation, see orc_core.h)
    std::vector<WPoint3D> InitializeFrame(const std::vector<Vec3> &xyz, const std::vector<double> &ts,
                                          const FrameInfo &info) {
        double sample_size = info.registered_fid < options_.init_num_frames ? options_.init_voxel_size
                                                                            : options_.voxel_size;
        std::vector<WPoint3D> frame(xyz.size());
        for (size_t i = 0; i < frame.size(); ++i) {
            frame[i].raw = xyz[i];
            frame[i].timestamp = ts[i];
            frame[i].world = xyz[i];
            frame[i].index_frame = info.frame_id;
        }
        const int k = info.registered_fid;
        ShuffleInPlace(frame, options_.shuffle_seed, ShuffleCounter(k, 0));
        sub_sample_frame(frame, sample_size);
        if (k <= 1)
            for (auto &p : frame) p.timestamp = info.end_timestamp;
        ShuffleInPlace(frame, options_.shuffle_seed, ShuffleCounter(k, 1));

        const auto &tr = trajectory_[k];
        if (k > 1 && options_.motion_compensation == CTICP_MC_CONSTANT_VELOCITY) {
            // DistortFrame, odometry.cpp:161-168
            SE3 end_inv = tr.end_pose.Inverse().pose;
            for (auto &p : frame) {
                SE3 interp = tr.begin_pose.InterpolatePose(tr.end_pose, p.timestamp).pose;
                p.raw = end_inv * (interp * p.raw);
            }
        }
        for (auto &p : frame) TransformPoint(p, tr.begin_pose, tr.end_pose);
        for (auto &p : frame) p.index_frame = info.frame_id;
        return frame;
    }

    // TryRegister, odometry.cpp:525-601
    void TryRegister(std::vector<WPoint3D> &frame, const FrameInfo &info, cticp_icp_options &options,
                     RegistrationSummary &rs, double sample_voxel_size, const MotionModel *motion_model) {
        const int k = info.registered_fid;
        const bool at_startup = k < options_.init_num_frames;
        auto start = clock::now();
        const int attempt_idx = try_register_calls_++;
        (void) attempt_idx;
        std::vector<WPoint3D> keypoints;
        if (options_.sampling == CTICP_SAMPLING_GRID)
            grid_sampling(frame, keypoints, sample_voxel_size);
        else if (options_.sampling == CTICP_SAMPLING_ADAPTIVE) {   // odometry.cpp:539-544 (RawPointConversion)
            std::vector<Vec3> raw(frame.size());
            for (size_t i = 0; i < frame.size(); ++i) raw[i] = frame[i].raw;
            for (auto idx : AdaptiveSampleIndices(raw, options_.adaptive_options)) keypoints.push_back(frame[idx]);
        }
        else
            keypoints = frame;
        if (!at_startup && options_.max_num_keypoints > 0 && (int) keypoints.size() > options_.max_num_keypoints) {
            ShuffleInPlace(keypoints, options_.shuffle_seed, ShuffleCounter(k, 2 + attempt_idx));
            keypoints.resize(options_.max_num_keypoints);
        }
        rs.sample_size = (int) keypoints.size();
        rs.logged_values["odometry_duration_sampling"] = ms(clock::now() - start);
        if (at_startup) {
            options.threshold_voxel_occupancy = 1;
            options.num_iters_icp = std::max(options.num_iters_icp, 15);
        }
        rs.icp_summary = Register(*map_, options, options_.neighborhood_strategy, keypoints, rs.frame, motion_model);
        rs.success = rs.icp_summary.success;
        rs.number_of_residuals = rs.icp_summary.num_residuals_used;
        if (!rs.success) return;
        for (auto &p : frame) TransformPoint(p, rs.frame.begin_pose, rs.frame.end_pose);
        rs.keypoints = keypoints;
    }

    // AssessRegistration, odometry.cpp:604-684
    bool AssessRegistration(RegistrationSummary &s) const {
        if (s.relative_distance > options_.distance_error_threshold) return false;
        if (s.relative_orientation > options_.orientation_error_threshold ||
            s.ego_orientation > options_.orientation_error_threshold)
            return false;
        bool success = s.success;
        if (options_.robust_registration) {
            if (s.robust_level == 0 && (s.relative_orientation > options_.robust_threshold_relative_orientation ||
                                        s.ego_orientation > options_.robust_threshold_ego_orientation)) {
                if (s.robust_level < options_.robust_num_attempts_when_rotation) {
                    s.error_message = "Large rotations require at a robust_level of at least 1 (got:" +
                                      std::to_string(s.robust_level) + ").";
                    return false;
                }
            }
            if (s.relative_distance > options_.robust_relative_trans_threshold) {
                s.error_message = "The relative distance is too important";
                return false;
            }
        }
        return success;
    }

    // RobustRegistrationAttempt, include/ct_icp/odometry.h:289-316, odometry.cpp:996-1050
    struct Attempt {
        int robust_level = 0;
        double sample_voxel_size;
        int index_frame;
        TrajectoryFrame previous_frame, initial_estimate;
        const cticp_odometry_options &opt;
        cticp_icp_options registration_options;
        RegistrationSummary summary;
        Attempt(int idx, const cticp_odometry_options &o, const TrajectoryFrame &init)
            : index_frame(idx), initial_estimate(init), opt(o), registration_options(o.ct_icp_options) {
            sample_voxel_size = idx < o.init_num_frames ? o.init_sample_voxel_size : o.sample_voxel_size;
        }
        void IncreaseRobustnessLevel() {   // odometry.cpp:996-1018
            sample_voxel_size = index_frame < opt.init_num_frames ? opt.init_sample_voxel_size
                                                                   : opt.sample_voxel_size;
            double min_voxel_size = std::min(opt.init_voxel_size, opt.voxel_size);
            previous_frame = summary.frame;
            summary.frame = initial_estimate;
            registration_options.ls_max_num_iters += 30;
            if (registration_options.max_num_residuals > 0)
                registration_options.max_num_residuals = registration_options.max_num_residuals * 2;
            registration_options.num_iters_icp = std::min(registration_options.num_iters_icp + 20, 50);
            registration_options.threshold_orientation_norm =
                std::max(registration_options.threshold_orientation_norm / 10, 1.e-5);
            registration_options.threshold_translation_norm =
                std::max(registration_options.threshold_orientation_norm / 10, 1.e-4);
            sample_voxel_size = std::max(opt.sample_voxel_size / 1.5, double(min_voxel_size));
            registration_options.ls_sigma *= 1.2;
            registration_options.max_dist_to_plane_ct_icp *= 1.5;
            robust_level++;
        }
        void SetRobustLevel(int level) {   // odometry.cpp:1021-1025
            while (robust_level < level) IncreaseRobustnessLevel();
        }
    };

    // RobustRegistration, odometry.cpp:780-852
    void RobustRegistration(std::vector<WPoint3D> &frame, const FrameInfo &info, RegistrationSummary &rs,
                            const MotionModel *motion_model) {
        Attempt attempt(info.registered_fid, options_, rs.frame);
        attempt.summary = rs;
        attempt.summary.number_of_attempts = 0;
        bool good_enough = false;
        if (next_robust_level_ > 0) attempt.SetRobustLevel(next_robust_level_);
        do {
            TryRegister(frame, info, attempt.registration_options, attempt.summary, attempt.sample_voxel_size,
                        motion_model);
            if (attempt.index_frame > 0) {
                int k = attempt.index_frame;
                attempt.summary.distance_correction =
                    (attempt.summary.frame.BeginTr() - trajectory_[k - 1].EndTr()).norm();
                attempt.summary.relative_orientation =
                    AngularDistance(trajectory_[k - 1].end_pose.pose, attempt.summary.frame.end_pose.pose);
                attempt.summary.ego_orientation = attempt.summary.frame.EgoAngularDistance();
            }
            attempt.summary.relative_distance =
                (attempt.summary.frame.EndTr() - attempt.summary.frame.BeginTr()).norm();
            // NB the reference never copies attempt.robust_level into summary.robust_level (stays 0; :622)
            good_enough = AssessRegistration(attempt.summary);
            attempt.summary.number_of_attempts++;
            if (!good_enough) {
                if (attempt.summary.number_of_attempts < options_.robust_num_attempts)
                    attempt.IncreaseRobustnessLevel();
                else
                    good_enough = true;
            }
        } while (!good_enough);
        rs = attempt.summary;
        if (rs.number_of_attempts > options_.robust_num_attempts)
            robust_num_consecutive_failures_++;
        else
            robust_num_consecutive_failures_ = 0;
    }

    // DoRegister, odometry.cpp:386-501
    RegistrationSummary DoRegister(const std::vector<Vec3> &xyz, const std::vector<double> &ts,
                                   const FrameInfo &info) {
        auto start = clock::now();
        cticp_icp_options ct_icp_options = options_.ct_icp_options;
        const int k = info.registered_fid;
        try_register_calls_ = 0;
        auto frame = InitializeFrame(xyz, ts, info);

        RegistrationSummary summary;
        summary.frame = trajectory_.back();
        summary.initial_frame = summary.frame;
        auto end_initialization = clock::now();
        if (k > 0) {
            const MotionModel *mm = nullptr;
            if (options_.with_default_motion_model) {   // :412-417
                default_motion_model_.present = true;
                default_motion_model_.options = options_.default_motion_model;
                default_motion_model_.previous_frame = trajectory_[k - 1];
                mm = &default_motion_model_;
            }
            if (options_.robust_registration) {
                RobustRegistration(frame, info, summary, mm);
            } else {
                double sample_voxel_size = k < options_.init_num_frames ? options_.init_sample_voxel_size
                                                                        : options_.sample_voxel_size;
                auto t0 = clock::now();
                TryRegister(frame, info, ct_icp_options, summary, sample_voxel_size, mm);
                summary.logged_values["odometry_try_register"] = ms(clock::now() - t0);
                // NB trajectory_[k] is still the INITIAL estimate here (:429-431)
                summary.relative_orientation =
                    AngularDistance(trajectory_[k - 1].end_pose.pose, trajectory_[k].end_pose.pose);
                summary.ego_orientation = summary.frame.EgoAngularDistance();
                summary.relative_distance = (summary.frame.EndTr() - summary.frame.BeginTr()).norm();
                if (!AssessRegistration(summary)) {
                    summary.success = false;
                    if (options_.quit_on_error) return summary;
                }
            }
            trajectory_[k] = summary.frame;
        }
        auto end = clock::now();

        summary.corrected_points = frame;   // :462-486
        summary.all_corrected_points.resize(xyz.size());
        const Pose &bp = summary.frame.begin_pose;
        const Pose &ep = summary.frame.end_pose;
        // odometry.cpp:469,480: num_threads(options_.ct_icp_options.ls_num_threads) — NOT the machine's core count
        const int transform_threads = std::max(1, options_.ct_icp_options.ls_num_threads);
#pragma omp parallel for num_threads(transform_threads)
        for (long i = 0; i < (long) summary.all_corrected_points.size(); ++i) {
            auto &p = summary.all_corrected_points[i];
            p.raw = xyz[i];
            p.timestamp = ts[i];
            p.index_frame = info.frame_id;
            p.world = bp.ContinuousTransform(p.raw, ep, p.timestamp);
        }
#pragma omp parallel for num_threads(transform_threads)
        for (long i = 0; i < (long) summary.corrected_points.size(); ++i) {
            auto &p = summary.corrected_points[i];
            p.world = bp.ContinuousTransform(p.raw, ep, p.timestamp);
        }
        auto end_transform = clock::now();
        ComputeSummaryMetrics(summary, k);
        UpdateMap(summary, k);
        auto end_map = clock::now();
        summary.logged_values["odometry_num_keypoints"] = (double) summary.keypoints.size();
        summary.logged_values["odometry_total_duration(ms)"] = ms(end - start);
        summary.logged_values["odometry_initialization(ms)"] = ms(end_initialization - start);
        summary.logged_values["odometry_map_update(ms)"] = ms(end_map - end_transform);
        summary.logged_values["odometry_transform(ms)"] = ms(end_transform - end);
        return summary;
    }

    // ComputeSummaryMetrics, odometry.cpp:978-988
    void ComputeSummaryMetrics(RegistrationSummary &s, int k) {
        if (k > 0) {
            auto &cur = trajectory_[k];
            auto &prev = trajectory_[k - 1];
            s.distance_correction = (cur.BeginTr() - prev.EndTr()).norm();
            s.relative_orientation = AngularDistance(prev.end_pose.pose, cur.end_pose.pose);
            s.relative_distance = (prev.EndTr() - cur.EndTr()).norm();
            s.ego_orientation = cur.EgoAngularDistance();
        }
    }

    // UpdateMap, odometry.cpp:855-953
    void UpdateMap(RegistrationSummary &s, int registered_fid) {
        bool add_points = true;
        if (options_.robust_registration) {
            suspect_registration_error_ = s.number_of_attempts >= options_.robust_num_attempts;
            if (s.ego_orientation > options_.robust_threshold_ego_orientation ||
                s.relative_orientation > options_.robust_threshold_relative_orientation)
                add_points = false;
            if (suspect_registration_error_) add_points |= (robust_num_consecutive_failures_ > 5);
            next_robust_level_ = add_points ? options_.robust_minimal_level : options_.robust_minimal_level + 1;
            if (!s.success)
                next_robust_level_ = options_.robust_minimal_level + 2;
            else {
                if (s.relative_orientation > options_.robust_threshold_relative_orientation ||
                    s.ego_orientation > options_.robust_threshold_ego_orientation)
                    next_robust_level_ = options_.robust_minimal_level + 1;
                if (s.number_of_attempts > 1) next_robust_level_ = options_.robust_minimal_level + 1;
            }
        } else {
            tracker_.cum_orientation += s.relative_orientation;
            tracker_.cum_distance += s.relative_distance;
            if (tracker_.total_insertions > 0) {
                if (s.ego_orientation > options_.insertion_ego_rotation_threshold)
                    add_points = tracker_.skipped_frames > options_.insertion_threshold_frames_skipped;
                else
                    add_points = true;
            }
        }
        s.points_added = add_points;
        if (options_.do_no_insert) add_points = false;
        if (options_.always_insert) add_points = true;

        const Vec3 location = trajectory_.back().EndTr();
        map_->RemoveElementsFarFromLocation(location, options_.max_distance);
        if (add_points) {
            std::vector<Vec3> world(s.corrected_points.size());
            for (size_t i = 0; i < world.size(); ++i) world[i] = s.corrected_points[i].world;
            map_->InsertPoints(world, s.frame.BeginTr());   // frame_poses = {begin_pose, end_pose}, odometry.cpp:949
            tracker_.skipped_frames = 0;
            tracker_.cum_orientation = 0;
            tracker_.cum_distance = 0;
            tracker_.total_insertions++;
            (void) registered_fid;
        } else
            tracker_.skipped_frames++;
    }

    struct FrameInsertionTracker {   // include/ct_icp/odometry.h:318-347
        double cum_distance = 0, cum_orientation = 0;
        int skipped_frames = 0, total_insertions = 0;
    } tracker_;

    cticp_odometry_options options_;
    std::vector<TrajectoryFrame> trajectory_;
    std::shared_ptr<VoxelMap> map_;
    MotionModel default_motion_model_;
    int registered_frames_ = 0;
    int robust_num_consecutive_failures_ = 0;
    bool suspect_registration_error_ = false;
    int next_robust_level_ = 0;
    int try_register_calls_ = 0;
    // Order contract: the counter of each permutation is a pure function of (registered frame, purpose) so that a
    // device pipeline never needs data-dependent host state: purpose 0/1 = the two shuffles of InitializeFrame
    // (odometry.cpp:349,361), 2+a = the keypoint truncation shuffle of the a-th TryRegister call (:550).
    static uint64_t ShuffleCounter(int registered_fid, int purpose) {
        return (uint64_t(uint32_t(registered_fid)) << 8) | uint64_t(purpose & 0xff);
    }
};

}  // namespace orc
