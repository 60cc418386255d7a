// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_math.h header). PARITY UNPINNED.
//
// orc_capi.cpp — C entry points (ctypes) mirroring include/cticp.h one-for-one with the `orc_` prefix, so the
// parity tests drive oracle and engine with identical calls.
#include <cstdio>
#include <cstring>
#include <string>

#include "orc_odometry.h"

using namespace orc;

namespace {
thread_local std::string g_last_error;
int Fail(int code, const std::string &msg) {
    g_last_error = msg;
    return code;
}
template <typename F>
int Guard(F &&f) {
    try {
        return f();
    } catch (const TimestampError &e) {
        return Fail(CTICP_ERR_TIMESTAMP, e.what());
    } catch (const std::exception &e) {
        return Fail(CTICP_ERR_INTERNAL, e.what());
    }
}
const double *StrideAt(const double *base, size_t stride, size_t i) {
    return reinterpret_cast<const double *>(reinterpret_cast<const char *>(base) + stride * i);
}
}  // namespace

struct orc_odometry {
    std::unique_ptr<Odometry> impl;
    RegistrationSummary last;
};
struct orc_map {
    std::shared_ptr<VoxelMap> impl;
};

extern "C" {

const char *orc_last_error(void) { return g_last_error.c_str(); }

/* ---- defaults (include/ct_icp/ct_icp.h:60-152, map.h:115-125, odometry.h:37-157, motion_model.h:42-58) ---- */
void orc_default_icp_options(cticp_icp_options *o) {
    std::memset(o, 0, sizeof(*o));
    o->num_iters_icp = 5;
    o->parametrization = CTICP_PARAM_CONTINUOUS_TIME;
    o->distance = CTICP_DIST_POINT_TO_PLANE;
    o->solver = CTICP_SOLVER_CERES;
    o->max_num_residuals = -1;
    o->min_num_residuals = 100;
    o->weighting_scheme = CTICP_WEIGHT_ALL;
    o->weight_alpha = 0.9;
    o->weight_neighborhood = 0.1;
    o->power_planarity = 2.0;
    o->max_number_neighbors = 20;
    o->min_number_neighbors = 20;
    o->threshold_voxel_occupancy = 1;
    o->num_closest_neighbors = 1;
    o->threshold_orientation_norm = 0.0001;
    o->threshold_translation_norm = 0.001;
    o->point_to_plane_with_distortion = 1;
    o->loss_function = CTICP_LOSS_CAUCHY;
    o->ls_max_num_iters = 1;
    o->ls_num_threads = 16;
    o->ls_sigma = 0.1;
    o->ls_tolerant_min_threshold = 0.05;
    o->max_dist_to_plane_ct_icp = 0.3;
    o->threshold_linearity = 0.8;
    o->threshold_planarity = 0.8;
    o->weight_point_to_point = 0.1;
    o->outlier_distance = 1.0;
    o->use_barycenter = 0;
    o->use_lines = 1;
    o->debug_print = 1;
}
void orc_default_map_options(cticp_map_options *o) {
    std::memset(o, 0, sizeof(*o));
    o->num_resolutions = 3;
    o->resolutions[0] = {0.2, 0.03, 50, 0};
    o->resolutions[1] = {0.5, 0.1, 40, 0};
    o->resolutions[2] = {1.5, 0.15, 40, 0};
    o->select_valid_normals_direction = 1;
    o->max_frames_to_keep = 100;
    o->default_radius = 0.8;
}
void orc_legacy_map_options(cticp_map_options *o, double size_voxel_map, int max_num_points_in_voxel,
                            double min_distance_points) {   // src/ct_icp/map.cpp:13-29
    orc_default_map_options(o);
    o->num_resolutions = 1;
    o->max_frames_to_keep = 1;
    o->resolutions[0].resolution = size_voxel_map;
    o->resolutions[0].max_num_points = max_num_points_in_voxel;
    o->resolutions[0].min_distance_between_points = min_distance_points;
}
void orc_default_adaptive_options(cticp_adaptive_options *a) {   // include/ct_icp/algorithm/sampling.h:14-27
    std::memset(a, 0, sizeof(*a));
    a->num_points_per_voxel = 1;
    a->max_num_points = -1;
    a->num_bands = 6;
    const double d[6] = {0.5, 2.0, 4., 8., 16., 200.}, v[6] = {0.1, 0.2, 0.4, 0.8, 1.6, -1.};
    for (int i = 0; i < 6; ++i) { a->distance[i] = d[i]; a->voxel_size[i] = v[i]; }
}
void orc_default_odometry_options(cticp_odometry_options *o) {
    std::memset(o, 0, sizeof(*o));
    orc_default_adaptive_options(&o->adaptive_options);
    orc_default_icp_options(&o->ct_icp_options);
    orc_default_map_options(&o->map_options);
    o->neighborhood_strategy = {0, 20, 8, 0, 60., 0.1, 2.0, 1.0};   // neighborhood_strategy.h:47-49,113-119
    o->default_motion_model.model = CTICP_MM_CONSTANT_VELOCITY;
    o->default_motion_model.beta_location_consistency = 0.001;
    o->default_motion_model.beta_constant_velocity = 0.001;
    o->default_motion_model.beta_small_velocity = 0.0;
    o->default_motion_model.beta_orientation_consistency = 0.0;
    o->default_motion_model.threshold_orientation_deg = 15;
    o->default_motion_model.threshold_translation_diff = 0.3;
    o->default_motion_model.log_if_invalid = 1;
    o->motion_compensation = CTICP_MC_CONTINUOUS;
    o->initialization = CTICP_INIT_CONSTANT_VELOCITY;
    o->init_voxel_size = 0.2;
    o->init_sample_voxel_size = 1.0;
    o->init_num_frames = 20;
    o->sample_voxel_size = 1.5;
    o->max_num_keypoints = -1;
    o->sampling = CTICP_SAMPLING_GRID;
    o->voxel_size = 0.5;
    o->max_distance = 100.0;
    o->distance_error_threshold = 5.0;
    o->orientation_error_threshold = 30.;
    o->quit_on_error = 1;
    o->robust_minimal_level = 0;
    o->robust_registration = 0;
    o->robust_full_voxel_threshold = 0.7;
    o->robust_empty_voxel_threshold = 0.1;
    o->robust_neighborhood_min_dist = 0.10;
    o->robust_neighborhood_min_orientation = 0.1;
    o->robust_relative_trans_threshold = 1.0;
    o->robust_fail_early = 0;
    o->robust_num_attempts = 6;
    o->robust_num_attempts_when_rotation = 2;
    o->robust_max_voxel_neighborhood = 3;
    o->robust_threshold_ego_orientation = 3;
    o->robust_threshold_relative_orientation = 3;
    o->insertion_ego_rotation_threshold = 3;
    o->insertion_threshold_frames_skipped = 5;
    o->insertion_cum_distance_threshold = 0.8;
    o->insertion_cum_orientation_threshold = 5;
    o->always_insert = 0;
    o->do_no_insert = 0;
    o->debug_print = 1;
    o->with_default_motion_model = 1;
    o->shuffle_seed = 0x5DEECE66Dull;
    o->max_points_per_frame = 0;
}
void orc_profile_default_driving(cticp_odometry_options *o) {   // src/ct_icp/odometry.cpp:30-36
    orc_default_odometry_options(o);
    o->ct_icp_options.solver = CTICP_SOLVER_CERES;
    o->ct_icp_options.ls_num_threads = 6;
    o->ct_icp_options.num_iters_icp = 5;
}
void orc_profile_robust_driving(cticp_odometry_options *o) {   // src/ct_icp/odometry.cpp:39-89
    orc_default_odometry_options(o);
    o->voxel_size = 0.5;
    o->sample_voxel_size = 1.5;
    o->max_distance = 200.0;
    o->init_num_frames = 40;
    o->distance_error_threshold = 5.0;
    o->motion_compensation = CTICP_MC_CONTINUOUS;
    o->initialization = CTICP_INIT_CONSTANT_VELOCITY;
    o->debug_print = 0;
    o->robust_registration = 1;
    o->robust_full_voxel_threshold = 0.5;
    o->robust_empty_voxel_threshold = 0.2;
    o->robust_num_attempts = 10;
    o->robust_max_voxel_neighborhood = 4;
    o->robust_threshold_relative_orientation = 5;
    o->robust_threshold_ego_orientation = 5;
    o->default_motion_model.beta_constant_velocity = 0.001;
    o->default_motion_model.beta_location_consistency = 0.001;
    o->default_motion_model.beta_small_velocity = 0.00;
    auto &c = o->ct_icp_options;
    c.debug_print = 0;
    c.max_number_neighbors = 20;
    c.min_number_neighbors = 20;
    c.num_iters_icp = 15;
    c.max_dist_to_plane_ct_icp = 0.5;
    c.threshold_orientation_norm = 0.01;
    c.point_to_plane_with_distortion = 1;
    c.distance = CTICP_DIST_POINT_TO_PLANE;
    c.parametrization = CTICP_PARAM_CONTINUOUS_TIME;
    c.num_closest_neighbors = 1;
    c.loss_function = CTICP_LOSS_CAUCHY;
    c.solver = CTICP_SOLVER_CERES;
    c.ls_max_num_iters = 20;
    c.ls_num_threads = 8;
    c.ls_sigma = 0.2;
    c.ls_tolerant_min_threshold = 0.05;
}
void orc_profile_robust_outdoor_low_inertia(cticp_odometry_options *o) {   // src/ct_icp/odometry.cpp:92-151
    orc_default_odometry_options(o);
    o->voxel_size = 0.3;
    o->sample_voxel_size = 1.5;
    o->max_distance = 200.0;
    o->init_num_frames = 20;
    o->distance_error_threshold = 5.0;
    o->motion_compensation = CTICP_MC_CONTINUOUS;
    o->initialization = CTICP_INIT_NONE;
    o->debug_print = 0;
    o->robust_registration = 1;
    o->robust_full_voxel_threshold = 0.5;
    o->robust_empty_voxel_threshold = 0.1;
    o->robust_num_attempts = 3;
    o->robust_max_voxel_neighborhood = 4;
    o->robust_threshold_relative_orientation = 2;
    o->robust_threshold_ego_orientation = 2;
    o->default_motion_model.beta_constant_velocity = 0.000;
    o->default_motion_model.beta_location_consistency = 0.000;
    o->default_motion_model.beta_small_velocity = 0.001;
    o->default_motion_model.beta_orientation_consistency = 0.000;
    auto &c = o->ct_icp_options;
    c.num_iters_icp = 30;
    c.threshold_voxel_occupancy = 5;
    c.max_number_neighbors = 20;
    c.min_number_neighbors = 20;
    c.max_dist_to_plane_ct_icp = 0.5;
    c.threshold_orientation_norm = 0.01;
    c.point_to_plane_with_distortion = 1;
    c.distance = CTICP_DIST_POINT_TO_PLANE;
    c.parametrization = CTICP_PARAM_CONTINUOUS_TIME;
    c.num_closest_neighbors = 1;
    c.loss_function = CTICP_LOSS_CAUCHY;
    c.solver = CTICP_SOLVER_CERES;
    c.ls_max_num_iters = 10;
    c.ls_num_threads = 8;
    c.ls_sigma = 0.2;
    c.ls_tolerant_min_threshold = 0.05;
    c.weight_neighborhood = 0.2;
    c.weight_alpha = 0.8;
    c.weighting_scheme = CTICP_WEIGHT_ALL;
    c.max_num_residuals = 600;
    c.min_num_residuals = 200;
}

/* ---- Odometry ------------------------------------------------------------------------------------------------ */
int orc_odometry_create(const cticp_odometry_options *options, int /*device*/, orc_odometry **out) {
    return Guard([&] {
        auto *h = new orc_odometry();
        h->impl = std::make_unique<Odometry>(*options);
        *out = h;
        return (int) CTICP_OK;
    });
}
void orc_odometry_destroy(orc_odometry *h) { delete h; }

static void FillSummary(const RegistrationSummary &s, cticp_summary *out) {
    std::memset(out, 0, sizeof(*out));
    out->frame = FrameToC(s.frame);
    out->initial_frame = FrameToC(s.initial_frame);
    out->icp_summary.success = s.icp_summary.success;
    out->icp_summary.num_residuals_used = s.icp_summary.num_residuals_used;
    out->icp_summary.num_iters = s.icp_summary.num_iters;
    out->sample_size = s.sample_size;
    out->number_of_residuals = s.number_of_residuals;
    out->robust_level = s.robust_level;
    out->success = s.success;
    out->points_added = s.points_added;
    out->number_of_attempts = s.number_of_attempts;
    out->distance_correction = s.distance_correction;
    out->relative_distance = s.relative_distance;
    out->relative_orientation = s.relative_orientation;
    out->ego_orientation = s.ego_orientation;
    out->num_corrected_points = s.corrected_points.size();
    out->num_all_corrected_points = s.all_corrected_points.size();
    out->num_keypoints = s.keypoints.size();
    auto lv = [&](const char *k) {
        auto it = s.logged_values.find(k);
        return it == s.logged_values.end() ? 0.0 : it->second;
    };
    out->odometry_total = lv("odometry_total");
    out->odometry_initialization = lv("odometry_initialization(ms)");
    out->odometry_try_register = lv("odometry_try_register");
    out->odometry_duration_sampling = lv("odometry_duration_sampling");
    out->odometry_map_update = lv("odometry_map_update(ms)");
    out->odometry_transform = lv("odometry_transform(ms)");
    std::string msg = s.error_message.empty() ? s.icp_summary.error_log : s.error_message;
    std::snprintf(out->error_message, sizeof(out->error_message), "%s", msg.c_str());
}

int orc_odometry_register_frame(orc_odometry *h, const double *xyz, size_t xyz_stride, const double *t,
                                size_t t_stride, size_t n, uint32_t frame_id, const cticp_frame *initial_estimate,
                                cticp_summary *out_summary) {
    return Guard([&] {
        std::vector<Vec3> pts(n);
        std::vector<double> ts(n);
        for (size_t i = 0; i < n; ++i) {
            const double *p = StrideAt(xyz, xyz_stride, i);
            pts[i] = Vec3(p[0], p[1], p[2]);
            ts[i] = *StrideAt(t, t_stride, i);
        }
        TrajectoryFrame init;
        if (initial_estimate) init = FrameFromC(*initial_estimate);
        h->last = h->impl->RegisterFrame(pts, ts, frame_id, initial_estimate ? &init : nullptr);
        if (out_summary) FillSummary(h->last, out_summary);
        return (int) CTICP_OK;
    });
}
int64_t orc_odometry_get_points(orc_odometry *h, int which, cticp_wpoint *dst, size_t cap) {
    const std::vector<WPoint3D> *src = nullptr;
    switch (which) {
        case CTICP_POINTS_CORRECTED: src = &h->last.corrected_points; break;
        case CTICP_POINTS_ALL_CORRECTED: src = &h->last.all_corrected_points; break;
        case CTICP_POINTS_KEYPOINTS: src = &h->last.keypoints; break;
        default: return Fail(CTICP_ERR_INVALID_ARGUMENT, "which");
    }
    size_t m = std::min(cap, src->size());
    for (size_t i = 0; i < m; ++i) dst[i] = WPointToC((*src)[i]);
    return (int64_t) src->size();
}
int64_t orc_odometry_trajectory(orc_odometry *h, cticp_frame *dst, size_t cap) {
    const auto &tr = h->impl->Trajectory();
    size_t m = std::min(cap, tr.size());
    for (size_t i = 0; i < m; ++i) dst[i] = FrameToC(tr[i]);
    return (int64_t) tr.size();
}
int64_t orc_odometry_map_size(orc_odometry *h) { return (int64_t) h->impl->MapSize(); }
int orc_odometry_reset(orc_odometry *h) {
    h->impl->Reset();
    return CTICP_OK;
}
// oracle-only: counters of the last ICP for the bench's algorithmic-bytes figure
void orc_odometry_last_counters(orc_odometry *h, uint64_t *keypoint_iterations, uint64_t *stencil_points) {
    *keypoint_iterations = h->last.icp_summary.keypoint_iterations;
    *stencil_points = h->last.icp_summary.stencil_points;
}

/* ---- Map ----------------------------------------------------------------------------------------------------- */
int orc_map_create(const cticp_map_options *options, int /*device*/, orc_map **out) {
    auto *m = new orc_map();
    m->impl = std::make_shared<VoxelMap>(*options);
    *out = m;
    return CTICP_OK;
}
void orc_map_destroy(orc_map *m) { delete m; }
// borrowed view on the odometry's map (caller must orc_map_destroy the wrapper only)
orc_map *orc_odometry_map(orc_odometry *h) {
    auto *m = new orc_map();
    m->impl = h->impl->GetMapPointer();
    return m;
}
int orc_map_insert_from(orc_map *m, const double *xyz, size_t stride, size_t n, const double origin[3]) {
    std::vector<Vec3> pts(n);
    for (size_t i = 0; i < n; ++i) {
        const double *p = StrideAt(xyz, stride, i);
        pts[i] = Vec3(p[0], p[1], p[2]);
    }
    m->impl->InsertPoints(pts, origin ? Vec3(origin[0], origin[1], origin[2]) : Vec3());
    return CTICP_OK;
}
int orc_map_insert(orc_map *m, const double *xyz, size_t stride, size_t n) {
    return orc_map_insert_from(m, xyz, stride, n, nullptr);
}
int orc_map_remove_far(orc_map *m, const double location[3], double distance) {
    m->impl->RemoveElementsFarFromLocation(Vec3(location[0], location[1], location[2]), distance);
    return CTICP_OK;
}
int64_t orc_map_num_points(orc_map *m, int map_idx) { return (int64_t) m->impl->NumPoints(map_idx); }
int64_t orc_map_num_voxels(orc_map *m, int map_idx) { return (int64_t) m->impl->NumVoxels(map_idx); }
int64_t orc_map_export(orc_map *m, int map_idx, double *dst_xyz, int32_t *dst_voxel, size_t cap) {
    std::vector<Vec3> pts;
    std::vector<Voxel> vox;
    m->impl->Export(map_idx, pts, vox);
    size_t k = std::min(cap, pts.size());
    for (size_t i = 0; i < k; ++i) {
        if (dst_xyz) { dst_xyz[3 * i] = pts[i].x; dst_xyz[3 * i + 1] = pts[i].y; dst_xyz[3 * i + 2] = pts[i].z; }
        if (dst_voxel) { dst_voxel[3 * i] = vox[i].x; dst_voxel[3 * i + 1] = vox[i].y; dst_voxel[3 * i + 2] = vox[i].z; }
    }
    return (int64_t) pts.size();
}
int orc_map_compute_neighborhoods(orc_map *m, const double *q, size_t n, int max_num_neighbors, double *out_points,
                                  int32_t *out_counts) {
    for (size_t i = 0; i < n; ++i) {
        Neighborhood nb;
        m->impl->ComputeNeighborhoodInPlace(Vec3(q[3 * i], q[3 * i + 1], q[3 * i + 2]), max_num_neighbors, nb);
        out_counts[i] = (int32_t) nb.points.size();
        for (size_t j = 0; j < nb.points.size(); ++j) {
            double *o = out_points + (i * max_num_neighbors + j) * 3;
            o[0] = nb.points[j].x; o[1] = nb.points[j].y; o[2] = nb.points[j].z;
        }
    }
    return CTICP_OK;
}
int orc_map_radius_search(orc_map *m, const double *q, const double *radiuses, size_t n, int max_num_neighbors,
                          const double *sensor_location, double *out_points, int32_t *out_counts) {
    Vec3 sensor;
    if (sensor_location) sensor = Vec3(sensor_location[0], sensor_location[1], sensor_location[2]);
    for (size_t i = 0; i < n; ++i) {
        Neighborhood nb;
        m->impl->RadiusSearchInPlace(Vec3(q[3 * i], q[3 * i + 1], q[3 * i + 2]), nb, radiuses[i], max_num_neighbors, nullptr,
                                     sensor_location ? &sensor : nullptr);
        out_counts[i] = (int32_t) nb.points.size();
        for (size_t j = 0; j < nb.points.size(); ++j) {
            double *o = out_points + (i * max_num_neighbors + j) * 3;
            o[0] = nb.points[j].x; o[1] = nb.points[j].y; o[2] = nb.points[j].z;
        }
    }
    return CTICP_OK;
}
int orc_map_clear(orc_map *m) {
    m->impl->Clear();
    return CTICP_OK;
}

/* ---- Registration -------------------------------------------------------------------------------------------- */
static MotionModel MakeMotionModel(const cticp_frame *previous_frame, const cticp_motion_model_options *mo) {
    MotionModel mm;
    if (previous_frame && mo) {
        mm.present = true;
        mm.options = *mo;
        mm.previous_frame = FrameFromC(*previous_frame);
    }
    return mm;
}
int orc_icp_register(orc_map *m, const cticp_icp_options *options, const cticp_strategy_options *strategy,
                     cticp_wpoint *keypoints, size_t n, cticp_frame *frame, const cticp_frame *previous_frame,
                     const cticp_motion_model_options *motion_options, cticp_icp_summary *out_summary) {
    return Guard([&] {
        std::vector<WPoint3D> kpts(n);
        for (size_t i = 0; i < n; ++i) kpts[i] = WPointFromC(keypoints[i]);
        TrajectoryFrame f = FrameFromC(*frame);
        MotionModel mm = MakeMotionModel(previous_frame, motion_options);
        cticp_strategy_options st = strategy ? *strategy : cticp_strategy_options{0, 20, 8, 0, 60., 0.1, 2.0, 1.0};
        ICPSummary s = Register(*m->impl, *options, st, kpts, f, mm.present ? &mm : nullptr);
        for (size_t i = 0; i < n; ++i) keypoints[i] = WPointToC(kpts[i]);
        *frame = FrameToC(f);
        if (out_summary) {
            std::memset(out_summary, 0, sizeof(*out_summary));
            out_summary->success = s.success;
            out_summary->num_residuals_used = s.num_residuals_used;
            out_summary->num_iters = s.num_iters;
        }
        if (!s.success) g_last_error = s.error_log;
        return (int) CTICP_OK;
    });
}
int orc_icp_gn_normal_equations(orc_map *m, const cticp_icp_options *options, const cticp_wpoint *keypoints, size_t n,
                                const cticp_frame *frame, const cticp_frame *previous_frame,
                                const cticp_motion_model_options *motion_options, double *out_A144, double *out_b12,
                                int32_t *out_num_used) {
    return Guard([&] {
        std::vector<WPoint3D> kpts(n);
        for (size_t i = 0; i < n; ++i) kpts[i] = WPointFromC(keypoints[i]);
        TrajectoryFrame f = FrameFromC(*frame);
        MotionModel mm = MakeMotionModel(previous_frame, motion_options);
        GNLinearSystem sys;
        GNBuildSystem(*m->impl, *options, kpts, f, mm.present ? &mm : nullptr, sys);
        for (int i = 0; i < 12; ++i) {
            out_b12[i] = sys.b[i];
            for (int j = 0; j < 12; ++j) out_A144[i * 12 + j] = sys.A[i][j];
        }
        *out_num_used = sys.num_used;
        return (int) CTICP_OK;
    });
}

/* ---- Sampling / order contract ------------------------------------------------------------------------------- */
int64_t orc_grid_sample_indices(int /*device*/, const double *xyz, size_t stride, size_t n, double voxel_size,
                                uint32_t *out_indices, size_t cap) {
    std::vector<WPoint3D> frame(n);
    for (size_t i = 0; i < n; ++i) {
        const double *p = StrideAt(xyz, stride, i);
        frame[i].raw = Vec3(p[0], p[1], p[2]);
    }
    auto kept = SubSampleIndices(frame, voxel_size);
    for (size_t i = 0; i < std::min(cap, kept.size()); ++i) out_indices[i] = kept[i];
    return (int64_t) kept.size();
}
int64_t orc_adaptive_sample_indices(int /*device*/, const cticp_adaptive_options *options, const double *xyz, size_t stride,
                                    size_t n, uint32_t *out_indices, size_t cap) {
    int64_t total = 0;
    int rc = Guard([&] {
        std::vector<Vec3> pts(n);
        for (size_t i = 0; i < n; ++i) {
            const double *p = StrideAt(xyz, stride, i);
            pts[i] = Vec3(p[0], p[1], p[2]);
        }
        auto kept = AdaptiveSampleIndices(pts, *options);
        for (size_t i = 0; i < std::min(cap, kept.size()); ++i) out_indices[i] = kept[i];
        total = (int64_t) kept.size();
        return (int) CTICP_OK;
    });
    return rc < 0 ? rc : total;
}
int orc_permutation(uint64_t seed, uint64_t counter, uint32_t n, uint32_t *out_perm) {
    Permutation perm(seed, counter, n);
    for (uint32_t i = 0; i < n; ++i) out_perm[i] = perm(i);
    return CTICP_OK;
}

/* ---- KAT taps for the reference's own property tests --------------------------------------------- */
// test/unit/SlamCore/test_neighborhood.cxx:40-53 — normal / a2D of a point set; returns is_valid
int orc_neighborhood_describe(const double *xyz, size_t n, double normal[3], double *a2D, double *planarity,
                              double *linearity, double cov9[9]) {
    Neighborhood nb;
    for (size_t i = 0; i < n; ++i) nb.points.emplace_back(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    nb.ComputeNeighborhood();
    if (!nb.is_valid) return 0;
    for (int d = 0; d < 3; ++d) normal[d] = nb.description.normal[d];
    *a2D = nb.description.a2D;
    *planarity = nb.description.planarity;
    *linearity = nb.description.linearity;
    if (cov9)
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) cov9[3 * i + j] = nb.description.covariance(i, j);
    return 1;
}
// include/SlamCore/types.h:455-470 + :354-357 — world = InterpolatePose(begin, end, t) * raw
int orc_pose_transform(const cticp_frame *frame, const double raw[3], double timestamp, double out_world[3]) {
    return Guard([&] {
        TrajectoryFrame f = FrameFromC(*frame);
        Vec3 w = f.begin_pose.InterpolatePose(f.end_pose, timestamp) * Vec3(raw[0], raw[1], raw[2]);
        out_world[0] = w.x; out_world[1] = w.y; out_world[2] = w.z;
        return (int) CTICP_OK;
    });
}
// include/SlamCore/types.h:327-351 — SE3 inverse / product for test/unit/SlamCore/test_types.cxx:7-86
void orc_se3_inverse(const double q[4], const double t[3], double oq[4], double ot[3]) {
    SE3 s;
    s.quat = Quat(q[0], q[1], q[2], q[3]);
    s.tr = Vec3(t[0], t[1], t[2]);
    SE3 r = s.Inverse();
    oq[0] = r.quat.x; oq[1] = r.quat.y; oq[2] = r.quat.z; oq[3] = r.quat.w;
    ot[0] = r.tr.x; ot[1] = r.tr.y; ot[2] = r.tr.z;
}
void orc_se3_mul(const double qa[4], const double ta[3], const double qb[4], const double tb[3], double oq[4],
                 double ot[3]) {
    SE3 a, b;
    a.quat = Quat(qa[0], qa[1], qa[2], qa[3]); a.tr = Vec3(ta[0], ta[1], ta[2]);
    b.quat = Quat(qb[0], qb[1], qb[2], qb[3]); b.tr = Vec3(tb[0], tb[1], tb[2]);
    SE3 r = a * b;
    oq[0] = r.quat.x; oq[1] = r.quat.y; oq[2] = r.quat.z; oq[3] = r.quat.w;
    ot[0] = r.tr.x; ot[1] = r.tr.y; ot[2] = r.tr.z;
}
// TSE3::Interpolate, include/SlamCore/types.h:361-366 (slerp NOT renormalised + lerp)
void orc_se3_interpolate(const double qa[4], const double ta[3], const double qb[4], const double tb[3], double weight,
                         double oq[4], double ot[3]) {
    SE3 a, b;
    a.quat = Quat(qa[0], qa[1], qa[2], qa[3]); a.tr = Vec3(ta[0], ta[1], ta[2]);
    b.quat = Quat(qb[0], qb[1], qb[2], qb[3]); b.tr = Vec3(tb[0], tb[1], tb[2]);
    SE3 r = a.Interpolate(b, weight);
    oq[0] = r.quat.x; oq[1] = r.quat.y; oq[2] = r.quat.z; oq[3] = r.quat.w;
    ot[0] = r.tr.x; ot[1] = r.tr.y; ot[2] = r.tr.z;
}
// TSE3::operator*(point), include/SlamCore/types.h:354-357
void orc_se3_apply(const double q[4], const double t[3], const double p[3], double out[3]) {
    SE3 s;
    s.quat = Quat(q[0], q[1], q[2], q[3]);
    s.tr = Vec3(t[0], t[1], t[2]);
    Vec3 r = s * Vec3(p[0], p[1], p[2]);
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
double orc_angular_distance(const double qa[4], const double qb[4]) {
    SE3 a, b;
    a.quat = Quat(qa[0], qa[1], qa[2], qa[3]);
    b.quat = Quat(qb[0], qb[1], qb[2], qb[3]);
    return AngularDistance(a, b);
}

}  // extern "C"

// test/unit/ct_icp/test_cost_functions.cxx:70-105 — CT point-to-plane residual (+ 12 local partials)
namespace orc {
double CTResidualForTest(double alpha, const double ref[3], const double raw[3], const double normal[3], double weight,
                         const double qb[4], const double tb[3], const double qe[4], const double te[3],
                         double *local_jac12);
double CTResidualKindForTest(int kind, double alpha, const double ref[3], const double raw[3], const double dir[3],
                             const double *covariance, double weight, const double qb[4], const double tb[3],
                             const double qe[4], const double te[3], double *local_jac12);
}
extern "C" double orc_ct_residual(int kind, double alpha, const double ref[3], const double raw[3], const double dir[3],
                                  const double *covariance, double weight, const double qb[4], const double tb[3],
                                  const double qe[4], const double te[3], double *local_jac12) {
    return orc::CTResidualKindForTest(kind, alpha, ref, raw, dir, covariance, weight, qb, tb, qe, te, local_jac12);
}
extern "C" double orc_ct_point_to_plane_residual(double alpha, const double ref[3], const double raw[3],
                                                 const double normal[3], double weight, const double qb[4],
                                                 const double tb[3], const double qe[4], const double te[3],
                                                 double *local_jac12) {
    return orc::CTResidualForTest(alpha, ref, raw, normal, weight, qb, tb, qe, te, local_jac12);
}

/* ---- taps for tests/test_math_pins.py: the out-of-tree arithmetic one function at a time ---- */
namespace orc {
void LossEvaluateForTest(const cticp_icp_options &o, double s, double rho[3]);
void CorrectorForTest(double s, const double rho[3], double *residual_scale, double *jacobian_scale);
void QuatPlusForTest(const double q[4], const double delta[3], double out[4]);
void SolveLMForTest(const cticp_icp_options &o, int n, const double *alpha, const double *ref, const double *raw,
                    const double *normal, const double *weight, double *x, double out[5]);
}
extern "C" void orc_loss_evaluate(const cticp_icp_options *o, double s, double rho[3]) { orc::LossEvaluateForTest(*o, s, rho); }
extern "C" void orc_corrector(double s, const double rho[3], double *residual_scale, double *jacobian_scale) {
    orc::CorrectorForTest(s, rho, residual_scale, jacobian_scale);
}
extern "C" void orc_quat_plus(const double q[4], const double delta[3], double out[4]) { orc::QuatPlusForTest(q, delta, out); }
extern "C" void orc_lm_solve_plane_blocks(const cticp_icp_options *o, int n, const double *alpha, const double *ref,
                                          const double *raw, const double *normal, const double *weight, double *x14,
                                          double out5[5]) {
    orc::SolveLMForTest(*o, n, alpha, ref, raw, normal, weight, x14, out5);
}
// Eigen Quaternion(Matrix3) (ct_icp.cpp:950-954), row-major 3x3 in
extern "C" void orc_quat_from_matrix(const double R[9], double q[4]) {
    orc::Mat3 m;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) m(i, j) = R[3 * i + j];
    const orc::Quat r = orc::Quat::fromRotationMatrix(m);
    q[0] = r.x; q[1] = r.y; q[2] = r.z; q[3] = r.w;
}
extern "C" void orc_quat_to_matrix(const double q[4], double R[9]) {
    const orc::Mat3 m = orc::Quat(q[0], q[1], q[2], q[3]).toRotationMatrix();
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) R[3 * i + j] = m(i, j);
}
// JacobiSVD<Matrix3d>(C, ComputeFullV) stand-in (neighborhood.h:293): singular values descending, V row-major
extern "C" void orc_symmetric_svd3(const double C9[9], double sv[3], double V9[9]) {
    orc::Mat3 c, v;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) c(i, j) = C9[3 * i + j];
    orc::SymmetricSVD3(c, sv, v);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) V9[3 * i + j] = v(i, j);
}
// Matrix<12,12>.ldlt().solve(b) stand-in (ct_icp.cpp:914)
extern "C" void orc_ldlt_solve12(const double A144[144], const double b12[12], double x12[12]) {
    double A[12][12];
    std::array<double, 12> b;
    for (int i = 0; i < 12; ++i) {
        b[i] = b12[i];
        for (int j = 0; j < 12; ++j) A[i][j] = A144[12 * i + j];
    }
    const std::array<double, 12> x = orc::LDLTSolve<12>(A, b);
    for (int i = 0; i < 12; ++i) x12[i] = x[i];
}
