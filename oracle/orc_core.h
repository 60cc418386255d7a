// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_math.h header). PARITY UNPINNED.
//
// orc_core.h — voxel map, neighborhood description, samplers, order-contract permutation.
//   reference: include/ct_icp/map.h, include/SlamCore/experimental/neighborhood.h,
//              src/ct_icp/ct_icp.cpp:65-101, include/SlamCore/types.h:65-86,610-623, src/SlamCore/types.cxx:13-20
#pragma once
#include <cstring>
#include <queue>
#include <set>
#include <tuple>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "../include/cticp.h"
#include "orc_math.h"

namespace orc {

// slam::WPoint3D, include/SlamCore/types.h:35-60
struct WPoint3D {
    Vec3 raw;
    double timestamp = -1;
    Vec3 world;
    uint32_t index_frame = uint32_t(-1);
};

// slam::Voxel, include/SlamCore/types.h:65-86
struct Voxel {
    int x = -1, y = -1, z = -1;
    bool operator==(const Voxel &o) const { return x == o.x && y == o.y && z == o.z; }
    bool operator<(const Voxel &o) const { return x < o.x || (x == o.x && (y < o.y || (y == o.y && z < o.z))); }
    // src/SlamCore/types.cxx:13-20 — C int() TRUNCATION toward zero, not floor
    static Voxel Coordinates(const Vec3 &p, double voxel_size) {
        Voxel v;
        v.x = int(p.x / voxel_size);
        v.y = int(p.y / voxel_size);
        v.z = int(p.z / voxel_size);
        return v;
    }
};
// std::hash<slam::Voxel>, include/SlamCore/types.h:610-623
struct VoxelHash {
    size_t operator()(const Voxel &v) const {
        const size_t kP1 = 73856093, kP2 = 19349669, kP3 = 83492791;
        return v.x * kP1 + v.y * kP2 + v.z * kP3;
    }
};

/* ------------------------------------------------------------------------------------------------------------ */
// Order contract.  The reference orders points with std::shuffle(std::mt19937_64) and tsl::robin_map iteration
// order (src/ct_icp/odometry.cpp:349,361,550; src/ct_icp/ct_icp.cpp:79-82).  Neither is reproducible on a GPU,
// so the engine AND this oracle use a counter-based bijection instead:
//   shuffled[perm(i)] = original[i],  perm = 4-round Feistel network over 2^(2h) >= n with cycle walking.
inline uint32_t Mix32(uint32_t h) {
    h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
    return h;
}
inline uint64_t SplitMix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
struct Permutation {
    uint32_t n = 0, half_bits = 1, half_mask = 1;
    uint32_t keys[4] = {0, 0, 0, 0};
    Permutation(uint64_t seed, uint64_t counter, uint32_t n_) : n(n_) {
        uint32_t bits = 2;
        while (bits < 32 && (uint64_t(1) << bits) < uint64_t(n)) bits += 2;
        half_bits = bits / 2;
        half_mask = (1u << half_bits) - 1u;
        uint64_t s = SplitMix64(seed ^ SplitMix64(counter));
        uint64_t s2 = SplitMix64(s);
        keys[0] = uint32_t(s); keys[1] = uint32_t(s >> 32); keys[2] = uint32_t(s2); keys[3] = uint32_t(s2 >> 32);
    }
    uint32_t operator()(uint32_t i) const {
        uint32_t v = i;
        do {
            uint32_t l = v >> half_bits, r = v & half_mask;
            for (int k = 0; k < 4; ++k) {
                uint32_t f = Mix32(r ^ keys[k]) & half_mask;
                uint32_t nl = r;
                r = l ^ f;
                l = nl;
            }
            v = (l << half_bits) | r;
        } while (v >= n);
        return v;
    }
};
template <typename T>
inline void ShuffleInPlace(std::vector<T> &v, uint64_t seed, uint64_t counter) {
    if (v.size() < 2) return;
    Permutation perm(seed, counter, uint32_t(v.size()));
    std::vector<T> out(v.size());
    for (uint32_t i = 0; i < v.size(); ++i) out[perm(i)] = v[i];
    v.swap(out);
}

/* ------------------------------------------------------------------------------------------------------------ */
// ct_icp::sub_sample_frame, src/ct_icp/ct_icp.cpp:65-83 — first-seen point per voxel of RAW coordinates, voxel
// coords cast to short (:70-72).  Output order: order of first appearance (order contract; the reference emits
// tsl::robin_map iteration order, which the second std::shuffle at odometry.cpp:361 scrambles anyway).
struct ShortVoxelHash {
    size_t operator()(const std::array<short, 3> &v) const {
        return VoxelHash()(Voxel{v[0], v[1], v[2]});
    }
};
inline std::vector<uint32_t> SubSampleIndices(const std::vector<WPoint3D> &frame, double size_voxel) {
    std::unordered_set<std::array<short, 3>, ShortVoxelHash> grid;
    grid.reserve(frame.size() / 4 + 1);
    std::vector<uint32_t> kept;
    for (uint32_t i = 0; i < frame.size(); ++i) {
        std::array<short, 3> key = {static_cast<short>(frame[i].raw.x / size_voxel),
                                    static_cast<short>(frame[i].raw.y / size_voxel),
                                    static_cast<short>(frame[i].raw.z / size_voxel)};
        if (grid.insert(key).second) kept.push_back(i);
    }
    return kept;
}
inline void sub_sample_frame(std::vector<WPoint3D> &frame, double size_voxel) {
    auto kept = SubSampleIndices(frame, size_voxel);
    std::vector<WPoint3D> out;
    out.reserve(kept.size());
    for (auto i : kept) out.push_back(frame[i]);
    frame.swap(out);
}
// ct_icp::AdaptiveSamplePointsInGrid, include/ct_icp/algorithm/sampling.h:55-110 (num_points_per_voxel == 1).
// Output: band by band; inside a band the order of first appearance (order contract; the reference emits
// std::unordered_map iteration order). max_num_points reproduces the reference's `size() > max` test (:96-105), which
// lets max + 1 indices through.
inline std::vector<uint32_t> AdaptiveSampleIndices(const std::vector<Vec3> &points, const cticp_adaptive_options &o) {
    if (o.num_points_per_voxel != 1) throw std::runtime_error("oracle: adaptive sampling with num_points_per_voxel != 1");
    const int nb = o.num_bands;
    std::vector<std::unordered_set<Voxel, VoxelHash>> seen(nb);
    std::vector<std::vector<uint32_t>> per_band(nb);
    for (uint32_t idx = 0; idx < points.size(); ++idx) {
        const Vec3 &p = points[idx];
        const double dist = p.norm();
        int lw = 0;   // std::lower_bound with comp(elem, v) = elem.first < v
        while (lw < nb && o.distance[lw] < dist) ++lw;
        if (dist >= o.distance[0] && dist < o.distance[nb - 1]) {
            const int band = lw - 1;
            // NB when dist == distance[0] exactly, lower_bound returns 0 and the reference indexes [-1] (UB); skipped here
            if (band < 0) continue;
            const Voxel v = Voxel::Coordinates(p, o.voxel_size[band]);
            if (seen[band].insert(v).second) per_band[band].push_back(idx);
        }
    }
    const size_t kMax = o.max_num_points > 0 ? (size_t) o.max_num_points : std::numeric_limits<size_t>::max();
    std::vector<uint32_t> out;
    for (auto &b : per_band)
        for (auto i : b) {
            if (out.size() > kMax) return out;
            out.push_back(i);
        }
    return out;
}

// ct_icp::grid_sampling, src/ct_icp/ct_icp.cpp:86-101
inline void grid_sampling(const std::vector<WPoint3D> &frame, std::vector<WPoint3D> &keypoints, double size_voxel) {
    keypoints = frame;
    sub_sample_frame(keypoints, size_voxel);
}

/* ------------------------------------------------------------------------------------------------------------ */
// slam::NeighborhoodDescription + TNeighborhood::ComputeNeighborhood,
// include/SlamCore/experimental/neighborhood.h:226-257, 286-316
struct NeighborhoodDescription {
    double planarity = -1, linearity = -1, a2D = -1;
    Vec3 line, normal, barycenter;
    Mat3 covariance;
};
struct Neighborhood {
    std::vector<Vec3> points;
    NeighborhoodDescription description;
    bool is_valid = false;
    void ComputeNeighborhood() {
        if (points.size() < 5) {   // MinNeighborhoodSize, neighborhood.h:184,227
            is_valid = false;
            return;
        }
        Vec3 bary;
        Mat3 cov;
        for (auto &p : points) {   // uncentered accumulation, neighborhood.h:237-241
            bary += p;
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) cov(i, j) += p[i] * p[j];
        }
        double n = double(points.size());
        bary = bary * (1.0 / n);
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) cov(i, j) = cov(i, j) / n - bary[i] * bary[j];
        double sv[3];
        Mat3 V;
        SymmetricSVD3(cov, sv, V);
        description.covariance = cov;
        description.barycenter = bary;
        description.line = Vec3(V(0, 0), V(1, 0), V(2, 0));
        description.normal = Vec3(V(0, 2), V(1, 2), V(2, 2));
        description.linearity = (sv[0] - sv[1]) / sv[0];
        description.planarity = (sv[1] - sv[2]) / sv[0];
        description.a2D = (std::sqrt(sv[1]) - std::sqrt(sv[2])) / std::sqrt(sv[0]);
        is_valid = true;
    }
};

/* ------------------------------------------------------------------------------------------------------------ */
// ct_icp::MultipleResolutionVoxelMap, include/ct_icp/map.h:99-606
class VoxelMap {
public:
    struct Block {
        std::vector<Vec3> points;
        // PointType::{normal, is_normal_computed (= is_normal_oriented here), frame_id}, map.h:545-554
        std::vector<Vec3> normals;
        std::vector<char> has_normal;
        std::vector<size_t> frame_ids;
    };
    struct SearchParams {
        double radius = 0.5, voxel_resolution = 0;
        size_t map_id = 0;
        int voxel_neighborhood = 1;
    };
    explicit VoxelMap(const cticp_map_options &o) : options_(o) { maps_.resize(o.num_resolutions); }

    const cticp_map_options &Options() const { return options_; }

    // InsertPointInVoxelMap, map.h:261-293
    bool InsertPointInVoxelMap(const Vec3 &point, size_t map_index, size_t frame_idx = 0, Voxel *out_voxel = nullptr) {
        const auto &rp = options_.resolutions[map_index];
        auto &hm = maps_[map_index];
        Voxel voxel = Voxel::Coordinates(point, rp.resolution);
        if (out_voxel) *out_voxel = voxel;
        auto push = [&](Block &blk) {
            blk.points.push_back(point);
            blk.normals.push_back(Vec3());
            blk.has_normal.push_back(0);
            blk.frame_ids.push_back(frame_idx);
            hm.num_points++;
        };
        auto it = hm.map.find(voxel);
        if (it == hm.map.end()) {
            auto &blk = hm.map[voxel];
            blk.points.reserve(rp.max_num_points);
            push(blk);
            return true;
        }
        auto &blk = it->second;
        if ((int) blk.points.size() < rp.max_num_points) {
            double sq_dist_min = std::numeric_limits<double>::max();
            for (auto &q : blk.points) {
                double sq = (q - point).squaredNorm();
                if (sq < sq_dist_min) sq_dist_min = sq;
            }
            if (sq_dist_min > rp.min_distance_between_points * rp.min_distance_between_points) {
                push(blk);
                return true;
            }
        }
        return false;
    }
    // InsertPointCloud, map.h:153-254: world points, sequential in the given order; then every voxel that received a
    // point and holds >= 5 gets ONE normal (V.col(2) of its points' covariance) copied to all its points and oriented,
    // point by point, against the begin position of the frame that point came from (:211-235). `origin` is
    // frame_poses.front().tr of the inserted frame.
    void InsertPoints(const std::vector<Vec3> &world_points, const Vec3 &origin = Vec3()) {
        const size_t fidx = frame_origins_.size();
        frame_origins_.push_back(origin);
        std::vector<std::set<Voxel>> touched(maps_.size());
        for (auto &p : world_points)
            for (size_t m = 0; m < maps_.size(); ++m) {
                Voxel v;
                if (InsertPointInVoxelMap(p, m, fidx, &v)) touched[m].insert(v);
            }
        for (size_t m = 0; m < maps_.size(); ++m)
            for (auto &v : touched[m]) {
                Block &blk = maps_[m].map[v];
                if (blk.points.size() < 5) continue;
                Neighborhood nb;
                nb.points = blk.points;
                nb.ComputeNeighborhood();
                for (size_t i = 0; i < blk.points.size(); ++i) {
                    Vec3 n = nb.description.normal;
                    if ((blk.points[i] - frame_origins_[blk.frame_ids[i]]).dot(n) > 0.) n = n * -1.0;
                    blk.normals[i] = n;
                    blk.has_normal[i] = 1;
                }
            }
    }
    // RemoveElementsFarFromLocation, map.h:305-322 — tests the voxel's FIRST stored point
    void RemoveElementsFarFromLocation(const Vec3 &location, double distance) {
        for (auto &hm : maps_) {
            std::vector<Voxel> to_remove;
            for (auto &[voxel, blk] : hm.map) {
                if (blk.points.empty() || (blk.points.front() - location).norm() > distance) to_remove.push_back(voxel);
            }
            for (auto &v : to_remove) {
                hm.num_points -= hm.map[v].points.size();
                hm.map.erase(v);
            }
        }
    }
    void Clear() {
        maps_.clear();
        maps_.resize(options_.num_resolutions);
        frame_origins_.clear();
    }
    size_t NumPoints(size_t map_idx = 0) const { return maps_[map_idx].num_points; }   // map.h:345
    size_t NumVoxels(size_t map_idx = 0) const { return maps_[map_idx].map.size(); }

    // SearchParamsFromRadiusSearch, map.h:416-432
    SearchParams SearchParamsFromRadiusSearch(double radius) const {
        SearchParams params;
        int it = 0;   // std::lower_bound with comp(elem, radius) = elem.resolution <= radius
        while (it < options_.num_resolutions && options_.resolutions[it].resolution <= radius) ++it;
        int idx = std::max(0, it - 1);
        params.radius = radius;
        params.map_id = idx;
        params.voxel_resolution = options_.resolutions[idx].resolution;
        params.voxel_neighborhood = (int) std::ceil(radius / params.voxel_resolution);
        return params;
    }

    // RadiusSearchInPlace, map.h:449-514. sensor_location == nullptr on the GN / ROBUST / Default-strategy paths
    // (:762, :1235, neighborhood_strategy.h:77-83); the DistanceBasedStrategy passes the current end translation.
    void RadiusSearchInPlace(const Vec3 &query, Neighborhood &nb, double radius, int max_num_neighbors,
                             size_t *stencil_points = nullptr, const Vec3 *sensor_location = nullptr) const {
        nb.points.resize(0);
        const SearchParams params = SearchParamsFromRadiusSearch(radius);
        const auto &hm = maps_[params.map_id].map;
        const int nbv = params.voxel_neighborhood;
        Voxel voxel = Voxel::Coordinates(query, params.voxel_resolution);
        const int kx = voxel.x, ky = voxel.y, kz = voxel.z;
        using entry_t = std::tuple<double, Vec3, Voxel>;
        struct Cmp {
            bool operator()(const entry_t &l, const entry_t &r) const { return std::get<0>(l) < std::get<0>(r); }
        };
        std::priority_queue<entry_t, std::vector<entry_t>, Cmp> pq;
        for (short kxx = kx - nbv; kxx < kx + nbv + 1; ++kxx)        // loop vars are `short`, map.h:470-472
            for (short kyy = ky - nbv; kyy < ky + nbv + 1; ++kyy)
                for (short kzz = kz - nbv; kzz < kz + nbv + 1; ++kzz) {
                    voxel.x = kxx; voxel.y = kyy; voxel.z = kzz;
                    auto search = hm.find(voxel);
                    if (search == hm.end()) continue;
                    const auto &blk = search->second;
                    if (stencil_points) *stencil_points += blk.points.size();
                    for (size_t i = 0; i < blk.points.size(); ++i) {
                        const Vec3 &nbp = blk.points[i];
                        if (options_.select_valid_normals_direction && sensor_location && blk.has_normal[i]) {   // :482-490
                            const double scalar = (*sensor_location - query).dot(blk.normals[i]);
                            if (scalar < 0.) continue;
                        }
                        double distance = (nbp - query).norm();
                        if (distance > params.radius) continue;
                        if ((int) pq.size() == max_num_neighbors) {
                            if (distance < std::get<0>(pq.top())) {
                                pq.pop();
                                pq.emplace(distance, nbp, voxel);
                            }
                        } else
                            pq.emplace(distance, nbp, voxel);
                    }
                }
        nb.points.reserve(pq.size());
        while (!pq.empty()) {   // farthest first, map.h:508-513
            nb.points.push_back(std::get<1>(pq.top()));
            pq.pop();
        }
    }
    // ComputeNeighborhoodInPlace, map.h:527-530
    void ComputeNeighborhoodInPlace(const Vec3 &query, int max_num_neighbors, Neighborhood &nb,
                                    size_t *stencil_points = nullptr) const {
        RadiusSearchInPlace(query, nb, options_.default_radius, max_num_neighbors, stencil_points);
    }

    // GetMapPoints(map_idx), map.h:354-376 — emitted in sorted voxel order for determinism
    void Export(size_t map_idx, std::vector<Vec3> &pts, std::vector<Voxel> &voxels) const {
        std::vector<Voxel> keys;
        for (auto &[v, b] : maps_[map_idx].map) keys.push_back(v);
        std::sort(keys.begin(), keys.end());
        for (auto &k : keys)
            for (auto &p : maps_[map_idx].map.at(k).points) {
                pts.push_back(p);
                voxels.push_back(k);
            }
    }

private:
    struct HashMap {
        size_t num_points = 0;
        std::unordered_map<Voxel, Block, VoxelHash> map;
    };
    cticp_map_options options_;
    std::vector<HashMap> maps_;
    std::vector<Vec3> frame_origins_;   // frame_id_to_frame[fidx].poses.front().tr (entries are never erased, :246-252)
};

// ANeighborhoodStrategy::ComputeNeighborhoodInPlace, neighborhood_strategy.h:77-83 (nearest) / :121-141 (distance based)
inline bool StrategyComputeNeighborhoodInPlace(const cticp_strategy_options &st, const VoxelMap &map, const Vec3 &raw_point,
                                               const Vec3 &world_point, Neighborhood &nb, const Vec3 *sensor_location,
                                               size_t *stencil_points = nullptr) {
    if (st.type == CTICP_STRATEGY_DISTANCE_BASED) {
        // ComputeRadius (:121-126): the ratio is taken against radius_max (sic — the documentation says distance_max)
        const double alpha = std::pow(std::min(std::abs(raw_point.norm()), st.radius_max) / st.radius_max, st.exponent);
        const double radius = alpha * st.radius_max + (1 - alpha) * st.radius_min;
        map.RadiusSearchInPlace(world_point, nb, radius, st.max_num_neighbors, stencil_points, sensor_location);
        return true;
    }
    map.ComputeNeighborhoodInPlace(world_point, st.max_num_neighbors, nb, stencil_points);
    return (int) nb.points.size() >= st.min_num_neighbors;
}

/* ------------------------------------------------------------------------------------------------------------ */
inline Pose PoseFromC(const cticp_pose &c) {
    Pose p;
    p.pose.quat = Quat(c.quat[0], c.quat[1], c.quat[2], c.quat[3]);
    p.pose.tr = Vec3(c.tr[0], c.tr[1], c.tr[2]);
    p.ref_timestamp = c.ref_timestamp;
    p.dest_timestamp = c.dest_timestamp;
    p.ref_frame_id = c.ref_frame_id;
    p.dest_frame_id = c.dest_frame_id;
    return p;
}
inline cticp_pose PoseToC(const Pose &p) {
    cticp_pose c;
    c.quat[0] = p.pose.quat.x; c.quat[1] = p.pose.quat.y; c.quat[2] = p.pose.quat.z; c.quat[3] = p.pose.quat.w;
    c.tr[0] = p.pose.tr.x; c.tr[1] = p.pose.tr.y; c.tr[2] = p.pose.tr.z;
    c.ref_timestamp = p.ref_timestamp;
    c.dest_timestamp = p.dest_timestamp;
    c.ref_frame_id = p.ref_frame_id;
    c.dest_frame_id = p.dest_frame_id;
    return c;
}
inline TrajectoryFrame FrameFromC(const cticp_frame &c) {
    TrajectoryFrame f;
    f.begin_pose = PoseFromC(c.begin_pose);
    f.end_pose = PoseFromC(c.end_pose);
    return f;
}
inline cticp_frame FrameToC(const TrajectoryFrame &f) {
    cticp_frame c;
    c.begin_pose = PoseToC(f.begin_pose);
    c.end_pose = PoseToC(f.end_pose);
    return c;
}
inline WPoint3D WPointFromC(const cticp_wpoint &c) {
    WPoint3D p;
    p.raw = Vec3(c.raw[0], c.raw[1], c.raw[2]);
    p.timestamp = c.timestamp;
    p.world = Vec3(c.world[0], c.world[1], c.world[2]);
    p.index_frame = c.index_frame;
    return p;
}
inline cticp_wpoint WPointToC(const WPoint3D &p) {
    cticp_wpoint c;
    std::memset(&c, 0, sizeof(c));
    c.raw[0] = p.raw.x; c.raw[1] = p.raw.y; c.raw[2] = p.raw.z;
    c.timestamp = p.timestamp;
    c.world[0] = p.world.x; c.world[1] = p.world.y; c.world[2] = p.world.z;
    c.index_frame = p.index_frame;
    return c;
}

}  // namespace orc
