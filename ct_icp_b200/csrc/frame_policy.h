// frame_policy.h — the tail of RegisterFrame decided on the device.
//
// The reference decides on the host, after the ICP, whether the registration is acceptable (AssessRegistration,
// src/ct_icp/odometry.cpp:604-684) and whether the frame is inserted into the map (UpdateMap, :855-953). Every input of
// those decisions is either on the device when the ICP loop ends (the pose pair, the failure flag, the frame's point count)
// or known to the host before the loop starts (thresholds, the insertion tracker, the previous pose). So the engine
// enqueues, right behind the ICP kernel and without waiting for it:
//
//     k_frame_policy      one warp: evaluates the decision, writes a FrameVerdict to device memory AND to mapped pinned host
//                         memory (state + counters + decision + sequence number: the only thing the host waits for)
//     k_map_update_fused  reads the pose pair and the decision from the device copy of the verdict
//
// and then spins on the sequence number. One stream, no copy-engine operation and no host round trip between the ICP loop
// and the map update; the host learns the poses while the map update is already running.
//
// Covers the plain registration path (robust_registration = 0, no callbacks, motion compensation CONTINUOUS, fused map
// update); every other configuration keeps the host-side tail (engine.cu UpdateMap).
#pragma once
#include <cstddef>

#include "icp.h"
#include "se3.cuh"

namespace cticp {

enum FrameAction : int {
    kFrameSkip = 0,       // no map update (quit_on_error after a failed assessment, or an ICP error the host turns into an exception)
    kFrameEvict = 1,      // transform + RemoveElementsFarFromLocation, no insertion
    kFrameInsert = 2,     // transform + eviction + InsertPointCloud
    kFrameDeferred = 3,   // nothing done: the frame has more points than the tables were prepared for — the host runs UpdateMap
};

struct FramePolicyIn {   // host → k_frame_policy (kernel parameter)
    double distance_error_threshold, orientation_error_threshold;
    double relative_orientation;   // previous end pose vs the INITIAL estimate's end pose (odometry.cpp:429-431): host-known
    double insertion_ego_rotation_threshold;
    int quit_on_error;
    int has_insertions;            // tracker: total_insertions > 0
    int skipped_enough;            // tracker: skipped_frames > insertion_threshold_frames_skipped
    int do_no_insert, always_insert;
    int room_for;                  // the map tables have room for this many new points (DeviceMap::EnsureRoomFor)
    unsigned seq;
};

struct FrameVerdict {
    IcpState state;          // final registration state (what the D2H copy of d_state_ used to bring back)
    int counts[4];           // N, F, K of the frame (FramePipeline counters)
    int assess_ok;           // AssessRegistration
    int add_points_policy;   // UpdateMap's decision before the do_no_insert / always_insert overrides (summary.points_added)
    int action;              // FrameAction
    int pad0;
    double ego_orientation, relative_distance;
    SlerpConsts sc;          // of the final pose pair
    unsigned seq;            // written last (host copy: after a system-wide fence)
    unsigned pad1;
};

// What a kernel that ends a registration is given to decide the frame's tail itself (k_gn_persistent: its solver CTA holds
// the final state in shared memory and writes the verdict before the kernel ends — no k_frame_policy launch, and the host
// sees the poses one kernel boundary earlier)
struct FrameTailArgs {
    FramePolicyIn in;
    const int *counts;    // FramePipeline counters (N, F, K, -)
    FrameVerdict *dv;     // device copy
    FrameVerdict *hv;     // mapped pinned host copy (device address)
    int enabled;
};

#ifdef __CUDACC__
// AssessRegistration (odometry.cpp:604-684, the branch without robust_registration) and UpdateMap's insertion policy
// (:903-925) on v.state / v.counts; fills the rest of the verdict. One thread.
__device__ __forceinline__ void frame_policy_decide(FrameVerdict &v, const FramePolicyIn &in) {
    const IcpState &S = v.state;
    const Q4 qb{S.qb[0], S.qb[1], S.qb[2], S.qb[3]}, qe{S.qe[0], S.qe[1], S.qe[2], S.qe[3]};
    const V3 tb{S.tb[0], S.tb[1], S.tb[2]}, te{S.te[0], S.te[1], S.te[2]};
    const double ego = angular_distance_deg(qb, qe);   // EgoAngularDistance(summary.frame)
    const double rel_dist = norm(te - tb);             // summary.relative_distance (odometry.cpp:433)
    bool ok;
    if (rel_dist > in.distance_error_threshold) ok = false;
    else if (in.relative_orientation > in.orientation_error_threshold || ego > in.orientation_error_threshold) ok = false;
    else ok = !S.failed;
    bool add = true;   // UpdateMap, the branch without robust_registration
    if (in.has_insertions) add = (ego > in.insertion_ego_rotation_threshold) ? (in.skipped_enough != 0) : true;
    v.add_points_policy = add ? 1 : 0;
    if (in.do_no_insert) add = false;
    if (in.always_insert) add = true;
    int action;
    if (S.failed == 2 || S.failed == 3) action = kFrameSkip;          // the host raises an exception
    else if (!ok && in.quit_on_error) action = kFrameSkip;            // early return (odometry.cpp:437-441)
    else if (add && v.counts[1] > in.room_for) action = kFrameDeferred;
    else action = add ? kFrameInsert : kFrameEvict;
    v.assess_ok = ok ? 1 : 0;
    v.action = action;
    v.pad0 = 0;
    v.ego_orientation = ego;
    v.relative_distance = rel_dist;
    v.sc = slerp_consts(qb, qe);
    v.seq = in.seq;
    v.pad1 = 0;
}
// One warp: the verdict `v` (shared memory, complete and visible to the warp) → device memory and mapped pinned host memory,
// the sequence number last, after a system-wide fence.
__device__ __forceinline__ void frame_verdict_publish(const FrameVerdict &v, FrameVerdict *dv, FrameVerdict *hv, int lane) {
    constexpr int kWords = (int) (sizeof(FrameVerdict) / sizeof(int));
    constexpr int kSeqWord = (int) (offsetof(FrameVerdict, seq) / sizeof(int));
    const int *src = reinterpret_cast<const int *>(&v);
    int *d = reinterpret_cast<int *>(dv);
    volatile int *h = reinterpret_cast<volatile int *>(hv);
    for (int i = lane; i < kWords; i += 32) {
        d[i] = src[i];
        if (i != kSeqWord) h[i] = src[i];
    }
    __threadfence_system();
    __syncwarp();
    if (lane == 0) h[kSeqWord] = src[kSeqWord];
}
#endif

}  // namespace cticp
