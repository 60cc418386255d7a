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

}  // namespace cticp
