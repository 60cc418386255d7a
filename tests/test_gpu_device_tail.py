"""The tail of RegisterFrame decided on the device (ct_icp_b200/csrc/frame_policy.h): AssessRegistration
(src/ct_icp/odometry.cpp:604-684) and UpdateMap's insertion policy (:855-953) evaluated by k_frame_policy, the map update
launched speculatively behind the ICP kernel. Checked against the host-side tail of the same engine (CTICP_DEVICE_TAIL=0,
bit-identical) and against the CPU oracle on every branch of the policy."""
import os

import numpy as np
import pytest

from conftest import frame_diff
from ct_icp_b200 import _abi as abi
from test_gpu_parity import POSE_TOL_M, POSE_TOL_RAD, _run_sequence, _sequence_options

pytestmark = pytest.mark.gpu


class _env:
    def __init__(self, **kv):
        self.kv = kv

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.kv}
        os.environ.update({k: str(v) for k, v in self.kv.items()})

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _pose(sm):
    f = sm.frame
    return np.array(list(f.begin_pose.tr) + list(f.begin_pose.quat) + list(f.end_pose.tr) + list(f.end_pose.quat))


def _same_run(a, b):
    for i, ((sa, ma), (sb, mb)) in enumerate(zip(a, b)):
        assert sa.success == sb.success and sa.points_added == sb.points_added, i
        assert sa.num_keypoints == sb.num_keypoints and sa.num_corrected_points == sb.num_corrected_points, i
        assert sa.number_of_residuals == sb.number_of_residuals, i
        assert ma == mb, i
        assert np.array_equal(_pose(sa), _pose(sb)), i   # same kernels, same inputs, fixed-order reductions: bit-identical


@pytest.mark.parametrize("solver", ["GN", "CERES"])
def test_device_tail_equals_host_tail(eng, seq_small, solver):
    """Poses, counters and map sizes of a sequence: device-side tail == host-side tail == deferred tail, bit for bit."""
    with _env(CTICP_DEVICE_TAIL=1):
        _, dev = _run_sequence(eng, seq_small, solver=solver, init_num_frames=3)
    with _env(CTICP_DEVICE_TAIL=0):
        _, host = _run_sequence(eng, seq_small, solver=solver, init_num_frames=3)
    with _env(CTICP_DEVICE_TAIL=1, CTICP_TAIL_ROOM=16):   # every frame has more than 16 points: always deferred to the host
        _, deferred = _run_sequence(eng, seq_small, solver=solver, init_num_frames=3)
    with _env(CTICP_DEVICE_TAIL=1, CTICP_TAIL_IN_KERNEL=1):   # solver GN: the verdict written by k_gn_persistent's solver CTA
        _, in_kernel = _run_sequence(eng, seq_small, solver=solver, init_num_frames=3)
    with _env(CTICP_SAMPLE_PRECLEAR=0):   # every k_sample_fused launch clears its own grid / flags
        _, own_clear = _run_sequence(eng, seq_small, solver=solver, init_num_frames=3)
    _same_run(dev, host)
    _same_run(dev, deferred)
    _same_run(dev, in_kernel)
    _same_run(dev, own_clear)
    assert all(s.success for s, _ in dev)


POLICIES = {
    # frames whose ego rotation exceeds the threshold are not inserted until enough frames were skipped (odometry.cpp:912-921)
    "skip_on_ego_rotation": dict(insertion_ego_rotation_threshold=1e-4, insertion_threshold_frames_skipped=2),
    "do_no_insert": dict(do_no_insert=1),
    "always_insert": dict(insertion_ego_rotation_threshold=1e-4, insertion_threshold_frames_skipped=100, always_insert=1),
    # a frame that moves more than 1 cm fails the assessment: success = false, the map is still updated (quit_on_error = 0)
    "assessment_fails": dict(distance_error_threshold=0.01, quit_on_error=0),
    "orientation_fails": dict(orientation_error_threshold=1e-3, quit_on_error=0),
}


@pytest.mark.parametrize("policy", sorted(POLICIES))
def test_device_tail_policy_branches_match_oracle(orc, eng, seq_small, policy):
    kw = dict(POLICIES[policy], init_num_frames=3)
    _, ro = _run_sequence(orc, seq_small, **kw)
    _, re_ = _run_sequence(eng, seq_small, **kw)
    with _env(CTICP_DEVICE_TAIL=0):
        _, rh = _run_sequence(eng, seq_small, **kw)
    _same_run(re_, rh)
    for i, ((so, mo), (se, me)) in enumerate(zip(ro, re_)):
        assert bool(so.success) == bool(se.success), i
        assert so.points_added == se.points_added, i
        assert so.num_keypoints == se.num_keypoints, i
        assert mo == me, i
        dt, dr = frame_diff(so.frame, se.frame)
        assert dt < POSE_TOL_M and dr < POSE_TOL_RAD, (i, dt, dr)
    added = [bool(s.points_added) for s, _ in re_]
    ok = [bool(s.success) for s, _ in re_]
    if policy == "skip_on_ego_rotation":
        assert not all(added[1:]) and any(added[1:])      # both outcomes of the policy occurred
    if policy in ("assessment_fails", "orientation_fails"):
        assert not all(ok[1:])                            # the assessment did fail somewhere
    if policy == "do_no_insert":
        sizes = [m for _, m in re_]
        assert sizes[-1] <= sizes[0]


def test_device_tail_quit_on_error(orc, eng, seq_small):
    """A failed assessment with quit_on_error: no map update, the trajectory keeps the initial estimate — and the call
    reports the failure on both arms."""
    kw = dict(distance_error_threshold=0.01, quit_on_error=1, init_num_frames=3)
    runs = []
    for b in (orc, eng):
        od = b.odometry(_sequence_options(b, "GN", **kw))
        sizes, oks = [], []
        for s in seq_small[:5]:
            sm = od.RegisterFrame(s["xyz"], s["t"], s["frame_idx"])
            sizes.append(od.MapSize())
            oks.append(bool(sm.success))
        runs.append((sizes, oks))
    assert runs[0] == runs[1]
    assert not all(runs[1][1][1:])


def test_sampler_preclear_with_varying_scan_sizes(orc, eng, seq_small):
    """k_sample_fused leaves its hash grid and flag arrays clean for the NEXT frame (frame_pipeline.cu): scans whose size
    jumps up and down between frames take both paths (pre-cleared / own clear, a larger grid than the one left clean) and
    must keep selecting exactly the oracle's points."""
    strides = [1, 3, 1, 2, 1, 4, 1, 1]
    runs = []
    for b in (orc, eng):
        od = b.odometry(_sequence_options(b, "GN", init_num_frames=3))
        out = []
        for s, k in zip(seq_small, strides):
            sm = od.RegisterFrame(np.ascontiguousarray(s["xyz"][::k]), np.ascontiguousarray(s["t"][::k]), s["frame_idx"])
            out.append((sm, od.MapSize()))
        runs.append(out)
    for i, ((so, mo), (se, me)) in enumerate(zip(*runs)):
        assert bool(so.success) == bool(se.success), i
        assert so.num_corrected_points == se.num_corrected_points, i
        assert so.num_keypoints == se.num_keypoints, i
        assert mo == me, i
        if so.success:
            dt, dr = frame_diff(so.frame, se.frame)
            assert dt < POSE_TOL_M and dr < POSE_TOL_RAD, (i, dt, dr)
