"""GPU parity tests: the CUDA engine (through the C ABI) against the CPU oracle on identical seeded inputs.

Tolerances: integer / index work (sampling, voxel assignment, neighbor sets) is compared exactly; stored map
points within the fp32 voxel-local storage quantum (2e-7 m); normal equations 1e-6 relative; poses within the
north-star bound 1e-4 m / 1e-4 rad per frame (observed: ~1e-7).
"""
import ctypes as C

import numpy as np
import pytest

from conftest import frame_diff, get_sequence
from ct_icp_b200 import _abi as abi

pytestmark = pytest.mark.gpu

POS_TOL = 2e-7        # fp32 offset from the voxel origin, |offset| < 1.5 m → ulp 1.2e-7
POSE_TOL_M = 1e-4     # BASELINE.json north_star
POSE_TOL_RAD = 1e-4


def small_map_options(b, res=1.0, max_pts=20, min_dist=0.1, cap=1 << 15):
    o = b.legacy_map_options(res, max_pts, min_dist)
    o.capacity_voxels = cap
    return o


def random_cloud(rng, n, extent=30.0):
    # planar-ish structures + noise so voxels fill up and the min-distance rule matters
    pts = rng.uniform(-extent, extent, size=(n, 3))
    pts[: n // 2, 2] = rng.normal(0.0, 0.02, size=n // 2)            # ground
    pts[n // 2: 3 * n // 4, 1] = 10.0 + rng.normal(0.0, 0.02, size=n // 4)   # wall
    return pts.astype(np.float32).astype(np.float64)


# ---------------------------------------------------------------------------------------------------------------
def test_order_contract_permutation(orc, eng):
    for n in (1, 2, 3, 17, 1000, 131072, 132481):
        a = orc.permutation(0x5DEECE66D, 513, n)
        b = eng.permutation(0x5DEECE66D, 513, n)
        assert np.array_equal(a, b)
        assert np.array_equal(np.sort(a), np.arange(n, dtype=np.uint32))


@pytest.mark.parametrize("voxel", [0.2, 0.5, 1.5])
def test_grid_sampling_indices(orc, eng, voxel):
    s = get_sequence("hdl64", 2)[1]
    a = orc.grid_sample_indices(s["xyz"], voxel)
    b = eng.grid_sample_indices(s["xyz"], voxel)
    assert len(a) == len(b) and np.array_equal(a, b)


def test_grid_sampling_edge_cases(orc, eng):
    one = np.array([[1.0, 2.0, 3.0]])
    assert np.array_equal(eng.grid_sample_indices(one, 0.5), [0])
    dup = np.repeat(one, 1000, axis=0)
    assert np.array_equal(eng.grid_sample_indices(dup, 0.5), [0])
    # negative coordinates: truncation toward zero merges (-v, +v) around the origin like the reference's cast
    pts = np.array([[-0.3, 0, 0], [0.3, 0, 0], [-0.7, 0, 0], [0.7, 0, 0]])
    assert np.array_equal(eng.grid_sample_indices(pts, 0.5), orc.grid_sample_indices(pts, 0.5))


def test_map_insert_matches_oracle(orc, eng):
    rng = np.random.default_rng(7)
    mo, me = orc.voxel_map(small_map_options(orc)), eng.voxel_map(small_map_options(eng))
    for batch in range(4):
        pts = random_cloud(rng, 20000)
        mo.insert(pts)
        me.insert(pts)
        assert mo.num_points() == me.num_points()
        assert mo.num_voxels() == me.num_voxels()
    xo, vo = mo.export()
    xe, ve = me.export()
    assert np.array_equal(vo, ve)
    assert np.abs(xo - xe).max() < POS_TOL


def test_map_multi_resolution_insert(orc, eng):
    rng = np.random.default_rng(11)
    oo, oe = orc.default_map_options(), eng.default_map_options()
    oe.capacity_voxels = 1 << 17
    mo, me = orc.voxel_map(oo), eng.voxel_map(oe)
    pts = random_cloud(rng, 30000, extent=12.0)
    mo.insert(pts)
    me.insert(pts)
    for lvl in range(3):
        assert mo.num_points(lvl) == me.num_points(lvl)
        xo, vo = mo.export(lvl)
        xe, ve = me.export(lvl)
        assert np.array_equal(vo, ve)
        assert np.abs(xo - xe).max() < POS_TOL


def test_map_remove_far(orc, eng):
    rng = np.random.default_rng(3)
    mo, me = orc.voxel_map(small_map_options(orc)), eng.voxel_map(small_map_options(eng))
    pts = random_cloud(rng, 30000)
    mo.insert(pts); me.insert(pts)
    for loc, dist in (((5.0, -3.0, 0.5), 25.0), ((20.0, 10.0, 0.0), 12.0)):
        mo.remove_far(loc, dist); me.remove_far(loc, dist)
        assert mo.num_points() == me.num_points() and mo.num_voxels() == me.num_voxels()
        assert np.array_equal(mo.export()[1], me.export()[1])
    # insert again after eviction (tombstones in the table must not break lookups / re-insertion)
    pts2 = random_cloud(rng, 20000)
    mo.insert(pts2); me.insert(pts2)
    xo, vo = mo.export(); xe, ve = me.export()
    assert np.array_equal(vo, ve) and np.abs(xo - xe).max() < POS_TOL
    mo.clear(); me.clear()
    assert me.num_points() == 0 and me.num_voxels() == 0


def test_map_self_nearest_and_full_recall(eng):
    # reference test/unit/SlamCore/test_map.cxx:5-38: kNN=1 returns the point itself; a huge radius returns all
    rng = np.random.default_rng(5)
    pts = rng.uniform(-1, 1, size=(100, 3)).astype(np.float32).astype(np.float64)
    o = eng.legacy_map_options(0.01, 20, 0.0)
    o.capacity_voxels = 1 << 12
    m = eng.voxel_map(o)
    m.insert(pts)
    nb, cnt = m.compute_neighborhoods(pts, 1)
    assert np.all(cnt == 1)
    assert np.abs(nb[:, 0, :] - pts).max() < 1e-5
    o2 = eng.legacy_map_options(1.0, 64, 0.0)
    o2.default_radius = 4.0     # ceil(4/1) = 4 → 729-voxel stencil covers the whole cube
    o2.capacity_voxels = 1 << 12
    m2 = eng.voxel_map(o2)
    m2.insert(pts[:30])
    nb, cnt = m2.compute_neighborhoods(pts[:5], 32)
    assert np.all(cnt == 30)


def test_neighborhoods_match_oracle(orc, eng):
    rng = np.random.default_rng(9)
    mo, me = orc.voxel_map(small_map_options(orc)), eng.voxel_map(small_map_options(eng))
    pts = random_cloud(rng, 60000)
    mo.insert(pts); me.insert(pts)
    q = random_cloud(rng, 4000) + rng.normal(0, 0.05, size=(4000, 3))
    no, co = mo.compute_neighborhoods(q, 20)
    ne, ce = me.compute_neighborhoods(q, 20)
    assert np.array_equal(co, ce)
    assert (co >= 5).sum() > 500
    assert np.abs(no - ne).max() < POS_TOL     # same neighbors in the same (farthest-first) order


def test_neighborhoods_radius2_stencil(orc, eng):
    # nclt-like: resolution 0.5, radius 0.8 → r = 2 (125-voxel stencil, several 32-voxel rounds)
    rng = np.random.default_rng(13)
    oo, oe = small_map_options(orc, res=0.5, max_pts=40, min_dist=0.05), small_map_options(eng, res=0.5, max_pts=40, min_dist=0.05)
    oo.default_radius = oe.default_radius = 0.8
    mo, me = orc.voxel_map(oo), eng.voxel_map(oe)
    pts = random_cloud(rng, 60000, extent=15.0)
    mo.insert(pts); me.insert(pts)
    q = random_cloud(rng, 2000, extent=15.0)
    no, co = mo.compute_neighborhoods(q, 20)
    ne, ce = me.compute_neighborhoods(q, 20)
    assert np.array_equal(co, ce)
    assert np.abs(no - ne).max() < POS_TOL


def _registration_case(orc, eng, solver):
    """Map from frames 0..k-1 at the oracle's poses; register frame k's keypoints from a perturbed initial guess."""
    seq = get_sequence("hdl64", 4)
    o = orc.default_odometry_options()
    o.ct_icp_options.solver = abi.SOLVER["GN"]
    o.ct_icp_options.min_number_neighbors = 10
    o.map_options = orc.legacy_map_options(1.0, 20, 0.1)
    o.debug_print = 0
    od = orc.odometry(o)
    for s in seq[:3]:
        sm = od.RegisterFrame(s["xyz"], s["t"], s["frame_idx"])
        assert sm.success
    map_xyz, _ = od.GetMapPointer().export(0)
    # keypoints of the next frame: grid-sampled raw points, world from the initial estimate
    s = seq[3]
    idx = orc.grid_sample_indices(s["xyz"], 1.0)
    kp = np.zeros(len(idx), dtype=abi.wpoint_dtype())
    kp["raw"] = s["xyz"][idx]
    kp["timestamp"] = s["t"][idx]
    traj = od.Trajectory()
    frame = traj[-1].copy()
    frame.begin_pose = traj[-1].end_pose.copy()
    frame.begin_pose.dest_timestamp = float(s["t"].min())
    frame.end_pose.dest_timestamp = float(s["t"].max())
    return map_xyz, kp, frame, traj[-1]


def _fill_world(orc, kp, frame):
    f = frame
    out = (C.c_double * 3)()
    for i in range(len(kp)):
        raw = (C.c_double * 3)(*kp["raw"][i])
        orc.check(orc.fn("pose_transform")(C.byref(f), raw, float(kp["timestamp"][i]), out))
        kp["world"][i] = out[:]


def test_gn_normal_equations(orc, eng):
    map_xyz, kp, frame, prev = _registration_case(orc, eng, "GN")
    mo, me = orc.voxel_map(small_map_options(orc, cap=1 << 18)), eng.voxel_map(small_map_options(eng, cap=1 << 18))
    mo.insert(map_xyz); me.insert(map_xyz)
    assert mo.num_points() == me.num_points()
    io = orc.default_icp_options()
    io.solver = abi.SOLVER["GN"]
    io.min_number_neighbors = 10
    mm = orc.default_odometry_options().default_motion_model
    _fill_world(orc, kp, frame)
    Ao, bo, no = mo.gn_normal_equations(io, kp, frame, prev, mm)
    Ae, be, ne = me.gn_normal_equations(io, kp.copy(), frame, prev, mm)
    assert no == ne and no > 100
    scale = np.abs(Ao).max()
    assert np.abs(Ao - Ae).max() < 1e-6 * scale
    assert np.abs(bo - be).max() < 1e-6 * max(np.abs(bo).max(), 1e-3)
    assert np.allclose(Ae, Ae.T)


def test_gn_register_matches_oracle(orc, eng):
    map_xyz, kp, frame, prev = _registration_case(orc, eng, "GN")
    mo, me = orc.voxel_map(small_map_options(orc, cap=1 << 18)), eng.voxel_map(small_map_options(eng, cap=1 << 18))
    mo.insert(map_xyz); me.insert(map_xyz)
    io = orc.default_icp_options()
    io.solver = abi.SOLVER["GN"]
    io.min_number_neighbors = 10
    io.num_iters_icp = 8
    mm = orc.default_odometry_options().default_motion_model
    _fill_world(orc, kp, frame)
    kpo, kpe = kp.copy(), kp.copy()
    fo, fe = frame.copy(), frame.copy()
    so = mo.icp_register(io, kpo, fo, prev, mm)
    se = me.icp_register(io, kpe, fe, prev, mm)
    assert so.success and se.success
    assert so.num_residuals_used == se.num_residuals_used
    dt, dr = frame_diff(fo, fe)
    assert dt < POSE_TOL_M and dr < POSE_TOL_RAD, (dt, dr)
    assert dt < 1e-6 and dr < 1e-6, (dt, dr)     # observed level; tightens the north-star bound
    # the registration moved the pose (the test is not vacuous)
    assert frame_diff(fo, frame, log=False)[0] > 1e-3
    assert np.abs(kpo["world"] - kpe["world"]).max() < 1e-5


def test_gn_register_not_enough_keypoints(orc, eng):
    # reference: ct_icp.cpp:860-871 → success=false when fewer than 100 residuals
    me = eng.voxel_map(small_map_options(eng))
    rng = np.random.default_rng(1)
    me.insert(random_cloud(rng, 500))
    kp = np.zeros(50, dtype=abi.wpoint_dtype())
    kp["raw"] = random_cloud(rng, 50)
    kp["timestamp"] = 0.05
    frame = abi.Frame()
    frame.begin_pose = abi.Pose.make(dest_timestamp=0.0, dest_frame_id=1)
    frame.end_pose = abi.Pose.make(dest_timestamp=0.1, dest_frame_id=1)
    io = eng.default_icp_options()
    io.solver = abi.SOLVER["GN"]
    s = me.icp_register(io, kp, frame)
    assert not s.success


def test_timestamp_outside_pose_interval_is_an_error(eng):
    # reference: CHECK in TPose::InterpolatePose (types.h:456) aborts; the ABI returns CTICP_ERR_TIMESTAMP
    from ct_icp_b200 import CticpError
    me = eng.voxel_map(small_map_options(eng))
    kp = np.zeros(10, dtype=abi.wpoint_dtype())
    kp["timestamp"] = 0.5
    frame = abi.Frame()
    frame.begin_pose = abi.Pose.make(dest_timestamp=0.0)
    frame.end_pose = abi.Pose.make(dest_timestamp=0.1)
    io = eng.default_icp_options()
    io.solver = abi.SOLVER["GN"]
    with pytest.raises(CticpError) as e:
        me.icp_register(io, kp, frame)
    assert e.value.code == abi.ERR_TIMESTAMP


def _sequence_options(b, solver="GN", **overrides):
    o = b.default_odometry_options()
    o.ct_icp_options.solver = abi.SOLVER[solver]
    o.ct_icp_options.min_number_neighbors = 10
    o.ct_icp_options.ls_max_num_iters = 5
    o.ct_icp_options.ls_num_threads = 8
    o.map_options = b.legacy_map_options(1.0, 20, 0.1)
    o.debug_print = 0
    for k, v in overrides.items():
        setattr(o, k, v)
    return o


def _run_sequence(b, seq, solver="GN", **overrides):
    od = b.odometry(_sequence_options(b, solver, **overrides))
    out = []
    for s in seq:
        sm = od.RegisterFrame(s["xyz"], s["t"], s["frame_idx"])
        out.append((sm, od.MapSize()))
    return od, out


def test_odometry_sequence_small(orc, eng, seq_small):
    """config 1 stand-in (~10k-pt scans): full RegisterFrame pipeline, GN, frame by frame."""
    odo, ro = _run_sequence(orc, seq_small, init_num_frames=4)
    ode, re_ = _run_sequence(eng, seq_small, init_num_frames=4)
    for (so, mo), (se, me) in zip(ro, re_):
        assert so.success == se.success
        assert so.num_corrected_points == se.num_corrected_points
        assert so.num_keypoints == se.num_keypoints
        assert mo == me
        dt, dr = frame_diff(so.frame, se.frame)
        assert dt < POSE_TOL_M and dr < POSE_TOL_RAD, (dt, dr)
    assert np.array_equal(odo.GetMapPointer().export(0)[1], ode.GetMapPointer().export(0)[1])


def test_odometry_sequence_hdl64_gn(orc, eng, seq_hdl64):
    """config 2 (KITTI-shape 64-beam ~130k-pt scans, CT_ICP_GN point-to-plane): 26 frames incl. the switch from
    the start-up regime (init_num_frames = 20) to the steady state."""
    odo, ro = _run_sequence(orc, seq_hdl64)
    ode, re_ = _run_sequence(eng, seq_hdl64)
    worst_t = worst_r = 0.0
    for i, ((so, mo), (se, me)) in enumerate(zip(ro, re_)):
        assert so.success and se.success, i
        assert so.num_corrected_points == se.num_corrected_points, i
        assert so.num_keypoints == se.num_keypoints, i
        assert so.number_of_residuals == se.number_of_residuals, i
        assert mo == me, i
        dt, dr = frame_diff(so.frame, se.frame)
        worst_t, worst_r = max(worst_t, dt), max(worst_r, dr)
        assert dt < POSE_TOL_M and dr < POSE_TOL_RAD, (i, dt, dr)
    print("worst per-frame pose difference: %.3e m, %.3e rad" % (worst_t, worst_r))
    # corrected points of the last frame
    po = odo.corrected_points()
    pe = ode.corrected_points()
    assert len(po) == len(pe)
    assert np.abs(po["world"] - pe["world"]).max() < 1e-4
    assert np.abs(po["raw"] - pe["raw"]).max() == 0.0
    pa = ode.all_corrected_points()
    assert len(pa) == len(seq_hdl64[-1]["xyz"])
    ka, kb = odo.keypoints(), ode.keypoints()
    assert len(ka) == len(kb) and np.abs(ka["world"] - kb["world"]).max() < 1e-4


def _estimate_from(prev_frame, t_begin, t_end, frame_idx):
    """RegisterFrameWithEstimate input (odometry.cpp:236-248): a pose pair at the previous end pose whose timestamps
    bracket the scan a little wider than its own min / max (alpha is taken against the POSE timestamps)."""
    est = abi.Frame()
    for dst, ts in ((est.begin_pose, t_begin), (est.end_pose, t_end)):
        for i in range(4):
            dst.quat[i] = prev_frame.end_pose.quat[i]
        for i in range(3):
            dst.tr[i] = prev_frame.end_pose.tr[i]
        dst.ref_timestamp, dst.dest_timestamp = 0.0, ts
        dst.ref_frame_id, dst.dest_frame_id = 0, frame_idx
    return est


def _run_with_estimates(b, seq):
    od = b.odometry(_sequence_options(b, "GN", init_num_frames=4))
    out, prev = [], None
    for i, s in enumerate(seq):
        # frames 0 and 1 collapse every timestamp onto the end pose (odometry.cpp:355-359): with a wider pose interval
        # the begin pose would be all but unobservable there, so the estimates start at frame 2
        if i < 2:
            sm = od.RegisterFrame(s["xyz"], s["t"], s["frame_idx"])
        else:
            est = _estimate_from(prev.frame, float(s["t"].min()) - 1e-3, float(s["t"].max()) + 2e-3, s["frame_idx"])
            sm = od.RegisterFrameWithEstimate(s["xyz"], s["t"], est, s["frame_idx"])
        prev = sm
        out.append((sm, od.MapSize()))
    return od, out


def test_odometry_register_frame_with_estimate(orc, eng, seq_small):
    """RegisterFrameWithEstimate (odometry.cpp:236-248): the caller's pose pair replaces InitializeMotion and its
    timestamps — not the scan's min / max — define the per-point alpha."""
    _, ro = _run_with_estimates(orc, seq_small)
    _, re_ = _run_with_estimates(eng, seq_small)
    for i, ((so, mo), (se, me)) in enumerate(zip(ro, re_)):
        assert so.success == se.success, i
        assert so.num_corrected_points == se.num_corrected_points, i
        assert so.num_keypoints == se.num_keypoints, i
        assert mo == me, i
        dt, dr = frame_diff(so.frame, se.frame)
        assert dt < POSE_TOL_M and dr < POSE_TOL_RAD, (i, dt, dr)
        assert se.frame.begin_pose.dest_timestamp == so.frame.begin_pose.dest_timestamp


def test_estimate_not_covering_the_scan_is_an_error_and_recoverable(eng, seq_small):
    """A pose pair that does not bracket the scan's timestamps is CTICP_ERR_TIMESTAMP (the reference CHECK-aborts,
    types.h:456); the scan's upload is already in flight when that is detected, and the handle must stay usable."""
    from ct_icp_b200 import CticpError
    od = eng.odometry(_sequence_options(eng, "GN", init_num_frames=4))
    sms = [od.RegisterFrame(s["xyz"], s["t"], s["frame_idx"]) for s in seq_small[:3]]
    s = seq_small[3]
    bad = _estimate_from(sms[-1].frame, float(s["t"].min()) + 0.01, float(s["t"].max()) + 0.01, s["frame_idx"])
    with pytest.raises(CticpError) as e:
        od.RegisterFrameWithEstimate(s["xyz"], s["t"], bad, s["frame_idx"])
    assert e.value.code == abi.ERR_TIMESTAMP
    od.Reset()
    ref = eng.odometry(_sequence_options(eng, "GN", init_num_frames=4))
    for s in seq_small[:4]:
        a = od.RegisterFrame(s["xyz"], s["t"], s["frame_idx"])
        b = ref.RegisterFrame(s["xyz"], s["t"], s["frame_idx"])
        assert a.success and b.success
        assert frame_diff(a.frame, b.frame) == (0.0, 0.0)


def test_odometry_reset(eng, seq_small):
    od, r1 = _run_sequence(eng, seq_small[:4], init_num_frames=2)
    p1 = [list(s.frame.end_pose.tr) for s, _ in r1]
    od.Reset()
    assert od.MapSize() == 0 and len(od.Trajectory()) == 0
    p2 = []
    for s in seq_small[:4]:
        p2.append(list(od.RegisterFrame(s["xyz"], s["t"], s["frame_idx"]).frame.end_pose.tr))
    assert np.allclose(p1, p2, atol=0, rtol=0)      # deterministic: fixed-order reductions, counter-based shuffles


# ---- solver CERES (reproduced as device LM / IRLS) ---------------------------------------------------------------
def driving_config(b):
    """config/odometry/driving_config.yaml verbatim (BASELINE.json configs[2])."""
    o = b.profile("default_driving")
    o.debug_print = 0
    o.motion_compensation = abi.MOTION_COMPENSATION["CONTINUOUS"]
    o.initialization = abi.INITIALIZATION["INIT_CONSTANT_VELOCITY"]
    o.sample_voxel_size = 1.5
    o.voxel_size = 0.5
    o.max_distance = 100.0
    o.distance_error_threshold = 5.0
    o.neighborhood_strategy.max_num_neighbors = 20
    o.neighborhood_strategy.min_num_neighbors = 10
    m = b.default_map_options()
    m.num_resolutions = 1
    m.resolutions[0].resolution = 0.8
    m.resolutions[0].max_num_points = 30
    m.resolutions[0].min_distance_between_points = 0.1
    m.default_radius = 0.75
    o.map_options = m
    c = o.ct_icp_options
    c.debug_print = 0
    c.num_iters_icp = 5
    c.solver = abi.SOLVER["CERES"]
    c.max_num_residuals = 900
    c.min_num_residuals = 100
    c.weight_alpha = 0.9
    c.weight_neighborhood = 0.1
    c.min_number_neighbors = 20
    c.max_number_neighbors = 20
    c.num_closest_neighbors = 1
    c.power_planarity = 2
    c.threshold_orientation_norm = 0.1
    c.threshold_translation_norm = 0.01
    c.loss_function = abi.LOSS["CAUCHY"]
    c.ls_max_num_iters = 5
    c.ls_num_threads = 6
    c.ls_sigma = 0.1
    c.ls_tolerant_min_threshold = 0.05
    return o


@pytest.mark.parametrize("loss", ["CAUCHY", "HUBER", "STANDARD", "TOLERANT", "TRUNCATED"])
def test_ceres_register_matches_oracle(orc, eng, loss):
    map_xyz, kp, frame, prev = _registration_case(orc, eng, "CERES")
    mo, me = orc.voxel_map(small_map_options(orc, cap=1 << 18)), eng.voxel_map(small_map_options(eng, cap=1 << 18))
    mo.insert(map_xyz); me.insert(map_xyz)
    io = orc.default_icp_options()
    io.solver = abi.SOLVER["CERES"]
    io.loss_function = abi.LOSS[loss]
    io.min_number_neighbors = 10
    io.num_iters_icp = 4
    io.ls_max_num_iters = 5
    io.ls_num_threads = 4
    io.max_num_residuals = 700
    io.threshold_orientation_norm = 1e-9     # run every ICP iteration
    io.threshold_translation_norm = 1e-9
    mm = orc.default_odometry_options().default_motion_model
    mm.beta_small_velocity = 0.0005         # exercise every regulariser
    mm.beta_orientation_consistency = 0.0005
    st = abi.StrategyOptions(0, 20, 8, 0)
    _fill_world(orc, kp, frame)
    kpo, kpe = kp.copy(), kp.copy()
    fo, fe = frame.copy(), frame.copy()
    so = mo.icp_register(io, kpo, fo, prev, mm, st)
    se = me.icp_register(io, kpe, fe, prev, mm, st)
    assert so.success and se.success
    assert so.num_residuals_used == se.num_residuals_used == 700
    dt, dr = frame_diff(fo, fe)
    assert dt < POSE_TOL_M and dr < POSE_TOL_RAD, (loss, dt, dr)
    assert frame_diff(fo, frame, log=False)[0] > 1e-3
    print(loss, "pose diff %.3e m %.3e rad" % (dt, dr))


@pytest.mark.parametrize("ncn,max_res", [(2, 700), (3, 1000), (4, -1)])
def test_ceres_num_closest_neighbors(orc, eng, ncn, max_res):
    """num_closest_neighbors > 1 (src/ct_icp/ct_icp.cpp:554,593-601): every keypoint contributes ncn residual blocks,
    anchored on points[0..ncn) of its neighbor list (farthest first), block slots ncn * k + i; GetProblem keeps the first
    max_num_residuals of them in slot order."""
    map_xyz, kp, frame, prev = _registration_case(orc, eng, "CERES")
    mo, me = orc.voxel_map(small_map_options(orc, cap=1 << 18)), eng.voxel_map(small_map_options(eng, cap=1 << 18))
    mo.insert(map_xyz); me.insert(map_xyz)
    io = orc.default_icp_options()
    io.solver = abi.SOLVER["CERES"]
    io.loss_function = abi.LOSS["CAUCHY"]
    io.min_number_neighbors = 10
    io.num_closest_neighbors = ncn
    io.num_iters_icp = 3
    io.ls_max_num_iters = 5
    io.ls_num_threads = 4
    io.max_num_residuals = max_res
    io.threshold_orientation_norm = 1e-9
    io.threshold_translation_norm = 1e-9
    mm = orc.default_odometry_options().default_motion_model
    st = abi.StrategyOptions(0, 20, 8, 0)
    _fill_world(orc, kp, frame)
    kpo, kpe = kp.copy(), kp.copy()
    fo, fe = frame.copy(), frame.copy()
    so = mo.icp_register(io, kpo, fo, prev, mm, st)
    se = me.icp_register(io, kpe, fe, prev, mm, st)
    assert so.success and se.success
    assert so.num_residuals_used == se.num_residuals_used
    if max_res > 0:
        assert se.num_residuals_used == max_res
    else:
        assert se.num_residuals_used % ncn == 0 and se.num_residuals_used > len(kp)
    dt, dr = frame_diff(fo, fe)
    assert dt < POSE_TOL_M and dr < POSE_TOL_RAD, (ncn, dt, dr)
    # and it is a different problem from ncn = 1
    io.num_closest_neighbors = 1
    f1 = frame.copy()
    s1 = me.icp_register(io, kp.copy(), f1, prev, mm, st)
    assert s1.success and frame_diff(f1, fe, log=False)[0] > 1e-7


def test_ceres_num_closest_neighbors_out_of_range(orc, eng):
    map_xyz, kp, frame, prev = _registration_case(orc, eng, "CERES")
    me = eng.voxel_map(small_map_options(eng, cap=1 << 18))
    me.insert(map_xyz)
    io = eng.default_icp_options()
    io.solver = abi.SOLVER["CERES"]
    io.min_number_neighbors = 10
    io.num_closest_neighbors = 5
    mm = eng.default_odometry_options().default_motion_model
    st = abi.StrategyOptions(0, 20, 8, 0)
    with pytest.raises(Exception):
        me.icp_register(io, kp.copy(), frame.copy(), prev, mm, st)


def test_odometry_sequence_hdl64_ceres(orc, eng, seq_hdl64):
    """config 3: driving_config.yaml (solver CERES, Cauchy loss, 5 x 5 iterations, 900 residuals) as device LM/IRLS."""
    seq = seq_hdl64[:24]
    results = []
    for b in (orc, eng):
        od = b.odometry(driving_config(b))
        results.append([(od.RegisterFrame(s["xyz"], s["t"], s["frame_idx"]), od.MapSize()) for s in seq])
    worst_t = worst_r = 0.0
    for i, ((so, mo), (se, me)) in enumerate(zip(*results)):
        assert so.success and se.success, i
        assert so.num_keypoints == se.num_keypoints, i
        assert so.number_of_residuals == se.number_of_residuals, i
        dt, dr = frame_diff(so.frame, se.frame)
        worst_t, worst_r = max(worst_t, dt), max(worst_r, dr)
        assert dt < POSE_TOL_M and dr < POSE_TOL_RAD, (i, dt, dr)
        assert mo == me, i
    print("CERES worst per-frame pose difference: %.3e m, %.3e rad" % (worst_t, worst_r))


@pytest.mark.parametrize("p2p,mode", [("1", 2), ("0", 1)])
def test_multigpu_sharded_registration_matches_single_gpu(p2p, mode):
    """Keypoints sharded over 2 GPUs, JTJ/JTr summed over the ranks once per GN iteration / LM evaluation: inside the ICP
    kernels over NVLink peer mailboxes (sharding_mode 2) and through the ncclAllReduce fallback (CTICP_P2P=0, mode 1).
    Needs >= 2 devices; skipped otherwise."""
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, CTICP_CHECK_FRAMES="8", CTICP_P2P=p2p)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(29517 + mode),   # (the two parametrisations may run side by side under xdist)
                        os.path.join(root, "tools", "multigpu_check.py")], capture_output=True, text=True, env=env,
                       timeout=600)
    assert "MULTIGPU OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    assert "sharding_mode=%d" % mode in r.stdout, r.stdout[-2000:]


# ---- BASELINE.json configs[3] and configs[4] ------------------------------------------------------------------------
def nclt_config(b, solver="CERES"):
    """config/odometry/nclt_config.yaml (3 resolutions 0.5/1/2 → search level 0, r = ceil(0.8/0.5) = 2 → 125-voxel
    stencil; 20 ICP x 10 LS iterations; 1500 keypoints / residuals) with `sampling: GRID` (ADAPTIVE is SURVEY §8f-3)."""
    o = b.default_odometry_options()
    o.debug_print = 0
    o.sample_voxel_size = 0.8
    o.sampling = abi.SAMPLING["GRID"]
    o.voxel_size = 0.5
    o.max_distance = 100.0
    o.neighborhood_strategy.max_num_neighbors = 20
    o.neighborhood_strategy.min_num_neighbors = 10
    m = b.default_map_options()
    m.num_resolutions = 3
    for i, (res, md) in enumerate(((0.5, 0.05), (1.0, 0.1), (2.0, 0.2))):
        m.resolutions[i].resolution = res
        m.resolutions[i].max_num_points = 30
        m.resolutions[i].min_distance_between_points = md
    m.capacity_voxels = 1 << 19
    o.map_options = m
    o.max_num_keypoints = 1500
    c = o.ct_icp_options
    c.debug_print = 0
    c.num_iters_icp = 20
    c.solver = abi.SOLVER[solver]
    c.max_num_residuals = 1500
    c.min_number_neighbors = 10
    c.max_number_neighbors = 20
    c.threshold_orientation_norm = 0.1
    c.threshold_translation_norm = 0.01
    c.loss_function = abi.LOSS["CAUCHY"]
    c.ls_max_num_iters = 10
    c.ls_num_threads = 6
    c.ls_sigma = 0.1
    o.init_num_frames = 5
    return o


@pytest.mark.parametrize("solver", ["CERES", "GN"])
def test_odometry_sequence_nclt_shape(orc, eng, solver):
    """configs[3]: NCLT-shape HDL-32 (~65k-pt scans), multi-resolution map, 125-voxel stencil, max_num_keypoints."""
    from ct_icp_b200 import synthetic as syn
    seq = syn.make_sequence(9, syn.HDL32, seed=77, traj=syn.Trajectory(speed=2.0, sway=1.0, sway_rate=0.2, height=1.0,
                                                                       yaw_jerk=0.05))
    res = []
    for b in (orc, eng):
        od = b.odometry(nclt_config(b, solver))
        res.append([(od.RegisterFrame(s["xyz"], s["t"], s["frame_idx"]), od.MapSize()) for s in seq])
    worst = 0.0
    for i, ((so, mo), (se, me)) in enumerate(zip(*res)):
        assert so.success and se.success, (i, so.error_message, se.error_message)
        assert (so.num_keypoints, so.number_of_residuals, mo) == (se.num_keypoints, se.number_of_residuals, me), i
        if i >= 5:
            assert se.num_keypoints == 1500          # shuffle + resize (odometry.cpp:549-552)
        dt, dr = frame_diff(so.frame, se.frame)
        worst = max(worst, dt)
        assert dt < POSE_TOL_M and dr < POSE_TOL_RAD, (i, dt, dr)
    print("nclt-shape %s worst pose diff %.3e m" % (solver, worst))


def test_odometry_dense128_gn_20_iterations(orc, eng):
    """configs[4]: dense 128-beam scans (~290k pts), 20 GN iterations forced (thresholds 0), tens of thousands of
    keypoints per frame."""
    from ct_icp_b200 import synthetic as syn
    seq = syn.make_sequence(4, syn.DENSE128, seed=99)
    res = []
    for b in (orc, eng):
        o = b.default_odometry_options()
        o.debug_print = 0
        o.ct_icp_options.solver = abi.SOLVER["GN"]
        o.ct_icp_options.num_iters_icp = 20
        o.ct_icp_options.threshold_orientation_norm = 0.0
        o.ct_icp_options.min_number_neighbors = 10
        o.map_options = b.legacy_map_options(1.0, 20, 0.1)
        o.voxel_size = 0.25
        o.init_voxel_size = 0.25
        o.sample_voxel_size = 0.5
        o.init_sample_voxel_size = 0.5
        o.init_num_frames = 2
        od = b.odometry(o)
        res.append([od.RegisterFrame(s["xyz"], s["t"], s["frame_idx"]) for s in seq])
    for i, (so, se) in enumerate(zip(*res)):
        assert so.success and se.success, i
        assert so.num_keypoints == se.num_keypoints and so.number_of_residuals == se.number_of_residuals, i
        dt, dr = frame_diff(so.frame, se.frame)
        assert dt < POSE_TOL_M and dr < POSE_TOL_RAD, (i, dt, dr)
    assert res[1][-1].num_keypoints > 10000
    assert res[1][-1].icp_summary.num_iters == 20
    print("dense128: K = %d, pose diff %.3e m" % (res[1][-1].num_keypoints, frame_diff(res[0][-1].frame, res[1][-1].frame)[0]))


def test_adaptive_sampling_indices(orc, eng):
    s = get_sequence("hdl64", 2)[1]
    a = orc.adaptive_sample_indices(s["xyz"])
    b = eng.adaptive_sample_indices(s["xyz"])
    assert len(a) == len(b) > 100 and np.array_equal(a, b)
    o = abi.AdaptiveOptions()
    eng.fn("default_adaptive_options")(C.byref(o))
    o.max_num_points = 500
    assert np.array_equal(orc.adaptive_sample_indices(s["xyz"], o), eng.adaptive_sample_indices(s["xyz"], o))


@pytest.mark.parametrize("solver,init_frames", [("GN", 5), ("CERES", 2)])
def test_odometry_nclt_config_adaptive_sampling(orc, eng, solver, init_frames):
    """config/odometry/nclt_config.yaml including `sampling: ADAPTIVE`.

    The keypoints of the adaptive sampler come out band-major (nearest band first, sampling.h:92-108). While
    index_frame < init_num_frames the reference neither shuffles nor truncates them (odometry.cpp:549), so
    `max_num_residuals = 1500` keeps the 1500 NEAREST keypoints: in the synthetic street scene these all lie on the
    ground within a few metres of the sensor, yaw and x/y are unobservable, the 15 x 10 LM iterations wander along
    that null space (the begin-pose yaw changes sign from one ICP iteration to the next) and 1e-8 differences between
    two implementations flip a neighbor set and then diverge (measured with CTICP_DEBUG_LM / ORC_DEBUG_LM: identical
    accept/reject sequence and radii until one residual changes at ICP iteration 12). That case is ill-posed for the
    reference too, so the CERES variant leaves the init phase after 2 frames (then the 1500 keypoints are a random
    subset of all bands); the GN variant runs the yaml's init_num_frames."""
    from ct_icp_b200 import synthetic as syn
    seq = syn.make_sequence(8, syn.HDL32, seed=78, traj=syn.Trajectory(speed=2.0, sway=1.0, sway_rate=0.2, height=1.0))
    res = []
    for b in (orc, eng):
        o = nclt_config(b, solver)
        o.sampling = abi.SAMPLING["ADAPTIVE"]
        o.init_num_frames = init_frames
        od = b.odometry(o)
        res.append([od.RegisterFrame(s["xyz"], s["t"], s["frame_idx"]) for s in seq])
    worst = 0.0
    for i, (so, se) in enumerate(zip(*res)):
        assert so.success and se.success, i
        assert so.num_keypoints == se.num_keypoints and so.number_of_residuals == se.number_of_residuals, i
        dt, dr = frame_diff(so.frame, se.frame)
        worst = max(worst, dt)
        assert dt < POSE_TOL_M and dr < POSE_TOL_RAD, (i, dt, dr)
    print("nclt ADAPTIVE %s worst pose diff %.3e m" % (solver, worst))


# ---- solver ROBUST, src/ct_icp/ct_icp.cpp:1180-1370 ---------------------------------------------------
def robust_icp_options(b, use_barycenter, use_lines=1):
    """ct_icp_options of test/regression/regression_robust_config_short_drive.yaml:74-126."""
    c = b.default_icp_options()
    c.debug_print = 0
    c.num_iters_icp = 5
    c.solver = abi.SOLVER["ROBUST"]
    c.max_num_residuals = 1000
    c.min_num_residuals = 100
    c.weight_alpha = 0.9
    c.weight_neighborhood = 0.1
    c.min_number_neighbors = 8
    c.max_number_neighbors = 20
    c.power_planarity = 2
    c.threshold_orientation_norm = 0.001
    c.threshold_translation_norm = 0.001
    c.loss_function = abi.LOSS["CAUCHY"]
    c.ls_max_num_iters = 8
    c.ls_num_threads = 6
    c.ls_sigma = 0.1
    c.max_dist_to_plane_ct_icp = 0.5
    c.threshold_linearity = 0.9
    c.threshold_planarity = 0.8
    c.weight_point_to_point = 0.2
    c.outlier_distance = 0.8
    c.use_barycenter = use_barycenter
    c.use_lines = use_lines
    return c


@pytest.mark.parametrize("use_barycenter,use_lines,thr_lin", [(1, 1, 0.9), (0, 1, 0.6), (1, 0, 0.6)])
def test_robust_register_matches_oracle(orc, eng, use_barycenter, use_lines, thr_lin):
    """L3 Register with solver ROBUST: planar / linear / volumic classes, both anchors, all three functors."""
    map_xyz, kp, frame, prev = _registration_case(orc, eng, "ROBUST")
    mo, me = orc.voxel_map(small_map_options(orc, cap=1 << 18)), eng.voxel_map(small_map_options(eng, cap=1 << 18))
    mo.insert(map_xyz); me.insert(map_xyz)
    io = robust_icp_options(orc, use_barycenter, use_lines)
    io.threshold_linearity = thr_lin
    io.max_num_residuals = -1
    io.num_iters_icp = 4
    io.ls_num_threads = 4
    io.threshold_orientation_norm = 1e-9     # run every ICP iteration (the class memory across iterations matters)
    io.threshold_translation_norm = 1e-9
    mm = orc.default_odometry_options().default_motion_model
    st = abi.StrategyOptions(0, 20, 8, 0)
    _fill_world(orc, kp, frame)
    kpo, kpe = kp.copy(), kp.copy()
    fo, fe = frame.copy(), frame.copy()
    so = mo.icp_register(io, kpo, fo, prev, mm, st)
    se = me.icp_register(io, kpe, fe, prev, mm, st)
    assert so.success and se.success
    assert so.num_residuals_used == se.num_residuals_used > 500
    dt, dr = frame_diff(fo, fe)
    assert dt < POSE_TOL_M and dr < POSE_TOL_RAD, (dt, dr)
    assert frame_diff(fo, frame, log=False)[0] > 1e-3
    print("ROBUST residuals %d pose diff %.3e m %.3e rad" % (se.num_residuals_used, dt, dr))


def test_odometry_sequence_hdl64_robust(orc, eng, seq_hdl64):
    """test/regression/regression_robust_config_short_drive.yaml on the KITTI-shape scans."""
    seq = seq_hdl64[:16]
    results = []
    for b in (orc, eng):
        o = b.default_odometry_options()
        o.debug_print = 0
        o.motion_compensation = abi.MOTION_COMPENSATION["CONTINUOUS"]
        o.initialization = abi.INITIALIZATION["INIT_CONSTANT_VELOCITY"]
        o.sample_voxel_size = 1.5
        o.voxel_size = 0.5
        o.max_distance = 100.0
        o.distance_error_threshold = 5.0
        o.map_options = b.legacy_map_options(1.0, 20, 0.1)
        o.ct_icp_options = robust_icp_options(b, use_barycenter=1)
        o.default_motion_model.beta_location_consistency = 0.001
        o.default_motion_model.beta_constant_velocity = 0.001
        od = b.odometry(o)
        results.append([(od.RegisterFrame(s["xyz"], s["t"], s["frame_idx"]), od.MapSize()) for s in seq])
    worst_t = worst_r = 0.0
    for i, ((so, mo), (se, me)) in enumerate(zip(*results)):
        assert so.success and se.success, i
        assert (so.num_keypoints, so.number_of_residuals, mo) == (se.num_keypoints, se.number_of_residuals, me), i
        dt, dr = frame_diff(so.frame, se.frame)
        worst_t, worst_r = max(worst_t, dt), max(worst_r, dr)
        assert dt < POSE_TOL_M and dr < POSE_TOL_RAD, (i, dt, dr)
    print("ROBUST worst per-frame pose difference: %.3e m, %.3e rad" % (worst_t, worst_r))


# ---- PointCloud2-like record buffers in and out ------------------------------------------------------
@pytest.mark.parametrize("xyz_type,t_type", [("float32", "float32"), ("float32", "float64"), ("float64", "uint32")])
def test_register_cloud_records_match_oracle(orc, eng, seq_small, xyz_type, t_type):
    """cticp_odometry_register_cloud reads packed records in place (any PointField scalar type for the timestamp,
    float32/float64 coordinates, unaligned record size) and must behave like RegisterFrame on the converted arrays —
    the reference's XYZConst<double>() / TimestampsProxy<double>() views (odometry.cpp:335-336)."""
    rec_dtype = np.dtype({"names": ["intensity", "x", "y", "z", "ring", "t"],
                          "formats": ["u1", xyz_type, xyz_type, xyz_type, "u2", t_type],
                          "offsets": [0, 1, 1 + np.dtype(xyz_type).itemsize, 1 + 2 * np.dtype(xyz_type).itemsize,
                                      1 + 3 * np.dtype(xyz_type).itemsize, 3 + 3 * np.dtype(xyz_type).itemsize],
                          "itemsize": 3 + 3 * np.dtype(xyz_type).itemsize + np.dtype(t_type).itemsize + 2})
    odo = orc.odometry(_sequence_options(orc, init_num_frames=4))
    ode = eng.odometry(_sequence_options(eng, init_num_frames=4))
    for s in seq_small[:6]:
        rec = np.zeros(len(s["xyz"]), dtype=rec_dtype)
        rec["x"], rec["y"], rec["z"] = s["xyz"][:, 0], s["xyz"][:, 1], s["xyz"][:, 2]
        if t_type == "uint32":      # e.g. microseconds since the start of the sequence
            rec["t"] = np.round(s["t"] * 1e6).astype(np.uint32)
        else:
            rec["t"] = s["t"]
        xyz64 = np.stack([rec["x"], rec["y"], rec["z"]], axis=1).astype(np.float64)
        t64 = rec["t"].astype(np.float64)
        so = odo.RegisterFrame(xyz64, t64, s["frame_idx"])
        se = ode.RegisterCloud(rec, s["frame_idx"])
        assert so.success and se.success
        assert (so.num_keypoints, so.number_of_residuals) == (se.num_keypoints, se.number_of_residuals)
        dt, dr = frame_diff(so.frame, se.frame)
        assert dt < POSE_TOL_M and dr < POSE_TOL_RAD, (dt, dr)
    # egress into pcl::PointCloud<slam::XYZTPoint>-shaped records (float x, y, z, pad; double timestamp)
    ref = ode.corrected_points()
    out_dtype = np.dtype({"names": ["x", "y", "z", "t"], "formats": ["f4", "f4", "f4", "f8"], "offsets": [0, 4, 8, 16],
                          "itemsize": 32})
    out = np.zeros(len(ref), dtype=out_dtype)
    n = ode.write_points(abi.POINTS_CORRECTED, out, world=True)
    assert n == len(ref)
    assert np.array_equal(np.stack([out["x"], out["y"], out["z"]], 1), ref["world"].astype(np.float32))
    assert np.array_equal(out["t"], ref["timestamp"])


# ---- DistanceBasedStrategy, per-voxel normals, sensor-side filter ---------------------------------------
def _three_level_options(b, cap=1 << 16):
    m = b.default_map_options()
    m.num_resolutions = 3
    for i, (res, md) in enumerate(((0.5, 0.05), (1.0, 0.1), (2.0, 0.2))):
        m.resolutions[i].resolution = res
        m.resolutions[i].max_num_points = 30
        m.resolutions[i].min_distance_between_points = md
    m.capacity_voxels = cap
    m.select_valid_normals_direction = 1
    return m


def test_radius_search_per_query_radius_and_normal_filter(orc, eng):
    """ComputeNeighborhoods(queries, radiuses, k, true, sensor_location) (map.h:434-447): the radius picks the level and
    the stencil per query; with a sensor location, points of voxels whose oriented normal faces away are skipped.
    Three frames inserted from three origins so that the per-point orientation of a shared voxel normal differs."""
    seq = get_sequence("hdl64", 3)
    mo, me = orc.voxel_map(_three_level_options(orc)), eng.voxel_map(_three_level_options(eng))
    rng = np.random.default_rng(5)
    for i, s in enumerate(seq):
        world = s["xyz"][::7] + np.array([0.4 * i, 0.0, 0.0])
        origin = np.array([0.4 * i, 0.1 * i, 0.0])
        mo.insert(world, origin); me.insert(world, origin)
    assert [mo.num_points(l) for l in range(3)] == [me.num_points(l) for l in range(3)]
    q = seq[2]["xyz"][3::41][:1500]
    q = q + rng.normal(scale=0.05, size=q.shape)
    radiuses = rng.uniform(0.15, 2.4, len(q))
    total_unfiltered = total_filtered = 0
    for sensor in (None, (0.8, 0.2, 0.0), (30.0, -12.0, 4.0)):
        po, co = mo.radius_search(q, radiuses, 20, sensor)
        pe, ce = me.radius_search(q, radiuses, 20, sensor)
        assert np.array_equal(co, ce), (sensor, int((co != ce).sum()))
        assert np.abs(po - pe).max() < 1e-6          # fp32 storage quantum
        if sensor is None:
            total_unfiltered = int(co.sum())
        else:
            total_filtered = int(co.sum())
            assert total_filtered < total_unfiltered   # the filter removes something
    assert total_unfiltered > 10000


def _distance_based(b):
    st = abi.StrategyOptions(abi.STRATEGY["DISTANCE_BASED_STRATEGY"], 20, 8)
    st.radius_min, st.radius_max, st.exponent, st.distance_max = 0.1, 1.2, 1.0, 60.0
    return st


def test_odometry_sequence_distance_based_strategy(orc, eng, seq_hdl64):
    """driving_config.yaml with `neighborhood_strategy: DISTANCE_BASED_STRATEGY` (config.cpp:155-167) on a
    two-resolution map: per-keypoint radius → level, normals maintained by the map update, sensor-side filter."""
    seq = seq_hdl64[:14]
    results = []
    for b in (orc, eng):
        o = driving_config(b)
        m = b.default_map_options()
        m.num_resolutions = 2
        for i, (res, md) in enumerate(((0.6, 0.1), (1.2, 0.2))):
            m.resolutions[i].resolution = res
            m.resolutions[i].max_num_points = 30
            m.resolutions[i].min_distance_between_points = md
        m.select_valid_normals_direction = 1
        o.map_options = m
        o.neighborhood_strategy = _distance_based(b)
        o.ct_icp_options.min_number_neighbors = 10
        od = b.odometry(o)
        results.append([(od.RegisterFrame(s["xyz"], s["t"], s["frame_idx"]), od.MapSize()) for s in seq])
    worst = 0.0
    for i, ((so, mo), (se, me)) in enumerate(zip(*results)):
        assert so.success and se.success, i
        assert (so.num_keypoints, so.number_of_residuals, mo) == (se.num_keypoints, se.number_of_residuals, me), i
        dt, dr = frame_diff(so.frame, se.frame)
        worst = max(worst, dt)
        assert dt < POSE_TOL_M and dr < POSE_TOL_RAD, (i, dt, dr)
    print("DISTANCE_BASED_STRATEGY worst per-frame pose difference: %.3e m" % worst)
