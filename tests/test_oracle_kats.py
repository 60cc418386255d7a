"""CPU tests pinning the ORACLE against the reference's own property tests.

The reference holds no golden vectors for this path ("parity unpinned"); these are its unit-test properties restated
with seeded inputs, plus self-consistency checks of the restated Eigen/Ceres arithmetic (finite differences, SE3
identities). Everything here runs on the CPU.
"""
import ctypes as C

import numpy as np
import pytest

from ct_icp_b200 import _abi as abi


def _q(rng):
    q = rng.normal(size=4)
    return q / np.linalg.norm(q)


def _arr(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data


def se3_inverse(orc, q, t):
    oq, ot = np.zeros(4), np.zeros(3)
    orc.fn("se3_inverse")(_arr(q)[1], _arr(t)[1], oq.ctypes.data, ot.ctypes.data)
    return oq, ot


def se3_mul(orc, qa, ta, qb, tb):
    oq, ot = np.zeros(4), np.zeros(3)
    a, b, c, d = _arr(qa), _arr(ta), _arr(qb), _arr(tb)
    orc.fn("se3_mul")(a[1], b[1], c[1], d[1], oq.ctypes.data, ot.ctypes.data)
    return oq, ot


def se3_apply(orc, q, t, p):
    out = np.zeros(3)
    a, b, c = _arr(q), _arr(t), _arr(p)
    orc.fn("se3_apply")(a[1], b[1], c[1], out.ctypes.data)
    return out


def se3_interpolate(orc, qa, ta, qb, tb, w):
    oq, ot = np.zeros(4), np.zeros(3)
    a, b, c, d = _arr(qa), _arr(ta), _arr(qb), _arr(tb)
    orc.fn("se3_interpolate")(a[1], b[1], c[1], d[1], float(w), oq.ctypes.data, ot.ctypes.data)
    return oq, ot


def ct_residual(orc, alpha, ref, raw, normal, weight, qb, tb, qe, te, want_jac=False):
    jac = np.zeros(12)
    arrs = [_arr(x) for x in (ref, raw, normal, qb, tb, qe, te)]
    r = orc.fn("ct_point_to_plane_residual")(float(alpha), arrs[0][1], arrs[1][1], arrs[2][1], float(weight),
                                             arrs[3][1], arrs[4][1], arrs[5][1], arrs[6][1],
                                             jac.ctypes.data if want_jac else None)
    return (r, jac) if want_jac else r


def ct_residual_kind(orc, kind, alpha, ref, raw, direction, cov, weight, qb, tb, qe, te, want_jac=False):
    jac = np.zeros(12)
    arrs = [_arr(x) for x in (ref, raw, direction, qb, tb, qe, te)]
    c = _arr(np.asarray(cov, dtype=np.float64).reshape(9)) if cov is not None else (None, None)
    r = orc.fn("ct_residual")(abi.DISTANCE[kind], float(alpha), arrs[0][1], arrs[1][1], arrs[2][1], c[1], float(weight),
                              arrs[3][1], arrs[4][1], arrs[5][1], arrs[6][1], jac.ctypes.data if want_jac else None)
    return (r, jac) if want_jac else r


# ---- test/unit/ct_icp/test_cost_functions.cxx:70-105 -------------------------------------------------------------
@pytest.mark.parametrize("seed", range(8))
def test_ct_point_to_plane_residual_zero_on_plane(orc, seed):
    rng = np.random.default_rng(seed)
    normal = np.array([0.0, 0.0, 1.0])
    reference = rng.uniform(-1, 1, 3)
    world_point = rng.uniform(-1, 1, 3)
    world_in_plane = rng.uniform(-1, 1, 3)
    world_in_plane[2] = reference[2]
    qa, ta = _q(rng), rng.uniform(-1, 1, 3)
    qb, tb = _q(rng), rng.uniform(-1, 1, 3)
    if np.dot(qa, qb) < 0:
        pass    # slerp handles the antipodal case through the sign flip
    alpha = 0.3
    qi, ti = se3_interpolate(orc, qa, ta, qb, tb, alpha)
    qinv, tinv = se3_inverse(orc, qi, ti)
    raw = se3_apply(orc, qinv, tinv, world_point)
    raw_in_plane = se3_apply(orc, qinv, tinv, world_in_plane)
    r_perfect = ct_residual(orc, alpha, reference, raw_in_plane, normal, 1.0, qa, ta, qb, tb)
    r_error = ct_residual(orc, alpha, reference, raw, normal, 1.0, qa, ta, qb, tb)
    assert abs(r_perfect) <= 1e-12
    if abs(world_point[2] - reference[2]) > 1e-2:
        assert abs(r_error) >= 1e-3


# ---- test/unit/ct_icp/test_cost_functions.cxx:9-30 (FunctorPointToPlane on ONE pose) -----------------------------
@pytest.mark.parametrize("seed", range(6))
@pytest.mark.parametrize("alpha", [0.0, 0.37, 1.0])
def test_point_to_plane_functor_single_pose(orc, seed, alpha):
    """The reference's single-pose functor is what the CT functor evaluates when begin == end (any alpha): zero for a
    point in the plane through `reference`, clearly non-zero off the plane."""
    rng = np.random.default_rng(40 + seed)
    normal = np.array([0.0, 0.0, 1.0])
    reference = rng.uniform(-1, 1, 3)
    world_point = rng.uniform(-1, 1, 3)
    world_in_plane = rng.uniform(-1, 1, 3)
    world_in_plane[2] = reference[2]
    q, t = _q(rng), rng.uniform(-1, 1, 3)
    qinv, tinv = se3_inverse(orc, q, t)
    raw = se3_apply(orc, qinv, tinv, world_point)
    raw_in_plane = se3_apply(orc, qinv, tinv, world_in_plane)
    assert abs(ct_residual(orc, alpha, reference, raw_in_plane, normal, 1.0, q, t, q, t)) <= 1e-12
    if abs(world_point[2] - reference[2]) > 1e-2:
        assert abs(ct_residual(orc, alpha, reference, raw, normal, 1.0, q, t, q, t)) >= 1e-3


# ---- the autodiff restatement: tangent-space Jacobian vs central differences through Plus -----------------------
def _quat_plus(q, d):
    n = np.linalg.norm(d)
    if n == 0:
        return q.copy()
    s = np.sin(n) / n
    dq = np.array([s * d[0], s * d[1], s * d[2], np.cos(n)])
    x1, y1, z1, w1 = dq
    x2, y2, z2, w2 = q
    return np.array([w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2, w1 * y2 + y1 * w2 + z1 * x2 - x1 * z2,
                     w1 * z2 + z1 * w2 + x1 * y2 - y1 * x2, w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2])


@pytest.mark.parametrize("seed", range(5))
def test_ct_functor_jacobian_matches_finite_differences(orc, seed):
    rng = np.random.default_rng(100 + seed)
    qb, qe = _q(rng), None
    qe = _quat_plus(qb, rng.normal(scale=0.05, size=3))      # nearby orientation, like consecutive scan poses
    tb, te = rng.uniform(-5, 5, 3), rng.uniform(-5, 5, 3)
    raw = rng.uniform(-30, 30, 3)
    normal = rng.normal(size=3)
    normal /= np.linalg.norm(normal)
    ref = rng.uniform(-30, 30, 3)
    alpha, w = rng.uniform(0.05, 0.95), rng.uniform(0.2, 1.0)
    r0, J = ct_residual(orc, alpha, ref, raw, normal, w, qb, tb, qe, te, want_jac=True)
    h = 1e-6
    Jfd = np.zeros(12)
    for k in range(12):
        def f(step):
            qb2, qe2, tb2, te2 = qb.copy(), qe.copy(), tb.copy(), te.copy()
            d = np.zeros(3)
            d[k % 3] = step
            if k < 3:
                qb2 = _quat_plus(qb, d)
            elif k < 6:
                qe2 = _quat_plus(qe, d)
            elif k < 9:
                tb2 = tb + d
            else:
                te2 = te + d
            return ct_residual(orc, alpha, ref, raw, normal, w, qb2, tb2, qe2, te2)
        Jfd[k] = (f(h) - f(-h)) / (2 * h)
    assert np.abs(J - Jfd).max() < 1e-6 * max(1.0, np.abs(J).max())


# ---- test/unit/ct_icp/test_cost_functions.cxx:42-62 (PointToLine), wrapped in the CT functor like solver ROBUST ----
@pytest.mark.parametrize("seed", range(6))
def test_ct_point_to_line_residual_zero_on_line(orc, seed):
    rng = np.random.default_rng(300 + seed)
    line = np.array([0.0, 0.0, 1.0])
    reference = rng.uniform(-1, 1, 3)
    world_point = rng.uniform(-1, 1, 3)
    world_on_line = reference + rng.uniform(0.1, 1.0) * line
    qa, ta = _q(rng), rng.uniform(-1, 1, 3)
    qb, tb = _q(rng), rng.uniform(-1, 1, 3)
    alpha = 0.3
    qi, ti = se3_interpolate(orc, qa, ta, qb, tb, alpha)
    qinv, tinv = se3_inverse(orc, qi, ti)
    raw = se3_apply(orc, qinv, tinv, world_point)
    raw_on_line = se3_apply(orc, qinv, tinv, world_on_line)
    assert abs(ct_residual_kind(orc, "POINT_TO_LINE", alpha, reference, raw_on_line, 3.0 * line, None, 1.0, qa, ta, qb, tb)) <= 1e-12
    d = world_point - reference
    expect = np.linalg.norm(np.cross(line, d))
    got = ct_residual_kind(orc, "POINT_TO_LINE", alpha, reference, raw, 3.0 * line, None, 0.5, qa, ta, qb, tb)
    assert abs(got - 0.5 * expect) < 1e-12


def test_ct_point_to_distribution_residual_is_the_mahalanobis_form(orc):
    rng = np.random.default_rng(17)
    a = rng.normal(size=(3, 3))
    cov = a @ a.T * 0.05
    reference = rng.uniform(-1, 1, 3)
    world_point = reference + rng.uniform(-0.5, 0.5, 3)
    qa, ta = _q(rng), rng.uniform(-1, 1, 3)
    qb, tb = _q(rng), rng.uniform(-1, 1, 3)
    alpha = 0.6
    qi, ti = se3_interpolate(orc, qa, ta, qb, tb, alpha)
    qinv, tinv = se3_inverse(orc, qi, ti)
    raw = se3_apply(orc, qinv, tinv, world_point)
    d = world_point - reference
    expect = 0.1 * d @ np.linalg.inv(cov + 0.05 * np.eye(3)) @ d          # cost_functions.h:154-171
    got = ct_residual_kind(orc, "POINT_TO_DISTRIBUTION", alpha, reference, raw, np.zeros(3), cov, 0.1, qa, ta, qb, tb)
    assert abs(got - expect) < 1e-12
    raw0 = se3_apply(orc, qinv, tinv, reference)
    assert abs(ct_residual_kind(orc, "POINT_TO_DISTRIBUTION", alpha, reference, raw0, np.zeros(3), cov, 0.1, qa, ta, qb, tb)) < 1e-20


@pytest.mark.parametrize("kind", ["POINT_TO_LINE", "POINT_TO_DISTRIBUTION"])
@pytest.mark.parametrize("seed", range(3))
def test_robust_functor_jacobians_match_finite_differences(orc, kind, seed):
    rng = np.random.default_rng(400 + seed)
    qb = _q(rng)
    qe = _quat_plus(qb, rng.normal(scale=0.05, size=3))
    tb, te = rng.uniform(-5, 5, 3), rng.uniform(-5, 5, 3)
    raw = rng.uniform(-30, 30, 3)
    direction = rng.normal(size=3)
    a = rng.normal(size=(3, 3))
    cov = a @ a.T * 0.02
    alpha, w = rng.uniform(0.05, 0.95), rng.uniform(0.2, 1.0)
    qi, ti = se3_interpolate(orc, qb, tb, qe, te, alpha)
    ref = se3_apply(orc, qi, ti, raw) + rng.uniform(-0.4, 0.4, 3)           # an anchor a few decimetres away
    args = (kind, alpha, ref, raw, direction, cov, w)
    r0, J = ct_residual_kind(orc, *args, qb, tb, qe, te, want_jac=True)
    h = 1e-6
    Jfd = np.zeros(12)
    for k in range(12):
        def f(step):
            qb2, qe2, tb2, te2 = qb.copy(), qe.copy(), tb.copy(), te.copy()
            d = np.zeros(3)
            d[k % 3] = step
            if k < 3:
                qb2 = _quat_plus(qb, d)
            elif k < 6:
                qe2 = _quat_plus(qe, d)
            elif k < 9:
                tb2 = tb + d
            else:
                te2 = te + d
            return ct_residual_kind(orc, *args, qb2, tb2, qe2, te2)
        Jfd[k] = (f(h) - f(-h)) / (2 * h)
    assert np.abs(J - Jfd).max() < 2e-6 * max(1.0, np.abs(J).max())


# ---- test/unit/SlamCore/test_neighborhood.cxx:40-53 --------------------------------------------------------------
@pytest.mark.parametrize("seed", range(5))
def test_neighborhood_planar_normal(orc, seed):
    rng = np.random.default_rng(seed)
    pts = rng.uniform(-1, 1, size=(10, 3))
    pts[:, 2] = 1.0
    normal = np.zeros(3)
    a2d, plan, lin = C.c_double(), C.c_double(), C.c_double()
    ok = orc.fn("neighborhood_describe")(pts.ctypes.data, len(pts), normal.ctypes.data, C.byref(a2d), C.byref(plan),
                                         C.byref(lin), None)
    assert ok == 1
    assert abs(abs(normal[2]) - 1.0) < 1e-12      # the reference asserts == 1 exactly with JacobiSVD
    assert abs(normal[0]) < 1e-7 and abs(normal[1]) < 1e-7
    assert 0.0 <= a2d.value <= 1.0 + 1e-12


def test_neighborhood_needs_five_points(orc):
    pts = np.random.default_rng(0).uniform(-1, 1, size=(4, 3))
    normal = np.zeros(3)
    d = C.c_double()
    assert orc.fn("neighborhood_describe")(pts.ctypes.data, 4, normal.ctypes.data, C.byref(d), C.byref(d), C.byref(d), None) == 0


# ---- test/unit/SlamCore/test_types.cxx:7-86 -----------------------------------------------------------------------
@pytest.mark.parametrize("seed", range(5))
def test_se3_inverse_and_compose(orc, seed):
    rng = np.random.default_rng(seed)
    q, t = _q(rng), rng.uniform(-1, 1, 3)
    qi, ti = se3_inverse(orc, q, t)
    qid, tid = se3_mul(orc, qi, ti, q, t)
    assert np.abs(np.abs(qid) - np.array([0, 0, 0, 1])).max() < 1e-10
    assert np.abs(tid).max() < 1e-10
    p = rng.uniform(-1, 1, 3)
    back = se3_apply(orc, qi, ti, se3_apply(orc, q, t, p))
    assert np.abs(back - p).max() < 1e-10
    # angular distance of a pose with itself is 0; with a 90 degree turn about z it is 90
    qz = np.array([0, 0, np.sin(np.pi / 4), np.cos(np.pi / 4)])
    a, b = _arr(q), _arr(q)
    assert orc.fn("angular_distance")(a[1], b[1]) < 1e-6
    i4, z4 = _arr(np.array([0.0, 0, 0, 1])), _arr(qz)
    assert abs(orc.fn("angular_distance")(i4[1], z4[1]) - 90.0) < 1e-9


def test_pose_interpolation_endpoints_and_timestamp_check(orc):
    rng = np.random.default_rng(3)
    f = abi.Frame()
    f.begin_pose = abi.Pose.make(_q(rng), rng.uniform(-1, 1, 3), 10.0, 1)
    f.end_pose = abi.Pose.make(_q(rng), rng.uniform(-1, 1, 3), 10.1, 1)
    raw = (C.c_double * 3)(1.0, -2.0, 0.5)
    out = (C.c_double * 3)()
    orc.check(orc.fn("pose_transform")(C.byref(f), raw, 10.0, out))
    assert np.allclose(out[:], se3_apply(orc, np.array(f.begin_pose.quat), np.array(f.begin_pose.tr), raw[:]), atol=1e-12)
    orc.check(orc.fn("pose_transform")(C.byref(f), raw, 10.1, out))
    assert np.allclose(out[:], se3_apply(orc, np.array(f.end_pose.quat), np.array(f.end_pose.tr), raw[:]), atol=1e-12)
    # reference: CHECK(dest_timestamp <= t <= other.dest_timestamp) aborts (types.h:456) → error code
    assert orc.fn("pose_transform")(C.byref(f), raw, 10.2, out) == abi.ERR_TIMESTAMP


# ---- test/unit/SlamCore/test_map.cxx:5-38 --------------------------------------------------------------------------
def test_map_self_nearest_and_full_recall(orc):
    rng = np.random.default_rng(5)
    pts = rng.uniform(-1, 1, size=(100, 3))
    m = orc.voxel_map(orc.legacy_map_options(0.01, 20, 0.0))
    m.insert(pts)
    nb, cnt = m.compute_neighborhoods(pts, 1)
    assert np.all(cnt == 1) and np.abs(nb[:, 0, :] - pts).max() < 1e-5
    o2 = orc.legacy_map_options(1.0, 64, 0.0)
    o2.default_radius = 4.0
    m2 = orc.voxel_map(o2)
    m2.insert(pts[:30])
    nb, cnt = m2.compute_neighborhoods(pts[:5], 32)
    assert np.all(cnt == 30)


def test_map_insert_rules_and_eviction(orc):
    m = orc.voxel_map(orc.legacy_map_options(1.0, 3, 0.1))
    # same voxel: second point closer than min_distance is rejected, capacity 3 is enforced (map.h:276-291)
    m.insert(np.array([[0.5, 0.5, 0.5], [0.55, 0.5, 0.5], [0.9, 0.5, 0.5], [0.1, 0.5, 0.5], [0.5, 0.9, 0.5]]))
    assert m.num_points() == 3 and m.num_voxels() == 1
    # truncation toward zero merges the voxels around 0 (types.cxx:15-17)
    m.insert(np.array([[-0.5, 0.2, 0.2]]))
    assert m.num_voxels() == 1
    # neighbor list is farthest first (map.h:508-513)
    nb, cnt = m.compute_neighborhoods(np.array([[0.5, 0.5, 0.5]]), 3)
    d = np.linalg.norm(nb[0, :cnt[0]] - np.array([0.5, 0.5, 0.5]), axis=1)
    assert np.all(np.diff(d) <= 1e-12)
    # eviction tests the FIRST stored point of the voxel (map.h:313)
    m.insert(np.array([[50.2, 0.0, 0.0]]))
    m.remove_far((0.0, 0.0, 0.0), 10.0)
    assert m.num_voxels() == 1


# ---- test/unit/SlamCore/test_A_grid_sampling.cxx:7-23 + order contract ---------------------------------------------
def test_grid_sampling_properties(orc):
    rng = np.random.default_rng(0)
    pts = rng.uniform(-20, 20, size=(5000, 3))
    idx = orc.grid_sample_indices(pts, 1.5)
    assert 0 < len(idx) <= len(pts)
    keys = np.trunc(pts[idx] / 1.5).astype(np.int64)
    assert len({tuple(k) for k in keys}) == len(idx)          # one point per voxel
    assert np.all(np.diff(idx) > 0)                             # order of first appearance
    allkeys = {tuple(k) for k in np.trunc(pts / 1.5).astype(np.int64)}
    assert len(allkeys) == len(idx)                             # every occupied voxel represented


def test_permutation_is_a_bijection_and_seeded(orc):
    for n in (1, 2, 7, 1000, 65536, 100003):
        p = orc.permutation(42, 7, n)
        assert np.array_equal(np.sort(p), np.arange(n, dtype=np.uint32))
    assert not np.array_equal(orc.permutation(42, 7, 1000), orc.permutation(42, 8, 1000))
    assert np.array_equal(orc.permutation(42, 7, 1000), orc.permutation(42, 7, 1000))


# ---- test/integration/testint_odometry.cpp:57-114: end-to-end success on a synthetic scene ---------------------------
@pytest.mark.parametrize("solver", ["GN", "CERES", "ROBUST"])
def test_oracle_odometry_tracks_ground_truth(orc, seq_small, solver):
    from ct_icp_b200 import synthetic as syn
    o = orc.default_odometry_options()
    o.ct_icp_options.solver = abi.SOLVER[solver]
    o.ct_icp_options.min_number_neighbors = 10
    o.ct_icp_options.ls_max_num_iters = 5
    o.ct_icp_options.ls_num_threads = 4
    o.map_options = orc.legacy_map_options(1.0, 20, 0.1)
    o.init_num_frames = 4
    o.debug_print = 0
    od = orc.odometry(o)
    T0 = seq_small[0]["gt_begin"]
    for s in seq_small:
        sm = od.RegisterFrame(s["xyz"], s["t"], s["frame_idx"])
        assert sm.success, sm.error_message
    gt = syn.relative_pose(T0, seq_small[-1]["gt_end"])
    est = np.array(sm.frame.end_pose.tr)
    assert np.linalg.norm(est - gt[:3, 3]) < 0.15
    assert len(od.Trajectory()) == len(seq_small)


def test_gn_fails_with_too_few_keypoints(orc):
    # ct_icp.cpp:860-871: fewer than 100 residuals → success = false
    rng = np.random.default_rng(1)
    m = orc.voxel_map(orc.legacy_map_options(1.0, 20, 0.1))
    m.insert(rng.uniform(-5, 5, size=(500, 3)))
    kp = np.zeros(50, dtype=abi.wpoint_dtype())
    kp["raw"] = rng.uniform(-5, 5, size=(50, 3))
    kp["world"] = kp["raw"]
    kp["timestamp"] = 0.05
    frame = abi.Frame()
    frame.begin_pose = abi.Pose.make(dest_timestamp=0.0, dest_frame_id=1)
    frame.end_pose = abi.Pose.make(dest_timestamp=0.1, dest_frame_id=1)
    io = orc.default_icp_options()
    io.solver = abi.SOLVER["GN"]
    assert not m.icp_register(io, kp, frame).success


# ---- include/ct_icp/algorithm/sampling.h:55-110 (adaptive, distance-banded grid sampling) ---------------------------
def test_adaptive_sampling_properties(orc):
    rng = np.random.default_rng(4)
    pts = rng.normal(scale=12.0, size=(20000, 3))
    idx = orc.adaptive_sample_indices(pts)
    assert 0 < len(idx) < len(pts)
    d = np.linalg.norm(pts[idx], axis=1)
    assert d.min() >= 0.5 and d.max() < 200.0
    bounds = np.array([0.5, 2.0, 4.0, 8.0, 16.0, 200.0])
    vsize = np.array([0.1, 0.2, 0.4, 0.8, 1.6])
    band = np.searchsorted(bounds, d, side="left") - 1
    assert np.all(np.diff(band) >= 0)                       # band by band
    for b in range(5):
        sel = idx[band == b]
        assert np.all(np.diff(sel) > 0)                     # first appearance inside a band
        keys = np.trunc(pts[sel] / vsize[b]).astype(np.int64)
        assert len({tuple(k) for k in keys}) == len(sel)    # one point per voxel of the band
    o = abi.AdaptiveOptions()
    orc.fn("default_adaptive_options")(C.byref(o))
    o.max_num_points = 100
    assert len(orc.adaptive_sample_indices(pts, o)) == 101  # the reference's `size() > max` lets max + 1 through
