"""Pins for the out-of-tree arithmetic the oracle restates "from memory"  — independent of the
oracle's own C++: every routine is checked against numpy / scipy closed forms, and the trust-region schedule of
ceres::Solve against a separate Python transcription that shares no code with oracle/orc_ceres.cpp:

* Eigen slerp / Quaternion(Matrix3) / toRotationMatrix            vs scipy.spatial.transform
* JacobiSVD<Matrix3d> of a symmetric matrix                        vs numpy.linalg.eigh
* Matrix<12,12>::ldlt().solve                                      vs numpy.linalg.solve
* ceres Cauchy / Huber / Tolerant loss (+ ct_icp's TruncatedLoss)  vs their published formulas; rho', rho'' vs numerical
                                                                     derivatives of rho
* ceres::internal::Corrector                                       vs Triggs' identities (gradient exact; Gauss-Newton
                                                                     Hessian (rho' + 2 s rho'') J^T J where it applies)
* EigenQuaternionParameterization::Plus                            vs a rotation by 2|delta| composed on the left
* CTFunctor<FunctorPointToPlane> + Ceres autodiff                  vs a numpy restatement of include/ct_icp/cost_functions.h
                                                                     :186-222, 32-67 differentiated by complex steps
* ceres::Solve (TrustRegionMinimizer + LevenbergMarquardtStrategy) vs lm_reference() below on a 12-dof CT problem:
                                                                     same accepted / rejected step counts, same optimum

CPU only (the oracle is test infrastructure; the CUDA engine is compared with the oracle in the -m gpu tests).
"""
import ctypes as C

import mpmath
import numpy as np
import pytest
from scipy.spatial.transform import Rotation, Slerp

from ct_icp_b200 import _abi as abi

RNG = np.random.default_rng(20260923)


def dptr(a):
    return a.ctypes.data_as(C.c_void_p)


def rand_quat(n=None):
    q = RNG.normal(size=(4,) if n is None else (n, 4))
    return q / np.linalg.norm(q, axis=-1, keepdims=True)


# ------------------------------------------------------------------------------------------------ Eigen
def test_slerp_matches_scipy(orc):
    f = orc.fn("se3_interpolate")
    for trial in range(200):
        qa, qb = rand_quat(), rand_quat()
        if trial % 4 == 0:      # nearly equal orientations (the frame-to-frame case)
            qb = qa + 1e-3 * RNG.normal(size=4)
            qb /= np.linalg.norm(qb)
        if trial % 7 == 0:
            qb = -qb            # same rotation, opposite hemisphere: Eigen flips the sign of scale1
        ta, tb = RNG.normal(size=3), RNG.normal(size=3)
        w = float(RNG.uniform(0, 1))
        oq, ot = np.zeros(4), np.zeros(3)
        f(dptr(qa), dptr(ta), dptr(qb), dptr(tb), w, dptr(oq), dptr(ot))
        want = Slerp([0.0, 1.0], Rotation.from_quat([qa, qb]))([w]).as_quat()[0]
        got = oq / np.linalg.norm(oq)
        assert min(np.abs(got - want).max(), np.abs(got + want).max()) < 1e-12
        assert abs(np.linalg.norm(oq) - 1.0) < 1e-12      # unit inputs: the un-normalised result is unit too
        assert np.abs(ot - ((1 - w) * ta + w * tb)).max() < 1e-15


def test_quaternion_from_matrix_and_back(orc):
    fm, tm = orc.fn("quat_from_matrix"), orc.fn("quat_to_matrix")
    specials = [Rotation.from_euler("x", np.pi), Rotation.from_euler("y", np.pi), Rotation.from_euler("z", np.pi),
                Rotation.from_euler("zyx", [3.1, 0.01, -0.02]), Rotation.identity()]
    rots = specials + [Rotation.from_quat(rand_quat()) for _ in range(200)]
    for r in rots:
        R = np.ascontiguousarray(r.as_matrix())
        q = np.zeros(4)
        fm(dptr(R), dptr(q))
        want = r.as_quat()
        assert min(np.abs(q - want).max(), np.abs(q + want).max()) < 1e-12
        R2 = np.zeros((3, 3))
        tm(dptr(q), dptr(R2))
        assert np.abs(R2 - R).max() < 1e-12


def test_symmetric_svd_matches_eigh(orc):
    f = orc.fn("symmetric_svd3")
    for trial in range(200):
        pts = RNG.normal(size=(20, 3)) * np.array([1.0, 0.7, 0.02 if trial % 2 else 0.4])
        pts = pts @ Rotation.from_quat(rand_quat()).as_matrix().T + RNG.normal(size=3) * 50.0
        Cm = np.ascontiguousarray(np.cov(pts.T, bias=True))
        sv, V = np.zeros(3), np.zeros((3, 3))
        f(dptr(Cm), dptr(sv), dptr(V))
        w, U = np.linalg.eigh(Cm)
        assert np.abs(sv - np.abs(w)[::-1]).max() < 1e-12 * max(1.0, sv[0])
        for c in range(3):   # columns = eigenvectors of descending eigenvalue, sign free
            u = U[:, 2 - c]
            assert min(np.abs(V[:, c] - u).max(), np.abs(V[:, c] + u).max()) < 1e-8


def test_ldlt_solve_matches_numpy(orc):
    f = orc.fn("ldlt_solve12")
    for trial in range(100):
        J = RNG.normal(size=(40, 12)) * np.logspace(0, 3, 12)     # badly scaled columns, like rotation vs translation
        A = np.ascontiguousarray(J.T @ J / 40 + 1e-3 * np.eye(12))
        b = RNG.normal(size=12)
        x = np.zeros(12)
        f(dptr(A), dptr(b), dptr(x))
        want = np.linalg.solve(A, b)
        assert np.abs(x - want).max() < 1e-9 * max(1.0, np.abs(want).max())


# ------------------------------------------------------------------------------------------------ Ceres losses
def _opts(orc, loss, sigma=0.1, tol=0.05, iters=5):
    o = orc.default_icp_options()
    o.loss_function = abi.LOSS[loss]
    o.ls_sigma, o.ls_tolerant_min_threshold, o.ls_max_num_iters = sigma, tol, iters
    return o


def rho_published(loss, s, sigma, tol):
    """rho(s) as documented for ceres::LossFunction (s = squared residual norm) and ct_icp::TruncatedLoss."""
    if loss == "STANDARD":
        return s
    if loss == "CAUCHY":          # a^2 log(1 + s / a^2)
        return sigma * sigma * np.log1p(s / (sigma * sigma))
    if loss == "HUBER":           # s for s <= a^2, 2 a sqrt(s) - a^2 beyond
        return s if s <= sigma * sigma else 2 * sigma * np.sqrt(s) - sigma * sigma
    if loss == "TOLERANT":        # b log(1 + e^((s - a) / b)) - b log(1 + e^(-a / b)); ct_icp passes (a, b) = (tol, sigma)
        a, b = tol, sigma
        return b * np.logaddexp(0.0, (s - a) / b) - b * np.logaddexp(0.0, -a / b)
    if loss == "TRUNCATED":       # src/ct_icp/cost_function.cpp:5-15
        return min(s, sigma * sigma)
    raise ValueError(loss)


def rho_published_mp(loss, s, sigma, tol):
    """the same formulas in 40-digit arithmetic (mpmath), for numerical first / second derivatives"""
    mp = mpmath.mp
    sigma, tol = mpmath.mpf(sigma), mpmath.mpf(tol)
    if loss == "STANDARD":
        return s
    if loss == "CAUCHY":
        return sigma * sigma * mp.log(1 + s / (sigma * sigma))
    if loss == "HUBER":
        return s if s <= sigma * sigma else 2 * sigma * mp.sqrt(s) - sigma * sigma
    if loss == "TOLERANT":
        return sigma * mp.log(1 + mp.exp((s - tol) / sigma)) - sigma * mp.log(1 + mp.exp(-tol / sigma))
    return min(s, sigma * sigma)


@pytest.mark.parametrize("loss", ["STANDARD", "CAUCHY", "HUBER", "TOLERANT", "TRUNCATED"])
def test_loss_functions(orc, loss):
    mpmath.mp.dps = 40
    sigma, tol = 0.1, 0.05
    o = _opts(orc, loss, sigma, tol)
    f = orc.fn("loss_evaluate")
    for s in np.concatenate([[0.0, 1e-12], np.logspace(-6, 1, 60)]):
        rho = np.zeros(3)
        f(C.byref(o), float(s), dptr(rho))
        assert abs(rho[0] - rho_published(loss, s, sigma, tol)) < 1e-12 * max(1.0, abs(rho[0]))
        if s < 1e-6 or abs(s - sigma * sigma) < 1e-3 * sigma * sigma:
            continue    # derivative checks away from 0 and from the kinks of Huber / Truncated
        d1 = float(mpmath.diff(lambda t: rho_published_mp(loss, t, sigma, tol), mpmath.mpf(float(s)), 1))
        d2 = float(mpmath.diff(lambda t: rho_published_mp(loss, t, sigma, tol), mpmath.mpf(float(s)), 2))
        assert abs(rho[1] - d1) < 1e-10 * max(1.0, abs(d1)), (s, rho[1], d1)
        assert abs(rho[2] - d2) < 1e-8 * max(1.0, abs(d2)), (s, rho[2], d2)


@pytest.mark.parametrize("loss", ["CAUCHY", "HUBER", "TOLERANT", "TRUNCATED"])
def test_corrector_identities(orc, loss):
    """r_c = rs r, J_c = js J must reproduce the robustified gradient rho' J^T r exactly and, where Ceres applies the
    curvature correction (rho'' > 0: only TolerantLoss here), Triggs' Hessian (rho' + 2 s rho'') J^T J; elsewhere plain
    IRLS (both scaled by sqrt(rho'))."""
    o = _opts(orc, loss)
    fl, fc = orc.fn("loss_evaluate"), orc.fn("corrector")
    for s in np.logspace(-5, 0.5, 40):
        rho = np.zeros(3)
        fl(C.byref(o), float(s), dptr(rho))
        rs, js = C.c_double(), C.c_double()
        fc(float(s), dptr(rho), C.byref(rs), C.byref(js))
        rs, js = rs.value, js.value
        assert abs(rs * js - rho[1]) < 1e-13 * max(1.0, rho[1])                      # gradient: J_c^T r_c = rho' J^T r
        if rho[2] > 0:
            assert loss == "TOLERANT"
            assert abs(js * js - (rho[1] + 2 * s * rho[2])) < 1e-12                    # Gauss-Newton Hessian
        else:
            assert abs(rs - np.sqrt(rho[1])) < 1e-15 and abs(js - np.sqrt(rho[1])) < 1e-15


def test_quaternion_plus(orc):
    f = orc.fn("quat_plus")
    for _ in range(100):
        q, d = rand_quat(), RNG.normal(size=3) * RNG.choice([1e-8, 1e-3, 0.3])
        out = np.zeros(4)
        f(dptr(q), dptr(d), dptr(out))
        want = (Rotation.from_rotvec(2.0 * d) * Rotation.from_quat(q)).as_quat()
        assert min(np.abs(out - want).max(), np.abs(out + want).max()) < 1e-12
    q = rand_quat()
    out = np.zeros(4)
    f(dptr(q), dptr(np.zeros(3)), dptr(out))
    assert np.array_equal(out, q)


# ------------------------------------------------------------------------------- CT functor + ceres::Solve, transcribed
def _qmul(a, b):
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array([aw * bx + ax * bw + ay * bz - az * by, aw * by + ay * bw + az * bx - ax * bz,
                     aw * bz + az * bw + ax * by - ay * bx, aw * bw - ax * bx - ay * by - az * bz])


def _normalized(q):
    return q / np.sqrt(np.sum(q * q))          # no conjugate: analytic in the complex step


def _rotate(q, v):
    qv = q[:3]
    uv = 2.0 * np.cross(qv, v)
    return v + q[3] * uv + np.cross(qv, uv)


def ct_residual(x, blk):
    """CTFunctor<FunctorPointToPlane>::operator() (cost_functions.h:200-219 → :47-57); x = qb(4) qe(4) tb(3) te(3).
    Works on complex x (complex-step differentiation): branches look at real parts only."""
    alpha, ref, raw, normal, weight = blk
    qb, qe, tb, te = _normalized(x[0:4]), _normalized(x[4:8]), x[8:11], x[11:14]
    d = np.sum(qb * qe)
    ad = -d if d.real < 0 else d
    if ad.real >= 1.0 - np.finfo(float).eps:
        s0, s1 = 1.0 - alpha, alpha
    else:
        theta = np.arccos(ad)
        s0, s1 = np.sin((1.0 - alpha) * theta) / np.sin(theta), np.sin(alpha * theta) / np.sin(theta)
    if d.real < 0:
        s1 = -s1
    qi = _normalized(_normalized(s0 * qb + s1 * qe))       # quat_inter.normalize(), then quat.normalized() in the functor
    tr = (1.0 - alpha) * tb + alpha * te
    return weight * np.sum((ref - (_rotate(qi, raw) + tr)) * normal)


def plus(x, delta):
    """ProductParameterization of two EigenQuaternionParameterization and two identity blocks (ct_icp.cpp:221-232)."""
    out = np.array(x, dtype=float)
    for o, d in ((0, delta[0:3]), (4, delta[3:6])):
        n = np.linalg.norm(d)
        if n > 0:
            out[o:o + 4] = _qmul(np.concatenate([np.sin(n) / n * d, [np.cos(n)]]), x[o:o + 4])
    out[8:11] = x[8:11] + delta[6:9]
    out[11:14] = x[11:14] + delta[9:12]
    return out


def plus_jacobian(q):
    x, y, z, w = q
    return np.array([[w, z, -y], [-z, w, x], [y, -x, w], [-x, -y, -z]])


def evaluate(x, blocks, loss, sigma, tol):
    """cost, corrected residuals, corrected LOCAL Jacobian (m x 12) — residual_block.cc + corrector.cc"""
    m = len(blocks)
    r, J = np.zeros(m), np.zeros((m, 12))
    Jb, Je = plus_jacobian(x[0:4]), plus_jacobian(x[4:8])
    cost = 0.0
    h = 1e-30
    for i, blk in enumerate(blocks):
        res = ct_residual(x, blk)
        g = np.zeros(14)
        for k in range(14):
            xc = np.array(x, dtype=complex)
            xc[k] += 1j * h
            g[k] = ct_residual(xc, blk).imag / h
        row = np.concatenate([g[0:4] @ Jb, g[4:8] @ Je, g[8:11], g[11:14]])
        s = res * res
        if loss == "STANDARD":
            rs = js = 1.0
            cost += 0.5 * s
        else:
            hh = max(1e-7 * s, 1e-300)
            rho0 = rho_published(loss, s, sigma, tol)
            # analytic first / second derivatives of the published rho (independent of the oracle's Evaluate)
            if loss == "CAUCHY":
                rho1, rho2 = 1.0 / (1.0 + s / sigma ** 2), -1.0 / sigma ** 2 / (1.0 + s / sigma ** 2) ** 2
            elif loss == "HUBER":
                rho1, rho2 = (1.0, 0.0) if s <= sigma ** 2 else (sigma / np.sqrt(s), -sigma / (2.0 * s ** 1.5))
            elif loss == "TOLERANT":
                e = np.exp((s - tol) / sigma)
                rho1, rho2 = e / (1.0 + e), e / (sigma * (1.0 + e) ** 2)
            else:
                rho1, rho2 = (1.0, 0.0) if s < sigma ** 2 else (0.0, 0.0)
            del hh
            cost += 0.5 * rho0
            sq = np.sqrt(rho1)
            if s == 0.0 or rho2 <= 0.0:
                rs = js = sq
            else:
                D = 1.0 + 2.0 * s * rho2 / rho1
                a = 1.0 - np.sqrt(D)
                rs, js = sq / (1.0 - a), sq * (1.0 - a)
        r[i], J[i] = rs * res, js * row
    return cost, r, J


def lm_reference(x0, blocks, loss, sigma, tol, max_num_iterations):
    """ceres::Solve with default options (trust_region_minimizer.cc, levenberg_marquardt_strategy.cc,
    trust_region_step_evaluator.cc): Jacobi scaling from iteration 0, LM diagonal clamped to [1e-6, 1e32], radius 1e4,
    accept if cost change / model cost change > 1e-3, radius /= max(1/3, 1 - (2 rho - 1)^3) on success and /= 2, 4, 8 …
    on consecutive failures, tolerances 1e-6 / 1e-10 / 1e-8; `parameters` = the lowest-cost point visited."""
    x = np.array(x0, dtype=float)
    best = x.copy()
    x_cost, r, J = evaluate(x, blocks, loss, sigma, tol)
    g = J.T @ r
    scale = 1.0 / (1.0 + np.sqrt(np.sum(J * J, axis=0)))
    J = J * scale
    gmax = np.abs(x - plus(x, -g)).max()
    radius, decrease, reuse, diag = 1e4, 2.0, False, None
    iteration, ok, bad, invalid = 0, 0, 0, 0
    step_ok, min_cost, usable = True, np.inf, True
    x_norm = np.linalg.norm(x)
    while True:
        if step_ok:
            ok += 1
            if x_cost < min_cost:
                min_cost, best = x_cost, x.copy()
        else:
            bad += 1
        if iteration >= max_num_iterations or (step_ok and gmax <= 1e-10) or radius <= 1e-32:
            break
        iteration += 1
        if not reuse:
            diag = np.clip(np.sum(J * J, axis=0), 1e-6, 1e32)
        H = J.T @ J + np.diag(diag / radius)
        step = -np.linalg.solve(H, J.T @ r)
        reuse = True
        Js = J @ step
        model_change = -np.sum(Js * (r + Js / 2.0))
        if not (np.all(np.isfinite(step)) and model_change > 0.0):
            invalid += 1
            if invalid >= 5:
                usable = False
                break
            radius *= 0.5
            step_ok = False
            continue
        invalid = 0
        cand = plus(x, step * scale)
        cand_cost, _, _ = evaluate(cand, blocks, loss, sigma, tol)
        if np.linalg.norm(x - cand) <= 1e-8 * (x_norm + 1e-8):
            break
        change = x_cost - cand_cost
        if abs(change) <= 1e-6 * x_cost:
            break
        rel = change / model_change
        if rel > 1e-3:
            x = cand
            x_norm = np.linalg.norm(x)
            x_cost, r, J = evaluate(x, blocks, loss, sigma, tol)
            g = J.T @ r
            J = J * scale
            gmax = np.abs(x - plus(x, -g)).max()
            step_ok = True
            radius = min(1e16, radius / max(1.0 / 3.0, 1.0 - (2.0 * rel - 1.0) ** 3))
            decrease, reuse = 2.0, False
        else:
            step_ok = False
            radius /= decrease
            decrease *= 2.0
            reuse = True
    return best, min_cost, ok, bad, usable


def _toy_problem(n, noise, seed):
    """n point-to-plane blocks of a box room seen from a sensor that moved by a known pose pair; x0 = identity."""
    rng = np.random.default_rng(seed)
    planes = [(np.array([1.0, 0, 0]), 8.0), (np.array([-1.0, 0, 0]), 7.0), (np.array([0, 1.0, 0]), 5.0),
              (np.array([0, -1.0, 0]), 6.0), (np.array([0, 0, 1.0]), 1.5), (np.array([0, 0, -1.0]), 2.5)]
    qb_t = Rotation.from_rotvec([0.01, -0.02, 0.03]).as_quat()
    qe_t = Rotation.from_rotvec([0.015, -0.01, 0.06]).as_quat()
    tb_t, te_t = np.array([0.10, -0.05, 0.02]), np.array([0.35, -0.02, 0.03])
    blocks, arrays = [], {k: [] for k in ("alpha", "ref", "raw", "normal", "weight")}
    for i in range(n):
        nrm, dist = planes[i % 6]
        alpha = float(rng.uniform(0, 1))
        # world point on the plane n.p = dist, seen from the interpolated true pose
        u, v = rng.uniform(-4, 4, size=2)
        basis = np.linalg.svd(nrm[None, :])[2][1:]
        world = nrm * dist + u * basis[0] + v * basis[1]
        qi = Slerp([0, 1], Rotation.from_quat([qb_t, qe_t]))([alpha])
        ti = (1 - alpha) * tb_t + alpha * te_t
        raw = qi.inv().apply(world - ti)[0] + rng.normal(0, noise, size=3)
        if i % 17 == 0:
            raw += rng.normal(0, 0.5, size=3)         # outliers: the loss functions have something to do
        ref = world + rng.normal(0, noise, size=3) + 0.3 * (u * basis[0])      # another point of the same plane
        ref = ref - nrm * (nrm @ ref - dist)
        weight = float(rng.uniform(0.3, 1.0))
        blk = (alpha, ref, raw, nrm.copy(), weight)
        blocks.append(blk)
        for k, val in zip(("alpha", "ref", "raw", "normal", "weight"), blk):
            arrays[k].append(val)
    return blocks, {k: np.ascontiguousarray(np.array(v, dtype=np.float64)) for k, v in arrays.items()}


@pytest.mark.parametrize("loss,iters", [("STANDARD", 12), ("CAUCHY", 12), ("HUBER", 8), ("TOLERANT", 8), ("TRUNCATED", 5),
                                        ("CAUCHY", 1)])
def test_ceres_solve_schedule_matches_independent_transcription(orc, loss, iters):
    sigma, tol = 0.1, 0.05
    blocks, arr = _toy_problem(60, 0.01, seed=iters * 31 + len(loss))
    x0 = np.array([0, 0, 0, 1.0, 0, 0, 0, 1.0, 0, 0, 0, 0, 0, 0])
    o = _opts(orc, loss, sigma, tol, iters)
    x = x0.copy()
    out = np.zeros(5)
    orc.fn("lm_solve_plane_blocks")(C.byref(o), len(blocks), dptr(arr["alpha"]), dptr(arr["ref"]), dptr(arr["raw"]),
                                    dptr(arr["normal"]), dptr(arr["weight"]), dptr(x), dptr(out))
    best, min_cost, ok, bad, usable = lm_reference(x0, blocks, loss, sigma, tol, iters)
    c0, _, _ = evaluate(x0, blocks, loss, sigma, tol)
    assert abs(out[0] - c0) < 1e-12 * max(1.0, c0)                 # initial cost: residuals + loss
    assert (int(out[2]), int(out[3]), bool(out[4])) == (ok, bad, usable)     # same accept / reject decisions
    assert abs(out[1] - min_cost) < 1e-10 * max(1.0, min_cost)
    assert np.abs(x - best).max() < 1e-9
    assert out[1] < out[0]


def test_ct_functor_jacobian_matches_complex_step(orc):
    """orc_ct_residual's tangent-space Jacobian (Jet autodiff + Plus-Jacobian) vs complex steps on ct_residual()."""
    f = orc.fn("ct_residual")
    for trial in range(50):
        blocks, _ = _toy_problem(1, 0.01, seed=trial)
        alpha, ref, raw, normal, weight = blocks[0]
        x = np.concatenate([rand_quat(), rand_quat(), RNG.normal(size=6)])
        if trial % 2:
            x[4:8] = x[0:4] + 1e-2 * RNG.normal(size=4)
            x[4:8] /= np.linalg.norm(x[4:8])
        jac = np.zeros(12)
        qb, qe, tb, te = (np.ascontiguousarray(v) for v in (x[0:4], x[4:8], x[8:11], x[11:14]))
        val = f(0, alpha, dptr(ref), dptr(raw), dptr(normal), None, weight, dptr(qb), dptr(tb), dptr(qe), dptr(te), dptr(jac))
        _, r, J = evaluate(x, blocks, "STANDARD", 0.1, 0.05)
        assert abs(val - r[0]) < 1e-12
        assert np.abs(jac - J[0]).max() < 1e-10, (jac, J[0])
