// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_math.h header). PARITY UNPINNED.
//
// orc_ceres.cpp — DoRegisterCeres (src/ct_icp/ct_icp.cpp:460-706) with the out-of-tree Ceres arithmetic restated:
//   * ceres::AutoDiffCostFunction<CTFunctor<FunctorPointToPlane>,1,4,3,4,3>   → forward-mode Jet<14>
//       (include/ct_icp/cost_functions.h:32-67, 186-222)
//   * ceres::EigenQuaternionParameterization (Plus / ComputeJacobian)           (ct_icp.cpp:221-232)
//   * ceres::{Cauchy,Huber,Tolerant}Loss, ct_icp::TruncatedLoss, Corrector      (ct_icp.cpp:171-187)
//   * ceres::Solve with TRUST_REGION / LEVENBERG_MARQUARDT and default options  (ct_icp.cpp:489-492, 632-633)
//   * PreviousFrameMotionModel::AddConstraintsToCeresProblem                     (src/ct_icp/motion_model.cpp:12-61)
// Ceres is NOT vendored in the reference (superbuild fetches `master`, unpinned); the routines below restate the
// published Ceres 2.x algorithms (trust_region_minimizer.cc, levenberg_marquardt_strategy.cc, corrector.cc,
// loss_function.cc, local_parameterization.cc).
#include <cfloat>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

#include "orc_icp.h"

namespace orc {
namespace {

constexpr int NG = 14;   // global params: qb(4) tb(3) qe(4) te(3)  [functor argument order]
constexpr int NL = 12;   // local (tangent) params in PROGRAM order: qb(3) qe(3) tb(3) te(3)  (ct_icp.cpp:229-232)

struct Jet {
    double a = 0;
    double v[NG];
    Jet() { for (double &d : v) d = 0; }
    Jet(double s) : a(s) { for (double &d : v) d = 0; }   // NOLINT implicit
    static Jet Var(double s, int k) {
        Jet j(s);
        j.v[k] = 1.0;
        return j;
    }
};
inline Jet operator+(const Jet &x, const Jet &y) { Jet r; r.a = x.a + y.a; for (int i = 0; i < NG; ++i) r.v[i] = x.v[i] + y.v[i]; return r; }
inline Jet operator-(const Jet &x, const Jet &y) { Jet r; r.a = x.a - y.a; for (int i = 0; i < NG; ++i) r.v[i] = x.v[i] - y.v[i]; return r; }
inline Jet operator-(const Jet &x) { Jet r; r.a = -x.a; for (int i = 0; i < NG; ++i) r.v[i] = -x.v[i]; return r; }
inline Jet operator*(const Jet &x, const Jet &y) { Jet r; r.a = x.a * y.a; for (int i = 0; i < NG; ++i) r.v[i] = x.a * y.v[i] + x.v[i] * y.a; return r; }
inline Jet operator/(const Jet &x, const Jet &y) {
    Jet r;
    const double inv = 1.0 / y.a;
    r.a = x.a * inv;
    for (int i = 0; i < NG; ++i) r.v[i] = (x.v[i] - r.a * y.v[i]) * inv;
    return r;
}
inline Jet jsqrt(const Jet &x) { Jet r; r.a = std::sqrt(x.a); const double d = 1.0 / (2.0 * r.a); for (int i = 0; i < NG; ++i) r.v[i] = x.v[i] * d; return r; }
inline Jet jsin(const Jet &x) { Jet r; r.a = std::sin(x.a); const double c = std::cos(x.a); for (int i = 0; i < NG; ++i) r.v[i] = c * x.v[i]; return r; }
inline Jet jacos(const Jet &x) { Jet r; r.a = std::acos(x.a); const double d = -1.0 / std::sqrt(1.0 - x.a * x.a); for (int i = 0; i < NG; ++i) r.v[i] = d * x.v[i]; return r; }
inline Jet jabs(const Jet &x) { return x.a < 0.0 ? -x : x; }

struct JQuat {
    Jet x, y, z, w;
    Jet squaredNorm() const { return x * x + y * y + z * z + w * w; }
    // Eigen MatrixBase::normalized(): if (z > 0) n / sqrt(z)
    JQuat normalized() const {
        Jet z2 = squaredNorm();
        if (z2.a > 0) {
            Jet n = jsqrt(z2);
            return {x / n, y / n, z / n, w / n};
        }
        return *this;
    }
    Jet dot(const JQuat &o) const { return x * o.x + y * o.y + z * o.z + w * o.w; }
    // Eigen QuaternionBase::slerp on Jets (branches on the scalar part)
    JQuat slerp(const Jet &t, const JQuat &other) const {
        const double one = 1.0 - std::numeric_limits<double>::epsilon();
        Jet d = dot(other);
        Jet absD = jabs(d);
        Jet scale0, scale1;
        if (absD.a >= one) {
            scale0 = Jet(1.0) - t;
            scale1 = t;
        } else {
            Jet theta = jacos(absD);
            Jet sinTheta = jsin(theta);
            scale0 = jsin((Jet(1.0) - t) * theta) / sinTheta;
            scale1 = jsin(t * theta) / sinTheta;
        }
        if (d.a < 0) scale1 = -scale1;
        return {scale0 * x + scale1 * other.x, scale0 * y + scale1 * other.y, scale0 * z + scale1 * other.z,
                scale0 * w + scale1 * other.w};
    }
};

struct ResidualBlock {   // CTFunctor<FunctorT> state, cost_functions.h:186-222
    double alpha;
    Vec3 reference, raw, normal;   // normal: reference_normal_ (plane) or direction_ (line)
    double weight;
    int kind = CTICP_DIST_POINT_TO_PLANE;   // inner functor: POINT_TO_PLANE / POINT_TO_LINE / POINT_TO_DISTRIBUTION
    Mat3 information;                       // FunctorPointToDistribution::neighborhood_information_
    bool simple = false;                    // POSE_PARAMETRIZATION SIMPLE: the inner functor on (end_quat, end_t) alone
};

// Eigen::Matrix3d::inverse() (compute_inverse_size3: cofactors, det from the first column)
inline Mat3 Inverse3(const Mat3 &m) {
    auto cof = [&](int i, int j) {
        const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
        return m(i1, j1) * m(i2, j2) - m(i1, j2) * m(i2, j1);
    };
    const double c00 = cof(0, 0), c10 = cof(1, 0), c20 = cof(2, 0);
    const double det = c00 * m(0, 0) + c10 * m(1, 0) + c20 * m(2, 0);
    const double invdet = 1.0 / det;
    Mat3 r;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) r(i, j) = cof(j, i) * invdet;
    return r;
}

// residual and (optionally) its 14 global partials
inline double EvalCTResidual(const ResidualBlock &rb, const double *qb, const double *tb, const double *qe,
                             const double *te, double *global_jac /*14 or null*/) {
    JQuat Qb{Jet::Var(qb[0], 0), Jet::Var(qb[1], 1), Jet::Var(qb[2], 2), Jet::Var(qb[3], 3)};
    Jet Tb[3] = {Jet::Var(tb[0], 4), Jet::Var(tb[1], 5), Jet::Var(tb[2], 6)};
    JQuat Qe{Jet::Var(qe[0], 7), Jet::Var(qe[1], 8), Jet::Var(qe[2], 9), Jet::Var(qe[3], 10)};
    Jet Te[3] = {Jet::Var(te[0], 11), Jet::Var(te[1], 12), Jet::Var(te[2], 13)};

    Jet alpha_m(1.0 - rb.alpha), alpha(rb.alpha);
    JQuat qi = Qe;
    Jet tr[3] = {Te[0], Te[1], Te[2]};
    if (!rb.simple) {   // CTFunctor; with parametrization SIMPLE the functor sees the end pose's blocks directly
        qi = Qb.normalized().slerp(alpha, Qe.normalized());   // cost_functions.h:208-209
        qi = qi.normalized();                                   // :210 quat_inter.normalize()
        for (int k = 0; k < 3; ++k) tr[k] = alpha_m * Tb[k] + alpha * Te[k];
    }

    // FunctorPointToPlane / FunctorPointToDistribution: quat.normalized() * raw + t (cost_functions.h:47-51,163-167);
    // FunctorPointToLine: quat * raw + t (:121-125)
    JQuat q = rb.kind == CTICP_DIST_POINT_TO_LINE ? qi : qi.normalized();
    Jet vx(rb.raw.x), vy(rb.raw.y), vz(rb.raw.z);
    // Eigen _transformVector: uv = q.vec × v; uv += uv; v + w uv + q.vec × uv
    Jet uvx = q.y * vz - q.z * vy, uvy = q.z * vx - q.x * vz, uvz = q.x * vy - q.y * vx;
    uvx = uvx + uvx; uvy = uvy + uvy; uvz = uvz + uvz;
    Jet px = vx + q.w * uvx + (q.y * uvz - q.z * uvy) + tr[0];
    Jet py = vy + q.w * uvy + (q.z * uvx - q.x * uvz) + tr[1];
    Jet pz = vz + q.w * uvz + (q.x * uvy - q.y * uvx) + tr[2];
    Jet res;
    if (rb.kind == CTICP_DIST_POINT_TO_PLANE) {
        Jet product = (Jet(rb.reference.x) - px) * Jet(rb.normal.x) + (Jet(rb.reference.y) - py) * Jet(rb.normal.y) +
                      (Jet(rb.reference.z) - pz) * Jet(rb.normal.z);
        res = Jet(rb.weight) * product;
    } else if (rb.kind == CTICP_DIST_POINT_TO_LINE) {   // cost_functions.h:127-129
        Vec3 dir = rb.normal;
        const double z = dir.squaredNorm();
        if (z > 0) dir = dir * (1.0 / std::sqrt(z));
        Jet dx = px - Jet(rb.reference.x), dy = py - Jet(rb.reference.y), dz = pz - Jet(rb.reference.z);
        Jet cx = Jet(dir.y) * dz - Jet(dir.z) * dy, cy = Jet(dir.z) * dx - Jet(dir.x) * dz, cz = Jet(dir.x) * dy - Jet(dir.y) * dx;
        res = Jet(rb.weight) * jsqrt(cx * cx + cy * cy + cz * cz);
    } else {   // POINT_TO_DISTRIBUTION, cost_functions.h:169-171 : w * diff^T M diff
        Jet d[3] = {px - Jet(rb.reference.x), py - Jet(rb.reference.y), pz - Jet(rb.reference.z)};
        Jet acc(0.0);
        for (int j = 0; j < 3; ++j) {
            Jet row = d[0] * Jet(rb.information(0, j)) + d[1] * Jet(rb.information(1, j)) + d[2] * Jet(rb.information(2, j));
            acc = acc + row * d[j];
        }
        res = Jet(rb.weight) * acc;
    }
    if (global_jac)
        for (int i = 0; i < NG; ++i) global_jac[i] = res.v[i];
    return res.a;
}
inline double EvalCTPointToPlane(const ResidualBlock &rb, const double *qb, const double *tb, const double *qe,
                                 const double *te, double *global_jac) {
    return EvalCTResidual(rb, qb, tb, qe, te, global_jac);
}

// ceres::EigenQuaternionParameterization::ComputeJacobian (4x3, rows x,y,z,w)
inline void QuatLocalJacobian(const double *x, double J[4][3]) {
    J[0][0] = x[3];  J[0][1] = x[2];  J[0][2] = -x[1];
    J[1][0] = -x[2]; J[1][1] = x[3];  J[1][2] = x[0];
    J[2][0] = x[1];  J[2][1] = -x[0]; J[2][2] = x[3];
    J[3][0] = -x[0]; J[3][1] = -x[1]; J[3][2] = -x[2];
}
// ceres::EigenQuaternionParameterization::Plus : q+ = [sin|d|/|d| d, cos|d|] ⊗ q
inline void QuatPlus(const double *x, const double *delta, double *out) {
    const double n = std::sqrt(delta[0] * delta[0] + delta[1] * delta[1] + delta[2] * delta[2]);
    if (n > 0.0) {
        const double s = std::sin(n) / n;
        Quat dq(s * delta[0], s * delta[1], s * delta[2], std::cos(n));
        Quat r = dq * Quat(x[0], x[1], x[2], x[3]);
        out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
    } else {
        for (int i = 0; i < 4; ++i) out[i] = x[i];
    }
}

// Loss functions: rho[0..2] at s (ceres loss_function.cc; TruncatedLoss src/ct_icp/cost_function.cpp:5-15)
struct Loss {
    int type;
    double a, b, c;   // meaning per type
    Loss(const cticp_icp_options &o) : type(o.loss_function), a(0), b(0), c(0) {
        switch (type) {
            case CTICP_LOSS_CAUCHY: b = o.ls_sigma * o.ls_sigma; c = 1.0 / b; break;
            case CTICP_LOSS_HUBER: a = o.ls_sigma; b = a * a; break;
            case CTICP_LOSS_TOLERANT:   // TolerantLoss(a = ls_tolerant_min_threshold, b = ls_sigma), ct_icp.cpp:181-182
                a = o.ls_tolerant_min_threshold; b = o.ls_sigma; c = b * std::log(1.0 + std::exp(-a / b)); break;
            case CTICP_LOSS_TRUNCATED: b = o.ls_sigma * o.ls_sigma; break;
            default: break;
        }
    }
    bool present() const { return type != CTICP_LOSS_STANDARD; }
    void Evaluate(double s, double rho[3]) const {
        switch (type) {
            case CTICP_LOSS_CAUCHY: {
                const double sum = 1.0 + s * c, inv = 1.0 / sum;
                rho[0] = b * std::log(sum);
                rho[1] = std::max(DBL_MIN, inv);
                rho[2] = -c * (inv * inv);
                break;
            }
            case CTICP_LOSS_HUBER:
                if (s > b) {
                    const double r = std::sqrt(s);
                    rho[0] = 2.0 * a * r - b;
                    rho[1] = std::max(DBL_MIN, a / r);
                    rho[2] = -rho[1] / (2.0 * s);
                } else {
                    rho[0] = s; rho[1] = 1.0; rho[2] = 0.0;
                }
                break;
            case CTICP_LOSS_TOLERANT: {
                const double x = (s - a) / b;
                const double kLog2Pow53 = 36.7;
                if (x > kLog2Pow53) {
                    rho[0] = s - a - c; rho[1] = 1.0; rho[2] = 0.0;
                } else {
                    const double e_x = std::exp(x);
                    rho[0] = b * std::log(1.0 + e_x) - c;
                    rho[1] = std::max(DBL_MIN, e_x / (1.0 + e_x));
                    rho[2] = 0.5 / (b * (1.0 + std::cosh(x)));
                }
                break;
            }
            case CTICP_LOSS_TRUNCATED:
                if (s < b) { rho[0] = s; rho[1] = 1.0; rho[2] = 0.0; }
                else { rho[0] = b; rho[1] = 0.0; rho[2] = 0.0; }
                break;
            default:
                rho[0] = s; rho[1] = 1.0; rho[2] = 0.0;
        }
    }
};

// ceres::internal::Corrector for a 1-dimensional residual: r_c = rs * r, J_c = js * J
inline void Corrector1D(double sq_norm, const double rho[3], double &residual_scale, double &jacobian_scale) {
    const double sqrt_rho1 = std::sqrt(rho[1]);
    if (sq_norm == 0.0 || rho[2] <= 0.0) {
        residual_scale = sqrt_rho1;
        jacobian_scale = sqrt_rho1;
        return;
    }
    const double D = 1.0 + 2.0 * sq_norm * rho[2] / rho[1];
    const double alpha = 1.0 - std::sqrt(D);
    residual_scale = sqrt_rho1 / (1.0 - alpha);
    // J_c = sqrt_rho1 (J - alpha/s r r^T J) = sqrt_rho1 (1 - alpha) J for a scalar residual
    jacobian_scale = sqrt_rho1 * (1.0 - alpha);
}

struct Problem {
    std::vector<ResidualBlock> blocks;
    Loss loss;
    // regularisers (no loss), motion_model.cpp:12-61
    bool has_location = false, has_orientation = false, has_cv = false, has_small = false;
    double w_location = 0, w_orientation = 0, w_cv = 0, w_small = 0;
    Vec3 prev_end_tr, prev_velocity;
    Quat prev_orientation;
    int num_threads = 1;
    bool simple = false;   // parametrization SIMPLE: only end_quat / end_t are parameter blocks (ct_icp.cpp:234-237)
    explicit Problem(const cticp_icp_options &o) : loss(o) {}

    int NumRegResiduals() const { return (has_location ? 3 : 0) + (has_orientation ? 1 : 0) + (has_cv ? 3 : 0) + (has_small ? 3 : 0); }

    // params: x[14] = qb(4) qe(4) tb(3) te(3)  (program order)
    // Evaluates cost; if residuals/jacobian are given, fills CORRECTED residuals r (m) and local Jacobian J (m x 12)
    double Evaluate(const double *x, std::vector<double> *r, std::vector<double> *J) const {
        const double *qb = x, *qe = x + 4, *tb = x + 8, *te = x + 11;
        const int nb = (int) blocks.size();
        const int m = nb + NumRegResiduals();
        if (r) r->assign(m, 0.0);
        if (J) J->assign(size_t(m) * NL, 0.0);
        double Jqb[4][3], Jqe[4][3];
        QuatLocalJacobian(qb, Jqb);
        QuatLocalJacobian(qe, Jqe);
        double cost = 0;
#pragma omp parallel for num_threads(num_threads) reduction(+ : cost)
        for (int i = 0; i < nb; ++i) {
            double g[NG];
            double res = EvalCTResidual(blocks[i], qb, tb, qe, te, J ? g : nullptr);
            double sq = res * res;
            double rs = 1.0, js = 1.0;
            if (loss.present()) {
                double rho[3];
                loss.Evaluate(sq, rho);
                cost += 0.5 * rho[0];
                if (r || J) Corrector1D(sq, rho, rs, js);
            } else
                cost += 0.5 * sq;
            if (J) {
                double *row = J->data() + size_t(i) * NL;
                for (int k = 0; k < 3; ++k) {
                    row[k] = js * (g[0] * Jqb[0][k] + g[1] * Jqb[1][k] + g[2] * Jqb[2][k] + g[3] * Jqb[3][k]);
                    row[3 + k] = js * (g[7] * Jqe[0][k] + g[8] * Jqe[1][k] + g[9] * Jqe[2][k] + g[10] * Jqe[3][k]);
                    row[6 + k] = js * g[4 + k];
                    row[9 + k] = js * g[11 + k];
                }
            }
            if (r) (*r)[i] = rs * res;
        }
        int row = nb;
        if (has_location) {   // LocationConsistencyFunctor on begin_t, cost_functions.h:271-292
            for (int k = 0; k < 3; ++k) {
                double res = w_location * (tb[k] - prev_end_tr[k]);
                cost += 0.5 * res * res;
                if (r) (*r)[row + k] = res;
                if (J) (*J)[size_t(row + k) * NL + 6 + k] = w_location;
            }
            row += 3;
        }
        if (has_orientation) {   // OrientationConsistencyFunctor on begin_quat, cost_functions.h:295-314
            double s = qb[0] * prev_orientation.x + qb[1] * prev_orientation.y + qb[2] * prev_orientation.z +
                       qb[3] * prev_orientation.w;
            double res = w_orientation * (1.0 - s * s);
            cost += 0.5 * res * res;
            if (r) (*r)[row] = res;
            if (J) {
                double gq[4] = {-2.0 * w_orientation * s * prev_orientation.x, -2.0 * w_orientation * s * prev_orientation.y,
                                -2.0 * w_orientation * s * prev_orientation.z, -2.0 * w_orientation * s * prev_orientation.w};
                for (int k = 0; k < 3; ++k)
                    (*J)[size_t(row) * NL + k] = gq[0] * Jqb[0][k] + gq[1] * Jqb[1][k] + gq[2] * Jqb[2][k] + gq[3] * Jqb[3][k];
            }
            row += 1;
        }
        if (has_cv) {   // ConstantVelocityFunctor(begin_t, end_t), cost_functions.h:317-337
            for (int k = 0; k < 3; ++k) {
                double res = w_cv * (te[k] - tb[k] - prev_velocity[k]);
                cost += 0.5 * res * res;
                if (r) (*r)[row + k] = res;
                if (J) {
                    (*J)[size_t(row + k) * NL + 6 + k] = -w_cv;
                    (*J)[size_t(row + k) * NL + 9 + k] = w_cv;
                }
            }
            row += 3;
        }
        if (has_small) {   // SmallVelocityFunctor, cost_functions.h:340-354
            for (int k = 0; k < 3; ++k) {
                double res = w_small * (tb[k] - te[k]);
                cost += 0.5 * res * res;
                if (r) (*r)[row + k] = res;
                if (J) {
                    (*J)[size_t(row + k) * NL + 6 + k] = w_small;
                    (*J)[size_t(row + k) * NL + 9 + k] = -w_small;
                }
            }
            row += 3;
        }
        return cost;
    }
    static void Plus(const double *x, const double *delta, double *out) {
        QuatPlus(x, delta, out);
        QuatPlus(x + 4, delta + 3, out + 4);
        for (int k = 0; k < 3; ++k) out[8 + k] = x[8 + k] + delta[6 + k];
        for (int k = 0; k < 3; ++k) out[11 + k] = x[11 + k] + delta[9 + k];
    }
};

struct SolveSummary {
    bool usable = true;
    int num_successful = 0, num_unsuccessful = 0;
    double initial_cost = 0, final_cost = 0;
};

// ceres::Solve — TrustRegionMinimizer::Minimize with LevenbergMarquardtStrategy, default options except
// max_num_iterations (trust_region_minimizer.cc, levenberg_marquardt_strategy.cc)
SolveSummary SolveLM(const Problem &problem, double *parameters, int max_num_iterations) {
    const double initial_radius = 1e4, max_radius = 1e16, min_radius = 1e-32;
    const double min_lm_diagonal = 1e-6, max_lm_diagonal = 1e32;
    const double min_relative_decrease = 1e-3;
    const double function_tolerance = 1e-6, gradient_tolerance = 1e-10, parameter_tolerance = 1e-8;
    const int max_num_consecutive_invalid_steps = 5;

    SolveSummary summary;
    double x[14], candidate_x[14];
    for (int i = 0; i < 14; ++i) x[i] = parameters[i];
    // (parametrization SIMPLE: the problem's parameter vector is end_quat + end_t only; the begin blocks stay in x with
    //  all-zero Jacobian columns, which decouple in the damped normal equations and get a zero step)
    const bool simple = problem.simple;
    auto is_active = [simple](int i) { return !simple || (i >= 4 && i < 8) || i >= 11; };
    auto norm14 = [&](const double *v) { double s = 0; for (int i = 0; i < 14; ++i) if (is_active(i)) s += v[i] * v[i]; return std::sqrt(s); };
    double x_norm = norm14(x);
    double x_cost = 0, minimum_cost = std::numeric_limits<double>::max();
    std::vector<double> residuals, jacobian;
    double gradient[NL], jacobian_scaling[NL];
    double gradient_max_norm = 0;
    int iteration = 0;

    // EvaluateGradientAndJacobian
    auto evaluate = [&](bool first) {
        x_cost = problem.Evaluate(x, &residuals, &jacobian);
        const size_t m = residuals.size();
        for (int j = 0; j < NL; ++j) gradient[j] = 0;
        for (size_t i = 0; i < m; ++i)
            for (int j = 0; j < NL; ++j) gradient[j] += jacobian[i * NL + j] * residuals[i];
        if (first) {   // jacobi_scaling computed once at iteration 0
            for (int j = 0; j < NL; ++j) {
                double s = 0;
                for (size_t i = 0; i < m; ++i) s += jacobian[i * NL + j] * jacobian[i * NL + j];
                jacobian_scaling[j] = 1.0 / (1.0 + std::sqrt(s));
            }
        }
        for (size_t i = 0; i < m; ++i)
            for (int j = 0; j < NL; ++j) jacobian[i * NL + j] *= jacobian_scaling[j];
        double neg_g[NL], proj[14];
        for (int j = 0; j < NL; ++j) neg_g[j] = -gradient[j];
        Problem::Plus(x, neg_g, proj);
        gradient_max_norm = 0;
        for (int i = 0; i < 14; ++i) gradient_max_norm = std::max(gradient_max_norm, std::abs(x[i] - proj[i]));
    };

    evaluate(true);   // IterationZero
    summary.initial_cost = x_cost;
    bool step_is_successful = true;

    // LevenbergMarquardtStrategy state
    double radius = initial_radius, decrease_factor = 2.0;
    bool reuse_diagonal = false;
    double diagonal[NL];
    int num_consecutive_invalid_steps = 0;

    while (true) {
        // FinalizeIterationAndCheckIfMinimizerCanContinue
        if (step_is_successful) {
            summary.num_successful++;
            if (x_cost < minimum_cost) {
                minimum_cost = x_cost;
                for (int i = 0; i < 14; ++i) parameters[i] = x[i];
            }
        } else
            summary.num_unsuccessful++;
        if (iteration >= max_num_iterations) break;
        if (step_is_successful && gradient_max_norm <= gradient_tolerance) break;
        if (radius <= min_radius) break;
        iteration++;

        // ComputeTrustRegionStep → LevenbergMarquardtStrategy::ComputeStep
        const size_t m = residuals.size();
        if (!reuse_diagonal) {
            for (int j = 0; j < NL; ++j) {
                double s = 0;
                for (size_t i = 0; i < m; ++i) s += jacobian[i * NL + j] * jacobian[i * NL + j];
                diagonal[j] = std::min(std::max(s, min_lm_diagonal), max_lm_diagonal);
            }
        }
        double H[NL][NL];
        std::array<double, NL> rhs{};
        for (int a = 0; a < NL; ++a) {
            for (int b = 0; b < NL; ++b) H[a][b] = 0;
        }
        for (size_t i = 0; i < m; ++i) {
            const double *row = &jacobian[i * NL];
            for (int a = 0; a < NL; ++a) {
                rhs[a] += row[a] * residuals[i];
                for (int b = a; b < NL; ++b) H[a][b] += row[a] * row[b];
            }
        }
        for (int a = 0; a < NL; ++a)
            for (int b = 0; b < a; ++b) H[a][b] = H[b][a];
        for (int a = 0; a < NL; ++a) H[a][a] += diagonal[a] / radius;   // lm_diagonal^2
        std::array<double, NL> step = LDLTSolve<NL>(H, rhs);              // solves (J'J + D^2) y = J' r
        bool step_valid_numbers = true;
        for (double &s : step) {
            if (!std::isfinite(s)) step_valid_numbers = false;
            s = -s;
        }
        reuse_diagonal = true;

        // model_cost_change = -(J step)'(f + J step / 2)
        double model_cost_change = 0;
        if (step_valid_numbers) {
            for (size_t i = 0; i < m; ++i) {
                double js = 0;
                const double *row = &jacobian[i * NL];
                for (int a = 0; a < NL; ++a) js += row[a] * step[a];
                model_cost_change -= js * (residuals[i] + js / 2.0);
            }
        }
        const bool step_is_valid = step_valid_numbers && model_cost_change > 0.0;
        if (!step_is_valid) {   // HandleInvalidStep
            if (++num_consecutive_invalid_steps >= max_num_consecutive_invalid_steps) {
                summary.usable = false;
                break;
            }
            radius *= 0.5;   // StepIsInvalid
            reuse_diagonal = true;
            step_is_successful = false;
            continue;
        }
        num_consecutive_invalid_steps = 0;
        double delta[NL];
        for (int j = 0; j < NL; ++j) delta[j] = step[j] * jacobian_scaling[j];

        // ComputeCandidatePointAndEvaluateCost
        Problem::Plus(x, delta, candidate_x);
        double candidate_cost = problem.Evaluate(candidate_x, nullptr, nullptr);

        // ParameterToleranceReached
        double step_norm = 0;
        for (int i = 0; i < 14; ++i) step_norm += (x[i] - candidate_x[i]) * (x[i] - candidate_x[i]);
        step_norm = std::sqrt(step_norm);
        if (step_norm <= parameter_tolerance * (x_norm + parameter_tolerance)) break;
        // FunctionToleranceReached
        const double cost_change = x_cost - candidate_cost;
        if (std::fabs(cost_change) <= function_tolerance * x_cost) break;

        const double relative_decrease = cost_change / model_cost_change;
        if (getenv("ORC_DEBUG_LM"))
            fprintf(stderr, "[orc-lm] x_cost %.12g cand %.12g x %.12g %.12g %.12g %.12g | %.12g %.12g %.12g %s\n", x_cost,
                    candidate_cost, x[0], x[1], x[2], x[3], x[11], x[12], x[13], relative_decrease > min_relative_decrease ? "ACCEPT" : "reject");
        if (relative_decrease > min_relative_decrease) {   // HandleSuccessfulStep
            for (int i = 0; i < 14; ++i) x[i] = candidate_x[i];
            x_norm = norm14(x);
            evaluate(false);
            step_is_successful = true;
            radius = radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * relative_decrease - 1.0, 3));   // StepAccepted
            radius = std::min(max_radius, radius);
            decrease_factor = 2.0;
            reuse_diagonal = false;
        } else {
            step_is_successful = false;
            radius = radius / decrease_factor;   // StepRejected
            decrease_factor *= 2.0;
            reuse_diagonal = true;
        }
    }
    summary.final_cost = minimum_cost;
    return summary;
}

}  // namespace

// KAT tap: residual of one CTFunctor<FunctorT> (kind = CTICP_DIST_*) and its 12 tangent-space partials (program order).
// `dir` is the plane normal or the line direction; `covariance` (row-major 3x3) feeds the distribution functor.
double CTResidualKindForTest(int kind, double alpha, const double ref[3], const double raw[3], const double dir[3],
                             const double *covariance, double weight, const double qb[4], const double tb[3],
                             const double qe[4], const double te[3], double *local_jac12) {
    ResidualBlock rb;
    rb.kind = kind;
    rb.alpha = alpha;
    rb.reference = Vec3(ref[0], ref[1], ref[2]);
    rb.raw = Vec3(raw[0], raw[1], raw[2]);
    rb.normal = Vec3(dir[0], dir[1], dir[2]);
    rb.weight = weight;
    if (covariance) {
        Mat3 m;
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) m(i, j) = covariance[3 * i + j] + (i == j ? 0.05 : 0.0);
        rb.information = Inverse3(m);
    }
    double g[NG];
    double r = EvalCTResidual(rb, qb, tb, qe, te, g);
    if (local_jac12) {
        double Jqb[4][3], Jqe[4][3];
        QuatLocalJacobian(qb, Jqb);
        QuatLocalJacobian(qe, Jqe);
        for (int k = 0; k < 3; ++k) {
            local_jac12[k] = g[0] * Jqb[0][k] + g[1] * Jqb[1][k] + g[2] * Jqb[2][k] + g[3] * Jqb[3][k];
            local_jac12[3 + k] = g[7] * Jqe[0][k] + g[8] * Jqe[1][k] + g[9] * Jqe[2][k] + g[10] * Jqe[3][k];
            local_jac12[6 + k] = g[4 + k];
            local_jac12[9 + k] = g[11 + k];
        }
    }
    return r;
}
double CTResidualForTest(double alpha, const double ref[3], const double raw[3], const double normal[3], double weight,
                         const double qb[4], const double tb[3], const double qe[4], const double te[3],
                         double *local_jac12) {
    return CTResidualKindForTest(CTICP_DIST_POINT_TO_PLANE, alpha, ref, raw, normal, nullptr, weight, qb, tb, qe, te,
                                 local_jac12);
}

// KAT taps for tests/test_math_pins.py (checked there against numpy / scipy and an independent Python transcription of
// ceres::Solve): the loss functions, the Corrector, the quaternion Plus, and SolveLM on explicit point-to-plane blocks.
void LossEvaluateForTest(const cticp_icp_options &o, double s, double rho[3]) { Loss(o).Evaluate(s, rho); }
void CorrectorForTest(double s, const double rho[3], double *residual_scale, double *jacobian_scale) {
    Corrector1D(s, rho, *residual_scale, *jacobian_scale);
}
void QuatPlusForTest(const double q[4], const double delta[3], double out[4]) { QuatPlus(q, delta, out); }
// x: 14 parameters in program order qb(4) qe(4) tb(3) te(3), optimised in place; out = {initial cost, final cost,
// successful steps, unsuccessful steps, usable}
void SolveLMForTest(const cticp_icp_options &o, int n, const double *alpha, const double *ref, const double *raw,
                    const double *normal, const double *weight, double *x, double out[5]) {
    Problem problem(o);
    problem.num_threads = 1;
    for (int i = 0; i < n; ++i) {
        ResidualBlock rb;
        rb.kind = CTICP_DIST_POINT_TO_PLANE;
        rb.alpha = alpha[i];
        rb.reference = Vec3(ref[3 * i], ref[3 * i + 1], ref[3 * i + 2]);
        rb.raw = Vec3(raw[3 * i], raw[3 * i + 1], raw[3 * i + 2]);
        rb.normal = Vec3(normal[3 * i], normal[3 * i + 1], normal[3 * i + 2]);
        rb.weight = weight[i];
        problem.blocks.push_back(rb);
    }
    const SolveSummary sum = SolveLM(problem, x, o.ls_max_num_iters);
    out[0] = sum.initial_cost; out[1] = sum.final_cost; out[2] = sum.num_successful; out[3] = sum.num_unsuccessful;
    out[4] = sum.usable ? 1.0 : 0.0;
}

// One ICP iteration's tail shared by solvers CERES and ROBUST: GetProblem (ct_icp.cpp:409-424), the motion model's
// regularisers, ceres::Solve, the stop criterion (:613-672 / :1291-1336). Returns 0 = continue, 1 = converged,
// 2 = not enough residuals (`failed` filled).
static int AssembleAndSolve(const cticp_icp_options &options, const std::vector<ResidualBlock> &all_blocks,
                            const std::vector<char> &has_block, const MotionModel *motion_model, TrajectoryFrame &frame,
                            SE3 &previous_begin_pose, SE3 &previous_end_pose, int &number_of_residuals, int num_threads,
                            ICPSummary &failed) {
    Problem problem(options);   // GetProblem, :409-424 : first max_num_residuals non-null functors
    problem.num_threads = num_threads;
    problem.simple = options.parametrization == CTICP_PARAM_SIMPLE;
    number_of_residuals = 0;
    for (size_t i = 0; i < all_blocks.size(); ++i) {
        if (!has_block[i]) continue;
        if (options.max_num_residuals <= 0 || number_of_residuals < options.max_num_residuals) {
            problem.blocks.push_back(all_blocks[i]);
            number_of_residuals++;
        }
    }
    if (motion_model && motion_model->present && options.parametrization == CTICP_PARAM_CONTINUOUS_TIME) {   // :613
        const auto &mo = motion_model->options;
        const auto &prev = motion_model->previous_frame;
        problem.prev_velocity = prev.EndTr() - prev.BeginTr();
        problem.prev_orientation = prev.EndQuat();
        problem.prev_end_tr = prev.EndTr();
        if (mo.beta_location_consistency > 0.) {
            problem.has_location = true;
            problem.w_location = std::sqrt(number_of_residuals * mo.beta_location_consistency);
        }
        if (mo.beta_orientation_consistency > 0.) {
            problem.has_orientation = true;
            problem.w_orientation = std::sqrt(number_of_residuals * mo.beta_orientation_consistency);
        }
        if (mo.beta_constant_velocity > 0.) {
            problem.has_cv = true;
            problem.w_cv = std::sqrt(number_of_residuals * mo.beta_constant_velocity);
        }
        if (mo.beta_small_velocity > 0.) {
            problem.has_small = true;
            problem.w_small = std::sqrt(number_of_residuals * mo.beta_small_velocity);
        }
    }
    if (number_of_residuals < options.min_number_neighbors) {   // :617 (sic: compares with min_number_neighbors)
        std::stringstream ss;
        ss << "[CT_ICP] Error : not enough keypoints selected in ct-icp !" << std::endl;
        ss << "[CT_ICP] number_of_residuals : " << number_of_residuals << std::endl;
        failed.success = false;
        failed.num_residuals_used = number_of_residuals;
        failed.error_log = ss.str();
        return 2;
    }

    double params[14];
    auto pack = [&]() {
        const Quat &qb = frame.begin_pose.pose.quat, &qe = frame.end_pose.pose.quat;
        params[0] = qb.x; params[1] = qb.y; params[2] = qb.z; params[3] = qb.w;
        params[4] = qe.x; params[5] = qe.y; params[6] = qe.z; params[7] = qe.w;
        for (int d = 0; d < 3; ++d) {
            params[8 + d] = frame.begin_pose.pose.tr[d];
            params[11 + d] = frame.end_pose.pose.tr[d];
        }
    };
    pack();
    SolveSummary ss = SolveLM(problem, params, options.ls_max_num_iters);
    frame.begin_pose.pose.quat = Quat(params[0], params[1], params[2], params[3]);
    frame.end_pose.pose.quat = Quat(params[4], params[5], params[6], params[7]);
    frame.begin_pose.pose.tr = Vec3(params[8], params[9], params[10]);
    frame.end_pose.pose.tr = Vec3(params[11], params[12], params[13]);
    frame.begin_pose.pose.quat.normalize();
    frame.end_pose.pose.quat.normalize();
    if (!ss.usable) throw std::runtime_error("Error During Optimization");

    double diff_trans = (previous_begin_pose.tr - frame.BeginTr()).norm() +
                        (previous_end_pose.tr - frame.EndTr()).norm();
    double diff_rot = AngularDistance(frame.begin_pose.pose, previous_begin_pose) +
                      AngularDistance(frame.end_pose.pose, previous_end_pose);
    previous_begin_pose = frame.begin_pose.pose;
    previous_end_pose = frame.end_pose.pose;
    return (diff_rot < options.threshold_orientation_norm && diff_trans < options.threshold_translation_norm) ? 1 : 0;
}

// DoRegisterCeres, src/ct_icp/ct_icp.cpp:460-706 (POINT_TO_PLANE; parametrizations CONTINUOUS_TIME and SIMPLE)
ICPSummary DoRegisterCeres(const VoxelMap &map, const cticp_icp_options &options,
                           const cticp_strategy_options &strategy, std::vector<WPoint3D> &kpts,
                           TrajectoryFrame &frame, const MotionModel *motion_model) {
    if (options.distance != CTICP_DIST_POINT_TO_PLANE)
        throw std::runtime_error("oracle: only POINT_TO_PLANE is restated for the CERES solver");
    const bool simple = options.parametrization == CTICP_PARAM_SIMPLE;
    ICPSummary icp_summary;
    const size_t num_points = kpts.size();
    frame.begin_pose.pose.quat.normalize();
    frame.end_pose.pose.quat.normalize();
    const int kMinNumNeighbors = options.min_number_neighbors;
    const int num_threads = std::max(1, options.ls_num_threads);

    SE3 previous_begin_pose = frame.begin_pose.pose, previous_end_pose = frame.end_pose.pose;
    int number_of_residuals = 0;

    // ICPOptimizationBuilder::corrected_raw_points_ (= the raw points) and DistortFrame (:198-215): with parametrization
    // SIMPLE every raw point is moved into the coordinate frame of the END pose of the acquisition
    std::vector<Vec3> corrected_raw(num_points);
    for (size_t i = 0; i < num_points; ++i) corrected_raw[i] = kpts[i].raw;
    auto distort_frame = [&]() {
        if (!simple) return;
        const SE3 end_pose_I = frame.end_pose.Inverse().pose;
        for (size_t i = 0; i < num_points; ++i)
            corrected_raw[i] = end_pose_I * (frame.begin_pose.InterpolatePose(frame.end_pose, kpts[i].timestamp) * kpts[i].raw);
    };
    if (options.point_to_plane_with_distortion) distort_frame();   // :512-514
    auto transform_keypoints = [&]() {   // :516-531
        for (auto &kp : kpts) {
            if (options.point_to_plane_with_distortion || !simple)
                kp.world = frame.begin_pose.InterpolatePose(frame.end_pose, kp.timestamp) * kp.raw;
            else
                kp.world = frame.end_pose.pose * kp.raw;
        }
    };
    double lambda_weight = std::abs(options.weight_alpha);
    double lambda_neighborhood = std::abs(options.weight_neighborhood);
    const double kMaxPointToPlane = options.max_dist_to_plane_ct_icp;
    const double sum = lambda_weight + lambda_neighborhood;
    if (!(sum > 0.0)) throw std::runtime_error("Invalid requirement: weight_alpha + weight_neighborhood <= 0");
    lambda_weight /= sum;
    lambda_neighborhood /= sum;

    const int ncn = options.num_closest_neighbors;
    std::vector<Neighborhood> neighborhoods(num_points);
    std::vector<char> has_block;
    std::vector<ResidualBlock> all_blocks;
    int iter = 0;
    for (; iter < options.num_iters_icp; iter++) {
        transform_keypoints();
        has_block.assign(num_points * ncn, 0);
        all_blocks.assign(num_points * ncn, ResidualBlock());
        size_t kp_iters = 0, stencil_sum = 0;
#pragma omp parallel for num_threads(num_threads) reduction(+ : kp_iters, stencil_sum)
        for (long k = 0; k < (long) num_points; ++k) {   // :561-604
            const WPoint3D &pt = kpts[k];
            auto &neighborhood = neighborhoods[k];
            size_t st = 0;
            // const_strategy->ComputeNeighborhoodInPlace(voxels_map, pt, neighborhoods[k], &end_t), :571
            StrategyComputeNeighborhoodInPlace(strategy, map, pt.raw, pt.world, neighborhood, &frame.end_pose.pose.tr, &st);
            kp_iters++;
            stencil_sum += st;
            if ((int) neighborhood.points.size() < kMinNumNeighbors) continue;
            neighborhood.ComputeNeighborhood();
            // normal flip test at :578 is a no-op (BeginTr - BeginTr)
            double weight = std::pow(neighborhood.description.a2D, options.power_planarity);
            weight = lambda_weight * weight +
                     lambda_neighborhood *
                         std::exp(-(neighborhood.points[0] - pt.world).norm() / (kMaxPointToPlane * kMinNumNeighbors));
            double alpha = frame.begin_pose.GetAlphaTimestamp(pt.timestamp, frame.end_pose);
            for (int i = 0; i < ncn; ++i) {
                ResidualBlock rb;
                rb.alpha = alpha;
                rb.reference = neighborhood.points[i];
                rb.raw = corrected_raw[k];   // SetResidualBlock reads corrected_raw_points_ (:381)
                rb.simple = simple;
                rb.normal = neighborhood.description.normal;
                rb.weight = weight;
                all_blocks[ncn * k + i] = rb;
                has_block[ncn * k + i] = 1;
            }
        }
        icp_summary.keypoint_iterations += kp_iters;
        icp_summary.stencil_points += stencil_sum;

        ICPSummary failed;
        const int status = AssembleAndSolve(options, all_blocks, has_block, motion_model, frame, previous_begin_pose,
                                            previous_end_pose, number_of_residuals, num_threads, failed);
        if (status == 2) {
            failed.keypoint_iterations = icp_summary.keypoint_iterations;
            failed.stencil_points = icp_summary.stencil_points;
            return failed;
        }
        if (options.point_to_plane_with_distortion) distort_frame();   // :657-659 (before the stop test's break)
        if (status == 1) break;
    }
    transform_keypoints();
    icp_summary.success = true;
    icp_summary.num_residuals_used = number_of_residuals;
    icp_summary.num_iters = iter;
    frame.begin_pose.pose.quat.normalize();
    frame.end_pose.pose.quat.normalize();
    return icp_summary;
}

// DoRegisterRobust, src/ct_icp/ct_icp.cpp:1180-1370 (CONTINUOUS_TIME): neighborhoods classified planar / linear /
// other, one point-to-plane / point-to-line / point-to-distribution residual per keypoint.
ICPSummary DoRegisterRobust(const VoxelMap &map, const cticp_icp_options &options, std::vector<WPoint3D> &kpts,
                            TrajectoryFrame &frame, const MotionModel *motion_model) {
    if (options.parametrization != CTICP_PARAM_CONTINUOUS_TIME)
        throw std::runtime_error("oracle: only CONTINUOUS_TIME is restated for the ROBUST solver");
    enum { NONE = 0, LINEAR = 1, PLANAR = 2, VOLUMIC = 3 };   // slam::NEIGHBORHOOD_TYPE, neighborhood.h:138-143
    ICPSummary icp_summary;
    const size_t num_points = kpts.size();
    const int kMinNumNeighbors = options.min_number_neighbors;
    const int num_threads = std::max(1, options.ls_num_threads);
    SE3 previous_begin_pose = frame.begin_pose.pose, previous_end_pose = frame.end_pose.pose;   // OptimizationTracker
    int number_of_residuals = -1;
    auto transform_keypoints = [&]() {   // TransformKeyPoints, :1373-1393
        for (auto &kp : kpts) kp.world = frame.begin_pose.InterpolatePose(frame.end_pose, kp.timestamp) * kp.raw;
    };
    // `neighborhoods` lives across the ICP iterations (:1214): its classification and description are only
    // overwritten when the corresponding branch runs, so stale values leak from one iteration to the next.
    std::vector<Neighborhood> neighborhoods(num_points);
    std::vector<int> classes(num_points, NONE);
    std::vector<char> computed(num_points, 0);
    std::vector<char> has_block;
    std::vector<ResidualBlock> all_blocks;
    int iter = 0;
    for (; iter < options.num_iters_icp; iter++) {
        transform_keypoints();
        has_block.assign(num_points, 0);
        all_blocks.assign(num_points, ResidualBlock());
        size_t kp_iters = 0, stencil_sum = 0;
#pragma omp parallel for num_threads(num_threads) reduction(+ : kp_iters, stencil_sum)
        for (long k = 0; k < (long) num_points; ++k) {   // :1229-1289
            const WPoint3D &pt = kpts[k];
            auto &neighborhood = neighborhoods[k];
            size_t st = 0;
            map.ComputeNeighborhoodInPlace(pt.world, options.max_number_neighbors, neighborhood, &st);
            kp_iters++;
            stencil_sum += st;
            if ((int) neighborhood.points.size() < kMinNumNeighbors) continue;
            neighborhood.ComputeNeighborhood();   // ALL_BUT_KDTREE; a no-op below 5 points (neighborhood.h:227)
            if (neighborhood.is_valid) computed[k] = 1;
            const auto &desc = neighborhood.description;
            // ClassifyNeighborhood, neighborhood.h:268-282
            if (computed[k]) {
                if (desc.planarity > options.threshold_planarity) classes[k] = PLANAR;
                else if (desc.linearity > options.threshold_linearity) classes[k] = LINEAR;
            } else
                classes[k] = NONE;
            if (!options.use_lines && classes[k] == LINEAR)   // :1243-1248
                classes[k] = options.threshold_planarity < desc.planarity ? PLANAR : VOLUMIC;
            double weight;
            if (classes[k] == LINEAR) weight = std::pow(std::abs(desc.linearity), options.power_planarity);
            else if (classes[k] == PLANAR) weight = std::pow(std::abs(desc.planarity), options.power_planarity);
            else weight = options.weight_neighborhood;

            const Vec3 point = options.use_barycenter ? desc.barycenter : neighborhood.points.front();
            double distance;
            int kind = CTICP_DIST_POINT_TO_DISTRIBUTION;
            const Vec3 d = point - pt.world;
            if (classes[k] == LINEAR) {
                Vec3 dir = desc.line;
                const double z = dir.squaredNorm();
                if (z > 0) dir = dir * (1.0 / std::sqrt(z));
                distance = std::abs(d.cross(dir).norm());
                kind = CTICP_DIST_POINT_TO_LINE;
            } else if (classes[k] == PLANAR) {
                distance = std::abs(d.dot(desc.normal));
                kind = CTICP_DIST_POINT_TO_PLANE;
            } else
                distance = d.norm();
            if (distance < options.outlier_distance) {
                ResidualBlock rb;
                rb.alpha = frame.begin_pose.GetAlphaTimestamp(pt.timestamp, frame.end_pose);
                if (rb.alpha < 0 || rb.alpha > 1) throw std::runtime_error("BAD ALPHA TIMESTAMP !");
                rb.reference = point;
                rb.raw = pt.raw;
                rb.weight = weight;
                rb.kind = kind;
                if (kind == CTICP_DIST_POINT_TO_PLANE) rb.normal = desc.normal;
                else if (kind == CTICP_DIST_POINT_TO_LINE) rb.normal = desc.line;
                else {   // FunctorPointToDistribution ctor, cost_functions.h:147-158
                    Mat3 m = desc.covariance;
                    for (int i = 0; i < 3; ++i) m(i, i) += 0.05;
                    rb.information = Inverse3(m);
                }
                all_blocks[k] = rb;
                has_block[k] = 1;
            }
        }
        icp_summary.keypoint_iterations += kp_iters;
        icp_summary.stencil_points += stencil_sum;
        if (getenv("ORC_DEBUG_LM")) {
            int cnt[4] = {0, 0, 0, 0};
            for (size_t k = 0; k < num_points; ++k)
                if (has_block[k]) cnt[all_blocks[k].kind]++;
            fprintf(stderr, "[orc-robust] iter %d blocks: plane %d line %d distribution %d\n", iter, cnt[CTICP_DIST_POINT_TO_PLANE],
                    cnt[CTICP_DIST_POINT_TO_LINE], cnt[CTICP_DIST_POINT_TO_DISTRIBUTION]);
        }

        ICPSummary failed;
        const int status = AssembleAndSolve(options, all_blocks, has_block, motion_model, frame, previous_begin_pose,
                                            previous_end_pose, number_of_residuals, num_threads, failed);
        if (status == 2) {
            failed.keypoint_iterations = icp_summary.keypoint_iterations;
            failed.stencil_points = icp_summary.stencil_points;
            return failed;
        }
        // point_to_plane_with_distortion → DistortFrame only acts on parametrization SIMPLE (:200-215)
        if (status == 1) break;
    }
    transform_keypoints();
    icp_summary.success = true;
    icp_summary.num_residuals_used = number_of_residuals;
    icp_summary.num_iters = iter;
    frame.begin_pose.pose.quat.normalize();
    frame.end_pose.pose.quat.normalize();
    return icp_summary;
}

}  // namespace orc
