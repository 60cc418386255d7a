// lm_functor.cuh — the continuous-time point-to-plane residual of solver CERES and its derivatives, on the device.
//
// Reference: CTFunctor<FunctorPointToPlane> (include/ct_icp/cost_functions.h:186-222, 32-67) differentiated by
// ceres::AutoDiffCostFunction and ceres::EigenQuaternionParameterization (src/ct_icp/ct_icp.cpp:221-232), robustified
// by ceres::{Cauchy,Huber,Tolerant}Loss / ct_icp::TruncatedLoss through ceres::internal::Corrector (:171-187).
//
// Forward-mode differentiation with ONE tangent direction per lane: lane j (< 12) evaluates the functor on dual
// numbers seeded with the j-th column of the parameterisation's Plus-Jacobian (tangent order = Ceres program order:
// begin quaternion (3), end quaternion (3), begin translation (3), end translation (3)), so the 12 partials of a
// residual come out of one SIMT pass with the arithmetic of the reference's templated functor.
#pragma once
#include "se3.cuh"

namespace cticp {

struct Dual {
    double a, d;
};
__device__ __forceinline__ Dual mkd(double a, double d = 0.0) { return Dual{a, d}; }
__device__ __forceinline__ Dual operator+(Dual x, Dual y) { return {x.a + y.a, x.d + y.d}; }
__device__ __forceinline__ Dual operator-(Dual x, Dual y) { return {x.a - y.a, x.d - y.d}; }
__device__ __forceinline__ Dual operator-(Dual x) { return {-x.a, -x.d}; }
__device__ __forceinline__ Dual operator*(Dual x, Dual y) { return {x.a * y.a, x.a * y.d + x.d * y.a}; }
__device__ __forceinline__ Dual operator*(double s, Dual y) { return {s * y.a, s * y.d}; }
__device__ __forceinline__ Dual operator/(Dual x, Dual y) {
    const double inv = 1.0 / y.a;
    const double q = x.a * inv;
    return {q, (x.d - q * y.d) * inv};
}
__device__ __forceinline__ Dual dsqrt(Dual x) {
    const double r = sqrt(x.a);
    return {r, x.d / (2.0 * r)};
}
// (the slerp angles of one sweep are far below 0.5 rad: polynomials of se3.cuh, libm beyond)
__device__ __forceinline__ Dual dsin(Dual x) {
    const bool small = fabs(x.a) <= 0.5;
    return {small ? sin_upto_half(x.a) : sin(x.a), (small ? cos_upto_half(x.a) : cos(x.a)) * x.d};
}
__device__ __forceinline__ Dual dacos(Dual x) { return {acos(x.a), -x.d / sqrt(1.0 - x.a * x.a)}; }

struct DQuat {
    Dual x, y, z, w;
};
__device__ __forceinline__ Dual dq_dot(const DQuat &a, const DQuat &b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
// Eigen MatrixBase::normalized(): n / sqrt(squaredNorm) when squaredNorm > 0
__device__ __forceinline__ DQuat dq_normalized(const DQuat &q) {
    const Dual z2 = dq_dot(q, q);
    if (z2.a > 0) {
        const Dual n = dsqrt(z2);
        return {q.x / n, q.y / n, q.z / n, q.w / n};
    }
    return q;
}
// Eigen QuaternionBase::slerp on duals (branches on the value part)
__device__ __forceinline__ DQuat dq_slerp(const DQuat &a, const DQuat &b, double t) {
    const double one = 1.0 - 2.220446049250313e-16;
    const Dual d = dq_dot(a, b);
    const Dual ad = d.a < 0.0 ? -d : d;
    Dual s0, s1;
    if (ad.a >= one) {
        s0 = mkd(1.0 - t);
        s1 = mkd(t);
    } else {
        const Dual theta = dacos(ad);
        const Dual st = dsin(theta);
        s0 = dsin((1.0 - t) * theta) / st;
        s1 = dsin(t * theta) / st;
    }
    if (d.a < 0) s1 = -s1;
    return {s0 * a.x + s1 * b.x, s0 * a.y + s1 * b.y, s0 * a.z + s1 * b.z, s0 * a.w + s1 * b.w};
}

// ceres::EigenQuaternionParameterization::ComputeJacobian: column j of the 4x3 Plus-Jacobian at q (rows x,y,z,w)
__device__ __forceinline__ void quat_plus_column(const double q[4], int j, double col[4]) {
    const double x = q[0], y = q[1], z = q[2], w = q[3];
    if (j == 0) { col[0] = w;  col[1] = -z; col[2] = y;  col[3] = -x; }
    else if (j == 1) { col[0] = z;  col[1] = w;  col[2] = -x; col[3] = -y; }
    else { col[0] = -y; col[1] = x;  col[2] = w;  col[3] = -z; }
}
// ceres::EigenQuaternionParameterization::Plus: q+ = [sin|d|/|d| d, cos|d|] ⊗ q
CT_HD Q4 quat_plus(Q4 q, double dx, double dy, double dz) {
    const double n = sqrt(dx * dx + dy * dy + dz * dz);
    if (n > 0.0) {
#ifdef __CUDA_ARCH__
        // an LM step's rotation is far below 0.5 rad: polynomials (se3.cuh) instead of libm's argument reduction
        const bool small = n <= 0.5;
        const double s = (small ? sin_upto_half(n) : sin(n)) / n;
        const double c = small ? cos_upto_half(n) : cos(n);
#else
        const double s = sin(n) / n, c = cos(n);
#endif
        return qmul(Q4{s * dx, s * dy, s * dz, c}, q);
    }
    return q;
}

enum { kResPlane = 0, kResLine = 1, kResDistribution = 2 };
constexpr int kResSimple = 16;   // flag in ResidualBlock::kind: POSE_PARAMETRIZATION SIMPLE — the inner functor on the END pose alone

struct ResidualBlock {   // CTFunctor<FunctorT> state for one keypoint
    double ref[3];       // world_reference_ (the neighbor / barycenter the residual is anchored on)
    double normal[3];    // reference_normal_ (plane) or direction_ (line, not normalised)
    double weight;
    double alpha;
    double raw[3];       // raw_point_ (sensor frame; with parametrization SIMPLE + distortion: moved into the end pose's frame)
    double info[6];      // FunctorPointToDistribution::neighborhood_information_ (xx xy xz yy yz zz); solver ROBUST only
    int valid;
    int kind;            // kResPlane / kResLine / kResDistribution (solver CERES: always plane) [| kResSimple]
};

// Residual and its derivative along tangent direction `dir` (0..11; >= 12 → value only).
// params: qb[4], qe[4], tb[3], te[3]. kRobust = false compiles the point-to-plane functor only (solver CERES).
template <bool kRobust>
__device__ __forceinline__ Dual ct_residual(const ResidualBlock &rb, const double *qb, const double *qe,
                                            const double *tb, const double *te, int dir) {
    double sb[4] = {0, 0, 0, 0}, se[4] = {0, 0, 0, 0}, stb[3] = {0, 0, 0}, ste[3] = {0, 0, 0};
    if (dir < 3) quat_plus_column(qb, dir, sb);
    else if (dir < 6) quat_plus_column(qe, dir - 3, se);
    else if (dir < 9) stb[dir - 6] = 1.0;
    else if (dir < 12) ste[dir - 9] = 1.0;
    const DQuat Qb{{qb[0], sb[0]}, {qb[1], sb[1]}, {qb[2], sb[2]}, {qb[3], sb[3]}};
    const DQuat Qe{{qe[0], se[0]}, {qe[1], se[1]}, {qe[2], se[2]}, {qe[3], se[3]}};
    const double alpha = rb.alpha, alpha_m = 1.0 - rb.alpha;
    DQuat qi;
    Dual tx, ty, tz;
    if (rb.kind & kResSimple) {
        // parametrization SIMPLE (ct_icp.cpp:314-321, 352-358): FunctorPointToPlane on the end pose's blocks; the tangent
        // directions of the begin pose see a constant (derivative 0)
        qi = Qe;
        tx = mkd(te[0], ste[0]); ty = mkd(te[1], ste[1]); tz = mkd(te[2], ste[2]);
    } else {
        qi = dq_slerp(dq_normalized(Qb), dq_normalized(Qe), alpha);   // cost_functions.h:208-209
        qi = dq_normalized(qi);                                        // :210
        tx = alpha_m * mkd(tb[0], stb[0]) + alpha * mkd(te[0], ste[0]);
        ty = alpha_m * mkd(tb[1], stb[1]) + alpha * mkd(te[1], ste[1]);
        tz = alpha_m * mkd(tb[2], stb[2]) + alpha * mkd(te[2], ste[2]);
    }
    // FunctorPointToPlane / FunctorPointToDistribution: quat.normalized() (cost_functions.h:47-51, 163-167);
    // FunctorPointToLine rotates with the quaternion as is (:121-125)
    const int kind = rb.kind & 15;
    const DQuat q = (kRobust && kind == kResLine) ? qi : dq_normalized(qi);
    const Dual vx = mkd(rb.raw[0]), vy = mkd(rb.raw[1]), vz = mkd(rb.raw[2]);
    // Eigen _transformVector: uv = 2 (q.vec x v); v + w uv + q.vec x uv
    Dual uvx = q.y * vz - q.z * vy, uvy = q.z * vx - q.x * vz, uvz = q.x * vy - q.y * vx;
    uvx = uvx + uvx; uvy = uvy + uvy; uvz = uvz + uvz;
    const Dual px = vx + q.w * uvx + (q.y * uvz - q.z * uvy) + tx;
    const Dual py = vy + q.w * uvy + (q.z * uvx - q.x * uvz) + ty;
    const Dual pz = vz + q.w * uvz + (q.x * uvy - q.y * uvx) + tz;
    if (kRobust && kind == kResLine) {   // cost_functions.h:127-129
        double ux = rb.normal[0], uy = rb.normal[1], uz = rb.normal[2];
        const double z = ux * ux + uy * uy + uz * uz;
        if (z > 0) { const double inv = 1.0 / sqrt(z); ux *= inv; uy *= inv; uz *= inv; }
        const Dual dx = px - mkd(rb.ref[0]), dy = py - mkd(rb.ref[1]), dz = pz - mkd(rb.ref[2]);
        const Dual cx = uy * dz - uz * dy, cy = uz * dx - ux * dz, cz = ux * dy - uy * dx;
        return rb.weight * dsqrt(cx * cx + cy * cy + cz * cz);
    }
    if (kRobust && kind == kResDistribution) {   // cost_functions.h:169-171 : w * diff^T M diff
        const Dual dx = px - mkd(rb.ref[0]), dy = py - mkd(rb.ref[1]), dz = pz - mkd(rb.ref[2]);
        const Dual r0 = rb.info[0] * dx + rb.info[1] * dy + rb.info[2] * dz;
        const Dual r1 = rb.info[1] * dx + rb.info[3] * dy + rb.info[4] * dz;
        const Dual r2 = rb.info[2] * dx + rb.info[4] * dy + rb.info[5] * dz;
        return rb.weight * (r0 * dx + r1 * dy + r2 * dz);
    }
    const Dual prod = rb.normal[0] * (mkd(rb.ref[0]) - px) + rb.normal[1] * (mkd(rb.ref[1]) - py) +
                      rb.normal[2] * (mkd(rb.ref[2]) - pz);
    return rb.weight * prod;
}

// ---- loss functions (ceres loss_function.cc; ct_icp::TruncatedLoss src/ct_icp/cost_function.cpp:5-15) ---------
struct LossParams {
    int type;          // CTICP_LOSS_*
    double a, b, c;
};
CT_HD LossParams make_loss(int type, double ls_sigma, double ls_tolerant_min_threshold) {
    LossParams L{type, 0, 0, 0};
    switch (type) {
        case 1: L.b = ls_sigma * ls_sigma; L.c = 1.0 / L.b; break;                       // CauchyLoss(sigma)
        case 2: L.a = ls_sigma; L.b = L.a * L.a; break;                                  // HuberLoss(sigma)
        case 3:                                                                          // TolerantLoss(a = min_threshold, b = sigma)
            L.a = ls_tolerant_min_threshold; L.b = ls_sigma; L.c = L.b * log(1.0 + exp(-L.a / L.b)); break;
        case 4: L.b = ls_sigma * ls_sigma; break;                                        // TruncatedLoss(sigma)
        default: break;
    }
    return L;
}
CT_HD void loss_evaluate(const LossParams &L, double s, double rho[3]) {
    const double kDblMin = 2.2250738585072014e-308;
    switch (L.type) {
        case 1: {
            const double sum = 1.0 + s * L.c, inv = 1.0 / sum;
            rho[0] = L.b * log(sum);
            rho[1] = fmax(kDblMin, inv);
            rho[2] = -L.c * (inv * inv);
            break;
        }
        case 2:
            if (s > L.b) {
                const double r = sqrt(s);
                rho[0] = 2.0 * L.a * r - L.b;
                rho[1] = fmax(kDblMin, L.a / r);
                rho[2] = -rho[1] / (2.0 * s);
            } else {
                rho[0] = s; rho[1] = 1.0; rho[2] = 0.0;
            }
            break;
        case 3: {
            const double x = (s - L.a) / L.b;
            if (x > 36.7) {
                rho[0] = s - L.a - L.c; rho[1] = 1.0; rho[2] = 0.0;
            } else {
                const double e_x = exp(x);
                rho[0] = L.b * log(1.0 + e_x) - L.c;
                rho[1] = fmax(kDblMin, e_x / (1.0 + e_x));
                rho[2] = 0.5 / (L.b * (1.0 + cosh(x)));
            }
            break;
        }
        case 4:
            if (s < L.b) { rho[0] = s; rho[1] = 1.0; rho[2] = 0.0; }
            else { rho[0] = L.b; rho[1] = 0.0; rho[2] = 0.0; }
            break;
        default:
            rho[0] = s; rho[1] = 1.0; rho[2] = 0.0;
    }
}
// ceres::internal::Corrector for a 1-dimensional residual: r_c = rs r, J_c = js J
CT_HD void corrector_1d(double sq_norm, const double rho[3], double &rs, double &js) {
    const double sqrt_rho1 = sqrt(rho[1]);
    if (sq_norm == 0.0 || rho[2] <= 0.0) {
        rs = sqrt_rho1;
        js = sqrt_rho1;
        return;
    }
    const double D = 1.0 + 2.0 * sq_norm * rho[2] / rho[1];
    const double alpha = 1.0 - sqrt(D);
    rs = sqrt_rho1 / (1.0 - alpha);
    js = sqrt_rho1 * (1.0 - alpha);
}

}  // namespace cticp
