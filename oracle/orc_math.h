// ORACLE — TEST INFRASTRUCTURE ONLY.  Not linked into, imported by, or shipped with the product
// (ct_icp_b200/).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference legs use it.
//
// PARITY UNPINNED: the reference (jedeschaud/ct_icp @ d467813) cannot be compiled in this container
// (Eigen, Ceres, glog, yaml-cpp, tsl::robin_map absent) and holds no golden vectors for this path.
// This file is a dependency-free fp64 restatement; the Eigen/Ceres routines it depends on are restated
// from their published algorithms (one small function each, cited below).
//
// orc_math.h — SE3 algebra and small dense linear algebra.
//   reference: include/SlamCore/types.h:100-139, 313-470 (TSE3 / TPose), Eigen 3.x Quaternion/JacobiSVD/LDLT.
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <limits>
#include <stdexcept>

namespace orc {

struct Vec3 {
    double x = 0, y = 0, z = 0;
    Vec3() = default;
    Vec3(double x_, double y_, double z_) : x(x_), y(y_), z(z_) {}
    double &operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
    double operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
    Vec3 operator+(const Vec3 &o) const { return {x + o.x, y + o.y, z + o.z}; }
    Vec3 operator-(const Vec3 &o) const { return {x - o.x, y - o.y, z - o.z}; }
    Vec3 operator-() const { return {-x, -y, -z}; }
    Vec3 operator*(double s) const { return {x * s, y * s, z * s}; }
    Vec3 &operator+=(const Vec3 &o) { x += o.x; y += o.y; z += o.z; return *this; }
    double dot(const Vec3 &o) const { return x * o.x + y * o.y + z * o.z; }
    Vec3 cross(const Vec3 &o) const { return {y * o.z - z * o.y, z * o.x - x * o.z, x * o.y - y * o.x}; }
    double squaredNorm() const { return x * x + y * y + z * z; }
    double norm() const { return std::sqrt(squaredNorm()); }
};
inline Vec3 operator*(double s, const Vec3 &v) { return v * s; }

struct Mat3 {
    double m[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    double &operator()(int i, int j) { return m[i][j]; }
    double operator()(int i, int j) const { return m[i][j]; }
    Mat3 operator*(const Mat3 &o) const {
        Mat3 r;
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                double s = 0;
                for (int k = 0; k < 3; ++k) s += m[i][k] * o.m[k][j];
                r.m[i][j] = s;
            }
        return r;
    }
    Mat3 transpose() const {
        Mat3 r;
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) r.m[i][j] = m[j][i];
        return r;
    }
    double trace() const { return m[0][0] + m[1][1] + m[2][2]; }
};

// Eigen::Quaterniond, coefficients stored (x, y, z, w).
struct Quat {
    double x = 0, y = 0, z = 0, w = 1;
    Quat() = default;
    Quat(double x_, double y_, double z_, double w_) : x(x_), y(y_), z(z_), w(w_) {}
    Vec3 vec() const { return {x, y, z}; }
    double dot(const Quat &o) const { return x * o.x + y * o.y + z * o.z + w * o.w; }
    double squaredNorm() const { return dot(*this); }
    double norm() const { return std::sqrt(squaredNorm()); }
    void normalize() {
        double n = norm();
        x /= n; y /= n; z /= n; w /= n;
    }
    Quat normalized() const {
        Quat q = *this;
        q.normalize();
        return q;
    }
    // Eigen QuaternionBase::inverse(): conjugate / squaredNorm (zero quaternion → zero)
    Quat inverse() const {
        double n2 = squaredNorm();
        if (n2 > 0) return {-x / n2, -y / n2, -z / n2, w / n2};
        return {0, 0, 0, 0};
    }
    // Eigen quat product
    Quat operator*(const Quat &b) const {
        const Quat &a = *this;
        return {a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
                a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z,
                a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x,
                a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z};
    }
    // Eigen QuaternionBase::_transformVector: uv = 2 (q.vec × v); v + w uv + q.vec × uv
    Vec3 operator*(const Vec3 &v) const {
        Vec3 uv = vec().cross(v);
        uv += uv;
        return v + w * uv + vec().cross(uv);
    }
    // Eigen QuaternionBase::toRotationMatrix
    Mat3 toRotationMatrix() const {
        Mat3 res;
        const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
        const double twx = tx * w, twy = ty * w, twz = tz * w;
        const double txx = tx * x, txy = ty * x, txz = tz * x;
        const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
        res(0, 0) = 1 - (tyy + tzz);
        res(0, 1) = txy - twz;
        res(0, 2) = txz + twy;
        res(1, 0) = txy + twz;
        res(1, 1) = 1 - (txx + tzz);
        res(1, 2) = tyz - twx;
        res(2, 0) = txz - twy;
        res(2, 1) = tyz + twx;
        res(2, 2) = 1 - (txx + tyy);
        return res;
    }
    // Eigen quaternionbase_assign_impl<Matrix3> (trace based), used at src/ct_icp/ct_icp.cpp:950-954
    static Quat fromRotationMatrix(const Mat3 &mat) {
        Quat q;
        double t = mat.trace();
        if (t > 0) {
            t = std::sqrt(t + 1.0);
            q.w = 0.5 * t;
            t = 0.5 / t;
            q.x = (mat(2, 1) - mat(1, 2)) * t;
            q.y = (mat(0, 2) - mat(2, 0)) * t;
            q.z = (mat(1, 0) - mat(0, 1)) * t;
        } else {
            int i = 0;
            if (mat(1, 1) > mat(0, 0)) i = 1;
            if (mat(2, 2) > mat(i, i)) i = 2;
            int j = (i + 1) % 3;
            int k = (j + 1) % 3;
            t = std::sqrt(mat(i, i) - mat(j, j) - mat(k, k) + 1.0);
            double c[3];
            c[i] = 0.5 * t;
            t = 0.5 / t;
            q.w = (mat(k, j) - mat(j, k)) * t;
            c[j] = (mat(j, i) + mat(i, j)) * t;
            c[k] = (mat(k, i) + mat(i, k)) * t;
            q.x = c[0]; q.y = c[1]; q.z = c[2];
        }
        return q;
    }
    // Eigen QuaternionBase::slerp(t, other) — NOT renormalised (callers normalise where the reference does)
    Quat slerp(double t, const Quat &other) const {
        const double one = 1.0 - std::numeric_limits<double>::epsilon();
        double d = dot(other);
        double absD = std::abs(d);
        double scale0, scale1;
        if (absD >= one) {
            scale0 = 1.0 - t;
            scale1 = t;
        } else {
            double theta = std::acos(absD);
            double sinTheta = std::sin(theta);
            scale0 = std::sin((1.0 - t) * theta) / sinTheta;
            scale1 = std::sin(t * theta) / sinTheta;
        }
        if (d < 0) scale1 = -scale1;
        return {scale0 * x + scale1 * other.x, scale0 * y + scale1 * other.y, scale0 * z + scale1 * other.z,
                scale0 * w + scale1 * other.w};
    }
};

// slam::TSE3<double>, include/SlamCore/types.h:100-139, 313-366
struct SE3 {
    Quat quat;
    Vec3 tr;
    SE3() = default;
    SE3(const Quat &q, const Vec3 &t) : quat(q.normalized()), tr(t) {}
    SE3 Inverse() const {                    // types.h:327-332
        SE3 r;
        r.quat = quat.inverse();
        r.tr = -(r.quat * tr);
        return r;
    }
    SE3 operator*(const SE3 &rhs) const {    // types.h:344-351
        SE3 r;
        r.quat = quat * rhs.quat;
        r.quat.normalize();
        r.tr = quat.normalized() * rhs.tr + tr;
        return r;
    }
    Vec3 operator*(const Vec3 &p) const {    // types.h:354-357
        return quat.normalized() * p + tr;
    }
    SE3 Interpolate(const SE3 &rhs, double weight) const {   // types.h:361-366
        SE3 r;
        r.quat = quat.slerp(weight, rhs.quat);
        r.tr = (1.0 - weight) * tr + weight * rhs.tr;
        return r;
    }
    Mat3 Rotation() const { return quat.toRotationMatrix(); }   // types.h:315-317
};

struct TimestampError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

// slam::AngularDistance, include/SlamCore/types.h:141-156 (degrees)
inline double AngularDistance(const Mat3 &a, const Mat3 &b) {
    double norm = ((a * b.transpose()).trace() - 1.0) / 2.0;
    if (!(norm < 1.0 + 1e-8 && norm >= -1.0 - 1e-8)) throw std::runtime_error("Not a rotation matrix !");
    norm = std::fmax(std::fmin(norm, 1.0), -1.0);
    return std::acos(norm) * (180.0 / M_PI);
}
inline double AngularDistance(const SE3 &a, const SE3 &b) { return AngularDistance(a.Rotation(), b.Rotation()); }

// slam::TPose<double>, include/SlamCore/types.h:162-274, 434-470
struct Pose {
    SE3 pose;
    double ref_timestamp = 0;
    double dest_timestamp = -1;
    uint32_t ref_frame_id = 0;
    uint32_t dest_frame_id = uint32_t(-1);
    Pose() = default;
    Pose(const SE3 &p, double ts, uint32_t dest_fid) : pose(p), dest_timestamp(ts), dest_frame_id(dest_fid) {}

    double GetAlphaTimestamp(double mid, const Pose &other) const {   // types.h:192-219 (incl. the t>max → 0 quirk)
        double mn = std::min(dest_timestamp, other.dest_timestamp);
        double mx = std::max(dest_timestamp, other.dest_timestamp);
        if (mn > mid) return 0.0;
        if (mx < mid) return 0.0;
        if (mn == mx) return 1.0;
        return (mid - mn) / (mx - mn);
    }
    Pose InterpolatePoseAlpha(const Pose &other, double alpha) const {   // types.h:434-452
        Pose p;
        p.ref_frame_id = ref_frame_id;
        p.dest_frame_id = dest_frame_id;
        p.ref_timestamp = ref_timestamp;
        p.dest_timestamp = (1.0 - alpha) * dest_timestamp + alpha * other.dest_timestamp;
        p.pose = pose.Interpolate(other.pose, alpha);
        return p;
    }
    Pose InterpolatePose(const Pose &other, double timestamp) const {    // types.h:455-470 (CHECK → exception)
        if (!(dest_timestamp <= timestamp && timestamp <= other.dest_timestamp))
            throw TimestampError("The timestamp cannot be interpolated between the two poses");
        Pose p;
        p.ref_frame_id = ref_frame_id;
        p.dest_frame_id = dest_frame_id;
        p.ref_timestamp = ref_timestamp;
        p.dest_timestamp = timestamp;
        p.pose = pose.Interpolate(other.pose, GetAlphaTimestamp(timestamp, other));
        return p;
    }
    Vec3 ContinuousTransform(const Vec3 &p, const Pose &other, double timestamp) const {   // types.h:414-418
        return InterpolatePoseAlpha(other, GetAlphaTimestamp(timestamp, other)) * p;
    }
    Vec3 operator*(const Vec3 &p) const { return pose * p; }
    Pose Inverse() const {                                                // types.h:421-430
        Pose p;
        p.ref_frame_id = dest_frame_id;
        p.ref_timestamp = dest_timestamp;
        p.dest_frame_id = ref_frame_id;
        p.dest_timestamp = ref_timestamp;
        p.pose = pose.Inverse();
        return p;
    }
    Mat3 Rotation() const { return pose.quat.normalized().toRotationMatrix(); }
};

// ct_icp::TrajectoryFrame, include/ct_icp/types.h:31-61
struct TrajectoryFrame {
    Pose begin_pose, end_pose;
    double EgoAngularDistance() const { return AngularDistance(begin_pose.pose, end_pose.pose); }
    double TranslationDistance(const TrajectoryFrame &o) const {
        return (begin_pose.pose.tr - o.begin_pose.pose.tr).norm() + (end_pose.pose.tr - o.end_pose.pose.tr).norm();
    }
    double RotationDistance(const TrajectoryFrame &o) const {
        return AngularDistance(begin_pose.pose, o.begin_pose.pose) + AngularDistance(end_pose.pose, o.end_pose.pose);
    }
    const Vec3 &BeginTr() const { return begin_pose.pose.tr; }
    const Vec3 &EndTr() const { return end_pose.pose.tr; }
    const Quat &BeginQuat() const { return begin_pose.pose.quat; }
    const Quat &EndQuat() const { return end_pose.pose.quat; }
};

// Stand-in for Eigen::JacobiSVD<Matrix3d>(C, ComputeFullV) on a SYMMETRIC matrix
// (include/SlamCore/experimental/neighborhood.h:293): singular values = |eigenvalues| sorted descending,
// V columns = eigenvectors (sign arbitrary). Cyclic Jacobi eigenvalue iteration in fp64.
inline void SymmetricSVD3(const Mat3 &C, double sv[3], Mat3 &V) {
    double a[3][3];
    double v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) a[i][j] = 0.5 * (C(i, j) + C(j, i));
    for (int sweep = 0; sweep < 64; ++sweep) {
        double off = a[0][1] * a[0][1] + a[0][2] * a[0][2] + a[1][2] * a[1][2];
        double diag = a[0][0] * a[0][0] + a[1][1] * a[1][1] + a[2][2] * a[2][2];
        if (off <= 1e-32 * diag || off == 0.0) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                if (a[p][q] == 0.0) continue;
                double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
                double t = (theta >= 0 ? 1.0 : -1.0) / (std::abs(theta) + std::sqrt(theta * theta + 1.0));
                double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < 3; ++k) {   // A <- A J
                    double akp = a[k][p], akq = a[k][q];
                    a[k][p] = c * akp - s * akq;
                    a[k][q] = s * akp + c * akq;
                }
                for (int k = 0; k < 3; ++k) {   // A <- J^T A
                    double apk = a[p][k], aqk = a[q][k];
                    a[p][k] = c * apk - s * aqk;
                    a[q][k] = s * apk + c * aqk;
                }
                for (int k = 0; k < 3; ++k) {
                    double vkp = v[k][p], vkq = v[k][q];
                    v[k][p] = c * vkp - s * vkq;
                    v[k][q] = s * vkp + c * vkq;
                }
            }
    }
    int idx[3] = {0, 1, 2};
    double ev[3] = {std::abs(a[0][0]), std::abs(a[1][1]), std::abs(a[2][2])};
    std::sort(idx, idx + 3, [&](int i, int j) { return ev[i] > ev[j]; });
    for (int c = 0; c < 3; ++c) {
        sv[c] = ev[idx[c]];
        for (int r = 0; r < 3; ++r) V(r, c) = v[r][idx[c]];
    }
}

// Stand-in for Eigen::Matrix<double,12,12>::ldlt().solve(b) (src/ct_icp/ct_icp.cpp:914):
// LDL^T with symmetric diagonal pivoting (largest remaining |diagonal|), then the two triangular solves.
template <int N>
inline std::array<double, N> LDLTSolve(const double (&Ain)[N][N], const std::array<double, N> &b) {
    double A[N][N];
    int perm[N];
    for (int i = 0; i < N; ++i) {
        perm[i] = i;
        for (int j = 0; j < N; ++j) A[i][j] = Ain[i][j];
    }
    double D[N];
    for (int k = 0; k < N; ++k) {
        int p = k;
        double best = std::abs(A[k][k]);
        for (int i = k + 1; i < N; ++i)
            if (std::abs(A[i][i]) > best) { best = std::abs(A[i][i]); p = i; }
        if (p != k) {
            for (int j = 0; j < N; ++j) std::swap(A[k][j], A[p][j]);
            for (int i = 0; i < N; ++i) std::swap(A[i][k], A[i][p]);
            std::swap(perm[k], perm[p]);
        }
        D[k] = A[k][k];
        if (D[k] == 0.0) {
            for (int i = k + 1; i < N; ++i) A[i][k] = 0.0;
            continue;
        }
        for (int i = k + 1; i < N; ++i) A[i][k] /= D[k];
        for (int i = k + 1; i < N; ++i)
            for (int j = k + 1; j <= i; ++j) {
                A[i][j] -= A[i][k] * D[k] * A[j][k];
                A[j][i] = A[i][j];
            }
    }
    std::array<double, N> y{};
    for (int i = 0; i < N; ++i) y[i] = b[perm[i]];
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < i; ++j) y[i] -= A[i][j] * y[j];
    double dmax = 0;
    for (int i = 0; i < N; ++i) dmax = std::max(dmax, std::abs(D[i]));
    const double tol = std::numeric_limits<double>::min();   // Eigen: pseudo-inverse of D with tiny tolerance
    for (int i = 0; i < N; ++i) y[i] = (std::abs(D[i]) > tol) ? y[i] / D[i] : 0.0;
    for (int i = N - 1; i >= 0; --i)
        for (int j = i + 1; j < N; ++j) y[i] -= A[j][i] * y[j];
    std::array<double, N> x{};
    for (int i = 0; i < N; ++i) x[perm[i]] = y[i];
    return x;
}

}  // namespace orc
