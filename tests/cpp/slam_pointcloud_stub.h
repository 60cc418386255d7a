// slam_pointcloud_stub.h — the part of slam::PointCloud (include/SlamCore/pointcloud.h, data/view.h:99-120) that
// ct_icp::Odometry::RegisterFrame(const slam::PointCloud&, ...) touches (src/ct_icp/odometry.cpp:202,335-336,464-465):
// size(), HasTimestamps(), XYZConst<double>() and TimestampsProxy<double>() — proxy views = base pointer + stride + source
// scalar type, converting on access. Here: interleaved records {float x, y, z; float intensity; double t} like a
// PointCloud2 wrapped by ROSCloud2ToSlamPointCloudShallow (ros/roscore/src/pc2_conversion.cxx:86-96).
#pragma once
#include <array>
#include <cstddef>
#include <cstring>
#include <vector>

namespace slam {
template <typename DestT>
struct Vec3Proxy {   // ProxyView<Eigen::Matrix<DestT,3,1>> over float sources
    const char *base;
    size_t stride;
    std::array<DestT, 3> operator[](size_t i) const {
        float v[3];
        std::memcpy(v, base + i * stride, sizeof(v));
        return {{(DestT) v[0], (DestT) v[1], (DestT) v[2]}};
    }
};
template <typename DestT>
struct ScalarProxy {   // ProxyView<DestT> over a double source
    const char *base;
    size_t stride;
    DestT operator[](size_t i) const {
        double v;
        std::memcpy(&v, base + i * stride, sizeof(v));
        return (DestT) v;
    }
};
class PointCloud {
public:
    struct Record {
        float x, y, z, intensity;
        double t;
    };
    void push_back(double x, double y, double z, double t) { records_.push_back({(float) x, (float) y, (float) z, 0.f, t}); }
    size_t size() const { return records_.size(); }
    bool HasTimestamps() const { return true; }
    template <typename T> Vec3Proxy<T> XYZConst() const { return {reinterpret_cast<const char *>(records_.data()), sizeof(Record)}; }
    template <typename T> ScalarProxy<T> TimestampsProxy() const {
        return {reinterpret_cast<const char *>(records_.data()) + offsetof(Record, t), sizeof(Record)};
    }

private:
    std::vector<Record> records_;
};
}  // namespace slam
