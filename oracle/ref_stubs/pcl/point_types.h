// Stand-in for PCL point types — TEST INFRASTRUCTURE (oracle/_ref build only; see ref_stubs/Eigen/Dense).
// Only containers are real; PCL algorithms (template matching, outside the hot path and disabled by default
// in the reference, external_settings.h:17) are declared so the headers compile and abort if executed.
#pragma once
#include <Eigen/Dense>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <memory>
#include <queue>
#include <string>
#include <vector>

namespace pcl {

[[noreturn]] inline void stub_unreachable(const char* what) {
    std::fprintf(stderr, "ref_stubs/pcl: %s is outside the hot path and not implemented\n", what);
    std::abort();
}

struct PointXYZ {
    float x = 0.f, y = 0.f, z = 0.f;
    PointXYZ() = default;
    PointXYZ(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
};

struct PointXYZRGB {
    float x = 0.f, y = 0.f, z = 0.f;
    uint8_t b = 0, g = 0, r = 0, a = 255;
};

struct PointIndices {
    std::vector<int> indices;
};

template <typename PointT>
class PointCloud {
public:
    typedef std::shared_ptr<PointCloud<PointT>> Ptr;
    typedef std::shared_ptr<const PointCloud<PointT>> ConstPtr;
    std::vector<PointT> points;
    uint32_t width = 0, height = 0;
    bool is_dense = true;
    size_t size() const { return points.size(); }
    bool empty() const { return points.empty(); }
    void clear() { points.clear(); width = height = 0; }
    void push_back(const PointT& p) { points.push_back(p); }
    void reserve(size_t n) { points.reserve(n); }
    Ptr makeShared() const { return Ptr(new PointCloud<PointT>(*this)); }
    PointT& operator[](size_t i) { return points[i]; }
    const PointT& operator[](size_t i) const { return points[i]; }
    typename std::vector<PointT>::iterator begin() { return points.begin(); }
    typename std::vector<PointT>::iterator end() { return points.end(); }
    typename std::vector<PointT>::const_iterator begin() const { return points.begin(); }
    typename std::vector<PointT>::const_iterator end() const { return points.end(); }
};

template <typename PointT>
void copyPointCloud(const PointCloud<PointT>& in, PointCloud<PointT>& out) { out = in; }

template <typename PointT, typename Transform>
void transformPointCloud(const PointCloud<PointT>&, PointCloud<PointT>&, const Transform&) { stub_unreachable("transformPointCloud"); }

template <typename PointT>
class KdTreeFLANN {
public:
    void setInputCloud(const typename PointCloud<PointT>::Ptr&) { stub_unreachable("KdTreeFLANN"); }
    int nearestKSearch(const PointT&, int, std::vector<int>&, std::vector<float>&) const { stub_unreachable("KdTreeFLANN"); }
    int radiusSearch(const PointT&, double, std::vector<int>&, std::vector<float>&) const { stub_unreachable("KdTreeFLANN"); }
};

template <typename PointT>
class VoxelGrid {
public:
    void setInputCloud(const typename PointCloud<PointT>::Ptr&) { stub_unreachable("VoxelGrid"); }
    void setLeafSize(float, float, float) {}
    void filter(PointCloud<PointT>&) { stub_unreachable("VoxelGrid"); }
};

template <typename PointT>
class StatisticalOutlierRemoval {
public:
    void setInputCloud(const typename PointCloud<PointT>::Ptr&) { stub_unreachable("StatisticalOutlierRemoval"); }
    void setMeanK(int) {}
    void setStddevMulThresh(double) {}
    void filter(PointCloud<PointT>&) { stub_unreachable("StatisticalOutlierRemoval"); }
};

template <typename PointT>
class EuclideanClusterExtraction {
public:
    void setClusterTolerance(double) {}
    void setMinClusterSize(int) {}
    void setMaxClusterSize(int) {}
    void setInputCloud(const typename PointCloud<PointT>::Ptr&) { stub_unreachable("EuclideanClusterExtraction"); }
    void extract(std::vector<PointIndices>&) { stub_unreachable("EuclideanClusterExtraction"); }
};

namespace io {
template <typename PointT>
int loadPCDFile(const std::string&, PointCloud<PointT>&) { return -1; }
}  // namespace io

}  // namespace pcl
