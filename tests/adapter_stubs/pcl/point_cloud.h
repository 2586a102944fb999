// Minimal stand-in for pcl::PointCloud (TEST ONLY).
#pragma once
#include <memory>
#include <vector>
namespace pcl {
template <class P>
struct PointCloud {
    using Ptr = std::shared_ptr<PointCloud<P>>;
    std::vector<P> points;
    void push_back(const P& p) { points.push_back(p); }
    size_t size() const { return points.size(); }
};
}  // namespace pcl
