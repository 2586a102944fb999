#pragma once
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include <sensor_msgs/PointCloud2.h>
namespace pcl {
inline void toROSMsg(const PointCloud<PointXYZRGB>& cloud, sensor_msgs::PointCloud2& msg) {
    msg.points.clear();
    for (const auto& p : cloud.points) msg.points.push_back(sensor_msgs::PointCloud2::Point{p.x, p.y, p.z, p.r, p.g, p.b});
}
}  // namespace pcl
