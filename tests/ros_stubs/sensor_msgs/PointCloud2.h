#pragma once
#include <ros/ros.h>
namespace sensor_msgs {
struct PointCloud2 {
    std_msgs::Header header;
    struct Point { float x, y, z; uint8_t r, g, b; };
    std::vector<Point> points;          // the stand-in keeps the points unpacked
};
}  // namespace sensor_msgs
namespace ros { namespace replay_detail {
template <> struct Sink<sensor_msgs::PointCloud2> {
    static void publish(const std::string& topic, const sensor_msgs::PointCloud2& m) {
        std::ofstream& f = out_file();
        const uint32_t tl = static_cast<uint32_t>(topic.size()), n = static_cast<uint32_t>(m.points.size());
        f.write(reinterpret_cast<const char*>(&tl), 4); f.write(topic.data(), tl); f.write(reinterpret_cast<const char*>(&n), 4);
        for (const auto& p : m.points) { f.write(reinterpret_cast<const char*>(&p.x), 12); f.write(reinterpret_cast<const char*>(&p.r), 3); }
        f.flush();
    }
};
}}  // namespace ros::replay_detail
