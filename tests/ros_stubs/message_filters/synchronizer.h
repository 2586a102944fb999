// The synchroniser stand-in: registerCallback arms ros::spin() with the replay of DSPMAP_REPLAY_IN — uint32-length-prefixed
// serialised messages in triples (sensor_msgs/Image, geometry_msgs/PoseStamped, mask_kpts_msgs/MaskGroup), ROS 1 wire format.
#pragma once
#include <geometry_msgs/PoseStamped.h>
#include <mask_kpts_msgs/MaskGroup.h>
#include <message_filters/subscriber.h>
#include <sensor_msgs/Image.h>
namespace message_filters {
namespace replay_detail {
class Reader {
public:
    Reader(const std::vector<uint8_t>& b) : b_(b) {}
    template <typename T> T take() { if (pos_ + sizeof(T) > b_.size()) throw std::runtime_error("replay: buffer underfill"); T v; std::memcpy(&v, b_.data() + pos_, sizeof(T)); pos_ += sizeof(T); return v; }
    std::string string() { const uint32_t n = take<uint32_t>(); if (pos_ + n > b_.size()) throw std::runtime_error("replay: buffer underfill"); std::string s(reinterpret_cast<const char*>(b_.data() + pos_), n); pos_ += n; return s; }
    std_msgs::Header header() { std_msgs::Header h; h.seq = take<uint32_t>(); h.stamp.sec = take<uint32_t>(); h.stamp.nsec = take<uint32_t>(); h.frame_id = string(); return h; }
    sensor_msgs::Image image() {
        sensor_msgs::Image m;
        m.header = header(); m.height = take<uint32_t>(); m.width = take<uint32_t>(); m.encoding = string(); m.is_bigendian = take<uint8_t>(); m.step = take<uint32_t>();
        const uint32_t n = take<uint32_t>();
        if (pos_ + n > b_.size()) throw std::runtime_error("replay: buffer underfill");
        m.data.assign(b_.begin() + pos_, b_.begin() + pos_ + n); pos_ += n;
        return m;
    }
    mask_kpts_msgs::Keypoint keypoint() { mask_kpts_msgs::Keypoint k; k.x = take<float>(); k.y = take<float>(); k.z = take<float>(); return k; }
private:
    const std::vector<uint8_t>& b_;
    size_t pos_ = 0;
};
inline bool read_blob(std::ifstream& f, std::vector<uint8_t>& out) {
    uint32_t n = 0;
    if (!f.read(reinterpret_cast<char*>(&n), 4)) return false;
    out.resize(n);
    return static_cast<bool>(f.read(reinterpret_cast<char*>(out.data()), n));
}
}  // namespace replay_detail

template <typename Policy> class Synchronizer {
public:
    template <typename S0, typename S1, typename S2> Synchronizer(const Policy&, S0&, S1&, S2&) {}
    void registerCallback(const std::function<void(const sensor_msgs::ImageConstPtr&, const geometry_msgs::PoseStampedConstPtr&, const mask_kpts_msgs::MaskGroup::ConstPtr&)>& cb) {
        ros::replay_detail::spin_body() = [cb]() {
            const char* path = std::getenv("DSPMAP_REPLAY_IN");
            std::ifstream f(path ? path : "replay_in.bin", std::ios::binary);
            std::vector<uint8_t> a, b, c;
            while (replay_detail::read_blob(f, a) && replay_detail::read_blob(f, b) && replay_detail::read_blob(f, c)) {
                replay_detail::Reader ra(a), rb(b), rc(c);
                auto depth = std::make_shared<sensor_msgs::Image>(ra.image());
                auto pose = std::make_shared<geometry_msgs::PoseStamped>();
                pose->header = rb.header();
                pose->pose.position.x = rb.take<double>(); pose->pose.position.y = rb.take<double>(); pose->pose.position.z = rb.take<double>();
                pose->pose.orientation.x = rb.take<double>(); pose->pose.orientation.y = rb.take<double>(); pose->pose.orientation.z = rb.take<double>(); pose->pose.orientation.w = rb.take<double>();
                auto group = std::make_shared<mask_kpts_msgs::MaskGroup>();
                group->header = rc.header();
                const uint32_t n = rc.take<uint32_t>();
                for (uint32_t i = 0; i < n; ++i) {
                    mask_kpts_msgs::MaskKpts o;
                    o.label = rc.string(); o.track_id = rc.take<uint16_t>(); o.mask = rc.image(); o.bbox_tl = rc.keypoint(); o.bbox_br = rc.keypoint();
                    for (auto* list : {&o.kpts_curr, &o.kpts_last}) { const uint32_t m = rc.take<uint32_t>(); for (uint32_t k = 0; k < m; ++k) list->push_back(rc.keypoint()); }
                    group->objects.push_back(o);
                }
                cb(depth, pose, group);
            }
        };
    }
};
}  // namespace message_filters
