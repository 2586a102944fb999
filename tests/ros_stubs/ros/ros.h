// Stand-in for the slice of ROS 1 that the reference's node (src/mapping.cpp) uses — TEST INFRASTRUCTURE.
// It lets the UNMODIFIED src/mapping.cpp be compiled against adapter/semantic_dsp_map.h and run without ROS: ros::spin()
// replays a file of recorded (depth Image, PoseStamped, MaskGroup) triples in ROS 1 wire format (semantic_dsp_map_b200/wire.py
// writes them) through the callback the node registered on its synchroniser, and every PointCloud2 the node publishes is
// appended to an output file.  Environment: DSPMAP_REPLAY_PACKAGE_PATH (what ros::package::getPath returns: the directory
// holding cfg/), DSPMAP_REPLAY_IN (record file), DSPMAP_REPLAY_OUT (published clouds: per cloud a u32 topic length, the
// topic, a u32 point count, then x, y, z as f32 and r, g, b as u8 per point).
#pragma once
#include <chrono>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <functional>
#include <iostream>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

// The node reports its own timing through ROS_INFO_THROTTLE (src/mapping.cpp:279-280: time cost of the last update() and the
// running average).  The stand-in keeps the LAST message per call site and prints them when the process ends.
#define ROS_INFO_THROTTLE(period, ...) do { static ros::replay_detail::LastMessage _m; _m.set(__VA_ARGS__); } while (0)
#define ROS_INFO(...) do { std::printf(__VA_ARGS__); std::printf("\n"); } while (0)
#define ROS_ERROR(...) do { std::fprintf(stderr, __VA_ARGS__); std::fprintf(stderr, "\n"); } while (0)

namespace ros { namespace replay_detail {
struct LastMessage {
    char text[256] = {0};
    void set(const char* fmt, ...) __attribute__((format(printf, 2, 3))) {
        va_list ap;
        va_start(ap, fmt);
        std::vsnprintf(text, sizeof text, fmt, ap);
        va_end(ap);
    }
    ~LastMessage() { if (text[0]) std::fprintf(stderr, "[last ROS_INFO_THROTTLE] %s\n", text); }
};
}}  // namespace ros::replay_detail

namespace boost {
template <typename T> using shared_ptr = std::shared_ptr<T>;
namespace replay_detail { struct Placeholder {}; }
template <typename C, typename A1, typename A2, typename A3>
std::function<void(A1, A2, A3)> bind(void (C::*fn)(A1, A2, A3), C* obj, replay_detail::Placeholder, replay_detail::Placeholder, replay_detail::Placeholder) {
    return [fn, obj](A1 a, A2 b, A3 c) { (obj->*fn)(a, b, c); };
}
}  // namespace boost
static const boost::replay_detail::Placeholder _1{}, _2{}, _3{};

namespace ros {

struct Time {
    uint32_t sec = 0, nsec = 0;
    double toSec() const { return sec + 1e-9 * nsec; }
    static Time now() {
        const auto ns = std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
        Time t;
        t.sec = static_cast<uint32_t>(ns / 1000000000LL);
        t.nsec = static_cast<uint32_t>(ns % 1000000000LL);
        return t;
    }
};

inline void init(int&, char**, const std::string&) {}
namespace package {
inline std::string getPath(const std::string&) {
    const char* p = std::getenv("DSPMAP_REPLAY_PACKAGE_PATH");
    return p ? p : ".";
}
}  // namespace package

namespace replay_detail {
inline std::ofstream& out_file() {
    static std::ofstream f(std::getenv("DSPMAP_REPLAY_OUT") ? std::getenv("DSPMAP_REPLAY_OUT") : "replay_out.bin", std::ios::binary);
    return f;
}
template <typename M> struct Sink { static void publish(const std::string&, const M&) {} };
inline std::function<void()>& spin_body() { static std::function<void()> f; return f; }
}  // namespace replay_detail

class Publisher {
public:
    Publisher() = default;
    explicit Publisher(std::string topic) : topic_(std::move(topic)) {}
    template <typename M> void publish(const M& m) const { replay_detail::Sink<M>::publish(topic_, m); }
private:
    std::string topic_;
};

class NodeHandle {
public:
    template <typename M> Publisher advertise(const std::string& topic, uint32_t) { return Publisher(topic); }
};

inline void spin() { if (replay_detail::spin_body()) replay_detail::spin_body()(); }

}  // namespace ros

namespace std_msgs {
struct Header { uint32_t seq = 0; ros::Time stamp; std::string frame_id; };
}
