// mask_group_wire.h — ROS-free decoding of the node's input messages into the adapter's types.
//
// src/mapping.cpp receives a synchronised triple — sensor_msgs/Image (depth, 32FC1), geometry_msgs/PoseStamped and
// mask_kpts_msgs/MaskGroup (script/mmseg/mask_kpts_msgs/msg/_MaskGroup.py:17-18, _MaskKpts.py:17-24, _Keypoint.py:14-16) —
// and converts it in syncCallback (src/mapping.cpp:165-260).  This header decodes the ROS 1 serialisation of those
// messages (little-endian scalars, uint32 length-prefixed strings and arrays) and applies the same conversion rules, so a
// recorded stream can be replayed through SemanticDSPMap::update without ROS, cv_bridge or generated message headers.
// The Python twin is semantic_dsp_map_b200/wire.py; both are checked against bytes written by the reference's own
// generated serialiser (tests/golden/wire_maskgroup.bin).
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "semantic_dsp_map.h"

namespace dspwire {

struct Header { uint32_t seq = 0, secs = 0, nsecs = 0; std::string frame_id; };

class Reader {
public:
    Reader(const uint8_t* p, size_t n) : p_(p), n_(n) {}
    template <typename T>
    T take() {
        if (pos_ + sizeof(T) > n_) throw std::runtime_error("dspwire: buffer underfill");
        T v;
        std::memcpy(&v, p_ + pos_, sizeof(T));
        pos_ += sizeof(T);
        return v;
    }
    std::string string() {
        const uint32_t len = take<uint32_t>();
        if (pos_ + len > n_) throw std::runtime_error("dspwire: buffer underfill");
        std::string s(reinterpret_cast<const char*>(p_ + pos_), len);
        pos_ += len;
        return s;
    }
    Header header() {
        Header h;
        h.seq = take<uint32_t>(); h.secs = take<uint32_t>(); h.nsecs = take<uint32_t>(); h.frame_id = string();
        return h;
    }
    // sensor_msgs/Image -> cv::Mat of the given type; the node reads depth as 32FC1 (:180) and masks as MONO8 (:212)
    cv::Mat image(int cv_type, size_t elem, const char* enc_a, const char* enc_b, Header* hdr = nullptr) {
        const Header h = header();
        if (hdr) *hdr = h;
        const uint32_t height = take<uint32_t>(), width = take<uint32_t>();
        const std::string enc = string();
        const uint8_t big = take<uint8_t>();
        const uint32_t step = take<uint32_t>(), len = take<uint32_t>();
        if (enc != enc_a && enc != enc_b) throw std::runtime_error("dspwire: image encoding " + enc + " is not what the node reads here");
        if (big) throw std::runtime_error("dspwire: big-endian image data");
        if (step < width * elem || (size_t)step * height > len || pos_ + len > n_) throw std::runtime_error("dspwire: image data shorter than step * height");
        cv::Mat m(static_cast<int>(height), static_cast<int>(width), cv_type);
        for (uint32_t i = 0; i < height; ++i) std::memcpy(m.data + (size_t)i * m.step, p_ + pos_ + (size_t)i * step, width * elem);
        pos_ += len;
        return m;
    }
    bool done() const { return pos_ == n_; }

private:
    const uint8_t* p_;
    size_t n_, pos_ = 0;
};

struct Frame {
    cv::Mat depth;
    std::vector<MaskKpts> tracking_result;
    Eigen::Vector3d position;
    Eigen::Quaterniond orientation;
    double time_stamp = 0.0;
    uint32_t seq = 0;
};

// src/mapping.cpp:165-260.  Returns false where the node returns without calling update().
inline bool toFrame(const std::vector<uint8_t>& depth_msg, const std::vector<uint8_t>& pose_msg, const std::vector<uint8_t>& mask_group_msg, Frame& out) {
    Reader rp(pose_msg.data(), pose_msg.size());
    const Header ph = rp.header();
    out.time_stamp = ph.secs + 1e-9 * ph.nsecs;                                   // :169
    const double px = rp.take<double>(), py = rp.take<double>(), pz = rp.take<double>();
    const double qx = rp.take<double>(), qy = rp.take<double>(), qz = rp.take<double>(), qw = rp.take<double>();
    Reader rd(depth_msg.data(), depth_msg.size());
    Header dh;
    cv::Mat depth = rd.image(CV_32FC1, 4, "32FC1", "32FC1", &dh);
    if (dh.seq < 2) return false;                                                 // :172
    out.depth = depth;                                                            // :180
    out.seq = dh.seq;
    out.position = Eigen::Vector3d(px, py, pz);                                   // :185
    out.orientation = Eigen::Quaterniond(qw, qx, qy, qz);                         // :186
    Reader rg(mask_group_msg.data(), mask_group_msg.size());
    const Header gh = rg.header();
    if (gh.seq != dh.seq) return false;                                           // :194-200
    out.tracking_result.clear();
    const uint32_t n = rg.take<uint32_t>();
    for (uint32_t i = 0; i < n; ++i) {                                            // :203-251 (a uint16 id is always in [0, 65535])
        MaskKpts mk;
        mk.label = rg.string();
        mk.track_id = rg.take<uint16_t>();
        mk.mask = rg.image(CV_8UC1, 1, "mono8", "8UC1");
        float tl[3], br[3];
        for (float& v : tl) v = rg.take<float>();
        for (float& v : br) v = rg.take<float>();
        std::vector<Eigen::Vector3d> kp[2];
        for (auto& list : kp) {
            const uint32_t m = rg.take<uint32_t>();
            for (uint32_t k = 0; k < m; ++k) {
                const float x = rg.take<float>(), y = rg.take<float>(), z = rg.take<float>();
                list.push_back(Eigen::Vector3d(x, y, z));
            }
        }
        if (mk.label != "static") {                                               // :214-247
            mk.bbox = BBox2D{(int)tl[0], (int)tl[1], (int)br[0], (int)br[1]};
            mk.kpts_current = kp[0];
            mk.kpts_previous = kp[1];
            if (mk.kpts_current.size() != mk.kpts_previous.size()) continue;
        }
        out.tracking_result.push_back(mk);
    }
    if (!rg.done()) throw std::runtime_error("dspwire: trailing bytes after MaskGroup");
    return true;
}

}  // namespace dspwire
