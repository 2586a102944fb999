#pragma once
#include <opencv2/opencv.hpp>
#include <sensor_msgs/Image.h>
namespace cv_bridge {
struct Exception : std::runtime_error { using std::runtime_error::runtime_error; };
struct CvImage { cv::Mat image; };
typedef std::shared_ptr<CvImage> CvImagePtr;
// the two conversions the node asks for: depth as 32FC1, masks as MONO8 (src/mapping.cpp:180, 212); no colour conversion
inline CvImagePtr toCvCopy(const sensor_msgs::Image& m, const std::string& want) {
    const bool f32 = want == sensor_msgs::image_encodings::TYPE_32FC1;
    if (f32 ? m.encoding != "32FC1" : (m.encoding != "mono8" && m.encoding != "8UC1")) throw Exception("cv_bridge stand-in: cannot convert " + m.encoding + " to " + want);
    const size_t elem = f32 ? 4 : 1;
    if (m.is_bigendian || m.step < m.width * elem || m.data.size() < static_cast<size_t>(m.step) * m.height) throw Exception("cv_bridge stand-in: malformed image");
    CvImagePtr out(new CvImage);
    out->image = cv::Mat(static_cast<int>(m.height), static_cast<int>(m.width), f32 ? CV_32FC1 : CV_8UC1);
    for (uint32_t i = 0; i < m.height; ++i) std::memcpy(out->image.data + i * out->image.step, m.data.data() + static_cast<size_t>(i) * m.step, m.width * elem);
    return out;
}
inline CvImagePtr toCvCopy(const sensor_msgs::ImageConstPtr& m, const std::string& want) { return toCvCopy(*m, want); }
}  // namespace cv_bridge
