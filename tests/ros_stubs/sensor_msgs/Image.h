#pragma once
#include <ros/ros.h>
namespace sensor_msgs {
struct Image {
    std_msgs::Header header;
    uint32_t height = 0, width = 0;
    std::string encoding;
    uint8_t is_bigendian = 0;
    uint32_t step = 0;
    std::vector<uint8_t> data;
    typedef std::shared_ptr<Image> Ptr;
    typedef std::shared_ptr<const Image> ConstPtr;
};
typedef std::shared_ptr<const Image> ImageConstPtr;
namespace image_encodings {
const std::string TYPE_32FC1 = "32FC1";
const std::string MONO8 = "mono8";
}  // namespace image_encodings
}  // namespace sensor_msgs
