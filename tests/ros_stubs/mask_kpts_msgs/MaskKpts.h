#pragma once
#include <mask_kpts_msgs/Keypoint.h>
#include <sensor_msgs/Image.h>
namespace mask_kpts_msgs {
struct MaskKpts {      // script/mmseg/mask_kpts_msgs/msg/_MaskKpts.py:17-24
    std::string label;
    uint16_t track_id = 0;
    sensor_msgs::Image mask;
    Keypoint bbox_tl, bbox_br;
    std::vector<Keypoint> kpts_curr, kpts_last;
};
}  // namespace mask_kpts_msgs
