#pragma once
#include <mask_kpts_msgs/MaskKpts.h>
namespace mask_kpts_msgs {
struct MaskGroup {     // script/mmseg/mask_kpts_msgs/msg/_MaskGroup.py:17-18
    std_msgs::Header header;
    std::vector<MaskKpts> objects;
    typedef std::shared_ptr<const MaskGroup> ConstPtr;
};
}  // namespace mask_kpts_msgs
