#pragma once
#include <cstdint>
namespace pcl {
struct PointXYZRGB { float x = 0, y = 0, z = 0; uint8_t r = 0, g = 0, b = 0; };
}  // namespace pcl
