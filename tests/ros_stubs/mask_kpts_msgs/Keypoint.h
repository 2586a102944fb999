#pragma once
#include <ros/ros.h>
namespace mask_kpts_msgs { struct Keypoint { float x = 0, y = 0, z = 0; }; }
