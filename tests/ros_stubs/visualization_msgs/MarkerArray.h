#pragma once
#include <ros/ros.h>
namespace visualization_msgs { struct MarkerArray {}; }
