#pragma once
#include <ros/ros.h>
namespace std_msgs { struct Int32 { int32_t data = 0; }; }
