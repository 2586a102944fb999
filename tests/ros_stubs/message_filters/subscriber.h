#pragma once
#include <ros/ros.h>
namespace message_filters {
template <typename M> class Subscriber {
public:
    Subscriber(ros::NodeHandle&, const std::string& topic, uint32_t) : topic_(topic) {}
    std::string topic_;
};
}  // namespace message_filters
