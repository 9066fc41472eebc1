#pragma once
#include "ros/ros.h"
namespace jack_msgs {
struct JackAudio {
  std::vector<float> data;
  unsigned short size;
  std_msgs::Header header;
  typedef std::shared_ptr<const JackAudio> ConstPtr;
};
}
