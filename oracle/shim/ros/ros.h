// TEST INFRASTRUCTURE: offline stand-in for roscpp, just enough for the reference nodes to compile
// unmodified.  Parameters come from a text file, topics from an event file, and ros::spin() runs the
// offline driver defined in shim_driver.inl.
#pragma once
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <iostream>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

namespace bfshim {
struct ParamServer {
  std::map<std::string, std::string> scalars;                    // key -> text
  std::map<std::string, std::map<std::string, double> > maps;    // micN -> {id,x,y}
  static ParamServer& get() { static ParamServer p; return p; }
};
struct Topics {
  std::function<void(float)> theta;
  std::function<void(unsigned short, float)> interf;
  static Topics& get() { static Topics t; return t; }
};
inline std::string leaf(const std::string& name) {
  size_t p = name.rfind('/');
  return p == std::string::npos ? name : name.substr(p + 1);
}
inline bool verbose_log() { static int v = getenv("BFREF_VERBOSE") ? 1 : 0; return v != 0; }
void load_params();   // shim_driver.inl
void run_offline();   // shim_driver.inl
}   // namespace bfshim

#define ROS_INFO(...) do { if (bfshim::verbose_log()) { printf(__VA_ARGS__); printf("\n"); } } while (0)
#define ROS_WARN(...) do { if (bfshim::verbose_log()) { printf(__VA_ARGS__); printf("\n"); } } while (0)
#define ROS_ERROR(...) do { fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); } while (0)

namespace std_msgs {
struct Float32 {
  float data;
  typedef std::shared_ptr<const Float32> ConstPtr;
};
struct Header {};
}   // namespace std_msgs
namespace beamform {
struct InterfTheta {
  unsigned short id;   // msg/InterfTheta.msg: uint16 id
  float angle;         //                      float32 angle
  typedef std::shared_ptr<const InterfTheta> ConstPtr;
};
}   // namespace beamform

namespace ros {
inline void init(int&, char**, const char*) { bfshim::load_params(); }
inline void shutdown() {}
inline void spin() { bfshim::run_offline(); }
inline bool ok() { return true; }
namespace this_node {
inline std::string getName() { return "/beamform"; }
}
struct Subscriber {};
struct Publisher {};
class NodeHandle {
 public:
  bool getParam(const std::string& key, bool& v) {
    auto& s = bfshim::ParamServer::get().scalars;
    auto it = s.find(bfshim::leaf(key));
    if (it == s.end()) return false;
    v = (it->second == "true" || it->second == "True" || it->second == "1");
    return true;
  }
  bool getParam(const std::string& key, double& v) {
    auto& s = bfshim::ParamServer::get().scalars;
    auto it = s.find(bfshim::leaf(key));
    if (it == s.end()) return false;
    v = atof(it->second.c_str());
    return true;
  }
  bool getParam(const std::string& key, int& v) {
    // roscpp only converts an XmlRpc int to int; launch files write e.g. "smooth_size: 3"
    auto& s = bfshim::ParamServer::get().scalars;
    auto it = s.find(bfshim::leaf(key));
    if (it == s.end()) return false;
    v = atoi(it->second.c_str());
    return true;
  }
  bool getParam(const std::string& key, std::map<std::string, double>& v) {
    auto& m = bfshim::ParamServer::get().maps;
    auto it = m.find(bfshim::leaf(key));
    if (it == m.end()) return false;
    v = it->second;
    return true;
  }
  Subscriber subscribe(const char* topic, int, void (*cb)(const std_msgs::Float32::ConstPtr&)) {
    if (std::string(topic) == "theta")
      bfshim::Topics::get().theta = [cb](float a) {
        auto m = std::make_shared<std_msgs::Float32>();
        m->data = a;
        cb(m);
      };
    return Subscriber();
  }
  Subscriber subscribe(const char* topic, int, void (*cb)(const beamform::InterfTheta::ConstPtr&)) {
    if (std::string(topic) == "theta_interference")
      bfshim::Topics::get().interf = [cb](unsigned short id, float a) {
        auto m = std::make_shared<beamform::InterfTheta>();
        m->id = id;
        m->angle = a;
        cb(m);
      };
    return Subscriber();
  }
};
}   // namespace ros
