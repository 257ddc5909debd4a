/*
 * ros_shim.h — the ROS surface the reference's hot-path translation units touch (ros::Time, Rate, ok,
 * NodeHandle/Publisher/Subscriber, the CameraInfo / Marker / Point / ColorRGBA message structs), so that
 * the UNMODIFIED reference sources compile without ROS.  TEST INFRASTRUCTURE ONLY (see eigen_shim.h).
 * Nothing here computes anything on the path: time is std::chrono, publishing a marker stores it in a
 * process-wide slot the bridge (oracle/ref_bridge.cpp) reads back, ros::ok() counts down a budget so
 * that SDF::visualize's loop (sdf.cpp:324-389) runs a chosen number of times.
 */
#ifndef TSDF_ORACLE_ROS_SHIM_H_
#define TSDF_ORACLE_ROS_SHIM_H_

#include <chrono>
#include <cstdint>
#include <cstdlib>
#include <memory>
#include <string>
#include <vector>
#include "boost_shim.h"

namespace std_msgs {
struct Header { uint32_t seq = 0; struct { double t = 0; } stamp_; std::string frame_id; };
struct ColorRGBA { float r = 0, g = 0, b = 0, a = 0; };
}
namespace geometry_msgs {
struct Point { double x = 0, y = 0, z = 0; };
struct Quaternion { double x = 0, y = 0, z = 0, w = 0; };
struct Vector3 { double x = 0, y = 0, z = 0; };
struct Pose { Point position; Quaternion orientation; };
}

namespace ros {
struct Duration {
    double s;
    explicit Duration(double v = 0) : s(v) {}
    double toSec() const { return s; }
};
struct Time {
    double s;
    Time() : s(0) {}
    explicit Time(double v) : s(v) {}
    static Time now() {
        return Time(std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count());
    }
    double toSec() const { return s; }
    Duration operator-(const Time& o) const { return Duration(s - o.s); }
};
struct Rate {
    explicit Rate(double) {}
    bool sleep() { return true; }
};
namespace shim {
inline int& ok_budget() { static int b = 0; return b; }
inline void*& last_marker_slot() { static void* p = nullptr; return p; }
}
inline bool ok() { return shim::ok_budget()-- > 0; }
struct Subscriber { void shutdown() {} };
}

namespace visualization_msgs {
struct Marker {
    enum { ARROW = 0, CUBE = 1, SPHERE = 2, CYLINDER = 3, LINE_STRIP = 4, LINE_LIST = 5, CUBE_LIST = 6,
           SPHERE_LIST = 7, POINTS = 8, TEXT_VIEW_FACING = 9, MESH_RESOURCE = 10, TRIANGLE_LIST = 11 };
    enum { ADD = 0, MODIFY = 0, DELETE = 2 };
    struct { uint32_t seq = 0; ros::Time stamp; std::string frame_id; } header;
    std::string ns;
    int32_t id = 0, type = 0, action = 0;
    geometry_msgs::Pose pose;
    geometry_msgs::Vector3 scale;
    std_msgs::ColorRGBA color;
    std::vector<geometry_msgs::Point> points;
    std::vector<std_msgs::ColorRGBA> colors;
};
struct MarkerArray { std::vector<Marker> markers; };
}

namespace ros {
struct Publisher {
    /* the bridge reads the last published marker back; nothing leaves the process */
    void publish(const visualization_msgs::Marker& m) const {
        auto*& slot = reinterpret_cast<visualization_msgs::Marker*&>(shim::last_marker_slot());
        if (!slot) slot = new visualization_msgs::Marker();
        *slot = m;
    }
};
struct NodeHandle {
    template <typename M> Publisher advertise(const std::string&, uint32_t) { return Publisher(); }
};
}

namespace sensor_msgs {
struct CameraInfo { uint32_t height = 0, width = 0; boost::array<double, 9> K; };
typedef boost::shared_ptr<CameraInfo const> CameraInfoConstPtr;
struct Image {};
}

#endif
