// oracle/rosshim/rosshim.hpp — MINIMAL ROS stand-in (TEST INFRASTRUCTURE ONLY): just enough of roscpp /
// sensor_msgs / message_filters for the reference's src/Imu.cpp + include/Imu.hpp to compile UNMODIFIED, so its
// Imu::initializate / estimate chain can run in-process.  The "imu/data_raw" -> imu_filter_madgwick -> "imu/data" round
// trip becomes a function call: publish() hands the raw message to a hook installed by the harness
// (oracle/cvshim/ref_imu_capi.cpp), which fills in the orientation (the external filter's job), and callAvailable()
// delivers the result to the registered subscriber callback.
#ifndef VSO_ROSSHIM_HPP
#define VSO_ROSSHIM_HPP
#include <unistd.h>

#include <cstdint>
#include <functional>
#include <memory>
#include <string>
#include <vector>

namespace boost {
using std::shared_ptr;
}

namespace sensor_msgs {
struct Imu {
    struct Stamp { uint32_t sec = 0, nsec = 0; };
    struct Header { std::string frame_id; Stamp stamp; };
    struct Vector3 { double x = 0, y = 0, z = 0; };
    struct Quaternion { double x = 0, y = 0, z = 0, w = 1; };
    typedef std::shared_ptr<const Imu> ConstPtr;
    Header header;
    Quaternion orientation;
    double orientation_covariance[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    Vector3 angular_velocity;
    Vector3 linear_acceleration;
};
}  // namespace sensor_msgs

namespace rosshim {
typedef void (*FilterHook)(void* user, const sensor_msgs::Imu& raw, sensor_msgs::Imu& fused);
struct Bus {
    FilterHook hook = nullptr;
    void* user = nullptr;
    std::vector<sensor_msgs::Imu> queue;                                   // fused messages waiting for callAvailable
    std::vector<std::function<void(const sensor_msgs::Imu::ConstPtr&)>> subscribers;
};
inline Bus& bus() { static Bus b; return b; }
}  // namespace rosshim

namespace ros {
inline bool ok() { return true; }
inline void init(int&, char**, const std::string&) {}   // VISystem(int, char**) (VISystem.cpp:15-17)
struct WallDuration { explicit WallDuration(double) {} };
struct Rate { explicit Rate(double) {} void sleep() {} };
namespace names { inline std::string resolve(const std::string& n) { return "/" + n; } }

class Publisher {
public:
    int getNumSubscribers() const { return 1; }
    template <typename M> void publish(const M& raw) const {
        M fused = raw;
        rosshim::Bus& b = rosshim::bus();
        if (b.hook) b.hook(b.user, raw, fused);
        b.queue.push_back(fused);
    }
};
class NodeHandle {
public:
    template <typename M> Publisher advertise(const std::string&, int) { return Publisher(); }
};
class CallbackQueue {
public:
    void callAvailable(WallDuration) {
        rosshim::Bus& b = rosshim::bus();
        std::vector<sensor_msgs::Imu> q;
        q.swap(b.queue);
        for (const sensor_msgs::Imu& m : q) {
            sensor_msgs::Imu::ConstPtr p = std::make_shared<const sensor_msgs::Imu>(m);
            for (auto& cb : b.subscribers) cb(p);
        }
    }
};
inline CallbackQueue* getGlobalCallbackQueue() { static CallbackQueue q; return &q; }
}  // namespace ros

#define ROS_WARN_ONCE(...) do {} while (0)

namespace std_msgs { struct Int32 { int data = 0; }; }
namespace geometry_msgs {
struct Vector3 { double x = 0, y = 0, z = 0; };
struct Quaternion { double x = 0, y = 0, z = 0, w = 1; };
}

namespace message_filters {
template <typename M>
class Subscriber {
public:
    Subscriber(ros::NodeHandle&, const std::string&, int) {}
    template <typename C> void registerCallback(void (C::*fn)(const typename M::ConstPtr&), C* obj) {
        rosshim::bus().subscribers.push_back([fn, obj](const typename M::ConstPtr& m) { (obj->*fn)(m); });
    }
};
}  // namespace message_filters
#endif
