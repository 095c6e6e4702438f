// Imu.hpp — the IMU prior of the tracking loop with the reference's class interface (include/Imu.hpp:11-105,
// src/Imu.cpp), without ROS.  Upstream, ImuFilterNode publishes every raw sample on a ROS topic and
// reads the fused orientation back from the external `imu_filter_madgwick` node (launch/vi_slam.launch:19-22:
// use_mag false, gain 0.1, world frame NWU, stateless false, constant_dt 0.005) — the only inter-process hop of the
// system.  Here the node is an in-process filter object (MadgwickFilter, a restatement of that package's published
// algorithm); UpdatePublisher / UpdateSubscriber keep their names and meaning.  Imu::estimate() produces
// residual_rotationMatrix, which VISystem turns into the initial rotation of the GN pose solve (VISystem.cpp:1135).
#ifndef VISLAM_IMU_HPP_
#define VISLAM_IMU_HPP_
#include <string>
#include <vector>

#include "Plus.hpp"
#include "compat.hpp"

namespace vi {

// Madgwick's gradient-descent orientation filter, IMU-only variant, as the imu_filter_madgwick node runs it:
// q (w, x, y, z) rotates the sensor frame into a z-up world frame (NWU / ENU), the accelerometer reference direction is
// (0, 0, 1), one update is  qdot = 1/2 q (x) (0, w) - gain * normalised(J^T f),  q += qdot dt,  q normalised.
// With stateless == false the first sample initialises q from the accelerometer alone (zero yaw) and is then filtered.
class MadgwickFilter {
public:
    explicit MadgwickFilter(double gain = 0.1, double constant_dt = 0.005);
    void reset();
    void setOrientation(double w, double x, double y, double z);
    void getOrientation(double& w, double& x, double& y, double& z) const;
    void update(double gx, double gy, double gz, double ax, double ay, double az);   // one sample, dt = constant_dt
    bool initialized() const { return initialized_; }

private:
    double gain_, dt_;
    double q0_, q1_, q2_, q3_;
    bool initialized_;
};

// sensor_msgs/Imu as far as this path reads it
struct ImuMsg {
    struct Vec3 { double x, y, z; Vec3() : x(0), y(0), z(0) {} };
    struct Quat { double x, y, z, w; Quat() : x(0), y(0), z(0), w(1) {} };
    Quat orientation;
    Vec3 angular_velocity;
    Vec3 linear_acceleration;
};

}  // namespace vi

class ImuFilterNode {
public:
    ImuFilterNode();
    explicit ImuFilterNode(int rate);
    void createROSPublisher(int rate);
    void createROSSubscriber();
    void UpdatePublisher(cv::Point3d w_measure, cv::Point3d a_measure);   // "publish" = feed the in-process filter
    void UpdateSubscriber();                                              // "spin" = fetch the fused sample
    std::string getNodeName();
    double getRateHZ();
    // test / integration hook: take the orientation from this callback instead of the built-in filter
    typedef void (*OrientationSource)(void* user, const vi::ImuMsg& raw, vi::ImuMsg& fused);
    void setOrientationSource(OrientationSource fn, void* user);

    vi::ImuMsg imuFusedData;
    unsigned int timeNs;
    unsigned int timeS;

private:
    int rateHZ;
    vi::MadgwickFilter filter_;
    vi::ImuMsg pending_;
    bool has_pending_;
    OrientationSource source_;
    void* source_user_;
};

class Imu : public ImuFilterNode {
public:
    Imu();
    explicit Imu(double timestep);
    void createPublisher(double _timeStep);
    void setImuData(std::vector<cv::Point3d>& w_measure, std::vector<cv::Point3d>& a_measure);
    void setImuBias(cv::Point3d acc_Bias, cv::Point3d ang_Bias);
    void setImuInitialVelocity(cv::Point3d initial_velocity);
    void setImuInitialPosition();
    void initializate(double gt_yaw, cv::Point3d gt_velocity, std::vector<cv::Point3d>& w_measure,
                      std::vector<cv::Point3d>& a_measure);
    void estimate();
    void estimateOrientation();
    void computeGravity();
    void computePosition();
    void computeVelocity();
    void computeAcceleration();
    void computeAngularVelocity();
    void computeAngularPosition();
    void calibrateAng(int axis);
    void calibrateAcc(int axis);
    void detectAngBias();
    void detectAccBias();
    void printStatistics();
    void clearData();
    cv::Point3d transform2World(cv::Point3d acc, cv::Point3d localAngles);

    cv::Point3d angularPosition;
    cv::Point3d angularVelocity;
    cv::Point3d initialVelocity;
    double initialYawGt;
    double initialYawFilter;
    double YawGt;
    cv::Point3d position;
    cv::Point3d velocity;
    cv::Point3d accBias;
    cv::Point3d angBias;
    std::vector<cv::Point3d> angularVelocityIMUFilter;
    std::vector<Quaterniond> quaternionWorld;
    std::vector<cv::Point3d> rpyAnglesWorld;
    std::vector<cv::Point3d> accelerationWorld;
    cv::Matx33f init_rotationMatrix;
    cv::Matx33f final_rotationMatrix;
    cv::Matx33f residual_rotationMatrix;
    std::vector<cv::Matx33f> world2imuRotation;
    cv::Point3d residualRPY;
    cv::Point3d residualPosition;
    cv::Point3d residualVelocity;
    double timeStep;

private:
    double elapsed_filter;
    std::vector<cv::Point3d> angularVelocityMeasure;
    std::vector<cv::Point3d> accelerationMeasure;
    int n;
    int n_total;
    double currentTimeMs;
};

#endif
