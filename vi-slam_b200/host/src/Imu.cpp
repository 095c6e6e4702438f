// Imu.cpp — see Imu.hpp.  The Imu methods follow src/Imu.cpp statement by statement where a statement fixes the result
// (float vs double arithmetic, which sample a loop keeps, which field feeds which) — the harness in
// oracle/cvshim/ref_imu_capi.cpp runs the reference's own Imu.cpp on the same samples and the same filter output, and
// tests/test_imu.py compares every public field bit for bit.
#include "vislam/Imu.hpp"

#include <chrono>
#include <cmath>
#include <iomanip>
#include <iostream>

// ------------------------------------------------------------------------------------------------ MadgwickFilter
namespace vi {

MadgwickFilter::MadgwickFilter(double gain, double constant_dt)
    : gain_(gain), dt_(constant_dt), q0_(1), q1_(0), q2_(0), q3_(0), initialized_(false) {}

void MadgwickFilter::reset() { q0_ = 1; q1_ = q2_ = q3_ = 0; initialized_ = false; }

void MadgwickFilter::setOrientation(double w, double x, double y, double z) {
    const double n = std::sqrt(w * w + x * x + y * y + z * z);
    q0_ = w / n; q1_ = x / n; q2_ = y / n; q3_ = z / n;
    initialized_ = true;
}

void MadgwickFilter::getOrientation(double& w, double& x, double& y, double& z) const { w = q0_; x = q1_; y = q2_; z = q3_; }

void MadgwickFilter::update(double gx, double gy, double gz, double ax, double ay, double az) {
    if (!initialized_) {
        // stateless orientation from the accelerometer: the measured specific force points along world +z; zero yaw
        const double na = std::sqrt(ax * ax + ay * ay + az * az);
        if (na > 0) {
            const double roll = std::atan2(ay, az), pitch = std::atan2(-ax, std::sqrt(ay * ay + az * az));
            const Quaterniond q = toQuaternion(roll, pitch, 0.0);
            setOrientation(q.w, q.x, q.y, q.z);
        } else {
            initialized_ = true;
        }
    }
    const double q0 = q0_, q1 = q1_, q2 = q2_, q3 = q3_;
    // rate of change of the quaternion from the gyroscope: 1/2 q (x) (0, w)
    double d0 = 0.5 * (-q1 * gx - q2 * gy - q3 * gz);
    double d1 = 0.5 * (q0 * gx + q2 * gz - q3 * gy);
    double d2 = 0.5 * (q0 * gy - q1 * gz + q3 * gx);
    double d3 = 0.5 * (q0 * gz + q1 * gy - q2 * gx);
    const double na = std::sqrt(ax * ax + ay * ay + az * az);
    if (na > 0) {
        ax /= na; ay /= na; az /= na;
        // f = R(q)^T (0, 0, 1) - a: the world's up direction seen from the sensor minus the measured one
        const double f0 = 2.0 * (q1 * q3 - q0 * q2) - ax;
        const double f1 = 2.0 * (q0 * q1 + q2 * q3) - ay;
        const double f2 = 1.0 - 2.0 * (q1 * q1 + q2 * q2) - az;
        // s = J^T f with J = d(R^T (0, 0, 1)) / dq
        double s0 = -2.0 * q2 * f0 + 2.0 * q1 * f1;
        double s1 = 2.0 * q3 * f0 + 2.0 * q0 * f1 - 4.0 * q1 * f2;
        double s2 = -2.0 * q0 * f0 + 2.0 * q3 * f1 - 4.0 * q2 * f2;
        double s3 = 2.0 * q1 * f0 + 2.0 * q2 * f1;
        const double ns = std::sqrt(s0 * s0 + s1 * s1 + s2 * s2 + s3 * s3);
        if (ns > 0) {
            d0 -= gain_ * (s0 / ns); d1 -= gain_ * (s1 / ns); d2 -= gain_ * (s2 / ns); d3 -= gain_ * (s3 / ns);
        }
    }
    const double r0 = q0 + d0 * dt_, r1 = q1 + d1 * dt_, r2 = q2 + d2 * dt_, r3 = q3 + d3 * dt_;
    const double nq = std::sqrt(r0 * r0 + r1 * r1 + r2 * r2 + r3 * r3);
    q0_ = r0 / nq; q1_ = r1 / nq; q2_ = r2 / nq; q3_ = r3 / nq;
}

}  // namespace vi

// ------------------------------------------------------------------------------------------------ ImuFilterNode
ImuFilterNode::ImuFilterNode() : timeNs(0), timeS(0), rateHZ(0), has_pending_(false), source_(nullptr), source_user_(nullptr) {}
ImuFilterNode::ImuFilterNode(int rate) : timeNs(0), timeS(0), rateHZ(0), has_pending_(false), source_(nullptr), source_user_(nullptr) {
    createROSPublisher(rate);
}
void ImuFilterNode::createROSPublisher(int rate) { rateHZ = rate; }                         // Imu.cpp:462-468
void ImuFilterNode::createROSSubscriber() { timeNs = 0; timeS = 0; }                       // :471-480
void ImuFilterNode::setOrientationSource(OrientationSource fn, void* user) { source_ = fn; source_user_ = user; }

// Imu.cpp:500-541: one raw sample to the filter (message stamp advances 5000 ns per sample; no sleep, no topic)
void ImuFilterNode::UpdatePublisher(cv::Point3d w, cv::Point3d a) {
    timeNs = timeNs + 5000;
    vi::ImuMsg raw;
    raw.angular_velocity.x = w.x; raw.angular_velocity.y = w.y; raw.angular_velocity.z = w.z;
    raw.linear_acceleration.x = a.x; raw.linear_acceleration.y = a.y; raw.linear_acceleration.z = a.z;
    pending_ = raw;
    if (source_) {
        source_(source_user_, raw, pending_);
    } else {
        filter_.update(w.x, w.y, w.z, a.x, a.y, a.z);
        double qw, qx, qy, qz;
        filter_.getOrientation(qw, qx, qy, qz);
        pending_.orientation.w = qw; pending_.orientation.x = qx; pending_.orientation.y = qy; pending_.orientation.z = qz;
    }
    has_pending_ = true;
}

// Imu.cpp:543-546 + imuCallback :482-497: the fused message becomes visible to the caller
void ImuFilterNode::UpdateSubscriber() {
    if (!has_pending_) return;
    imuFusedData.angular_velocity = pending_.angular_velocity;
    imuFusedData.linear_acceleration = pending_.linear_acceleration;
    imuFusedData.orientation = pending_.orientation;
    has_pending_ = false;
}

std::string ImuFilterNode::getNodeName() { return "ImuFilter"; }
double ImuFilterNode::getRateHZ() { return rateHZ; }

// ------------------------------------------------------------------------------------------------ Imu
using cv::Matx33f;
using cv::Point3d;
using cv::Point3f;

Imu::Imu() : initialYawGt(0), initialYawFilter(0), YawGt(0), timeStep(0), elapsed_filter(0), n(0), n_total(0), currentTimeMs(0) {}

Imu::Imu(double timestep)
    : initialYawGt(0), initialYawFilter(0), YawGt(0), timeStep(0), elapsed_filter(0), n(0), n_total(0), currentTimeMs(0) {
    createPublisher(timestep);
}

void Imu::createPublisher(double _timeStep) {                                               // :14-20
    timeStep = _timeStep;
    createROSPublisher(static_cast<int>((1.0 / timeStep) * 20));
    createROSSubscriber();
}

void Imu::setImuData(std::vector<Point3d>& w_measure, std::vector<Point3d>& a_measure) {    // :22-31
    angularVelocityMeasure.assign(w_measure.begin(), w_measure.end());
    accelerationMeasure.assign(a_measure.begin(), a_measure.end());
    n = (int)angularVelocityMeasure.size();
}

void Imu::setImuBias(Point3d acc_bias, Point3d ang_bias) { accBias = acc_bias; angBias = ang_bias; }
void Imu::setImuInitialVelocity(Point3d initial_velocity) { initialVelocity = initial_velocity; }
void Imu::setImuInitialPosition() {}
void Imu::computeGravity() {}

namespace {
Quaterniond fused_orientation(const vi::ImuMsg& m) {
    Quaterniond q;
    q.x = m.orientation.x; q.y = m.orientation.y; q.z = m.orientation.z; q.w = m.orientation.w;
    return q;
}
}  // namespace

void Imu::initializate(double gt_yaw, Point3d gt_velocity, std::vector<Point3d>& w_measure,
                       std::vector<Point3d>& a_measure) {                                   // :42-91
    initialYawGt = gt_yaw;
    setImuInitialVelocity(gt_velocity);
    setImuData(w_measure, a_measure);
    angBias = accBias = Point3d(0.0, 0.0, 0.0);
    accelerationWorld.clear();
    rpyAnglesWorld.clear();
    quaternionWorld.clear();
    angularVelocityIMUFilter.clear();
    calibrateAng(3);
    for (int i = 0; i < n; i++) {
        UpdatePublisher(angularVelocityMeasure[i] - angBias, accelerationMeasure[i]);
        UpdateSubscriber();
        rpyAnglesWorld.push_back(toRPY(fused_orientation(imuFusedData)));
    }
    initialYawFilter = rpyAnglesWorld.back().z;
    for (int i = 0; i < n; i++) {
        rpyAnglesWorld[i].z = initialYawGt;                         // yaw aligned with the ground truth's initial yaw
        quaternionWorld.push_back(toQuaternion(rpyAnglesWorld[i].x, rpyAnglesWorld[i].y, rpyAnglesWorld[i].z));
        world2imuRotation.push_back(RPY2rotationMatrix(rpyAnglesWorld[i]));
    }
    n_total = n;
    currentTimeMs = n_total * timeStep * 1000;
}

void Imu::calibrateAng(int axis) {                                                          // :93-124
    Point3d sum(0.0, 0.0, 0.0);
    for (int i = 0; i < n; i++) sum = sum + angularVelocityMeasure[i];
    if (axis == 0) angBias.x = sum.x / n;
    if (axis == 1) angBias.y = sum.y / n;
    if (axis == 2) angBias.z = sum.z / n;
    if (axis == 3) angBias = sum / n;
}

void Imu::calibrateAcc(int axis) {                                                          // :127-163 (float arithmetic)
    const Point3f gravityInWorld(0.0f, 0.0f, 9.68f);
    Point3f sum(0.0f, 0.0f, 0.0f);
    for (int i = 0; i < n; i++) {
        const Point3f gravityInImu = world2imuRotation[i].t() * gravityInWorld;
        sum = sum + Point3f(accelerationMeasure[i]) - gravityInImu;
    }
    if (axis == 0) accBias.x = sum.x / n;
    if (axis == 1) accBias.y = sum.y / n;
    if (axis == 2) accBias.z = sum.z / n;
    if (axis == 3) accBias = sum / n;
}

void Imu::detectAngBias() {                                                                 // :165-210
    int cx = 0, cy = 0, cz = 0;
    const Point3d thr(std::fabs(angBias.x) * (1 + 0.2), std::fabs(angBias.y) * (1 + 0.2), std::fabs(angBias.z) * (1 + 0.2));
    for (int i = 0; i < n; i++) {
        if (std::fabs(angularVelocityMeasure[i].x) < thr.x) cx++;
        if (std::fabs(angularVelocityMeasure[i].y) < thr.y) cy++;
        if (std::fabs(angularVelocityMeasure[i].z) < thr.z) cz++;
    }
    if (cx >= 9 && cy >= 9 && cz >= 9) {
        calibrateAng(3);
    } else {
        if (cx >= 9) calibrateAng(0);
        if (cy >= 9) calibrateAng(1);
        if (cz >= 9) calibrateAng(2);
    }
}

void Imu::detectAccBias() {                                                                 // :212-272
    int cx = 0, cy = 0, cz = 0;
    const Point3f thr(0.3f, 0.3f, 0.1f), gravityInWorld(0.0f, 0.0f, 9.68f);
    for (int i = 0; i < n; i++) {
        const Point3f g = world2imuRotation[i].t() * gravityInWorld;
        const Point3f a = Point3f(accelerationMeasure[i]) - g;
        if (std::fabs(a.x) < thr.x) cx++;
        if (std::fabs(a.y) < thr.y) cy++;
        if (std::fabs(a.z) < thr.z) cz++;
    }
    if (cx >= 9 && cy >= 9 && cz >= 9) {
        calibrateAcc(3);
    } else {
        if (cx >= 9) calibrateAcc(0);
        if (cy >= 9) calibrateAcc(1);
        if (cz >= 9) calibrateAcc(2);
    }
}

void Imu::estimateOrientation() {                                                           // :279-317
    clearData();
    elapsed_filter = 0.0;
    for (int i = 0; i < n; i++) {
        const auto t1 = std::chrono::steady_clock::now();
        UpdatePublisher(angularVelocityMeasure[i] - angBias, accelerationMeasure[i]);
        UpdateSubscriber();
        elapsed_filter += std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count();
        rpyAnglesWorld.push_back(toRPY(fused_orientation(imuFusedData)));
        rpyAnglesWorld[i].z = rpyAnglesWorld[i].z - initialYawFilter + initialYawGt;
        quaternionWorld.push_back(toQuaternion(rpyAnglesWorld[i].x, rpyAnglesWorld[i].y, rpyAnglesWorld[i].z));
        world2imuRotation.push_back(RPY2rotationMatrix(rpyAnglesWorld[i]));
        angularVelocityIMUFilter.push_back(Point3d(imuFusedData.angular_velocity.x, imuFusedData.angular_velocity.y,
                                                   imuFusedData.angular_velocity.z));
    }
    elapsed_filter = elapsed_filter / n;
}

void Imu::clearData() {                                                                     // :319-325
    rpyAnglesWorld.clear();
    quaternionWorld.clear();
    angularVelocityIMUFilter.clear();
    world2imuRotation.clear();
}

void Imu::computeAcceleration() {                                                           // :327-343
    accelerationWorld.clear();
    for (int i = 0; i < n; i++) {
        accelerationMeasure[i] = accelerationMeasure[i] + accBias;           // the stored sample itself is modified
        Point3d accWorld(world2imuRotation[i] * Point3f(accelerationMeasure[i]));   // float product, widened
        accWorld.z = accWorld.z - 9.68;
        accelerationWorld.push_back(accWorld);
    }
}

void Imu::computeVelocity() {                                                               // :345-357
    velocity = Point3d(0, 0, 0);
    for (int i = 0; i < n; i++) {
        velocity.x = velocity.x + (accelerationWorld[i].x) * timeStep;
        velocity.y = velocity.y + (accelerationWorld[i].y) * timeStep;
        velocity.z = velocity.z + (accelerationWorld[i].z) * timeStep;
    }
}

void Imu::computePosition() {                                                               // :359-374: only the LAST sample's term survives
    Point3d t;
    for (int i = 0; i < n; i++) {
        t.x = 0.5 * (accelerationWorld[i].x) * timeStep * timeStep;
        t.y = 0.5 * (accelerationWorld[i].y) * timeStep * timeStep;
        t.z = 0.5 * (accelerationWorld[i].z) * timeStep * timeStep;
    }
    position.x = initialVelocity.x * n * timeStep + t.x;
    position.y = initialVelocity.y * n * timeStep + t.y;
    position.z = initialVelocity.z * n * timeStep + t.z;
}

void Imu::computeAngularVelocity() {                                                        // :376-390
    angularVelocity = Point3d(0.0, 0.0, 0.0);
    for (int i = 0; i < n; i++) {
        angularVelocity.x = angularVelocity.x + angularVelocityMeasure[i].x;
        angularVelocity.y = angularVelocity.y + angularVelocityMeasure[i].y;
        angularVelocity.z = angularVelocity.z + angularVelocityMeasure[i].z;
    }
    angularVelocity.x = angularVelocity.x / n;
    angularVelocity.y = angularVelocity.y / n;
    angularVelocity.z = angularVelocity.z / n;
}

void Imu::computeAngularPosition() {                                                        // :392-399: all three from .x upstream
    angularPosition.x = angularVelocity.x * timeStep * n;
    angularPosition.y = angularVelocity.x * timeStep * n;
    angularPosition.z = angularVelocity.x * timeStep * n;
}

void Imu::estimate() {                                                                      // :402-433
    estimateOrientation();
    computeAcceleration();
    computeVelocity();
    computePosition();
    computeAngularVelocity();
    computeAngularPosition();
    init_rotationMatrix = RPY2rotationMatrix(rpyAnglesWorld[0]);
    final_rotationMatrix = RPY2rotationMatrix(rpyAnglesWorld.back());
    residual_rotationMatrix = init_rotationMatrix.t() * final_rotationMatrix;
    residualRPY = rotationMatrix2RPY(residual_rotationMatrix);
    residualVelocity = velocity;
    residualPosition = position;
    initialVelocity = initialVelocity + residualVelocity;
    n_total = n_total + n;
    currentTimeMs = n_total * timeStep * 1000;
    if (currentTimeMs < 2500) calibrateAng(3);
}

void Imu::printStatistics() {                                                               // :447-451
    std::cout << "\nESTADISTICAS IMU" << "\nTiempo de filtrado : " << std::fixed << std::setprecision(3)
              << elapsed_filter * 1000 << " ms" << std::endl;
}

Point3d Imu::transform2World(Point3d acc, Point3d angl) {                                   // :553-575
    const double c1 = std::cos(angl.x), s1 = std::sin(angl.x), c2 = std::cos(angl.y), s2 = std::sin(angl.y);
    const double c3 = std::cos(angl.z), s3 = std::sin(angl.z);
    Point3d w;
    w.x = c3 * c2 * acc.x + (c3 * s2 * s1 - s3 * c1) * acc.y + (c3 * s2 * c1 + s3 * s1) * acc.z;
    w.y = s3 * c2 * acc.x + (s3 * s2 * s1 + c3 * c1) * acc.y + (s3 * s2 * c1 - c3 * s1) * acc.z;
    w.z = -s2 * acc.x + c2 * s1 * acc.y + c2 * c1 * acc.z;
    return w;
}
