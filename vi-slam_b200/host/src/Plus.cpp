// Plus.cpp — see Plus.hpp.  Every function evaluates the same double expressions as the reference
// (src/Plus.cpp:3-351) and narrows to float exactly where the reference stores into Matx33f / CV_32F, so the outputs
// are bit-identical to the reference's own build (tests/test_dataset_io.py checks that against oracle/_ref/libref_io.so).
#include "vislam/Plus.hpp"

#include <cmath>

namespace {
const double kPi = 3.14159265358979323846;

struct Trig {
    double c1, s1, c2, s2, c3, s3;   // roll, pitch, yaw
    explicit Trig(const cv::Point3d& rpy)
        : c1(std::cos(rpy.x)), s1(std::sin(rpy.x)), c2(std::cos(rpy.y)), s2(std::sin(rpy.y)), c3(std::cos(rpy.z)),
          s3(std::sin(rpy.z)) {}
    // R = Rz(yaw) Ry(pitch) Rx(roll), element (r, c) as a double expression (Plus.cpp:200-210, 301-311)
    double at(int r, int c) const {
        switch (3 * r + c) {
            case 0: return c3 * c2;
            case 1: return c3 * s2 * s1 - s3 * c1;
            case 2: return c3 * s2 * c1 + s3 * s1;
            case 3: return s3 * c2;
            case 4: return s3 * s2 * s1 + c3 * c1;
            case 5: return s3 * s2 * c1 - c3 * s1;
            case 6: return -s2;
            case 7: return c2 * s1;
            default: return c2 * c1;
        }
    }
};

cv::Point3d rpy_of(double r11, double r21, double r31, double r32, double r33) {   // Plus.cpp:72-74, 103-105
    cv::Point3d a;
    a.z = std::atan2(r21, r11);
    a.y = std::atan2(-r31, std::sqrt(r32 * r32 + r33 * r33));
    a.x = std::atan2(r32, r33);
    return a;
}
}  // namespace

Quaterniond toQuaternion(double roll, double pitch, double yaw) {            // Plus.cpp:3-20
    const double cy = std::cos(yaw * 0.5), sy = std::sin(yaw * 0.5);
    const double cr = std::cos(roll * 0.5), sr = std::sin(roll * 0.5);
    const double cp = std::cos(pitch * 0.5), sp = std::sin(pitch * 0.5);
    Quaterniond q;
    q.w = cy * cr * cp + sy * sr * sp;
    q.x = cy * sr * cp - sy * cr * sp;
    q.y = cy * cr * sp + sy * sr * cp;
    q.z = sy * cr * cp - cy * sr * sp;
    return q;
}

cv::Point3d toRPY(const Quaterniond& q) {                                     // Plus.cpp:24-54
    cv::Point3d a;
    a.x = std::atan2(+2.0 * (q.w * q.x + q.y * q.z), +1.0 - 2.0 * (q.x * q.x + q.y * q.y));
    const double sinp = +2.0 * (q.w * q.y - q.z * q.x);
    a.y = std::fabs(sinp) >= 1 ? std::copysign(kPi / 2, sinp) : std::asin(sinp);
    a.z = std::atan2(+2.0 * (q.w * q.z + q.x * q.y), +1.0 - 2.0 * (q.y * q.y + q.z * q.z));
    return a;
}

cv::Point3d toRPY360(cv::Point3d angles) {                                    // Plus.cpp:161-179
    cv::Point3d out = angles;
    if (angles.x < 0.0) out.x = angles.x + 2 * kPi;
    if (angles.y < 0.0) out.y = angles.y + 2 * kPi;
    if (angles.z < 0.0) out.z = angles.z + 2 * kPi;
    return out;
}

double computeDiff(double angle_ref, double angle2) {                         // Plus.cpp:128-159
    double mag = std::fabs(angle_ref - angle2);
    if (mag > kPi) mag = std::fabs(mag - 2 * kPi);
    if (angle_ref < 0.0) angle_ref = angle_ref + 2 * kPi;
    if (angle2 < 0.0) angle2 = angle2 + 2 * kPi;
    const double diff = angle_ref - angle2;
    if ((diff < kPi && diff > 0.0) || diff < -kPi) return -mag;
    return mag;
}

cv::Point3d rotationMatrix2RPY(cv::Matx33f m) {                               // Plus.cpp:56-84
    return rpy_of(m(0, 0), m(1, 0), m(2, 0), m(2, 1), m(2, 2));
}

// Plus.cpp:86-114.  Upstream assigns the two atan2 the other way round here (x = atan2(r21, r11), z = atan2(r32, r33)):
// the 4x4 variant returns (yaw, pitch, roll) in a Point3d whose fields every caller reads as (roll, pitch, yaw).  Kept.
cv::Point3d transformationMatrix2RPY(cv::Mat t) {
    const cv::Point3d a = rpy_of(t.at<float>(0, 0), t.at<float>(1, 0), t.at<float>(2, 0), t.at<float>(2, 1), t.at<float>(2, 2));
    return cv::Point3d(a.z, a.y, a.x);
}

cv::Point3d transformationMatrix2position(cv::Mat t) {                        // Plus.cpp:116-126
    return cv::Point3d(t.at<float>(0, 3), t.at<float>(1, 3), t.at<float>(2, 3));
}

cv::Matx33f RPY2rotationMatrix(cv::Point3d rpy) {                             // Plus.cpp:182-220
    const Trig g(rpy);
    cv::Matx33f m;
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) m(r, c) = (float)g.at(r, c);
    return m;
}

cv::Mat RPYAndPosition2transformationMatrix(cv::Point3d rpy, cv::Point3d position) {   // Plus.cpp:284-323
    const Trig g(rpy);
    cv::Mat t = cv::Mat::zeros(4, 4, CV_32FC1);
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) t.at<float>(r, c) = (float)g.at(r, c);
    t.at<float>(0, 3) = (float)position.x;
    t.at<float>(1, 3) = (float)position.y;
    t.at<float>(2, 3) = (float)position.z;
    t.at<float>(3, 3) = 1.0f;
    return t;
}

cv::Mat transformationMatrix2rotationMatrix(cv::Mat t) {                      // Plus.cpp:222-241
    cv::Mat r = cv::Mat::zeros(3, 3, CV_32FC1);
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) r.at<float>(i, j) = t.at<float>(i, j);
    return r;
}

cv::Mat point2MatPlusOne(cv::Point3d p) {                                     // Plus.cpp:243-256
    cv::Mat m = cv::Mat::zeros(4, 1, CV_32FC1);
    m.at<float>(0, 0) = (float)p.x; m.at<float>(1, 0) = (float)p.y; m.at<float>(2, 0) = (float)p.z;
    m.at<float>(3, 0) = 1.0f;
    return m;
}

cv::Mat point2Mat(cv::Point3d p) {                                            // Plus.cpp:258-269
    cv::Mat m = cv::Mat::zeros(3, 1, CV_32FC1);
    m.at<float>(0, 0) = (float)p.x; m.at<float>(1, 0) = (float)p.y; m.at<float>(2, 0) = (float)p.z;
    return m;
}

cv::Point3d Mat2point(cv::Mat m) {                                            // Plus.cpp:271-282
    return cv::Point3d(m.at<float>(0, 0), m.at<float>(1, 0), m.at<float>(2, 0));
}

cv::Mat RPYWorld2ResidualAngImu(cv::Point3d rpy) {                            // Plus.cpp:325-351
    const double c1 = std::cos(rpy.x), s1 = std::sin(rpy.x), c2 = std::cos(rpy.y), s2 = std::sin(rpy.y);
    cv::Mat m = cv::Mat::zeros(3, 3, CV_32FC1);
    m.at<float>(0, 0) = 1.0f;
    m.at<float>(0, 1) = (float)(s2 * s1 / c1);
    m.at<float>(0, 2) = (float)(c2 * s1 / c1);
    m.at<float>(1, 1) = (float)c2;
    m.at<float>(1, 2) = (float)(-s2);
    m.at<float>(2, 1) = (float)(s2 / c1);
    m.at<float>(2, 2) = (float)(c2 / c1);
    return m;
}
