// oracle/refshim/sophus/se3.hpp — Sophus::SE3f stand-in (TEST INFRASTRUCTURE ONLY).
// The reference vendors Sophus (thirdparty/sophus) but not the Eigen it is built on, and Eigen is not in this image, so
// the vendored headers cannot be compiled.  This class offers the handful of members VISystem.cpp / Camera.hpp use and
// forwards the arithmetic to the oracle's restatement of those Sophus/Eigen routines (oracle/visystem.c: vso_se3_exp,
// vso_se3_mul, vso_se3_matrix, vso_rot_to_quat — each citing se3.hpp / so3.hpp lines, and checked against scipy in
// tests/test_oracle_props.py).  Consequently the reference-vs-oracle comparison pins everything in
// VISystem::EstimatePoseFeatures EXCEPT the SE3 exp / compose arithmetic itself, which both sides share.
#ifndef VSO_REFSHIM_SOPHUS_SE3_HPP
#define VSO_REFSHIM_SOPHUS_SE3_HPP
#include "Eigen/Core"
extern "C" {
void vso_se3_exp(const float delta[6], float pose[7]);
void vso_se3_mul(const float a[7], const float b[7], float out[7]);
void vso_se3_matrix(const float pose[7], float m[16]);
void vso_rot_to_quat(const float r[9], float q[4]);
}
namespace Sophus {
template <typename T, int N> using Vector = Eigen::Matrix<T, N, 1>;
template <typename T>
class SO3 {
public:
    typedef T Scalar;
};
template <typename T>
class SE3 {
public:
    typedef T Scalar;
    typedef Eigen::Matrix<T, 3, 1> Point;
    typedef Eigen::Matrix<T, 6, 1> Tangent;
    typedef Eigen::Matrix<T, 4, 4> Transformation;
    static const int DoF = 6;
    float p[7];                                      // qx qy qz qw tx ty tz
    SE3() { p[0] = p[1] = p[2] = 0; p[3] = 1; p[4] = p[5] = p[6] = 0; }
    SE3(const Eigen::Matrix<T, 3, 3>& R, const Point& t) {
        float r[9];
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r[3 * i + j] = R(i, j);
        vso_rot_to_quat(r, p);
        p[4] = t(0); p[5] = t(1); p[6] = t(2);
    }
    SE3(const Eigen::Quaternion<T>& q, const Point& t) {
        p[0] = q.x(); p[1] = q.y(); p[2] = q.z(); p[3] = q.w();
        p[4] = t(0); p[5] = t(1); p[6] = t(2);
    }
    static SE3 exp(const Tangent& a) {
        float d[6];
        for (int i = 0; i < 6; i++) d[i] = a(i);
        SE3 s;
        vso_se3_exp(d, s.p);
        return s;
    }
    SE3 operator*(const SE3& o) const { SE3 s; vso_se3_mul(p, o.p, s.p); return s; }
    Transformation matrix() const {
        float m[16];
        vso_se3_matrix(p, m);
        Transformation t;
        for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) t(i, j) = m[4 * i + j];
        return t;
    }
    Point translation() const { return Point(p[4], p[5], p[6]); }
    Eigen::Quaternion<T> unit_quaternion() const { return Eigen::Quaternion<T>(p[3], p[0], p[1], p[2]); }
};
typedef SE3<float> SE3f;
typedef SO3<float> SO3f;
}  // namespace Sophus
#endif
