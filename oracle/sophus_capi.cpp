// oracle/sophus_capi.cpp — C entry points around the reference's OWN vendored Sophus (thirdparty/sophus/se3.hpp, so3.hpp,
// compiled unmodified; TEST INFRASTRUCTURE).  Eigen, which Sophus is built on, is neither vendored by the reference nor
// installed here: oracle/eigenshim/Eigen/Core supplies the fixed-size matrix / quaternion operations, each following the
// evaluation order of Eigen's generic (non-vectorised) code path.  So what these entries pin is Sophus' own arithmetic
// and control flow — SO3::expAndTheta with its Taylor branch below epsilon (so3.hpp:534-568), SO3::operator*= with the
// first-order renormalisation (so3.hpp:338-353), SE3::exp's V matrix (se3.hpp:723-742), SE3::operator*= (se3.hpp:285-321),
// SE3::matrix (se3.hpp:253-268), SO3(R) (so3.hpp:422-427) — with Eigen's formulas restated.
// Poses are {qx, qy, qz, qw, tx, ty, tz}.
#include "sophus/se3.hpp"

namespace {
Sophus::SE3f load(const float* p) {
    // the quaternion constructor normalises (so3.hpp:433-440); internal state is restored exactly through the data pointer
    Sophus::SE3f s;
    float* d = s.data();                       // so3 quaternion coefficients x y z w, then the translation
    for (int i = 0; i < 7; i++) d[i] = p[i];
    return s;
}
void store(const Sophus::SE3f& s, float* p) {
    const float* d = s.data();
    for (int i = 0; i < 7; i++) p[i] = d[i];
}
}  // namespace

extern "C" {
void sph_se3_exp(const float delta[6], float pose[7]) {
    Eigen::Matrix<float, 6, 1> a;
    for (int i = 0; i < 6; i++) a(i) = delta[i];
    store(Sophus::SE3f::exp(a), pose);
}
void sph_se3_mul(const float a[7], const float b[7], float out[7]) { store(load(a) * load(b), out); }
void sph_se3_matrix(const float pose[7], float m[16]) {
    const Eigen::Matrix<float, 4, 4> t = load(pose).matrix();
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) m[4 * i + j] = t(i, j);
}
// returns 0, or 1 when Sophus' orthogonality / determinant precondition would abort (so3.hpp:422-427)
int sph_se3_from_rt(const float r[9], const float t[3], float pose[7]) {
    Eigen::Matrix<float, 3, 3> R;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) R(i, j) = r[3 * i + j];
    if (!Sophus::isOrthogonal(R) || !(R.determinant() > 0.f)) return 1;
    store(Sophus::SE3f(R, Eigen::Matrix<float, 3, 1>(t[0], t[1], t[2])), pose);
    return 0;
}
}
