// oracle/refshim/ceres/ceres.h — compile-only stand-in (TEST INFRASTRUCTURE): the reference's Options.hpp specialises
// Eigen::internal::cast_impl for ceres::Jet (Options.hpp:31-42); nothing on the tracked path uses ceres.
#ifndef VSO_REFSHIM_CERES_H
#define VSO_REFSHIM_CERES_H
namespace ceres {
template <typename T, int N> struct Jet { T a; T v[N]; };
}
#endif
