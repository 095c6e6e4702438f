// oracle/refshim/sophus/sim3.hpp — include/Options.hpp names Sophus::Sim3f in one typedef and never uses it; the real
// sim3.hpp would pull in rxso3.hpp and more of Eigen than oracle/eigenshim supplies.  se3.hpp / so3.hpp are NOT shimmed:
// the reference's own thirdparty/sophus headers are compiled (oracle/Makefile puts $(REFROOT)/thirdparty on the path).
#ifndef VSO_REFSHIM_SOPHUS_SIM3_HPP
#define VSO_REFSHIM_SOPHUS_SIM3_HPP
namespace Sophus { template <typename T> class Sim3 { public: typedef T Scalar; }; typedef Sim3<float> Sim3f; }
#endif
