#include "se3.hpp"
namespace Sophus { template <typename T> class Sim3 { public: typedef T Scalar; }; typedef Sim3<float> Sim3f; }
