// VISystemGPU.hpp — forwarding header with the reference's file name (include/VISystemGPU.hpp); the class lives in vislam/VISystem.hpp.
#ifndef VISLAM_REFNAMES_VISystemGPU_HPP_
#define VISLAM_REFNAMES_VISystemGPU_HPP_
#include "vislam/VISystem.hpp"
using namespace cv;
using namespace std;
#endif
