// CameraGPU.hpp — forwarding header with the reference's file name (include/CameraGPU.hpp); the class lives in vislam/Camera.hpp.
#ifndef VISLAM_REFNAMES_CameraGPU_HPP_
#define VISLAM_REFNAMES_CameraGPU_HPP_
#include "vislam/Camera.hpp"
using namespace cv;
using namespace std;
#endif
