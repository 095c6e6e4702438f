// CameraModel.hpp — forwarding header with the reference's file name (include/CameraModel.hpp): a caller written against the reference
// includes "CameraModel.hpp" and gets the B200 class mirror.  Like the reference's headers, it opens cv and std.
#ifndef VISLAM_REFNAMES_CameraModel_HPP_
#define VISLAM_REFNAMES_CameraModel_HPP_
#include "vislam/CameraModel.hpp"
using namespace cv;
using namespace std;
#endif
