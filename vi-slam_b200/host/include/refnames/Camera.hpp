// Camera.hpp — forwarding header with the reference's file name (include/Camera.hpp): a caller written against the reference
// includes "Camera.hpp" and gets the B200 class mirror.  Like the reference's headers, it opens cv and std.
#ifndef VISLAM_REFNAMES_Camera_HPP_
#define VISLAM_REFNAMES_Camera_HPP_
#include "vislam/Camera.hpp"
using namespace cv;
using namespace std;
#endif
