// Imu.hpp — forwarding header with the reference's file name (include/Imu.hpp): a caller written against the reference
// includes "Imu.hpp" and gets the B200 class mirror.  Like the reference's headers, it opens cv and std.
#ifndef VISLAM_REFNAMES_Imu_HPP_
#define VISLAM_REFNAMES_Imu_HPP_
#include "vislam/Imu.hpp"
using namespace cv;
using namespace std;
#endif
