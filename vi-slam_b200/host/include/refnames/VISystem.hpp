// VISystem.hpp — forwarding header with the reference's file name (include/VISystem.hpp): a caller written against the reference
// includes "VISystem.hpp" and gets the B200 class mirror.  Like the reference's headers, it opens cv and std.
#ifndef VISLAM_REFNAMES_VISystem_HPP_
#define VISLAM_REFNAMES_VISystem_HPP_
#include "vislam/VISystem.hpp"
using namespace cv;
using namespace std;
#endif
