// Plus.hpp — forwarding header with the reference's file name (include/Plus.hpp): a caller written against the reference
// includes "Plus.hpp" and gets the B200 class mirror.  Like the reference's headers, it opens cv and std.
#ifndef VISLAM_REFNAMES_Plus_HPP_
#define VISLAM_REFNAMES_Plus_HPP_
#include "vislam/Plus.hpp"
using namespace cv;
using namespace std;
#endif
