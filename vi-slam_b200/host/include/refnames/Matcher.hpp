// Matcher.hpp — forwarding header with the reference's file name (include/Matcher.hpp): a caller written against the reference
// includes "Matcher.hpp" and gets the B200 class mirror.  Like the reference's headers, it opens cv and std.
#ifndef VISLAM_REFNAMES_Matcher_HPP_
#define VISLAM_REFNAMES_Matcher_HPP_
#include "vislam/Matcher.hpp"
using namespace cv;
using namespace std;
#endif
