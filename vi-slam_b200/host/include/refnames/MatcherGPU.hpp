// MatcherGPU.hpp — forwarding header with the reference's file name (include/MatcherGPU.hpp); the class lives in vislam/Matcher.hpp.
#ifndef VISLAM_REFNAMES_MatcherGPU_HPP_
#define VISLAM_REFNAMES_MatcherGPU_HPP_
#include "vislam/Matcher.hpp"
using namespace cv;
using namespace std;
#endif
