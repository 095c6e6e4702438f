// Options.hpp — the reference's type vocabulary (include/Options.hpp:53-75): vi::SE3 is declared in vislam/compat.hpp.
#ifndef VISLAM_REFNAMES_OPTIONS_HPP_
#define VISLAM_REFNAMES_OPTIONS_HPP_
#include "vislam/compat.hpp"
#endif
