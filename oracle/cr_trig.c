/* oracle/cr_trig.c — correctly rounded single-precision sin / cos for the reference-derived checker libraries (TEST
 * INFRASTRUCTURE ONLY).  The reference's vendored Sophus calls std::sin / std::cos on floats (so3.hpp:541-560,
 * se3.hpp:724-738), i.e. the platform libm's sinf / cosf — glibc's differ from the correctly rounded value in about 1.3 %
 * of arguments (by one unit in the last place), other libms differently, so "the reference's bits" depend on the machine
 * it was built on.  The oracle and the kernels define sin / cos of a float as the float rounding of the double function;
 * linking this file (with -Wl,-Bsymbolic-functions) makes the reference-derived libraries evaluate them the same way, so
 * that everything ELSE in Sophus can be compared bit for bit.  libref_sophus_libm.so is the same code on the native libm. */
#include <math.h>
float sinf(float x) { return (float)sin((double)x); }
float cosf(float x) { return (float)cos((double)x); }
void sincosf(float x, float* s, float* c) { *s = (float)sin((double)x); *c = (float)cos((double)x); }
