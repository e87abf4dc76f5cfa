// oracle/ref_shim/fake_include/mitsuba/mitsuba.h -- TEST INFRASTRUCTURE ONLY.
// Stand-in for the reference's <mitsuba/mitsuba.h> (which needs boost) so that
// /root/reference/src/emitters/sunsky/skymodel.cpp compiles unmodified from where it lies.
// In the reference's single-precision build M_PI is the fp32 literal
// (include/mitsuba/core/constants.h:42-44,63,80), which skymodel.cpp then promotes to double.
#pragma once
#include <cmath>
#include <cassert>
#include <cstdlib>
#ifdef M_PI
#undef M_PI
#endif
#define M_PI 3.14159265358979323846f
