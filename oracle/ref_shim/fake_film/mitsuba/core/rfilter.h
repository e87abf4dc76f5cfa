// TEST INFRASTRUCTURE ONLY: stands in for <mitsuba/core/rfilter.h> when src/libcore/rfilter.cpp and src/rfilters/{tent,box,gaussian}.cpp are
// compiled unmodified for oracle/_ref.  The class declaration itself is the reference's (include/mitsuba/core/rfilter.h:30-80, cut out at
// build time into oracle/_ref/ref_rfilter_class.inc); only its base class comes from the scaffolding in ../../../fake_bsdf/mitsuba_shim.h.
#pragma once
#include "mitsuba_shim.h"
namespace mitsuba {
#define MTS_FILTER_RESOLUTION 31
#include "ref_rfilter_class.inc"
}
