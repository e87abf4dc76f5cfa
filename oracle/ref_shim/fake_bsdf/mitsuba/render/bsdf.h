// TEST INFRASTRUCTURE ONLY: stands in for <mitsuba/render/bsdf.h> when the reference BSDF plugins are compiled for oracle/_ref (see ../../mitsuba_shim.h)
#include "mitsuba_shim.h"
