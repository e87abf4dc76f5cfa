// TEST INFRASTRUCTURE ONLY: stands in for <mitsuba/core/fstream.h> when the reference BSDF plugins are compiled for oracle/_ref (see ../../mitsuba_shim.h)
#include "mitsuba_shim.h"
