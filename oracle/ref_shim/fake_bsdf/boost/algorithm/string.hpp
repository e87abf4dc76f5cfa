// TEST INFRASTRUCTURE ONLY: stands in for <boost/algorithm/string.hpp> (to_lower_copy lives in ../../mitsuba_shim.h)
#include "mitsuba_shim.h"
