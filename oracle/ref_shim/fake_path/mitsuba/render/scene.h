// TEST INFRASTRUCTURE ONLY: stands in for <mitsuba/render/scene.h> when src/integrators/path/path.cpp is compiled for oracle/_ref (see ../../path_shim.h)
#include "path_shim.h"
