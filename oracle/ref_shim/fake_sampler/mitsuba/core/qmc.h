#include "sampler_shim.h"   /* TEST INFRASTRUCTURE: forwards to the stand-ins */
