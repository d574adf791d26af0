// TEST STAND-IN for the reference header src/topology_provider_glc.h: see reference_stubs.h
#include "reference_stubs.h"
