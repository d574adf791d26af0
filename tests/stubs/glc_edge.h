// TEST STAND-IN for the reference header src/glc_edge.h: see reference_stubs.h
#include "reference_stubs.h"
