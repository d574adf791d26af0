// TEST STAND-IN for the reference header src/se2_compatibility.h: see reference_stubs.h
#include "reference_stubs.h"
