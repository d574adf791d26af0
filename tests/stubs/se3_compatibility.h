// TEST STAND-IN for the reference header src/se3_compatibility.h: see reference_stubs.h
#include "reference_stubs.h"
