// TEST STAND-IN for the reference header src/glc_reparam_se3.h: see reference_stubs.h
#include "reference_stubs.h"
