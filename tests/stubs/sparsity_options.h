// TEST STAND-IN for the reference header src/sparsity_options.h: see reference_stubs.h
#include "reference_stubs.h"
