// TEST STAND-IN for the reference header src/multi_edge_correlated.h: see reference_stubs.h
#include "reference_stubs.h"
