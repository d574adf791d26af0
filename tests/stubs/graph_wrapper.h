// TEST STAND-IN for the reference header src/graph_wrapper.h: see reference_stubs.h
#include "reference_stubs.h"
