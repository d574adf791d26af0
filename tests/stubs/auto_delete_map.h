// TEST STAND-IN for the reference header src/auto_delete_map.h: see reference_stubs.h
#include "reference_stubs.h"
