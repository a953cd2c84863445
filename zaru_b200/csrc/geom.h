// The f32 rectangle / view algebra lives in include/ so that the C++ host mirror (include/zaru_b200.hpp) and the
// kernels share ONE definition.
#pragma once
#include "../../include/zaru_b200_geom.h"
