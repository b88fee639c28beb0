// alpaka::executeForEachAccTag lives in alpaka/b200/Tags.hpp; this header exists because reference drivers include it
// by this path (example/heatEquation2D/src/heatEquation2D.cpp).
#pragma once
#include <alpaka/alpaka.hpp>
