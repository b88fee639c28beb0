// getExtents / getWidth / getHeight / getExtentProduct live in alpaka/b200/Vec.hpp; this header exists because
// reference code includes it by this path (example/heatEquation2D/src/analyticalSolution.hpp).
#pragma once
#include <alpaka/alpaka.hpp>
