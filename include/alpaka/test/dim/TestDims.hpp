// The dimensionalities the test suite instantiates (reference: include/alpaka/test/dim/TestDims.hpp; CUDA grids are 3-D).
#pragma once
#include <alpaka/test/acc/TestAccs.hpp>
