// The index types the test suite instantiates (reference: include/alpaka/test/idx/TestIdxs.hpp).
#pragma once
#include <alpaka/test/acc/TestAccs.hpp>
