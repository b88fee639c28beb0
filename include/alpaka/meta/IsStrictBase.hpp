// Compatibility path: alpaka::meta::IsStrictBase lives in include/alpaka/b200/Meta.hpp (reference:
// include/alpaka/meta/IsStrictBase.hpp; pinned by test/unit/meta/src/IsStrictBaseTest.cpp).
#pragma once
#include <alpaka/alpaka.hpp>
