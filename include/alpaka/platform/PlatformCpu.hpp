// Compatibility path: the reference splits its API over many headers and user code includes some of them directly
// (here: <alpaka/platform/PlatformCpu.hpp>). In this implementation the whole API comes from the umbrella header.
#pragma once
#include <alpaka/alpaka.hpp>
