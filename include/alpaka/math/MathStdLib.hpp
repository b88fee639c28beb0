// Compatibility path (reference: math/MathStdLib.hpp). Host-side std math needs no accelerator here.
#pragma once
#include <alpaka/alpaka.hpp>
