// Compatibility path: the reference splits its API over many headers and user code includes some of them directly
// (here: <alpaka/core/DemangleTypeNames.hpp>). In this implementation the whole API comes from the umbrella header; accelerators of other
// back-ends exist as names only.
#pragma once
#include <alpaka/alpaka.hpp>
