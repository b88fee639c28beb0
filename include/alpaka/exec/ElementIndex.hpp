// Compatibility path: the reference splits its API over many headers and user code includes some of them directly
// (here: <alpaka/exec/ElementIndex.hpp>). In this implementation the whole API comes from the umbrella header
// (the ranges themselves live in include/alpaka/b200/Exec.hpp).
#pragma once
#include <alpaka/alpaka.hpp>
