// include/alpaka/alpaka.hpp -- umbrella header of the B200-native alpaka API layer.
//
// User code written against the reference (`#include <alpaka/alpaka.hpp>`, reference: include/alpaka/alpaka.hpp)
// compiles against this header with nvcc (>= 12.8, -std=c++20 --expt-relaxed-constexpr --extended-lambda
// -gencode arch=compute_100a,code=sm_100a -cudart shared) and links libalpaka_b200.so. One accelerator exists:
// AccGpuB200<TDim,TIdx> / TagGpuB200 (also reachable under the reference's AccGpuCudaRt / TagGpuCudaRt names).
// See INTEGRATION.md for the build line and DESIGN.md for the layering:
//
//   user kernels + drivers            (unchanged alpaka source)
//   include/alpaka/b200/*.hpp         C++20 API layer: Vec, WorkDiv, Dev/Queue/Event, Buf/memcpy, Acc, Kernel
//   include/b200/b200.h               C ABI of libalpaka_b200.so (runtime + hand-written sm_100a kernels)
#pragma once

#include "b200/Config.hpp"
#include "b200/Vec.hpp"
#include "b200/Tags.hpp"
#include "b200/Dev.hpp"
#include "b200/Mem.hpp"
#include "b200/WorkDiv.hpp"
#include "b200/Acc.hpp"
#include "b200/Global.hpp"
#include "b200/Exec.hpp"
#include "b200/Kernel.hpp"
#include "b200/Native.hpp"
#include "b200/Heat2D.hpp"
#include "b200/Meta.hpp"
