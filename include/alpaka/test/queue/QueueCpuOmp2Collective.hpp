// include/alpaka/test/queue/QueueCpuOmp2Collective.hpp -- the reference's collective OpenMP host queue
// (include/alpaka/test/queue/QueueCpuOmp2Collective.hpp) belongs to its AccCpuOmp2Blocks back-end. This framework has
// no CPU accelerator, ALPAKA_ACC_CPU_B_OMP2_T_SEQ_ENABLED is never defined, and the tests that use the type are
// compiled out by that macro; the header exists so that their include line resolves.
#pragma once

#include <alpaka/alpaka.hpp>
