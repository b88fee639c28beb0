// alpaka::omp::Schedule and trait::OmpSchedule (reference: include/alpaka/core/OmpSchedule.hpp, kernel/Traits.hpp:126-161).
// The B200 back-end has no OpenMP loop to schedule: the type and the trait exist so that kernels carrying an
// `ompScheduleKind` member or a trait::OmpSchedule specialisation keep compiling; the setting is ignored.
#pragma once

#include <alpaka/alpaka.hpp>

namespace alpaka::omp
{
    struct Schedule
    {
        enum Kind
        {
            NoSchedule,
            Static = 1u,
            Dynamic = 2u,
            Guided = 3u,
            Auto = 4u,
            Runtime = 5u
        };
        Kind kind;
        int chunkSize;
        ALPAKA_FN_HOST constexpr Schedule(Kind myKind = NoSchedule, int myChunkSize = 0) : kind(myKind), chunkSize(myChunkSize)
        {
        }
    };
    ALPAKA_FN_HOST inline auto getSchedule()
    {
        return Schedule{};
    }
    ALPAKA_FN_HOST inline void setSchedule(Schedule)
    {
    }
} // namespace alpaka::omp

namespace alpaka::trait
{
    template<typename TKernelFnObj, typename TAcc, typename TSfinae = void>
    struct OmpSchedule
    {
        struct TraitNotSpecialized
        {
        };
        template<typename TDim, typename... TArgs>
        ALPAKA_FN_HOST static auto getOmpSchedule(
            TKernelFnObj const&,
            Vec<TDim, Idx<TAcc>> const&,
            Vec<TDim, Idx<TAcc>> const&,
            TArgs const&...) -> TraitNotSpecialized
        {
            return TraitNotSpecialized{};
        }
    };
} // namespace alpaka::trait

namespace alpaka
{
    template<typename TAcc, typename TKernelFnObj, typename TDim, typename... TArgs>
    ALPAKA_FN_HOST auto getOmpSchedule(
        TKernelFnObj const& kernelFnObj,
        Vec<TDim, Idx<TAcc>> const& blockThreadExtent,
        Vec<TDim, Idx<TAcc>> const& threadElemExtent,
        TArgs const&... args)
    {
        return trait::OmpSchedule<TKernelFnObj, TAcc>::getOmpSchedule(kernelFnObj, blockThreadExtent, threadElemExtent, args...);
    }
} // namespace alpaka
