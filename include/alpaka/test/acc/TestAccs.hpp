// include/alpaka/test/acc/TestAccs.hpp -- the accelerator type lists test and benchmark drivers iterate over.
//
// API parity with the reference's test/acc/TestAccs.hpp:104-182: EnabledAccs<TDim,TIdx> is a std::tuple of the
// accelerators enabled in this build -- here exactly one, AccGpuB200<TDim,TIdx> -- and TestAccs is the list of
// fully instantiated accelerators over the test dimensions and index types (reference: test/dim/TestDims.hpp,
// test/idx/TestIdxs.hpp; CUDA grids stop at 3 dimensions).
#pragma once

#include <alpaka/alpaka.hpp>

#include <iosfwd>
#include <tuple>

namespace alpaka::test
{
    namespace detail
    {
        //! The reference's per-back-end aliases "the accelerator if its back-end is enabled, else int"
        //! (test/acc/TestAccs.hpp:29-101), which its tag tests enumerate: one back-end is enabled here.
        template<typename TDim, typename TIdx>
        using AccGpuCudaRtIfAvailableElseInt = AccGpuCudaRt<TDim, TIdx>;
        template<typename TDim, typename TIdx>
        using AccCpuSerialIfAvailableElseInt = int;
        template<typename TDim, typename TIdx>
        using AccCpuThreadsIfAvailableElseInt = int;
        template<typename TDim, typename TIdx>
        using AccCpuTbbIfAvailableElseInt = int;
        template<typename TDim, typename TIdx>
        using AccCpuOmp2BlocksIfAvailableElseInt = int;
        template<typename TDim, typename TIdx>
        using AccCpuOmp2ThreadsIfAvailableElseInt = int;
        template<typename TDim, typename TIdx>
        using AccGpuHipRtIfAvailableElseInt = int;
        template<typename TDim, typename TIdx>
        using AccCpuSyclIfAvailableElseInt = int;
        template<typename TDim, typename TIdx>
        using AccFpgaSyclIntelIfAvailableElseInt = int;
        template<typename TDim, typename TIdx>
        using AccGpuSyclIntelIfAvailableElseInt = int;
    } // namespace detail

    //! A std::tuple containing all enabled accelerators for the given dimensionality and index type.
    template<typename TDim, typename TIdx>
    using EnabledAccs = std::tuple<AccGpuB200<TDim, TIdx>>;

    //! Writes the enabled accelerators to the given stream.
    template<typename TDim, typename TIdx>
    auto writeEnabledAccs(std::ostream& os) -> void
    {
        os << "Accelerators enabled: " << getAccName<AccGpuB200<TDim, TIdx>>() << " " << std::endl;
    }

    using TestDims = std::tuple<DimInt<0u>, DimInt<1u>, DimInt<2u>, DimInt<3u>>;
    using NonZeroTestDims = std::tuple<DimInt<1u>, DimInt<2u>, DimInt<3u>>;
    using TestIdxs = std::tuple<int, std::uint32_t, std::size_t, std::int64_t>;

    //! every enabled accelerator x {1,2,3} dimensions x index types
    using TestAccs = std::tuple<
        AccGpuB200<DimInt<1u>, int>,
        AccGpuB200<DimInt<2u>, int>,
        AccGpuB200<DimInt<3u>, int>,
        AccGpuB200<DimInt<1u>, std::uint32_t>,
        AccGpuB200<DimInt<2u>, std::uint32_t>,
        AccGpuB200<DimInt<3u>, std::uint32_t>,
        AccGpuB200<DimInt<1u>, std::size_t>,
        AccGpuB200<DimInt<2u>, std::size_t>,
        AccGpuB200<DimInt<3u>, std::size_t>,
        AccGpuB200<DimInt<1u>, std::int64_t>,
        AccGpuB200<DimInt<2u>, std::int64_t>,
        AccGpuB200<DimInt<3u>, std::int64_t>>;
} // namespace alpaka::test
