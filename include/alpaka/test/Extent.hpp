// Canonical N-d test extents (reference: include/alpaka/test/Extent.hpp): buffer extent (11, 10, 9, ...), sub-view
// extent (8, 6, 4, ...) and offset (2, 3, 4, ...), slowest dimension first.
#pragma once

#include <alpaka/alpaka.hpp>

namespace alpaka::test
{
    namespace detail
    {
        template<typename TDim, typename TVal>
        constexpr auto ramp(int first, int step) -> Vec<TDim, TVal>
        {
            Vec<TDim, TVal> v;
            for(std::size_t i = 0; i < TDim::value; ++i)
                v[i] = static_cast<TVal>(first + step * static_cast<int>(i));
            return v;
        }
    } // namespace detail

    template<typename TDim, typename TVal>
    inline constexpr auto extentBuf = detail::ramp<TDim, TVal>(11, -1);
    template<typename TDim, typename TVal>
    inline constexpr auto extentSubView = detail::ramp<TDim, TVal>(8, -2);
    template<typename TDim, typename TVal>
    inline constexpr auto offset = detail::ramp<TDim, TVal>(2, 1);
} // namespace alpaka::test
