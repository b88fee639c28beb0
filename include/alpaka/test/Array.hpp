// alpaka::test::Array<T,N> (reference: include/alpaka/test/Array.hpp): a trivially copyable fixed-size array usable as
// a block-shared variable type in test kernels.
#pragma once

#include <alpaka/alpaka.hpp>

#include <cstddef>

namespace alpaka::test
{
    template<typename TType, std::size_t TSize>
    struct Array
    {
        TType m_data[TSize];

        template<typename TIdx>
        ALPAKA_FN_HOST_ACC auto operator[](TIdx const idx) const -> TType const&
        {
            return m_data[idx];
        }
        template<typename TIdx>
        ALPAKA_FN_HOST_ACC auto operator[](TIdx const idx) -> TType&
        {
            return m_data[idx];
        }
    };
} // namespace alpaka::test
