// alpaka::meta::IsArrayOrVector<T> (reference: include/alpaka/meta/IsArrayOrVector.hpp; pinned by
// test/unit/meta/src/IsArrayOrVectorTest.cpp): true for C arrays, std::array, std::vector and alpaka::Vec.
#pragma once
#include <alpaka/alpaka.hpp>

#include <array>
#include <type_traits>
#include <vector>

namespace alpaka::meta
{
    namespace detail
    {
        template<typename T>
        struct IsArrayOrVectorImpl : std::is_array<T>
        {
        };
        template<typename T, std::size_t N>
        struct IsArrayOrVectorImpl<std::array<T, N>> : std::true_type
        {
        };
        template<typename T, typename A>
        struct IsArrayOrVectorImpl<std::vector<T, A>> : std::true_type
        {
        };
        template<typename TDim, typename TVal>
        struct IsArrayOrVectorImpl<alpaka::Vec<TDim, TVal>> : std::true_type
        {
        };
    } // namespace detail
    template<typename T>
    struct IsArrayOrVector : detail::IsArrayOrVectorImpl<std::remove_cv_t<std::remove_reference_t<T>>>
    {
    };
} // namespace alpaka::meta
