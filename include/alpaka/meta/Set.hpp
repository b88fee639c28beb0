// alpaka::meta::IsSet<TList> (reference: include/alpaka/meta/Set.hpp; pinned by test/unit/meta/src/SetTest.cpp): no type
// occurs twice in the list.
#pragma once
#include <alpaka/alpaka.hpp>

#include <type_traits>

namespace alpaka::meta
{
    template<typename TList>
    struct IsSet : std::is_same<TList, Unique<TList>>
    {
    };
} // namespace alpaka::meta
