// Type-list helpers (reference: include/alpaka/meta/TypeListOps.hpp; pinned by test/unit/meta/src/TypeListOpsTest.cpp).
// Front and Contains come from include/alpaka/b200/Meta.hpp through the umbrella header; isList, ToList and ToTuple are
// defined here.
#pragma once
#include <alpaka/alpaka.hpp>

#include <tuple>
#include <type_traits>

namespace alpaka::meta
{
    namespace detail
    {
        template<typename T>
        struct IsListImpl : std::false_type
        {
        };
        template<template<typename...> class TList, typename... Ts>
        struct IsListImpl<TList<Ts...>> : std::true_type
        {
        };
    } // namespace detail

    //! true for any instantiation of a variadic class template taking types only (std::tuple<...>, user type lists)
    template<typename T>
    inline constexpr bool isList = detail::IsListImpl<T>::value;

    //! TListType<Ts...>, or -- when the single argument already is an instantiation of TListType -- that list itself
    template<template<typename...> class TListType, typename... Ts>
    struct ToList
    {
        using type = TListType<Ts...>;
    };
    template<template<typename...> class TListType, typename... Ts>
    struct ToList<TListType, TListType<Ts...>>
    {
        using type = TListType<Ts...>;
    };

    template<typename... Ts>
    using ToTuple = typename ToList<std::tuple, Ts...>::type;
} // namespace alpaka::meta
