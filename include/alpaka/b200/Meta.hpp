// include/alpaka/b200/Meta.hpp -- the handful of type-list utilities user and test code reaches for.
//
// API parity with the reference's meta/ForEachType.hpp (forEachType), meta/Concatenate.hpp, meta/Filter.hpp,
// meta/Transform.hpp, meta/Apply.hpp, meta/CartesianProduct.hpp, meta/Unique.hpp, meta/TypeListOps.hpp (Front,
// Contains), meta/NdLoop.hpp (ndLoopIncIdx), meta/IsStrictBase.hpp, meta/InheritFromList.hpp. Lists are any variadic class template
// (std::tuple in practice). Written fresh with C++20 fold expressions.
#pragma once

#include "Vec.hpp"

#include <tuple>
#include <type_traits>
#include <utility>

namespace alpaka::meta
{
    // ---- forEachType<List>(f, args...): calls f.template operator()<T>(args...) for every T of the list
    namespace detail
    {
        template<typename TList>
        struct ForEachTypeHelper;
        template<template<typename...> class TList, typename... Ts>
        struct ForEachTypeHelper<TList<Ts...>>
        {
            ALPAKA_NO_HOST_ACC_WARNING
            template<typename TFnObj, typename... TArgs>
            ALPAKA_FN_HOST_ACC static auto forEachTypeHelper(TFnObj&& f, TArgs&&... args) -> void
            {
                (f.template operator()<Ts>(std::forward<TArgs>(args)...), ...);
            }
        };
    } // namespace detail

    ALPAKA_NO_HOST_ACC_WARNING
    template<typename TList, typename TFnObj, typename... TArgs>
    ALPAKA_FN_HOST_ACC auto forEachType(TFnObj&& f, TArgs&&... args) -> void
    {
        detail::ForEachTypeHelper<TList>::forEachTypeHelper(std::forward<TFnObj>(f), std::forward<TArgs>(args)...);
    }

    // ---- Concatenate<List...>
    namespace detail
    {
        template<typename... TLists>
        struct ConcatenateImpl;
        template<typename TList>
        struct ConcatenateImpl<TList>
        {
            using type = TList;
        };
        template<template<typename...> class TList, typename... As, typename... Bs, typename... TRest>
        struct ConcatenateImpl<TList<As...>, TList<Bs...>, TRest...>
        {
            using type = typename ConcatenateImpl<TList<As..., Bs...>, TRest...>::type;
        };
    } // namespace detail
    template<typename... TLists>
    using Concatenate = typename detail::ConcatenateImpl<TLists...>::type;

    // ---- Filter<List, Pred>
    namespace detail
    {
        template<template<typename...> class TList, template<typename> class TPred, typename... Ts>
        struct FilterImplHelper;
        template<template<typename...> class TList, template<typename> class TPred>
        struct FilterImplHelper<TList, TPred>
        {
            using type = TList<>;
        };
        template<template<typename...> class TList, template<typename> class TPred, typename T, typename... Ts>
        struct FilterImplHelper<TList, TPred, T, Ts...>
        {
            using type = std::conditional_t<
                TPred<T>::value,
                Concatenate<TList<T>, typename FilterImplHelper<TList, TPred, Ts...>::type>,
                typename FilterImplHelper<TList, TPred, Ts...>::type>;
        };
        template<typename TList, template<typename> class TPred>
        struct FilterImpl;
        template<template<typename...> class TList, template<typename> class TPred, typename... Ts>
        struct FilterImpl<TList<Ts...>, TPred>
        {
            using type = typename FilterImplHelper<TList, TPred, Ts...>::type;
        };
    } // namespace detail
    template<typename TList, template<typename> class TPred>
    using Filter = typename detail::FilterImpl<TList, TPred>::type;

    // ---- Transform<List, Op>, Apply<List, Applicee>
    namespace detail
    {
        template<typename TList, template<typename> class TOp>
        struct TransformImpl;
        template<template<typename...> class TList, typename... Ts, template<typename> class TOp>
        struct TransformImpl<TList<Ts...>, TOp>
        {
            using type = TList<TOp<Ts>...>;
        };
        template<typename TList, template<typename...> class TApplicee>
        struct ApplyImpl;
        template<template<typename...> class TList, template<typename...> class TApplicee, typename... Ts>
        struct ApplyImpl<TList<Ts...>, TApplicee>
        {
            using type = TApplicee<Ts...>;
        };
    } // namespace detail
    template<typename TList, template<typename> class TOp>
    using Transform = typename detail::TransformImpl<TList, TOp>::type;
    template<typename TList, template<typename...> class TApplicee>
    using Apply = typename detail::ApplyImpl<TList, TApplicee>::type;

    // ---- CartesianProduct<List, Lists...>: List<List<a,b,...>...>, first list varies FASTEST (the order the reference's
    //      test/unit/meta/src/CartesianProductTest.cpp:20-29 pins)
    namespace detail
    {
        template<template<typename...> class TList, typename TPrefixes, typename... TLists>
        struct CartesianImpl;
        template<template<typename...> class TList, typename... TPrefixes>
        struct CartesianImpl<TList, TList<TPrefixes...>>
        {
            using type = TList<TPrefixes...>;
        };
        template<typename TPrefix, typename T>
        struct Append;
        template<template<typename...> class TList, typename... Ps, typename T>
        struct Append<TList<Ps...>, T>
        {
            using type = TList<Ps..., T>;
        };
        template<template<typename...> class TList, typename... TPrefixes, typename... Ts, typename... TRest>
        struct CartesianImpl<TList, TList<TPrefixes...>, TList<Ts...>, TRest...>
        {
            template<typename T>
            using Expand = TList<typename Append<TPrefixes, T>::type...>;
            using type = typename CartesianImpl<TList, Concatenate<Expand<Ts>...>, TRest...>::type;
        };
    } // namespace detail
    template<template<typename...> class TList, typename... TLists>
    using CartesianProduct = typename detail::CartesianImpl<TList, TList<TList<>>, TLists...>::type;

    // ---- Unique<List>, Contains<List, T>, Front<List>
    template<typename TList, typename T>
    struct Contains;
    template<template<typename...> class TList, typename... Ts, typename T>
    struct Contains<TList<Ts...>, T> : std::bool_constant<(std::is_same_v<Ts, T> || ...)>
    {
    };
    namespace detail
    {
        template<typename TDone, typename TTodo>
        struct UniqueImpl;
        template<template<typename...> class TList, typename... Ds>
        struct UniqueImpl<TList<Ds...>, TList<>>
        {
            using type = TList<Ds...>;
        };
        template<template<typename...> class TList, typename... Ds, typename T, typename... Ts>
        struct UniqueImpl<TList<Ds...>, TList<T, Ts...>>
        {
            using type = typename UniqueImpl<
                std::conditional_t<(std::is_same_v<Ds, T> || ...), TList<Ds...>, TList<Ds..., T>>,
                TList<Ts...>>::type;
        };
        template<typename TList>
        struct EmptyOf;
        template<template<typename...> class TList, typename... Ts>
        struct EmptyOf<TList<Ts...>>
        {
            using type = TList<>;
        };
        template<typename TList>
        struct FrontImpl;
        template<template<typename...> class TList, typename T, typename... Ts>
        struct FrontImpl<TList<T, Ts...>>
        {
            using type = T;
        };
    } // namespace detail
    template<typename TList>
    using Unique = typename detail::UniqueImpl<typename detail::EmptyOf<TList>::type, TList>::type;
    template<typename TList>
    using Front = typename detail::FrontImpl<TList>::type;

    template<typename TBase, typename TDerived>
    using IsStrictBase = std::bool_constant<
        std::is_base_of_v<TBase, TDerived> && !std::is_same_v<TBase, std::decay_t<TDerived>>>;

    template<typename TList>
    class InheritFromList;
    template<template<typename...> class TList, typename... TBases>
    class InheritFromList<TList<TBases...>> : public TBases...
    {
    };

    // ---- ndLoopIncIdx(extent, f): calls f(idx) for every index of the N-d extent, slowest dimension outermost
    ALPAKA_NO_HOST_ACC_WARNING
    template<typename TExtentVec, typename TFnObj>
    ALPAKA_FN_HOST_ACC auto ndLoopIncIdx(TExtentVec const& extent, TFnObj const& f) -> void
    {
        constexpr std::size_t n = TExtentVec::size();
        using V = TExtentVec;
        if constexpr(n == 0u)
        {
            f(V{});
        }
        else
        {
            auto const total = extent.prod();
            for(std::decay_t<decltype(total)> lin = 0; lin < total; ++lin)
            {
                V idx;
                auto rest = lin;
                for(std::size_t d = n; d-- > 0u;)
                {
                    idx[d] = static_cast<typename V::value_type>(rest % extent[d]);
                    rest = static_cast<decltype(rest)>(rest / extent[d]);
                }
                f(idx);
            }
        }
    }
    template<typename TExtentVec, typename TFnObj>
    ALPAKA_FN_HOST_ACC auto ndLoop(std::index_sequence<>, TExtentVec const& extent, TFnObj const& f) -> void
    {
        ndLoopIncIdx(extent, f);
    }
} // namespace alpaka::meta
