// alpaka::meta integral-type relations (reference: include/alpaka/meta/Integral.hpp; pinned by
// test/unit/meta/src/IntegralTest.cpp). Used by the reference for extent / index type deduction; nothing on the B200
// hot path needs them, they exist so that user code and the reference's own meta tests keep compiling.
//   IsIntegralSuperset<TSuper, TSub>  TSuper is at least as wide as TSub when both have the same signedness, strictly
//       wider when they differ -- in EITHER direction, as the reference's test pins it (uint64 counts as a superset
//       of int32); false for non-integral types
//   HigherMax / LowerMax / HigherMin / LowerMin<T, U>  the type with the higher / lower maximum / minimum; when the two
//       limits are equal (two unsigned types both have minimum 0) the WIDER type, and T if they are equally wide
#pragma once
#include <alpaka/alpaka.hpp>

#include <limits>
#include <type_traits>

namespace alpaka::meta
{
    template<typename TSuper, typename TSub>
    struct IsIntegralSuperset
        : std::bool_constant<
              std::is_integral_v<TSuper> && std::is_integral_v<TSub>
              && ((std::is_signed_v<TSuper> == std::is_signed_v<TSub> && sizeof(TSuper) >= sizeof(TSub))
                  || (std::is_signed_v<TSuper> != std::is_signed_v<TSub> && sizeof(TSuper) > sizeof(TSub)))>
    {
    };

    namespace detail
    {
        //! -1: T wins, +1: U wins; `higher` selects the larger limit, ties go to the wider type, then to T
        template<typename T, typename U, bool kMax, bool kHigher>
        constexpr bool pickSecond()
        {
            static_assert(std::is_integral_v<T> && std::is_integral_v<U>);
            if constexpr(kMax)
            {
                auto const t = static_cast<unsigned long long>(std::numeric_limits<T>::max());
                auto const u = static_cast<unsigned long long>(std::numeric_limits<U>::max());
                if(t != u)
                    return kHigher ? (u > t) : (u < t);
            }
            else
            {
                auto const t = static_cast<long long>(std::numeric_limits<T>::min());
                auto const u = static_cast<long long>(std::numeric_limits<U>::min());
                if(t != u)
                    return kHigher ? (u > t) : (u < t);
            }
            return sizeof(U) > sizeof(T);
        }
    } // namespace detail

    template<typename T, typename U>
    using HigherMax = std::conditional_t<detail::pickSecond<T, U, true, true>(), U, T>;
    template<typename T, typename U>
    using LowerMax = std::conditional_t<detail::pickSecond<T, U, true, false>(), U, T>;
    template<typename T, typename U>
    using HigherMin = std::conditional_t<detail::pickSecond<T, U, false, true>(), U, T>;
    template<typename T, typename U>
    using LowerMin = std::conditional_t<detail::pickSecond<T, U, false, false>(), U, T>;
} // namespace alpaka::meta
