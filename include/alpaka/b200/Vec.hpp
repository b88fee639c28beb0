// include/alpaka/b200/Vec.hpp -- DimInt, Vec<TDim,TVal>, the Dim/Idx/Elem trait aliases, extent/offset getters and
// index mapping.
//
// API parity with the reference's vec/Vec.hpp:36-800 (constructors, all/ones/zeros, prod/sum/min/max, element-wise
// operators, stream output, structured bindings), dim/DimIntegralConst.hpp, dim/DimArithmetic.hpp:14-18 (arithmetic
// types are 1-D extents), extent/Traits.hpp:59-153, offset/Traits.hpp, idx/MapIdx.hpp:21-97 and core/Utility.hpp:27-62.
// Written fresh; index 0 is the SLOWEST dimension, the last component maps to CUDA x.
#pragma once

#include "Config.hpp"

#include <algorithm>
#include <array>
#include <cmath>
#include <concepts>
#include <functional>
#include <limits>
#include <ostream>
#include <tuple>
#include <type_traits>
#include <utility>

namespace alpaka
{
    template<std::size_t N>
    using DimInt = std::integral_constant<std::size_t, N>;

    namespace trait
    {
        //! customisation points: the dimensionality, index type and element type of T
        template<typename T, typename TSfinae = void>
        struct DimType;
        template<typename T, typename TSfinae = void>
        struct IdxType;
        template<typename T, typename TSfinae = void>
        struct ElemType;

        template<typename T>
        struct DimType<T, std::enable_if_t<std::is_arithmetic_v<T>>>
        {
            using type = DimInt<1u>;
        };

        template<typename T>
        struct IdxType<T, std::enable_if_t<std::is_arithmetic_v<T>>>
        {
            using type = std::decay_t<T>;
        };

        template<typename T>
        struct ElemType<T, std::enable_if_t<std::is_fundamental_v<T>>>
        {
            using type = T;
        };
    } // namespace trait

    template<typename T>
    using Dim = typename trait::DimType<std::remove_cv_t<std::remove_reference_t<T>>>::type;
    template<typename T>
    using Idx = typename trait::IdxType<std::remove_cv_t<std::remove_reference_t<T>>>::type;
    template<typename T>
    using Elem = std::remove_volatile_t<typename trait::ElemType<std::remove_cv_t<std::remove_reference_t<T>>>::type>;

    namespace core
    {
        //! ceil(a / b) for integers
        template<typename T, typename = std::enable_if_t<std::is_integral_v<T>>>
        [[nodiscard]] ALPAKA_FN_HOST_ACC constexpr auto divCeil(T a, T b) -> T
        {
            return (a + b - T{1}) / b;
        }

        //! base^n by squaring
        template<typename T, typename = std::enable_if_t<std::is_integral_v<T>>>
        [[nodiscard]] ALPAKA_FN_HOST_ACC constexpr auto intPow(T base, T n) -> T
        {
            T r{1};
            while(n != 0)
            {
                if(n & T{1})
                    r *= base;
                base *= base;
                n >>= 1;
            }
            return r;
        }

        //! floor(value^(1/n)) by bisection
        template<typename T, typename = std::enable_if_t<std::is_integral_v<T>>>
        [[nodiscard]] ALPAKA_FN_HOST_ACC constexpr auto nthRootFloor(T value, T n) -> T
        {
            T lo{0};
            T hi = value;
            while(lo < hi)
            {
                T const mid = lo + (hi - lo + T{1}) / T{2};
                // mid^n <= value without overflow: divide down
                T acc = value;
                bool le = true;
                T p{1};
                for(T k{0}; k < n; ++k)
                {
                    if(mid != 0 && p > acc / mid)
                    {
                        le = false;
                        break;
                    }
                    p *= mid;
                }
                if(le && p <= value)
                    lo = mid;
                else
                    hi = mid - T{1};
            }
            return lo;
        }

        //! saturating integral conversion
        template<typename T, typename V>
        [[nodiscard]] ALPAKA_FN_HOST_ACC constexpr auto clipCast(V const& val) -> T
        {
            static_assert(std::is_integral_v<T> && std::is_integral_v<V>, "clipCast is defined for integral types");
            constexpr auto tMax = std::numeric_limits<T>::max();
            constexpr auto tMin = std::numeric_limits<T>::min();
            if constexpr(std::is_signed_v<V>)
            {
                if(val < 0)
                {
                    if constexpr(std::is_signed_v<T>)
                        return static_cast<std::intmax_t>(val) < static_cast<std::intmax_t>(tMin) ? tMin
                                                                                                    : static_cast<T>(val);
                    else
                        return T{0};
                }
            }
            return static_cast<std::uintmax_t>(val) > static_cast<std::uintmax_t>(tMax) ? tMax : static_cast<T>(val);
        }
    } // namespace core

    template<typename TDim, typename TVal>
    class Vec;

    namespace detail
    {
        template<typename T>
        inline constexpr bool isVec = false;
        template<typename TDim, typename TVal>
        inline constexpr bool isVec<Vec<TDim, TVal>> = true;
    } // namespace detail

    //! N-dimensional value vector; component 0 is the slowest-varying dimension.
    template<typename TDim, typename TVal>
    class Vec final
    {
        static constexpr std::size_t kN = TDim::value;
        TVal m_v[kN == 0u ? 1u : kN];

    public:
        static_assert(kN <= 8u, "Vec supports up to 8 dimensions");
        using Dim = TDim;
        using Val = TVal;
        using value_type = TVal;
        using size_type = std::size_t;
        using iterator = TVal*;
        using const_iterator = TVal const*;

        ALPAKA_FN_HOST_ACC constexpr Vec() : m_v{}
        {
        }

        //! one value per dimension
        template<
            typename... TArgs,
            typename = std::enable_if_t<
                sizeof...(TArgs) == kN && (kN > 0u) && (std::is_convertible_v<std::decay_t<TArgs>, TVal> && ...)>>
        ALPAKA_FN_HOST_ACC constexpr Vec(TArgs&&... args) : m_v{static_cast<TVal>(std::forward<TArgs>(args))...}
        {
        }

        //! generator: f(std::integral_constant<size_t, i>) -> value of component i
        template<
            typename F,
            typename = std::enable_if_t<
                (kN > 0u) && std::is_invocable_v<F, std::integral_constant<std::size_t, 0u>>
                && !std::is_convertible_v<std::decay_t<F>, TVal>>,
            typename = void>
        ALPAKA_FN_HOST_ACC constexpr explicit Vec(F&& generator) : Vec(std::forward<F>(generator), std::make_index_sequence<kN>{})
        {
        }

    private:
        template<typename F, std::size_t... Is>
        ALPAKA_FN_HOST_ACC constexpr Vec(F&& generator, std::index_sequence<Is...>)
            : m_v{static_cast<TVal>(generator(std::integral_constant<std::size_t, Is>{}))...}
        {
        }

    public:
        [[nodiscard]] ALPAKA_FN_HOST_ACC static constexpr auto all(TVal const& val) -> Vec
        {
            Vec v;
            for(std::size_t i = 0; i < kN; ++i)
                v.m_v[i] = val;
            return v;
        }

        [[nodiscard]] ALPAKA_FN_HOST_ACC static constexpr auto zeros() -> Vec
        {
            return all(static_cast<TVal>(0));
        }

        [[nodiscard]] ALPAKA_FN_HOST_ACC static constexpr auto ones() -> Vec
        {
            return all(static_cast<TVal>(1));
        }

        [[nodiscard]] ALPAKA_FN_HOST_ACC constexpr auto begin() -> iterator
        {
            return m_v;
        }
        [[nodiscard]] ALPAKA_FN_HOST_ACC constexpr auto begin() const -> const_iterator
        {
            return m_v;
        }
        [[nodiscard]] ALPAKA_FN_HOST_ACC constexpr auto cbegin() const -> const_iterator
        {
            return m_v;
        }
        [[nodiscard]] ALPAKA_FN_HOST_ACC constexpr auto end() -> iterator
        {
            return m_v + kN;
        }
        [[nodiscard]] ALPAKA_FN_HOST_ACC constexpr auto end() const -> const_iterator
        {
            return m_v + kN;
        }
        [[nodiscard]] ALPAKA_FN_HOST_ACC constexpr auto cend() const -> const_iterator
        {
            return m_v + kN;
        }
        [[nodiscard]] ALPAKA_FN_HOST_ACC constexpr auto data() -> TVal*
        {
            return m_v;
        }
        [[nodiscard]] ALPAKA_FN_HOST_ACC constexpr auto data() const -> TVal const*
        {
            return m_v;
        }
        [[nodiscard]] ALPAKA_FN_HOST_ACC static constexpr auto size() -> std::size_t
        {
            return kN;
        }
        [[nodiscard]] ALPAKA_FN_HOST_ACC constexpr auto front() -> TVal&
        {
            return m_v[0];
        }
        [[nodiscard]] ALPAKA_FN_HOST_ACC constexpr auto front() const -> TVal const&
        {
            return m_v[0];
        }
        [[nodiscard]] ALPAKA_FN_HOST_ACC constexpr auto back() -> TVal&
        {
            return m_v[kN - 1u];
        }
        [[nodiscard]] ALPAKA_FN_HOST_ACC constexpr auto back() const -> TVal const&
        {
            return m_v[kN - 1u];
        }

        //! named access from the fastest dimension: x = last component
        [[nodiscard]] ALPAKA_FN_HOST_ACC constexpr auto x() const -> TVal
        {
            static_assert(kN >= 1u);
            return m_v[kN - 1u];
        }
        [[nodiscard]] ALPAKA_FN_HOST_ACC constexpr auto x() -> TVal&
        {
            static_assert(kN >= 1u);
            return m_v[kN - 1u];
        }
        [[nodiscard]] ALPAKA_FN_HOST_ACC constexpr auto y() const -> TVal
        {
            static_assert(kN >= 2u);
            return m_v[kN - 2u];
        }
        [[nodiscard]] ALPAKA_FN_HOST_ACC constexpr auto y() -> TVal&
        {
            static_assert(kN >= 2u);
            return m_v[kN - 2u];
        }
        [[nodiscard]] ALPAKA_FN_HOST_ACC constexpr auto z() const -> TVal
        {
            static_assert(kN >= 3u);
            return m_v[kN - 3u];
        }
        [[nodiscard]] ALPAKA_FN_HOST_ACC constexpr auto z() -> TVal&
        {
            static_assert(kN >= 3u);
            return m_v[kN - 3u];
        }
        [[nodiscard]] ALPAKA_FN_HOST_ACC constexpr auto w() const -> TVal
        {
            static_assert(kN >= 4u);
            return m_v[kN - 4u];
        }
        [[nodiscard]] ALPAKA_FN_HOST_ACC constexpr auto w() -> TVal&
        {
            static_assert(kN >= 4u);
            return m_v[kN - 4u];
        }

        template<typename TI, typename = std::enable_if_t<std::is_integral_v<TI>>>
        [[nodiscard]] ALPAKA_FN_HOST_ACC constexpr auto operator[](TI const i) -> TVal&
        {
            return m_v[static_cast<std::size_t>(i)];
        }

        template<typename TI, typename = std::enable_if_t<std::is_integral_v<TI>>>
        [[nodiscard]] ALPAKA_FN_HOST_ACC constexpr auto operator[](TI const i) const -> TVal const&
        {
            return m_v[static_cast<std::size_t>(i)];
        }

        template<typename F>
        [[nodiscard]] ALPAKA_FN_HOST_ACC constexpr auto foldrAll(F const& f, TVal init = TVal{}) const -> TVal
        {
            TVal r = init;
            for(std::size_t i = kN; i-- > 0u;)
                r = f(m_v[i], r);
            return r;
        }

        [[nodiscard]] ALPAKA_FN_HOST_ACC constexpr auto prod() const -> TVal
        {
            TVal r = static_cast<TVal>(1);
            for(std::size_t i = 0; i < kN; ++i)
                r = static_cast<TVal>(r * m_v[i]);
            return r;
        }
        [[nodiscard]] ALPAKA_FN_HOST_ACC constexpr auto product() const -> TVal
        {
            return prod();
        }

        [[nodiscard]] ALPAKA_FN_HOST_ACC constexpr auto sum() const -> TVal
        {
            TVal r = static_cast<TVal>(0);
            for(std::size_t i = 0; i < kN; ++i)
                r = static_cast<TVal>(r + m_v[i]);
            return r;
        }

        [[nodiscard]] ALPAKA_FN_HOST_ACC constexpr auto min() const -> TVal
        {
            TVal r = std::numeric_limits<TVal>::max();
            for(std::size_t i = 0; i < kN; ++i)
                r = m_v[i] < r ? m_v[i] : r;
            return r;
        }

        [[nodiscard]] ALPAKA_FN_HOST_ACC constexpr auto max() const -> TVal
        {
            TVal r = std::numeric_limits<TVal>::lowest();
            for(std::size_t i = 0; i < kN; ++i)
                r = m_v[i] > r ? m_v[i] : r;
            return r;
        }

        [[nodiscard]] ALPAKA_FN_HOST_ACC constexpr auto all() const -> bool
        {
            for(std::size_t i = 0; i < kN; ++i)
                if(!m_v[i])
                    return false;
            return true;
        }

        [[nodiscard]] ALPAKA_FN_HOST_ACC constexpr auto any() const -> bool
        {
            for(std::size_t i = 0; i < kN; ++i)
                if(m_v[i])
                    return true;
            return false;
        }

        [[nodiscard]] ALPAKA_FN_HOST_ACC constexpr auto none() const -> bool
        {
            return !any();
        }

        //! index of the smallest / largest component (first one on ties)
        [[nodiscard]] ALPAKA_FN_HOST_ACC constexpr auto minElem() const -> std::size_t
        {
            std::size_t r = 0;
            for(std::size_t i = 1; i < kN; ++i)
                if(m_v[i] < m_v[r])
                    r = i;
            return r;
        }
        [[nodiscard]] ALPAKA_FN_HOST_ACC constexpr auto maxElem() const -> std::size_t
        {
            std::size_t r = 0;
            for(std::size_t i = 1; i < kN; ++i)
                if(m_v[i] > m_v[r])
                    r = i;
            return r;
        }

        [[nodiscard]] ALPAKA_FN_HOST_ACC constexpr auto toArray() const -> std::array<TVal, kN>
        {
            std::array<TVal, kN> a{};
            for(std::size_t i = 0; i < kN; ++i)
                a[i] = m_v[i];
            return a;
        }

        template<std::size_t I>
        [[nodiscard]] ALPAKA_FN_HOST_ACC constexpr auto get() const -> TVal const&
        {
            static_assert(I < kN);
            return m_v[I];
        }
        template<std::size_t I>
        [[nodiscard]] ALPAKA_FN_HOST_ACC constexpr auto get() -> TVal&
        {
            static_assert(I < kN);
            return m_v[I];
        }

#define ALPAKA_B200_VEC_BINOP(op)                                                                                     \
    [[nodiscard]] ALPAKA_FN_HOST_ACC friend constexpr auto operator op(Vec const& a, Vec const& b) -> Vec            \
    {                                                                                                                 \
        Vec r;                                                                                                        \
        for(std::size_t i = 0; i < kN; ++i)                                                                           \
            r.m_v[i] = static_cast<TVal>(a.m_v[i] op b.m_v[i]);                                                       \
        return r;                                                                                                     \
    }                                                                                                                 \
    ALPAKA_FN_HOST_ACC friend constexpr auto operator op##=(Vec & a, Vec const& b) -> Vec&                            \
    {                                                                                                                 \
        for(std::size_t i = 0; i < kN; ++i)                                                                           \
            a.m_v[i] = static_cast<TVal>(a.m_v[i] op b.m_v[i]);                                                       \
        return a;                                                                                                     \
    }
        ALPAKA_B200_VEC_BINOP(+)
        ALPAKA_B200_VEC_BINOP(-)
        ALPAKA_B200_VEC_BINOP(*)
        ALPAKA_B200_VEC_BINOP(/)
        ALPAKA_B200_VEC_BINOP(%)
#undef ALPAKA_B200_VEC_BINOP

        [[nodiscard]] ALPAKA_FN_HOST_ACC friend constexpr auto operator==(Vec const& a, Vec const& b) -> bool
        {
            for(std::size_t i = 0; i < kN; ++i)
                if(!(a.m_v[i] == b.m_v[i]))
                    return false;
            return true;
        }
        [[nodiscard]] ALPAKA_FN_HOST_ACC friend constexpr auto operator!=(Vec const& a, Vec const& b) -> bool
        {
            return !(a == b);
        }

#define ALPAKA_B200_VEC_CMP(op)                                                                                       \
    [[nodiscard]] ALPAKA_FN_HOST_ACC friend constexpr auto operator op(Vec const& a, Vec const& b) -> Vec<TDim, bool> \
    {                                                                                                                 \
        Vec<TDim, bool> r;                                                                                            \
        for(std::size_t i = 0; i < kN; ++i)                                                                           \
            r[i] = a.m_v[i] op b.m_v[i];                                                                              \
        return r;                                                                                                     \
    }
        ALPAKA_B200_VEC_CMP(<)
        ALPAKA_B200_VEC_CMP(<=)
        ALPAKA_B200_VEC_CMP(>)
        ALPAKA_B200_VEC_CMP(>=)
        ALPAKA_B200_VEC_CMP(&&)
        ALPAKA_B200_VEC_CMP(||)
#undef ALPAKA_B200_VEC_CMP

        friend auto operator<<(std::ostream& os, Vec const& v) -> std::ostream&
        {
            os << "(";
            for(std::size_t i = 0; i < kN; ++i)
            {
                if constexpr(sizeof(TVal) == 1u)
                    os << static_cast<int>(v.m_v[i]);
                else
                    os << v.m_v[i];
                if(i + 1u != kN)
                    os << ", ";
            }
            return os << ")";
        }
    };

    // CTAD: Vec(a, b, c) -> Vec<DimInt<3>, decltype(a)>
    template<typename TFirst, typename... TRest>
    ALPAKA_FN_HOST_ACC Vec(TFirst&&, TRest&&...) -> Vec<DimInt<1u + sizeof...(TRest)>, std::decay_t<TFirst>>;

    template<typename T>
    inline constexpr bool isVec = detail::isVec<std::remove_cv_t<std::remove_reference_t<T>>>;

    namespace trait
    {
        template<typename TDim, typename TVal>
        struct DimType<Vec<TDim, TVal>>
        {
            using type = TDim;
        };
        template<typename TDim, typename TVal>
        struct IdxType<Vec<TDim, TVal>>
        {
            using type = TVal;
        };
    } // namespace trait

    //! element-wise static_cast
    template<typename TValNew, typename TDim, typename TVal>
    [[nodiscard]] ALPAKA_FN_HOST_ACC constexpr auto castVec(Vec<TDim, TVal> const& v) -> Vec<TDim, TValNew>
    {
        Vec<TDim, TValNew> r;
        for(std::size_t i = 0; i < TDim::value; ++i)
            r[i] = static_cast<TValNew>(v[i]);
        return r;
    }
    template<typename TValNew, typename TDim, typename TVal>
    [[nodiscard]] ALPAKA_FN_HOST_ACC constexpr auto cast(Vec<TDim, TVal> const& v) -> Vec<TDim, TValNew>
    {
        return castVec<TValNew>(v);
    }

    [[nodiscard]] ALPAKA_FN_HOST_ACC constexpr auto reverseVec(auto const& v)
    {
        using V = std::decay_t<decltype(v)>;
        V r;
        for(std::size_t i = 0; i < V::size(); ++i)
            r[i] = v[V::size() - 1u - i];
        return r;
    }
    template<typename TDim, typename TVal>
    [[nodiscard]] ALPAKA_FN_HOST_ACC constexpr auto reverse(Vec<TDim, TVal> const& v) -> Vec<TDim, TVal>
    {
        return reverseVec(v);
    }

    //! components picked by an index_sequence
    template<typename TDim, typename TVal, std::size_t... Is>
    [[nodiscard]] ALPAKA_FN_HOST_ACC constexpr auto subVecFromIndices(Vec<TDim, TVal> const& v, std::index_sequence<Is...>)
        -> Vec<DimInt<sizeof...(Is)>, TVal>
    {
        if constexpr(sizeof...(Is) == 0u)
            return Vec<DimInt<0u>, TVal>{};
        else
            return Vec<DimInt<sizeof...(Is)>, TVal>{v[Is]...};
    }

    //! the same with the sequence as explicit template argument: subVecFromIndices<std::index_sequence<0, 2>>(v)
    template<typename TIndexSequence, typename TDim, typename TVal>
    [[nodiscard]] ALPAKA_FN_HOST_ACC constexpr auto subVecFromIndices(Vec<TDim, TVal> const& v)
    {
        return subVecFromIndices(v, TIndexSequence{});
    }

    namespace detail
    {
        template<std::size_t Off, std::size_t... Is>
        constexpr auto shiftSeq(std::index_sequence<Is...>) -> std::index_sequence<(Off + Is)...>
        {
            return {};
        }
    } // namespace detail

    //! the first TSubDim components
    template<typename TSubDim, typename TDim, typename TVal>
    [[nodiscard]] ALPAKA_FN_HOST_ACC constexpr auto subVecBegin(Vec<TDim, TVal> const& v) -> Vec<TSubDim, TVal>
    {
        static_assert(TSubDim::value <= TDim::value);
        return subVecFromIndices(v, std::make_index_sequence<TSubDim::value>{});
    }

    //! the last TSubDim components
    template<typename TSubDim, typename TDim, typename TVal>
    [[nodiscard]] ALPAKA_FN_HOST_ACC constexpr auto subVecEnd(Vec<TDim, TVal> const& v) -> Vec<TSubDim, TVal>
    {
        static_assert(TSubDim::value <= TDim::value);
        return subVecFromIndices(
            v,
            detail::shiftSeq<TDim::value - TSubDim::value>(std::make_index_sequence<TSubDim::value>{}));
    }

    template<typename TDimA, typename TDimB, typename TVal>
    [[nodiscard]] ALPAKA_FN_HOST_ACC constexpr auto concatVec(Vec<TDimA, TVal> const& a, Vec<TDimB, TVal> const& b)
        -> Vec<DimInt<TDimA::value + TDimB::value>, TVal>
    {
        Vec<DimInt<TDimA::value + TDimB::value>, TVal> r;
        for(std::size_t i = 0; i < TDimA::value; ++i)
            r[i] = a[i];
        for(std::size_t i = 0; i < TDimB::value; ++i)
            r[TDimA::value + i] = b[i];
        return r;
    }

    //! component-wise minimum / maximum of one or more vectors
    template<typename TDim, typename TVal, typename... TVecs>
    [[nodiscard]] ALPAKA_FN_HOST_ACC constexpr auto elementwise_min(Vec<TDim, TVal> const& first, TVecs const&... rest)
        -> Vec<TDim, TVal>
    {
        static_assert((std::is_same_v<Vec<TDim, TVal>, TVecs> && ...), "elementwise_min: all vectors must have one type");
        Vec<TDim, TVal> r = first;
        auto const fold = [&r](Vec<TDim, TVal> const& v)
        {
            for(std::size_t i = 0; i < TDim::value; ++i)
                r[i] = v[i] < r[i] ? v[i] : r[i];
        };
        (fold(rest), ...);
        return r;
    }
    template<typename TDim, typename TVal, typename... TVecs>
    [[nodiscard]] ALPAKA_FN_HOST_ACC constexpr auto elementwise_max(Vec<TDim, TVal> const& first, TVecs const&... rest)
        -> Vec<TDim, TVal>
    {
        static_assert((std::is_same_v<Vec<TDim, TVal>, TVecs> && ...), "elementwise_max: all vectors must have one type");
        Vec<TDim, TVal> r = first;
        auto const fold = [&r](Vec<TDim, TVal> const& v)
        {
            for(std::size_t i = 0; i < TDim::value; ++i)
                r[i] = v[i] > r[i] ? v[i] : r[i];
        };
        (fold(rest), ...);
        return r;
    }

    template<std::size_t I, typename TDim, typename TVal>
    [[nodiscard]] ALPAKA_FN_HOST_ACC constexpr auto get(Vec<TDim, TVal> const& v) -> TVal const&
    {
        return v.template get<I>();
    }
    template<std::size_t I, typename TDim, typename TVal>
    [[nodiscard]] ALPAKA_FN_HOST_ACC constexpr auto get(Vec<TDim, TVal>& v) -> TVal&
    {
        return v.template get<I>();
    }

    // ---------------------------------------------------------------------------------------------------------------
    // extents / offsets
    namespace trait
    {
        //! T -> Vec<Dim<T>, Idx<T>> of extents. Specialised by buffers and views.
        template<typename T, typename TSfinae = void>
        struct GetExtents;
        template<typename T, typename TSfinae = void>
        struct GetOffsets;

        template<typename TDim, typename TVal>
        struct GetExtents<Vec<TDim, TVal>>
        {
            ALPAKA_FN_HOST_ACC constexpr auto operator()(Vec<TDim, TVal> const& v) const -> Vec<TDim, TVal>
            {
                return v;
            }
        };
        template<typename TDim, typename TVal>
        struct GetOffsets<Vec<TDim, TVal>>
        {
            ALPAKA_FN_HOST_ACC constexpr auto operator()(Vec<TDim, TVal> const& v) const -> Vec<TDim, TVal>
            {
                return v;
            }
        };
        template<typename T>
        struct GetExtents<T, std::enable_if_t<std::is_arithmetic_v<T>>>
        {
            ALPAKA_FN_HOST_ACC constexpr auto operator()(T const& v) const -> Vec<DimInt<1u>, T>
            {
                return Vec<DimInt<1u>, T>{v};
            }
        };
        template<typename T>
        struct GetOffsets<T, std::enable_if_t<std::is_arithmetic_v<T>>>
        {
            ALPAKA_FN_HOST_ACC constexpr auto operator()(T const& v) const -> Vec<DimInt<1u>, T>
            {
                return Vec<DimInt<1u>, T>{v};
            }
        };
    } // namespace trait

    ALPAKA_NO_HOST_ACC_WARNING
    template<typename T>
    [[nodiscard]] ALPAKA_FN_HOST_ACC constexpr auto getExtents(T const& object) -> Vec<Dim<T>, Idx<T>>
    {
        return trait::GetExtents<T>{}(object);
    }
    template<typename T>
    [[nodiscard]] ALPAKA_FN_HOST_ACC constexpr auto getExtentVec(T const& object) -> Vec<Dim<T>, Idx<T>>
    {
        return getExtents(object);
    }
    template<typename TSubDim, typename T>
    [[nodiscard]] ALPAKA_FN_HOST_ACC constexpr auto getExtentVecEnd(T const& object) -> Vec<TSubDim, Idx<T>>
    {
        return subVecEnd<TSubDim>(getExtents(object));
    }
    ALPAKA_NO_HOST_ACC_WARNING
    template<typename T>
    [[nodiscard]] ALPAKA_FN_HOST_ACC constexpr auto getOffsets(T const& object) -> Vec<Dim<T>, Idx<T>>
    {
        return trait::GetOffsets<T>{}(object);
    }
    template<typename T>
    [[nodiscard]] ALPAKA_FN_HOST_ACC constexpr auto getOffsetVec(T const& object) -> Vec<Dim<T>, Idx<T>>
    {
        return getOffsets(object);
    }

    //! extent of the fastest (last) dimension; 1 if T has fewer dimensions
    template<typename T>
    [[nodiscard]] ALPAKA_FN_HOST_ACC constexpr auto getWidth(T const& object) -> Idx<T>
    {
        if constexpr(Dim<T>::value >= 1u)
            return getExtents(object)[Dim<T>::value - 1u];
        else
            return Idx<T>{1};
    }
    template<typename T>
    [[nodiscard]] ALPAKA_FN_HOST_ACC constexpr auto getHeight(T const& object) -> Idx<T>
    {
        if constexpr(Dim<T>::value >= 2u)
            return getExtents(object)[Dim<T>::value - 2u];
        else
            return Idx<T>{1};
    }
    template<typename T>
    [[nodiscard]] ALPAKA_FN_HOST_ACC constexpr auto getDepth(T const& object) -> Idx<T>
    {
        if constexpr(Dim<T>::value >= 3u)
            return getExtents(object)[Dim<T>::value - 3u];
        else
            return Idx<T>{1};
    }
    template<typename T>
    [[nodiscard]] ALPAKA_FN_HOST_ACC constexpr auto getExtentProduct(T const& object) -> Idx<T>
    {
        return getExtents(object).prod();
    }
    template<typename T>
    [[nodiscard]] ALPAKA_FN_HOST_ACC constexpr auto getOffsetX(T const& object) -> Idx<T>
    {
        if constexpr(Dim<T>::value >= 1u)
            return getOffsets(object)[Dim<T>::value - 1u];
        else
            return Idx<T>{0};
    }
    template<typename T>
    [[nodiscard]] ALPAKA_FN_HOST_ACC constexpr auto getOffsetY(T const& object) -> Idx<T>
    {
        if constexpr(Dim<T>::value >= 2u)
            return getOffsets(object)[Dim<T>::value - 2u];
        else
            return Idx<T>{0};
    }
    template<typename T>
    [[nodiscard]] ALPAKA_FN_HOST_ACC constexpr auto getOffsetZ(T const& object) -> Idx<T>
    {
        if constexpr(Dim<T>::value >= 3u)
            return getOffsets(object)[Dim<T>::value - 3u];
        else
            return Idx<T>{0};
    }

    // ---------------------------------------------------------------------------------------------------------------
    // index mapping between an N-d index and its row-major linearisation (slowest dimension first)
    template<std::size_t TDimOut, std::size_t TDimIn, std::size_t TDimExtents, typename TElem>
    [[nodiscard]] ALPAKA_FN_HOST_ACC constexpr auto mapIdx(
        Vec<DimInt<TDimIn>, TElem> const& in,
        Vec<DimInt<TDimExtents>, TElem> const& extent) -> Vec<DimInt<TDimOut>, TElem>
    {
        if constexpr(TDimOut == 0u || TDimIn == 0u)
        {
            return Vec<DimInt<TDimOut>, TElem>::zeros();
        }
        else if constexpr(TDimOut == TDimIn)
        {
            return in;
        }
        else if constexpr(TDimOut == 1u)
        {
            static_assert(TDimIn == TDimExtents, "mapIdx<1>: the index and the extent must have the same dimension");
            TElem lin = in[0];
            for(std::size_t d = 1; d < TDimIn; ++d)
                lin = static_cast<TElem>(lin * extent[d] + in[d]);
            return Vec<DimInt<1u>, TElem>{lin};
        }
        else
        {
            static_assert(TDimIn == 1u, "mapIdx: only 1 -> N, N -> 1 and N -> N mappings exist");
            static_assert(TDimOut == TDimExtents, "mapIdx<N>: the extent must have N dimensions");
            Vec<DimInt<TDimOut>, TElem> out;
            TElem rest = in[0];
            for(std::size_t d = TDimOut; d-- > 1u;)
            {
                out[d] = static_cast<TElem>(rest % extent[d]);
                rest = static_cast<TElem>(rest / extent[d]);
            }
            out[0] = rest;
            return out;
        }
    }

    //! same mapping with a byte pitch vector instead of extents (pitch[d] = bytes between neighbours in dimension d)
    template<std::size_t TDimOut, std::size_t TDimIn, std::size_t TDimPitch, typename TElem>
    [[nodiscard]] ALPAKA_FN_HOST_ACC constexpr auto mapIdxPitchBytes(
        Vec<DimInt<TDimIn>, TElem> const& in,
        Vec<DimInt<TDimPitch>, TElem> const& pitches) -> Vec<DimInt<TDimOut>, TElem>
    {
        if constexpr(TDimOut == 0u || TDimIn == 0u)
        {
            return Vec<DimInt<TDimOut>, TElem>::zeros();
        }
        else if constexpr(TDimOut == TDimIn)
        {
            return in;
        }
        else if constexpr(TDimOut == 1u)
        {
            static_assert(TDimIn == TDimPitch);
            return Vec<DimInt<1u>, TElem>{(in * pitches).sum()};
        }
        else
        {
            static_assert(TDimIn == 1u && TDimOut == TDimPitch);
            Vec<DimInt<TDimOut>, TElem> out;
            TElem rest = in[0];
            for(std::size_t d = 0; d < TDimOut; ++d)
            {
                out[d] = static_cast<TElem>(rest / pitches[d]);
                rest = static_cast<TElem>(rest % pitches[d]);
            }
            return out;
        }
    }
} // namespace alpaka

// structured bindings: auto const [i] = getIdx<Grid, Threads>(acc);
namespace std
{
    template<typename TDim, typename TVal>
    struct tuple_size<alpaka::Vec<TDim, TVal>> : integral_constant<size_t, TDim::value>
    {
    };

    template<size_t I, typename TDim, typename TVal>
    struct tuple_element<I, alpaka::Vec<TDim, TVal>>
    {
        using type = TVal;
    };
} // namespace std
