// alpaka::interface::Implements / ImplementationBase (reference: include/alpaka/core/Interface.hpp; pinned by
// test/unit/core/src/InterfaceTest.cpp). The reference dispatches its traits on "the base class of the accelerator that
// implements concept X": a class says so by deriving from Implements<X, TheImplementingClass>, and
// ImplementationBase<X, TDerived> names that class, or TDerived itself when nothing in its hierarchy implements X.
// The B200 layer has ONE accelerator class and dispatches on it directly, so nothing here is on the hot path; the
// templates exist for user code that extends accelerators the reference's way.
#pragma once
#include <alpaka/alpaka.hpp>

#include <type_traits>
#include <utility>

namespace alpaka::interface
{
    //! Tag base: "TImplementer implements TInterface".
    template<typename TInterface, typename TImplementer>
    struct Implements
    {
    };

    namespace detail
    {
        // derived-to-base pointer conversion finds the (one) Implements<TInterface, X> base and deduces X
        template<typename TInterface, typename TImplementer>
        auto implementer(Implements<TInterface, TImplementer> const*) -> TImplementer;
        template<typename TInterface, typename TDerived>
        auto implementerOr(...) -> TDerived;
        template<typename TInterface, typename TDerived>
        auto implementerOr(int) -> decltype(implementer<TInterface>(std::declval<TDerived const*>()));
    } // namespace detail

    //! The class in TDerived's hierarchy that was declared to implement TInterface; TDerived if there is none.
    template<typename TInterface, typename TDerived>
    using ImplementationBase = decltype(detail::implementerOr<TInterface, TDerived>(0));
} // namespace alpaka::interface
