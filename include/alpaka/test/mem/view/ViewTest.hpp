// include/alpaka/test/mem/view/ViewTest.hpp -- the generic checks the reference's memory tests apply to every view type
// (interface parity with include/alpaka/test/mem/view/ViewTest.hpp:17-264): testViewImmutable, verifyBytesSet,
// verifyViewsEqual, iotaFillView, testViewMutable. The verification kernels walk the views with test::Iterator on the
// device, one thread, so padded rows are honoured.
#pragma once

#include <alpaka/alpaka.hpp>
#include <alpaka/test/KernelExecutionFixture.hpp>
#include <alpaka/test/mem/view/Iterator.hpp>

#include <catch2/catch_test_macros.hpp>

#include <cstdint>
#include <cstdio>
#include <numeric>
#include <type_traits>
#include <vector>

namespace alpaka::test
{
    //! the trait surface of a view that must hold without touching its memory
    template<typename TElem, typename TDim, typename TIdx, typename TDev, typename TView>
    auto testViewImmutable(TView const& view, TDev const& dev, Vec<TDim, TIdx> const& extent, Vec<TDim, TIdx> const& offset)
        -> void
    {
        static_assert(std::is_same_v<Dev<TView>, TDev>, "unexpected device type of the view");
        static_assert(Dim<TView>::value == TDim::value, "unexpected dimensionality of the view");
        static_assert(std::is_same_v<Elem<TView>, TElem>, "unexpected element type of the view");
        static_assert(std::is_same_v<Idx<TView>, TIdx>, "unexpected index type of the view");

        REQUIRE(dev == getDev(view));
        REQUIRE(extent == getExtents(view));
        REQUIRE(offset == getOffsets(view));

        // every pitch is at least the dense one
        auto const dense = alpaka::detail::calculatePitchesFromExtents<TElem>(extent);
        auto const pitch = getPitchesInBytes(view);
        for(std::size_t d = 0; d < TDim::value; ++d)
            REQUIRE(pitch[d] >= dense[d]);

        // a const view hands out a pointer to const; it must be non-null whenever there are elements
        using NativePtr = decltype(getPtrNative(view));
        static_assert(std::is_pointer_v<NativePtr>, "getPtrNative must return a pointer");
        static_assert(std::is_const_v<std::remove_pointer_t<NativePtr>>, "a const view must yield a pointer to const");
        NativePtr const p = getPtrNative(view);
        if(getExtentProduct(view) != static_cast<TIdx>(0u))
            REQUIRE(p != nullptr);
    }

    //! clears *success if any byte of [begin, end) differs from `byte`
    struct VerifyBytesSetKernel
    {
        ALPAKA_NO_HOST_ACC_WARNING
        template<typename TAcc, typename TIter>
        ALPAKA_FN_ACC void operator()(TAcc const&, bool* success, TIter const& begin, TIter const& end, std::uint8_t const& byte) const
        {
            for(auto it = begin; it != end; ++it)
            {
                auto const& elem = *it;
                auto const* const raw = reinterpret_cast<std::uint8_t const*>(&elem);
                for(unsigned i = 0; i < static_cast<unsigned>(sizeof(elem)); ++i)
                    if(raw[i] != byte)
                    {
                        printf("byte %u of an element is %u, expected %u\n", i, unsigned{raw[i]}, unsigned{byte});
                        *success = false;
                    }
            }
        }
    };

    template<typename TAcc, typename TView>
    auto verifyBytesSet(TView const& view, std::uint8_t const& byte) -> void
    {
        KernelExecutionFixture<TAcc> fixture(Vec<Dim<TView>, Idx<TView>>::ones());
        REQUIRE(fixture(VerifyBytesSetKernel{}, test::begin(view), test::end(view), byte));
    }

    //! clears *success if the ranges [beginA, endA) and [beginB, ...) differ element-wise
    struct VerifyViewsEqualKernel
    {
        ALPAKA_NO_HOST_ACC_WARNING
        template<typename TAcc, typename TIterA, typename TIterB>
        ALPAKA_FN_ACC void operator()(TAcc const&, bool* success, TIterA beginA, TIterA const& endA, TIterB beginB) const
        {
            for(; beginA != endA; ++beginA, ++beginB)
                ALPAKA_CHECK(*success, *beginA == *beginB);
        }
    };

    template<typename TAcc, typename TViewB, typename TViewA>
    auto verifyViewsEqual(TViewA const& viewA, TViewB const& viewB) -> void
    {
        static_assert(Dim<TViewA>::value == Dim<TViewB>::value, "both views must have the same dimensionality");
        static_assert(std::is_same_v<Idx<TViewA>, Idx<TViewB>>, "both views must have the same index type");
        KernelExecutionFixture<TAcc> fixture(Vec<Dim<TViewA>, Idx<TViewA>>::ones());
        REQUIRE(fixture(VerifyViewsEqualKernel{}, test::begin(viewA), test::end(viewA), test::begin(viewB)));
    }

    //! view[i] = i in row-major order, staged through a host vector
    template<typename TView, typename TQueue>
    auto iotaFillView(TQueue& queue, TView& view) -> void
    {
        using E = Elem<TView>;
        auto const devHost = getDevByIdx(PlatformCpu{}, 0);
        auto const extent = getExtents(view);
        std::vector<E> staging(static_cast<std::size_t>(extent.prod()));
        std::iota(staging.begin(), staging.end(), E{});
        auto hostView = createView(devHost, staging, extent);
        memcpy(queue, view, hostView);
        wait(queue); // the staging vector dies with this scope
    }

    //! memset, copy-into and copy-out-of a writable view
    template<typename TAcc, typename TView, typename TQueue>
    auto testViewMutable(TQueue& queue, TView& view) -> void
    {
        using NativePtr = decltype(getPtrNative(view));
        static_assert(std::is_pointer_v<NativePtr>, "getPtrNative must return a pointer");
        static_assert(!std::is_const_v<std::remove_pointer_t<NativePtr>>, "a mutable view must yield a pointer to non-const");

        auto const byte = static_cast<std::uint8_t>(42u);
        memset(queue, view, byte);
        wait(queue);
        verifyBytesSet<TAcc>(view, byte);

        using E = Elem<TView>;
        using I = Idx<TView>;
        auto const dev = getDev(view);
        auto const extent = getExtents(view);
        {
            auto src = allocBuf<E, I>(dev, extent);
            iotaFillView(queue, src);
            memcpy(queue, view, src);
            wait(queue);
            verifyViewsEqual<TAcc>(view, src);
        }
        {
            auto dst = allocBuf<E, I>(dev, extent);
            memcpy(queue, dst, view);
            wait(queue);
            verifyViewsEqual<TAcc>(dst, view);
        }
    }
} // namespace alpaka::test
