// include/alpaka/test/mem/view/Iterator.hpp -- element cursor over a (possibly padded) N-d view, usable inside kernels.
// Interface parity with the reference's test helper (include/alpaka/test/mem/view/Iterator.hpp:13-143): Iterator<TView>,
// begin(view), end(view), trait::{IteratorView,Begin,End}. Elements are visited in row-major order (slowest dimension
// first); the cursor is the linear element number and is decoded against the extents on every dereference, so the rows
// may be padded (device buffers of this back-end pad rows to B200_ROW_ALIGN).
#pragma once

#include <alpaka/alpaka.hpp>

#include <cstddef>
#include <type_traits>

namespace alpaka::test
{
    namespace trait
    {
        //! T with the const-ness of TSource
        template<typename T, typename TSource>
        using MimicConst = std::conditional_t<std::is_const_v<TSource>, std::add_const_t<T>, std::remove_const_t<T>>;

        template<typename TView, typename TSfinae = void>
        class IteratorView
        {
            using View = std::decay_t<TView>;
            using Dim = alpaka::Dim<View>;
            using Idx = alpaka::Idx<View>;
            using Elem = MimicConst<alpaka::Elem<View>, TView>;
            using Byte = MimicConst<unsigned char, Elem>;

        public:
            IteratorView(TView& view, Idx const pos)
                : m_base(reinterpret_cast<Byte*>(getPtrNative(view)))
                , m_pos(pos)
                , m_extent(getExtents(view))
                , m_pitch(getPitchesInBytes(view))
            {
            }
            explicit IteratorView(TView& view) : IteratorView(view, Idx{0})
            {
            }

            ALPAKA_FN_HOST_ACC auto operator++() -> IteratorView&
            {
                ++m_pos;
                return *this;
            }
            ALPAKA_FN_HOST_ACC auto operator--() -> IteratorView&
            {
                --m_pos;
                return *this;
            }
            ALPAKA_FN_HOST_ACC auto operator++(int) -> IteratorView
            {
                auto const before = *this;
                ++m_pos;
                return before;
            }
            ALPAKA_FN_HOST_ACC auto operator--(int) -> IteratorView
            {
                auto const before = *this;
                --m_pos;
                return before;
            }
            template<typename TOther>
            ALPAKA_FN_HOST_ACC auto operator==(TOther const& other) const -> bool
            {
                return m_pos == other.m_pos;
            }
            template<typename TOther>
            ALPAKA_FN_HOST_ACC auto operator!=(TOther const& other) const -> bool
            {
                return m_pos != other.m_pos;
            }

            ALPAKA_FN_HOST_ACC auto operator*() const -> Elem&
            {
                // peel the row-major digits of m_pos from the fastest dimension upwards
                std::size_t byteOffset = 0;
                Idx rest = m_pos;
                for(std::size_t d = Dim::value; d-- > 0u;)
                {
                    Idx const digit = d == 0u ? rest : static_cast<Idx>(rest % m_extent[d]);
                    rest = d == 0u ? rest : static_cast<Idx>(rest / m_extent[d]);
                    byteOffset += static_cast<std::size_t>(digit) * static_cast<std::size_t>(m_pitch[d]);
                }
                return *reinterpret_cast<Elem*>(m_base + byteOffset);
            }

            Byte* m_base;
            Idx m_pos;
            Vec<Dim, Idx> m_extent;
            Vec<Dim, Idx> m_pitch;
        };

        template<typename TView, typename TSfinae = void>
        struct Begin
        {
            static auto begin(TView& view) -> IteratorView<TView>
            {
                return IteratorView<TView>(view);
            }
        };
        template<typename TView, typename TSfinae = void>
        struct End
        {
            static auto end(TView& view) -> IteratorView<TView>
            {
                return IteratorView<TView>(view, getExtents(view).prod());
            }
        };
    } // namespace trait

    template<typename TView>
    using Iterator = trait::IteratorView<TView>;

    template<typename TView>
    auto begin(TView& view) -> Iterator<TView>
    {
        return trait::Begin<TView>::begin(view);
    }
    template<typename TView>
    auto end(TView& view) -> Iterator<TView>
    {
        return trait::End<TView>::end(view);
    }
} // namespace alpaka::test
