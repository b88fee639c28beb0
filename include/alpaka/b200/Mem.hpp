// include/alpaka/b200/Mem.hpp -- buffers, views, allocation, memcpy and memset.
//
// API parity with the reference's mem/buf/Traits.hpp:63-137 (allocBuf, allocAsyncBuf, allocAsyncBufIfSupported,
// allocMappedBuf), mem/view/Traits.hpp:207-482 (memset, memcpy, createTaskMemcpy, createView, createSubView,
// getPitchesInBytes, getPtrNative), mem/view/ViewAccessOps.hpp:35-150, mem/buf/BufCpu.hpp:195-221 and the CUDA
// buffer/copy implementations (mem/buf/BufUniformCudaHipRt.hpp:53-415, mem/buf/uniformCudaHip/Copy.hpp:32-495,
// Set.hpp). What changes underneath (north star: "buffer allocation (stream-ordered cudaMallocAsync pools)"):
//   * allocBuf AND allocAsyncBuf come from the per-device stream-ordered pool through b200_malloc_async /
//     b200_malloc_pitched_async; allocBuf orders the allocation on the legacy stream and waits for it, and its
//     deleter waits for the device before freeing -- the observable semantics of the reference's cudaMalloc/cudaFree;
//   * N-d device buffers (N >= 2) get rows padded to B200_ROW_ALIGN bytes (pools have no pitched API; the TMA stencil
//     needs 16-byte-multiple row strides); getPitchesInBytes reports the padded pitch exactly like the reference
//     reports cudaMallocPitch's (mem/buf/BufUniformCudaHipRt.hpp:183-202); host buffers are unpadded;
//   * asynchronous buffers exist for every dimensionality (the reference restricts them to 1-D, :286-288).
#pragma once

#include "Dev.hpp"

#include <array>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>

namespace alpaka
{
    namespace trait
    {
        template<typename TView, typename TSfinae = void>
        struct GetPtrNative;
        template<typename TView, typename TSfinae = void>
        struct GetPitchesInBytes;
        template<typename TDev, typename TElem, typename TDim, typename TIdx, typename TSfinae = void>
        struct BufType;
        template<typename TElem, typename TDim, typename TIdx, typename TDev, typename TSfinae = void>
        struct BufAlloc;
        template<typename TElem, typename TDim, typename TIdx, typename TDev, typename TSfinae = void>
        struct AsyncBufAlloc;
        template<typename TDim, typename TDev>
        struct HasAsyncBufSupport : std::false_type
        {
        };
        template<typename TPlatform>
        struct HasMappedBufSupport : std::false_type
        {
        };
    } // namespace trait

    //! The buffer type of device TDev holding TElem with TDim dimensions indexed by TIdx.
    template<typename TDev, typename TElem, typename TDim, typename TIdx>
    using Buf = typename trait::BufType<Dev<TDev>, TElem, TDim, TIdx>::type;

    namespace b200
    {
        //! byte pitches for a row pitch: [N-1] = sizeof(T), [N-2] = rowPitch, [N-3] = rowPitch * extent[N-2], ...
        template<typename TElem, typename TDim, typename TIdx>
        [[nodiscard]] constexpr auto pitchesFromRowPitch(Vec<TDim, TIdx> const& extent, std::size_t rowPitchBytes)
            -> Vec<TDim, TIdx>
        {
            Vec<TDim, TIdx> p;
            constexpr std::size_t n = TDim::value;
            if constexpr(n > 0u)
            {
                p[n - 1u] = static_cast<TIdx>(sizeof(TElem));
                if constexpr(n > 1u)
                {
                    p[n - 2u] = static_cast<TIdx>(rowPitchBytes);
                    for(std::size_t d = n - 2u; d-- > 0u;)
                        p[d] = static_cast<TIdx>(p[d + 1u] * extent[d + 1u]);
                }
            }
            return p;
        }

        //! data()/operator[]/at/begin/end shared by every view (reference: mem/view/ViewAccessOps.hpp)
        template<typename TView, typename TElem, typename TDim, typename TIdx>
        struct ViewOps
        {
            using value_type = TElem;
            [[nodiscard]] auto data() -> TElem*
            {
                return static_cast<TView*>(this)->nativePtr();
            }
            [[nodiscard]] auto data() const -> TElem const*
            {
                return static_cast<TView const*>(this)->nativePtr();
            }
            [[nodiscard]] auto operator*() -> TElem&
            {
                static_assert(TDim::value == 0u, "operator* is only valid for 0-dimensional views");
                return *data();
            }
            [[nodiscard]] auto operator*() const -> TElem const&
            {
                static_assert(TDim::value == 0u, "operator* is only valid for 0-dimensional views");
                return *data();
            }
            [[nodiscard]] auto operator->() -> TElem*
            {
                return data();
            }
            [[nodiscard]] auto operator->() const -> TElem const*
            {
                return data();
            }
            template<typename TI, typename = std::enable_if_t<std::is_integral_v<TI>>>
            [[nodiscard]] auto operator[](TI i) -> TElem&
            {
                static_assert(TDim::value == 1u, "operator[](i) is only valid for 1-dimensional views");
                return data()[i];
            }
            template<typename TI, typename = std::enable_if_t<std::is_integral_v<TI>>>
            [[nodiscard]] auto operator[](TI i) const -> TElem const&
            {
                static_assert(TDim::value == 1u, "operator[](i) is only valid for 1-dimensional views");
                return data()[i];
            }
            template<typename TI>
            [[nodiscard]] auto operator[](Vec<TDim, TI> const& idx) -> TElem&
            {
                return *elemPtr(idx);
            }
            template<typename TI>
            [[nodiscard]] auto operator[](Vec<TDim, TI> const& idx) const -> TElem const&
            {
                return *const_cast<ViewOps*>(this)->elemPtr(idx);
            }
            template<typename TI>
            [[nodiscard]] auto at(Vec<TDim, TI> const& idx) -> TElem&
            {
                auto const ext = static_cast<TView const*>(this)->extents();
                for(std::size_t d = 0; d < TDim::value; ++d)
                    if(static_cast<TIdx>(idx[d]) >= ext[d])
                        throw std::out_of_range("alpaka view index out of range");
                return *elemPtr(idx);
            }
            template<typename TI, typename = std::enable_if_t<std::is_integral_v<TI>>>
            [[nodiscard]] auto at(TI i) -> TElem&
            {
                return at(Vec<TDim, TI>{i});
            }
            template<typename TI>
            [[nodiscard]] auto at(Vec<TDim, TI> const& idx) const -> TElem const&
            {
                return const_cast<ViewOps*>(this)->at(idx);
            }
            template<typename TI, typename = std::enable_if_t<std::is_integral_v<TI>>>
            [[nodiscard]] auto at(TI i) const -> TElem const&
            {
                return const_cast<ViewOps*>(this)->at(Vec<TDim, TI>{i});
            }
            //! contiguous iteration is only meaningful for unpadded views
            [[nodiscard]] auto begin() -> TElem*
            {
                return data();
            }
            [[nodiscard]] auto end() -> TElem*
            {
                return data() + static_cast<TView const*>(this)->extents().prod();
            }
            [[nodiscard]] auto begin() const -> TElem const*
            {
                return data();
            }
            [[nodiscard]] auto end() const -> TElem const*
            {
                return data() + static_cast<TView const*>(this)->extents().prod();
            }

        private:
            template<typename TI>
            auto elemPtr(Vec<TDim, TI> const& idx) -> TElem*
            {
                auto const pitches = static_cast<TView const*>(this)->pitchesInBytes();
                std::size_t off = 0;
                for(std::size_t d = 0; d < TDim::value; ++d)
                    off += static_cast<std::size_t>(idx[d]) * static_cast<std::size_t>(pitches[d]);
                return reinterpret_cast<TElem*>(
                    reinterpret_cast<char*>(const_cast<std::remove_const_t<TElem>*>(data())) + off);
            }
        };
    } // namespace b200

    // -----------------------------------------------------------------------------------------------------------
    //! Host buffer (unpadded rows). Plain allocations are 64-byte aligned; mapped allocations are pinned.
    template<typename TElem, typename TDim, typename TIdx>
    class BufCpu : public b200::ViewOps<BufCpu<TElem, TDim, TIdx>, TElem, TDim, TIdx>
    {
        static_assert(!std::is_const_v<TElem>, "The elem type of the buffer must not be const");

    public:
        template<typename TDeleter>
        BufCpu(DevCpu const& dev, TElem* ptr, TDeleter deleter, Vec<TDim, TIdx> const& extent)
            : m_dev(dev)
            , m_extent(extent)
            , m_mem(ptr, std::move(deleter))
        {
        }
        [[nodiscard]] auto nativePtr() const -> TElem*
        {
            return m_mem.get();
        }
        [[nodiscard]] auto extents() const -> Vec<TDim, TIdx>
        {
            return m_extent;
        }
        [[nodiscard]] auto pitchesInBytes() const -> Vec<TDim, TIdx>
        {
            std::size_t const row = TDim::value > 0u ? sizeof(TElem) * static_cast<std::size_t>(m_extent[TDim::value - 1u]) : sizeof(TElem);
            return b200::pitchesFromRowPitch<TElem>(m_extent, row);
        }
        DevCpu m_dev;
        Vec<TDim, TIdx> m_extent;
        std::shared_ptr<TElem> m_mem;
    };

    //! Device buffer of the B200 back-end.
    template<typename TElem, typename TDim, typename TIdx>
    class BufB200 : public b200::ViewOps<BufB200<TElem, TDim, TIdx>, TElem, TDim, TIdx>
    {
        static_assert(!std::is_const_v<TElem>, "The elem type of the buffer must not be const");

    public:
        template<typename TDeleter>
        BufB200(DevB200 const& dev, TElem* ptr, TDeleter deleter, Vec<TDim, TIdx> const& extent, std::size_t rowPitchBytes)
            : m_dev(dev)
            , m_extent(extent)
            , m_rowPitchBytes(rowPitchBytes)
            , m_mem(ptr, std::move(deleter))
        {
        }
        [[nodiscard]] auto nativePtr() const -> TElem*
        {
            return m_mem.get();
        }
        [[nodiscard]] auto extents() const -> Vec<TDim, TIdx>
        {
            return m_extent;
        }
        [[nodiscard]] auto pitchesInBytes() const -> Vec<TDim, TIdx>
        {
            return b200::pitchesFromRowPitch<TElem>(m_extent, m_rowPitchBytes);
        }
        DevB200 m_dev;
        Vec<TDim, TIdx> m_extent;
        std::size_t m_rowPitchBytes;
        std::shared_ptr<TElem> m_mem;
    };
    template<typename TElem, typename TDim, typename TIdx>
    using BufCudaRt = BufB200<TElem, TDim, TIdx>;

    //! Non-owning view of memory on TDev.
    template<typename TDev, typename TElem, typename TDim, typename TIdx>
    class ViewPlainPtr : public b200::ViewOps<ViewPlainPtr<TDev, TElem, TDim, TIdx>, TElem, TDim, TIdx>
    {
    public:
        ViewPlainPtr(TElem* ptr, TDev dev, Vec<TDim, TIdx> const& extent)
            : m_ptr(ptr)
            , m_dev(std::move(dev))
            , m_extent(extent)
            , m_pitches(b200::pitchesFromRowPitch<TElem>(
                  extent,
                  TDim::value > 0u ? sizeof(TElem) * static_cast<std::size_t>(extent[TDim::value - 1u]) : sizeof(TElem)))
        {
        }
        ViewPlainPtr(TElem* ptr, TDev dev, Vec<TDim, TIdx> const& extent, Vec<TDim, TIdx> const& pitchesBytes)
            : m_ptr(ptr)
            , m_dev(std::move(dev))
            , m_extent(extent)
            , m_pitches(pitchesBytes)
        {
        }
        [[nodiscard]] auto nativePtr() const -> TElem*
        {
            return m_ptr;
        }
        [[nodiscard]] auto extents() const -> Vec<TDim, TIdx>
        {
            return m_extent;
        }
        [[nodiscard]] auto pitchesInBytes() const -> Vec<TDim, TIdx>
        {
            return m_pitches;
        }
        TElem* m_ptr;
        TDev m_dev;
        Vec<TDim, TIdx> m_extent;
        Vec<TDim, TIdx> m_pitches;
    };

    //! A window [offset, offset + extent) of another view; shares its pitches.
    template<typename TDev, typename TElem, typename TDim, typename TIdx>
    class ViewSubView : public b200::ViewOps<ViewSubView<TDev, TElem, TDim, TIdx>, TElem, TDim, TIdx>
    {
    public:
        template<typename TView>
        ViewSubView(TView& view, Vec<TDim, TIdx> const& extent, Vec<TDim, TIdx> const& offset = Vec<TDim, TIdx>::zeros())
            : m_dev(trait::GetDev<std::remove_const_t<TView>>::getDev(view))
            , m_extent(extent)
            , m_offset(offset)
            , m_pitches(view.pitchesInBytes())
        {
            static_assert(std::is_same_v<Dim<TView>, TDim>, "The sub-view must have the dimensionality of its parent");
            auto const parentExtent = view.extents();
            for(std::size_t d = 0; d < TDim::value; ++d)
                ALPAKA_ASSERT(offset[d] + extent[d] <= parentExtent[d]);
            std::size_t off = 0;
            for(std::size_t d = 0; d < TDim::value; ++d)
                off += static_cast<std::size_t>(offset[d]) * static_cast<std::size_t>(m_pitches[d]);
            m_ptr = reinterpret_cast<TElem*>(
                reinterpret_cast<char*>(const_cast<std::remove_const_t<TElem>*>(view.nativePtr())) + off);
        }
        template<typename TView>
        explicit ViewSubView(TView& view) : ViewSubView(view, view.extents())
        {
        }
        [[nodiscard]] auto nativePtr() const -> TElem*
        {
            return m_ptr;
        }
        [[nodiscard]] auto extents() const -> Vec<TDim, TIdx>
        {
            return m_extent;
        }
        [[nodiscard]] auto offsets() const -> Vec<TDim, TIdx>
        {
            return m_offset;
        }
        [[nodiscard]] auto pitchesInBytes() const -> Vec<TDim, TIdx>
        {
            return m_pitches;
        }
        TElem* m_ptr = nullptr;
        TDev m_dev;
        Vec<TDim, TIdx> m_extent;
        Vec<TDim, TIdx> m_offset;
        Vec<TDim, TIdx> m_pitches;
    };

    //! Read-only wrapper around another view (held by value): same memory, element type `Elem const`
    //! (reference: mem/view/ViewConst.hpp:22-60). Wrapping a ViewConst again yields the same type.
    template<typename TView>
    class ViewConst : public b200::ViewOps<ViewConst<TView>, std::add_const_t<Elem<TView>>, Dim<TView>, Idx<TView>>
    {
        static_assert(!std::is_const_v<TView>, "ViewConst must be instantiated with a non-const view type");
        static_assert(!std::is_reference_v<TView>, "ViewConst must be instantiated with a non-reference view type");

    public:
        ViewConst(TView const& view) : m_view(view)
        {
        }
        ViewConst(TView&& view) : m_view(std::move(view))
        {
        }
        [[nodiscard]] auto nativePtr() const -> Elem<TView> const*
        {
            return trait::GetPtrNative<TView>::getPtrNative(m_view);
        }
        [[nodiscard]] auto extents() const -> Vec<Dim<TView>, Idx<TView>>
        {
            return trait::GetExtents<TView>{}(m_view);
        }
        [[nodiscard]] auto pitchesInBytes() const -> Vec<Dim<TView>, Idx<TView>>
        {
            return trait::GetPitchesInBytes<TView>{}(m_view);
        }
        TView m_view;
    };
    template<typename TView>
    ViewConst(TView) -> ViewConst<std::decay_t<TView>>;
    template<typename TView>
    ViewConst(ViewConst<TView>) -> ViewConst<std::decay_t<TView>>;

    // -----------------------------------------------------------------------------------------------------------
    namespace detail
    {
        //! byte pitches of a dense (unpadded) row-major array of TElem (reference: mem/view/Traits.hpp:35-49)
        template<typename TElem, typename TDim, typename TIdx>
        [[nodiscard]] ALPAKA_FN_HOST_ACC constexpr auto calculatePitchesFromExtents(Vec<TDim, TIdx> const& extent)
            -> Vec<TDim, TIdx>
        {
            Vec<TDim, TIdx> p{};
            TIdx stride = static_cast<TIdx>(sizeof(TElem));
            for(std::size_t d = TDim::value; d-- > 0u;)
            {
                p[d] = stride;
                stride = static_cast<TIdx>(stride * extent[d]);
            }
            return p;
        }

        template<typename T>
        inline constexpr bool isB200View = false;
        template<typename E, typename D, typename I>
        inline constexpr bool isB200View<BufCpu<E, D, I>> = true;
        template<typename E, typename D, typename I>
        inline constexpr bool isB200View<BufB200<E, D, I>> = true;
        template<typename V, typename E, typename D, typename I>
        inline constexpr bool isB200View<ViewPlainPtr<V, E, D, I>> = true;
        template<typename V, typename E, typename D, typename I>
        inline constexpr bool isB200View<ViewSubView<V, E, D, I>> = true;
        template<typename V>
        inline constexpr bool isB200View<ViewConst<V>> = true;
    } // namespace detail

    namespace concepts
    {
        template<typename T>
        concept View = detail::isB200View<std::remove_cv_t<std::remove_reference_t<T>>>;
    }

    namespace trait
    {
#define ALPAKA_B200_VIEW_TRAITS(TPL, VIEW, DEV, ELEM, DIM, IDX)                                                       \
    template<TPL>                                                                                                     \
    struct DevType<VIEW>                                                                                              \
    {                                                                                                                 \
        using type = DEV;                                                                                             \
    };                                                                                                                \
    template<TPL>                                                                                                     \
    struct DimType<VIEW>                                                                                              \
    {                                                                                                                 \
        using type = DIM;                                                                                             \
    };                                                                                                                \
    template<TPL>                                                                                                     \
    struct IdxType<VIEW>                                                                                              \
    {                                                                                                                 \
        using type = IDX;                                                                                             \
    };                                                                                                                \
    template<TPL>                                                                                                     \
    struct ElemType<VIEW>                                                                                             \
    {                                                                                                                 \
        using type = ELEM;                                                                                            \
    };                                                                                                                \
    template<TPL>                                                                                                     \
    struct GetDev<VIEW>                                                                                               \
    {                                                                                                                 \
        static auto getDev(VIEW const& v) -> DEV                                                                      \
        {                                                                                                             \
            return v.m_dev;                                                                                           \
        }                                                                                                             \
    };                                                                                                                \
    template<TPL>                                                                                                     \
    struct GetExtents<VIEW>                                                                                           \
    {                                                                                                                 \
        auto operator()(VIEW const& v) const -> Vec<DIM, IDX>                                                         \
        {                                                                                                             \
            return v.extents();                                                                                       \
        }                                                                                                             \
    };                                                                                                                \
    template<TPL>                                                                                                     \
    struct GetPtrNative<VIEW>                                                                                         \
    {                                                                                                                 \
        static auto getPtrNative(VIEW const& v) -> ELEM const*                                                        \
        {                                                                                                             \
            return v.nativePtr();                                                                                     \
        }                                                                                                             \
        static auto getPtrNative(VIEW& v) -> ELEM*                                                                    \
        {                                                                                                             \
            return v.nativePtr();                                                                                     \
        }                                                                                                             \
    };                                                                                                                \
    template<TPL>                                                                                                     \
    struct GetPitchesInBytes<VIEW>                                                                                    \
    {                                                                                                                 \
        auto operator()(VIEW const& v) const -> Vec<DIM, IDX>                                                         \
        {                                                                                                             \
            return v.pitchesInBytes();                                                                                \
        }                                                                                                             \
    };

#define ALPAKA_B200_COMMA ,
        ALPAKA_B200_VIEW_TRAITS(
            typename E ALPAKA_B200_COMMA typename D ALPAKA_B200_COMMA typename I,
            BufCpu<E ALPAKA_B200_COMMA D ALPAKA_B200_COMMA I>,
            DevCpu,
            E,
            D,
            I)
        ALPAKA_B200_VIEW_TRAITS(
            typename E ALPAKA_B200_COMMA typename D ALPAKA_B200_COMMA typename I,
            BufB200<E ALPAKA_B200_COMMA D ALPAKA_B200_COMMA I>,
            DevB200,
            E,
            D,
            I)
        ALPAKA_B200_VIEW_TRAITS(
            typename V ALPAKA_B200_COMMA typename E ALPAKA_B200_COMMA typename D ALPAKA_B200_COMMA typename I,
            ViewPlainPtr<V ALPAKA_B200_COMMA E ALPAKA_B200_COMMA D ALPAKA_B200_COMMA I>,
            V,
            E,
            D,
            I)
        ALPAKA_B200_VIEW_TRAITS(
            typename V ALPAKA_B200_COMMA typename E ALPAKA_B200_COMMA typename D ALPAKA_B200_COMMA typename I,
            ViewSubView<V ALPAKA_B200_COMMA E ALPAKA_B200_COMMA D ALPAKA_B200_COMMA I>,
            V,
            E,
            D,
            I)
#undef ALPAKA_B200_VIEW_TRAITS

        template<typename V>
        struct DevType<ViewConst<V>> : DevType<V>
        {
        };
        template<typename V>
        struct DimType<ViewConst<V>> : DimType<V>
        {
        };
        template<typename V>
        struct IdxType<ViewConst<V>> : IdxType<V>
        {
        };
        template<typename V>
        struct ElemType<ViewConst<V>>
        {
            using type = std::add_const_t<typename ElemType<V>::type>;
        };
        template<typename V>
        struct GetDev<ViewConst<V>>
        {
            static auto getDev(ViewConst<V> const& v)
            {
                return alpaka::getDev(v.m_view);
            }
        };
        template<typename V>
        struct GetExtents<ViewConst<V>>
        {
            auto operator()(ViewConst<V> const& v) const
            {
                return v.extents();
            }
        };
        template<typename V>
        struct GetOffsets<ViewConst<V>>
        {
            auto operator()(ViewConst<V> const& v) const
            {
                return alpaka::getOffsets(v.m_view);
            }
        };
        //! there is no mutable access through a ViewConst
        template<typename V>
        struct GetPtrNative<ViewConst<V>>
        {
            using E = typename ElemType<V>::type;
            static auto getPtrNative(ViewConst<V> const& v) -> E const*
            {
                return v.nativePtr();
            }
        };
        template<typename V>
        struct GetPitchesInBytes<ViewConst<V>>
        {
            auto operator()(ViewConst<V> const& v) const
            {
                return v.pitchesInBytes();
            }
        };

        // ---- std::vector and std::array are 1-D host views (reference: mem/view/ViewStdVector.hpp, ViewStdArray.hpp)
        template<typename TElem, typename TAlloc>
        struct DevType<std::vector<TElem, TAlloc>>
        {
            using type = DevCpu;
        };
        template<typename TElem, typename TAlloc>
        struct DimType<std::vector<TElem, TAlloc>>
        {
            using type = DimInt<1u>;
        };
        template<typename TElem, typename TAlloc>
        struct IdxType<std::vector<TElem, TAlloc>>
        {
            using type = std::size_t;
        };
        template<typename TElem, typename TAlloc>
        struct ElemType<std::vector<TElem, TAlloc>>
        {
            using type = TElem;
        };
        template<typename TElem, typename TAlloc>
        struct GetDev<std::vector<TElem, TAlloc>>
        {
            static auto getDev(std::vector<TElem, TAlloc> const&) -> DevCpu
            {
                return DevCpu{};
            }
        };
        template<typename TElem, typename TAlloc>
        struct GetExtents<std::vector<TElem, TAlloc>>
        {
            auto operator()(std::vector<TElem, TAlloc> const& v) const -> Vec<DimInt<1u>, std::size_t>
            {
                return Vec<DimInt<1u>, std::size_t>{v.size()};
            }
        };
        template<typename TElem, typename TAlloc>
        struct GetOffsets<std::vector<TElem, TAlloc>>
        {
            auto operator()(std::vector<TElem, TAlloc> const&) const -> Vec<DimInt<1u>, std::size_t>
            {
                return Vec<DimInt<1u>, std::size_t>{std::size_t{0}};
            }
        };
        template<typename TElem, typename TAlloc>
        struct GetPtrNative<std::vector<TElem, TAlloc>>
        {
            static auto getPtrNative(std::vector<TElem, TAlloc> const& v) -> TElem const*
            {
                return v.data();
            }
            static auto getPtrNative(std::vector<TElem, TAlloc>& v) -> TElem*
            {
                return v.data();
            }
        };
        template<typename TElem, typename TAlloc>
        struct GetPitchesInBytes<std::vector<TElem, TAlloc>>
        {
            auto operator()(std::vector<TElem, TAlloc> const&) const -> Vec<DimInt<1u>, std::size_t>
            {
                return Vec<DimInt<1u>, std::size_t>{sizeof(TElem)};
            }
        };
        template<typename TElem, std::size_t N>
        struct DevType<std::array<TElem, N>>
        {
            using type = DevCpu;
        };
        template<typename TElem, std::size_t N>
        struct DimType<std::array<TElem, N>>
        {
            using type = DimInt<1u>;
        };
        template<typename TElem, std::size_t N>
        struct IdxType<std::array<TElem, N>>
        {
            using type = std::size_t;
        };
        template<typename TElem, std::size_t N>
        struct ElemType<std::array<TElem, N>>
        {
            using type = TElem;
        };
        template<typename TElem, std::size_t N>
        struct GetDev<std::array<TElem, N>>
        {
            static auto getDev(std::array<TElem, N> const&) -> DevCpu
            {
                return DevCpu{};
            }
        };
        template<typename TElem, std::size_t N>
        struct GetExtents<std::array<TElem, N>>
        {
            auto operator()(std::array<TElem, N> const&) const -> Vec<DimInt<1u>, std::size_t>
            {
                return Vec<DimInt<1u>, std::size_t>{N};
            }
        };
        template<typename TElem, std::size_t N>
        struct GetOffsets<std::array<TElem, N>>
        {
            auto operator()(std::array<TElem, N> const&) const -> Vec<DimInt<1u>, std::size_t>
            {
                return Vec<DimInt<1u>, std::size_t>{std::size_t{0}};
            }
        };
        template<typename TElem, std::size_t N>
        struct GetPtrNative<std::array<TElem, N>>
        {
            static auto getPtrNative(std::array<TElem, N> const& v) -> TElem const*
            {
                return v.data();
            }
            static auto getPtrNative(std::array<TElem, N>& v) -> TElem*
            {
                return v.data();
            }
        };
        template<typename TElem, std::size_t N>
        struct GetPitchesInBytes<std::array<TElem, N>>
        {
            auto operator()(std::array<TElem, N> const&) const -> Vec<DimInt<1u>, std::size_t>
            {
                return Vec<DimInt<1u>, std::size_t>{sizeof(TElem)};
            }
        };

        template<typename V, typename E, typename D, typename I>
        struct GetOffsets<ViewSubView<V, E, D, I>>
        {
            auto operator()(ViewSubView<V, E, D, I> const& v) const -> Vec<D, I>
            {
                return v.offsets();
            }
        };
        template<typename E, typename D, typename I>
        struct GetOffsets<BufCpu<E, D, I>>
        {
            auto operator()(BufCpu<E, D, I> const&) const -> Vec<D, I>
            {
                return Vec<D, I>::zeros();
            }
        };
        template<typename E, typename D, typename I>
        struct GetOffsets<BufB200<E, D, I>>
        {
            auto operator()(BufB200<E, D, I> const&) const -> Vec<D, I>
            {
                return Vec<D, I>::zeros();
            }
        };
        template<typename V, typename E, typename D, typename I>
        struct GetOffsets<ViewPlainPtr<V, E, D, I>>
        {
            auto operator()(ViewPlainPtr<V, E, D, I> const&) const -> Vec<D, I>
            {
                return Vec<D, I>::zeros();
            }
        };

        template<typename TElem, typename TDim, typename TIdx>
        struct BufType<DevCpu, TElem, TDim, TIdx>
        {
            using type = BufCpu<TElem, TDim, TIdx>;
        };
        template<typename TElem, typename TDim, typename TIdx>
        struct BufType<DevB200, TElem, TDim, TIdx>
        {
            using type = BufB200<TElem, TDim, TIdx>;
        };

        template<typename TDim>
        struct HasAsyncBufSupport<TDim, DevB200> : std::true_type
        {
        };
        template<typename TDim>
        struct HasAsyncBufSupport<TDim, DevCpu> : std::true_type
        {
        };
        template<>
        struct HasMappedBufSupport<PlatformB200> : std::true_type
        {
        };
    } // namespace trait

    template<typename TView>
    [[nodiscard]] auto getPtrNative(TView const& view) -> Elem<TView> const*
    {
        return trait::GetPtrNative<TView>::getPtrNative(view);
    }
    template<typename TView>
    [[nodiscard]] auto getPtrNative(TView& view) -> Elem<TView>*
    {
        return trait::GetPtrNative<TView>::getPtrNative(view);
    }
    //! pointer usable on `dev`: device memory and pinned-mapped host memory share one address space (UVA)
    template<typename TView, typename TDev>
    [[nodiscard]] auto getPtrDev(TView const& view, TDev const&) -> Elem<TView> const*
    {
        return getPtrNative(view);
    }
    template<typename TView, typename TDev>
    [[nodiscard]] auto getPtrDev(TView& view, TDev const&) -> Elem<TView>*
    {
        return getPtrNative(view);
    }
    //! pitch[d] = bytes between two neighbouring elements in dimension d; pitch[Dim-1] == sizeof(Elem)
    template<typename TView>
    [[nodiscard]] auto getPitchesInBytes(TView const& view) -> Vec<Dim<TView>, Idx<TView>>
    {
        return trait::GetPitchesInBytes<TView>{}(view);
    }

    // -----------------------------------------------------------------------------------------------------------
    // allocation
    namespace trait
    {
        template<typename TElem, typename TDim, typename TIdx>
        struct BufAlloc<TElem, TDim, TIdx, DevCpu>
        {
            template<typename TExtent>
            static auto allocBuf(DevCpu const& dev, TExtent const& extent) -> BufCpu<TElem, TDim, TIdx>
            {
                auto const ext = getExtents(extent);
                std::size_t const bytes = sizeof(TElem) * static_cast<std::size_t>(ext.prod());
                constexpr std::size_t alignment = alignof(TElem) > 64u ? alignof(TElem) : 64u;
                void* p = nullptr;
                if(bytes != 0u)
                {
                    p = std::aligned_alloc(alignment, (bytes + alignment - 1u) / alignment * alignment);
                    if(p == nullptr)
                        throw std::bad_alloc();
                }
                return BufCpu<TElem, TDim, TIdx>(dev, static_cast<TElem*>(p), [](TElem* q) { std::free(q); }, ext);
            }
        };
        template<typename TElem, typename TDim, typename TIdx>
        struct AsyncBufAlloc<TElem, TDim, TIdx, DevCpu>
        {
            template<typename TQueue, typename TExtent>
            static auto allocAsyncBuf(TQueue const& queue, TExtent const& extent) -> BufCpu<TElem, TDim, TIdx>
            {
                return BufAlloc<TElem, TDim, TIdx, DevCpu>::allocBuf(getDev(queue), extent);
            }
        };

        template<typename TElem, typename TDim, typename TIdx>
        struct BufAlloc<TElem, TDim, TIdx, DevB200>
        {
            template<typename TExtent>
            static auto allocBuf(DevB200 const& dev, TExtent const& extent) -> BufB200<TElem, TDim, TIdx>
            {
                auto const ext = getExtents(extent);
                int const d = dev.getNativeHandle();
                void* p = nullptr;
                std::size_t rowPitch = sizeof(TElem);
                if constexpr(TDim::value <= 1u)
                {
                    std::size_t const bytes = sizeof(TElem) * static_cast<std::size_t>(ext.prod());
                    rowPitch = bytes;
                    b200::check(b200_malloc_async(d, nullptr, bytes, &p));
                }
                else
                {
                    std::size_t const widthBytes = sizeof(TElem) * static_cast<std::size_t>(ext[TDim::value - 1u]);
                    std::size_t rows = 1;
                    for(std::size_t k = 0; k + 1u < TDim::value; ++k)
                        rows *= static_cast<std::size_t>(ext[k]);
                    b200::check(b200_malloc_pitched_async(d, nullptr, widthBytes, rows, &p, &rowPitch));
                }
                // allocation ordered on the legacy stream; make it visible to every (non-blocking) queue
                b200::check(b200_stream_sync(nullptr));
                auto deleter = [d](TElem* q)
                {
                    if(q == nullptr)
                        return;
                    // cudaFree semantics: outstanding work on any queue may still use the memory
                    b200::checkNoexcept(b200_device_sync(d));
                    b200::checkNoexcept(b200_free_async(d, nullptr, q));
                };
                return BufB200<TElem, TDim, TIdx>(dev, static_cast<TElem*>(p), deleter, ext, rowPitch);
            }
        };
        template<typename TElem, typename TDim, typename TIdx>
        struct AsyncBufAlloc<TElem, TDim, TIdx, DevB200>
        {
            template<typename TQueue, typename TExtent>
            static auto allocAsyncBuf(TQueue queue, TExtent const& extent) -> BufB200<TElem, TDim, TIdx>
            {
                auto const ext = getExtents(extent);
                DevB200 const dev = getDev(queue);
                int const d = dev.getNativeHandle();
                void* p = nullptr;
                std::size_t rowPitch = sizeof(TElem);
                if constexpr(TDim::value <= 1u)
                {
                    std::size_t const bytes = sizeof(TElem) * static_cast<std::size_t>(ext.prod());
                    rowPitch = bytes;
                    b200::check(b200_malloc_async(d, queue.getNativeHandle(), bytes, &p));
                }
                else
                {
                    std::size_t const widthBytes = sizeof(TElem) * static_cast<std::size_t>(ext[TDim::value - 1u]);
                    std::size_t rows = 1;
                    for(std::size_t k = 0; k + 1u < TDim::value; ++k)
                        rows *= static_cast<std::size_t>(ext[k]);
                    b200::check(b200_malloc_pitched_async(d, queue.getNativeHandle(), widthBytes, rows, &p, &rowPitch));
                }
                // the deleter owns a copy of the queue: the free is stream-ordered behind everything enqueued so far
                // (reference: mem/buf/BufUniformCudaHipRt.hpp:310-316)
                auto deleter = [d, queue](TElem* q)
                {
                    if(q != nullptr)
                        b200::checkNoexcept(b200_free_async(d, queue.getNativeHandle(), q));
                };
                return BufB200<TElem, TDim, TIdx>(dev, static_cast<TElem*>(p), deleter, ext, rowPitch);
            }
        };
    } // namespace trait

    //! Allocates memory on the given device.
    template<typename TElem, typename TIdx, typename TExtent, typename TDev>
    [[nodiscard]] auto allocBuf(TDev const& dev, TExtent const& extent = TExtent())
    {
        return trait::BufAlloc<TElem, Dim<TExtent>, TIdx, TDev>::allocBuf(dev, extent);
    }
    //! Allocates stream-ordered memory: usable by work enqueued to `queue` after this call.
    template<typename TElem, typename TIdx, typename TExtent, typename TQueue>
    [[nodiscard]] auto allocAsyncBuf(TQueue queue, TExtent const& extent = TExtent())
    {
        return trait::AsyncBufAlloc<TElem, Dim<TExtent>, TIdx, Dev<TQueue>>::allocAsyncBuf(queue, extent);
    }
    template<typename TDev, typename TDim>
    inline constexpr bool hasAsyncBufSupport = trait::HasAsyncBufSupport<TDim, TDev>::value;
    template<typename TElem, typename TIdx, typename TExtent, typename TQueue>
    [[nodiscard]] auto allocAsyncBufIfSupported(TQueue queue, TExtent const& extent = TExtent())
    {
        return allocAsyncBuf<TElem, TIdx>(queue, extent);
    }
    template<typename TPlatform>
    inline constexpr bool hasMappedBufSupport = trait::HasMappedBufSupport<TPlatform>::value;

    //! Pinned host memory, mapped into the address space of the platform's devices.
    template<typename TElem, typename TIdx, typename TExtent>
    [[nodiscard]] auto allocMappedBuf(DevCpu const& host, PlatformB200 const&, TExtent const& extent = TExtent())
        -> BufCpu<TElem, Dim<TExtent>, TIdx>
    {
        auto const ext = getExtents(extent);
        void* p = nullptr;
        b200::check(b200_host_alloc_pinned(sizeof(TElem) * static_cast<std::size_t>(ext.prod()), &p));
        return BufCpu<TElem, Dim<TExtent>, TIdx>(
            host,
            static_cast<TElem*>(p),
            [](TElem* q) { b200::checkNoexcept(b200_host_free_pinned(q)); },
            ext);
    }
    template<typename TElem, typename TIdx, typename TExtent, typename TPlatform>
    [[nodiscard]] auto allocMappedBufIfSupported(DevCpu const& host, TPlatform const& platform, TExtent const& extent = TExtent())
    {
        if constexpr(hasMappedBufSupport<TPlatform>)
            return allocMappedBuf<TElem, TIdx>(host, platform, extent);
        else
            return allocBuf<TElem, TIdx>(host, extent);
    }

    // ---- views
    template<typename TDev, typename TElem, typename TExtent>
    [[nodiscard]] auto createView(TDev const& dev, TElem* pMem, TExtent const& extent)
    {
        using D = Dim<TExtent>;
        using I = Idx<TExtent>;
        return ViewPlainPtr<TDev, TElem, D, I>(pMem, dev, getExtents(extent));
    }
    template<typename TDev, typename TElem, typename TExtent, typename TPitch>
    [[nodiscard]] auto createView(TDev const& dev, TElem* pMem, TExtent const& extent, TPitch pitch)
    {
        using D = Dim<TExtent>;
        using I = Idx<TExtent>;
        return ViewPlainPtr<TDev, TElem, D, I>(pMem, dev, getExtents(extent), castVec<I>(getExtents(pitch)));
    }
    template<typename TDev, typename TElem, typename TAlloc>
    [[nodiscard]] auto createView(TDev const& dev, std::vector<TElem, TAlloc>& con)
    {
        return createView(dev, con.data(), con.size());
    }
    template<typename TDev, typename TElem, std::size_t N>
    [[nodiscard]] auto createView(TDev const& dev, std::array<TElem, N>& con)
    {
        return createView(dev, con.data(), N);
    }
    template<typename TDev, typename TContainer, typename TExtent>
    [[nodiscard]] auto createView(TDev const& dev, TContainer& con, TExtent const& extent)
        -> decltype(createView(dev, std::data(con), extent))
    {
        return createView(dev, std::data(con), extent);
    }
    template<typename TView, typename TExtent, typename TOffsets>
    [[nodiscard]] auto createSubView(TView& view, TExtent const& extent, TOffsets const& offset = TExtent())
    {
        using D = Dim<TView>;
        using I = Idx<TView>;
        return ViewSubView<Dev<TView>, Elem<TView>, D, I>(view, castVec<I>(getExtents(extent)), castVec<I>(getOffsets(offset)));
    }

    // -----------------------------------------------------------------------------------------------------------
    // copy / set tasks
    namespace b200
    {
        template<typename TDev>
        inline constexpr bool isHost = std::is_same_v<TDev, DevCpu>;

        //! N-d strided region: base pointer, byte pitches, extent in elements
        struct Region
        {
            char* base = nullptr;
            std::size_t pitch[8] = {};
        };

        //! Copy of an N-d box between two views. Issued as one 1-D copy when rows are contiguous on both sides,
        //! otherwise as 2-D copies over the two fastest dimensions for each outer index.
        template<typename TDim>
        struct TaskCopy
        {
            Region dst, src;
            std::size_t extent[TDim::value == 0u ? 1u : TDim::value] = {};
            std::size_t elemBytes = 0;
            int kind = B200_COPY_DEFAULT;
            int devIssue = 0; //!< device whose context issues the copy (destination device, reference Copy.hpp:143)
            bool deviceInvolved = true;

            template<typename F1, typename F2>
            void forEachChunk(F1&& copy1d, F2&& copy2d) const
            {
                constexpr std::size_t n = TDim::value;
                if constexpr(n == 0u)
                {
                    copy1d(dst.base, src.base, elemBytes);
                }
                else
                {
                    for(std::size_t d = 0; d < n; ++d)
                        if(extent[d] == 0u)
                            return;
                    std::size_t const rowBytes = extent[n - 1u] * elemBytes;
                    if constexpr(n == 1u)
                    {
                        copy1d(dst.base, src.base, rowBytes);
                    }
                    else
                    {
                        std::size_t const rows = extent[n - 2u];
                        std::size_t outer = 1;
                        for(std::size_t d = 0; d + 2u < n; ++d)
                            outer *= extent[d];
                        for(std::size_t o = 0; o < outer; ++o)
                        {
                            std::size_t rest = o, offD = 0, offS = 0;
                            for(std::size_t d = n - 2u; d-- > 0u;)
                            {
                                std::size_t const i = rest % extent[d];
                                rest /= extent[d];
                                offD += i * dst.pitch[d];
                                offS += i * src.pitch[d];
                            }
                            if(dst.pitch[n - 2u] == rowBytes && src.pitch[n - 2u] == rowBytes)
                                copy1d(dst.base + offD, src.base + offS, rowBytes * rows);
                            else
                                copy2d(dst.base + offD, dst.pitch[n - 2u], src.base + offS, src.pitch[n - 2u], rowBytes, rows);
                        }
                    }
                }
            }

            void enqueueOn(b200_stream_t stream) const
            {
                forEachChunk(
                    [&](char* d, char const* s, std::size_t bytes)
                    { check(b200_memcpy_async(devIssue, d, s, bytes, kind, stream)); },
                    [&](char* d, std::size_t dp, char const* s, std::size_t sp, std::size_t w, std::size_t h)
                    { check(b200_memcpy2d_async(devIssue, d, dp, s, sp, w, h, kind, stream)); });
            }

            void runOnHost() const
            {
                forEachChunk(
                    [&](char* d, char const* s, std::size_t bytes) { std::memcpy(d, s, bytes); },
                    [&](char* d, std::size_t dp, char const* s, std::size_t sp, std::size_t w, std::size_t h)
                    {
                        for(std::size_t r = 0; r < h; ++r)
                            std::memcpy(d + r * dp, s + r * sp, w);
                    });
            }
        };

        template<typename TDim>
        struct TaskSet
        {
            Region dst;
            std::size_t extent[TDim::value == 0u ? 1u : TDim::value] = {};
            std::size_t elemBytes = 0;
            int byte = 0;
            int devIssue = 0;
            bool onHost = false;

            template<typename F1, typename F2>
            void forEachChunk(F1&& set1d, F2&& set2d) const
            {
                constexpr std::size_t n = TDim::value;
                if constexpr(n == 0u)
                {
                    set1d(dst.base, elemBytes);
                }
                else
                {
                    for(std::size_t d = 0; d < n; ++d)
                        if(extent[d] == 0u)
                            return;
                    std::size_t const rowBytes = extent[n - 1u] * elemBytes;
                    if constexpr(n == 1u)
                    {
                        set1d(dst.base, rowBytes);
                    }
                    else
                    {
                        std::size_t const rows = extent[n - 2u];
                        std::size_t outer = 1;
                        for(std::size_t d = 0; d + 2u < n; ++d)
                            outer *= extent[d];
                        for(std::size_t o = 0; o < outer; ++o)
                        {
                            std::size_t rest = o, off = 0;
                            for(std::size_t d = n - 2u; d-- > 0u;)
                            {
                                off += (rest % extent[d]) * dst.pitch[d];
                                rest /= extent[d];
                            }
                            if(dst.pitch[n - 2u] == rowBytes)
                                set1d(dst.base + off, rowBytes * rows);
                            else
                                set2d(dst.base + off, dst.pitch[n - 2u], rowBytes, rows);
                        }
                    }
                }
            }

            void enqueueOn(b200_stream_t stream) const
            {
                forEachChunk(
                    [&](char* d, std::size_t bytes) { check(b200_memset_async(devIssue, d, byte, bytes, stream)); },
                    [&](char* d, std::size_t p, std::size_t w, std::size_t h)
                    { check(b200_memset2d_async(devIssue, d, p, byte, w, h, stream)); });
            }

            void runOnHost() const
            {
                forEachChunk(
                    [&](char* d, std::size_t bytes) { std::memset(d, byte, bytes); },
                    [&](char* d, std::size_t p, std::size_t w, std::size_t h)
                    {
                        for(std::size_t r = 0; r < h; ++r)
                            std::memset(d + r * p, byte, w);
                    });
            }
        };
    } // namespace b200

    //! Creates the task that copies `extent` elements from viewSrc to viewDst (both may be padded).
    template<typename TExtent, typename TViewSrc, typename TViewDstFwd>
    [[nodiscard]] auto createTaskMemcpy(TViewDstFwd&& viewDst, TViewSrc const& viewSrc, TExtent const& extent)
    {
        using TViewDst = std::remove_reference_t<TViewDstFwd>;
        using D = Dim<TViewDst>;
        static_assert(!std::is_const_v<TViewDst>, "The destination view must not be const!");
        static_assert(!std::is_const_v<Elem<TViewDst>>, "The destination view's element type must not be const!");
        static_assert(
            D::value == Dim<TViewSrc>::value,
            "The source and the destination view are required to have the same dimensionality!");
        static_assert(
            D::value == Dim<TExtent>::value,
            "The views and the extent are required to have the same dimensionality!");
        static_assert(
            std::is_same_v<Elem<TViewDst>, std::remove_const_t<Elem<TViewSrc>>>,
            "The source and the destination view are required to have the same element type!");

        b200::TaskCopy<D> task;
        auto const ext = getExtents(extent);
        auto const pd = getPitchesInBytes(viewDst);
        auto const ps = getPitchesInBytes(viewSrc);
        auto const extDst = getExtents(viewDst);
        auto const extSrc = getExtents(viewSrc);
        for(std::size_t d = 0; d < D::value; ++d)
        {
            ALPAKA_ASSERT(static_cast<std::size_t>(ext[d]) <= static_cast<std::size_t>(extDst[d]));
            ALPAKA_ASSERT(static_cast<std::size_t>(ext[d]) <= static_cast<std::size_t>(extSrc[d]));
            task.extent[d] = static_cast<std::size_t>(ext[d]);
            task.dst.pitch[d] = static_cast<std::size_t>(pd[d]);
            task.src.pitch[d] = static_cast<std::size_t>(ps[d]);
        }
        task.dst.base = reinterpret_cast<char*>(getPtrNative(viewDst));
        task.src.base = const_cast<char*>(reinterpret_cast<char const*>(getPtrNative(viewSrc)));
        task.elemBytes = sizeof(Elem<TViewDst>);
        constexpr bool dstHost = b200::isHost<Dev<TViewDst>>;
        constexpr bool srcHost = b200::isHost<Dev<TViewSrc>>;
        task.kind = dstHost ? (srcHost ? B200_COPY_H2H : B200_COPY_D2H) : (srcHost ? B200_COPY_H2D : B200_COPY_D2D);
        task.deviceInvolved = !(dstHost && srcHost);
        if constexpr(!dstHost)
            task.devIssue = getDev(viewDst).getNativeHandle();
        else if constexpr(!srcHost)
            task.devIssue = getDev(viewSrc).getNativeHandle();
        return task;
    }

    template<typename TExtent, typename TViewFwd>
    [[nodiscard]] auto createTaskMemset(TViewFwd&& view, std::uint8_t const& byte, TExtent const& extent)
    {
        using TView = std::remove_reference_t<TViewFwd>;
        using D = Dim<TView>;
        static_assert(D::value == Dim<TExtent>::value, "The view and the extent are required to have the same dimensionality!");
        b200::TaskSet<D> task;
        auto const ext = getExtents(extent);
        auto const p = getPitchesInBytes(view);
        for(std::size_t d = 0; d < D::value; ++d)
        {
            task.extent[d] = static_cast<std::size_t>(ext[d]);
            task.dst.pitch[d] = static_cast<std::size_t>(p[d]);
        }
        task.dst.base = reinterpret_cast<char*>(getPtrNative(view));
        task.elemBytes = sizeof(Elem<TView>);
        task.byte = byte;
        task.onHost = b200::isHost<Dev<TView>>;
        if constexpr(!b200::isHost<Dev<TView>>)
            task.devIssue = getDev(view).getNativeHandle();
        return task;
    }

    namespace trait
    {
        template<typename TProperty, typename TDim>
        struct Enqueue<QueueB200<TProperty>, b200::TaskCopy<TDim>>
        {
            static void enqueue(QueueB200<TProperty>& q, b200::TaskCopy<TDim> const& task)
            {
                task.enqueueOn(q.getNativeHandle());
                q.afterEnqueue();
            }
        };
        template<typename TProperty, typename TDim>
        struct Enqueue<QueueCpu<TProperty>, b200::TaskCopy<TDim>>
        {
            static void enqueue(QueueCpu<TProperty>& q, b200::TaskCopy<TDim> const& task)
            {
                q.m_impl->run(
                    [task]
                    {
                        if(task.deviceInvolved)
                        {
                            // a host queue copying from/to device memory: synchronous copy on the legacy stream
                            task.enqueueOn(nullptr);
                            b200::check(b200_stream_sync(nullptr));
                        }
                        else
                            task.runOnHost();
                    });
            }
        };
        template<typename TProperty, typename TDim>
        struct Enqueue<QueueB200<TProperty>, b200::TaskSet<TDim>>
        {
            static void enqueue(QueueB200<TProperty>& q, b200::TaskSet<TDim> const& task)
            {
                if(task.onHost)
                {
                    // host memory set in stream order
                    auto copy = task;
                    alpaka::enqueue(q, [copy] { copy.runOnHost(); });
                    return;
                }
                task.enqueueOn(q.getNativeHandle());
                q.afterEnqueue();
            }
        };
        template<typename TProperty, typename TDim>
        struct Enqueue<QueueCpu<TProperty>, b200::TaskSet<TDim>>
        {
            static void enqueue(QueueCpu<TProperty>& q, b200::TaskSet<TDim> const& task)
            {
                q.m_impl->run(
                    [task]
                    {
                        if(task.onHost)
                            task.runOnHost();
                        else
                        {
                            task.enqueueOn(nullptr);
                            b200::check(b200_stream_sync(nullptr));
                        }
                    });
            }
        };
    } // namespace trait

    //! Copies `extent` elements from viewSrc to viewDst, in queue order.
    template<typename TExtent, typename TViewSrc, typename TViewDstFwd, typename TQueue>
    void memcpy(TQueue& queue, TViewDstFwd&& viewDst, TViewSrc const& viewSrc, TExtent const& extent)
    {
        enqueue(queue, createTaskMemcpy(std::forward<TViewDstFwd>(viewDst), viewSrc, extent));
    }
    //! Copies the whole destination extent.
    template<typename TViewSrc, typename TViewDstFwd, typename TQueue>
    void memcpy(TQueue& queue, TViewDstFwd&& viewDst, TViewSrc const& viewSrc)
    {
        enqueue(queue, createTaskMemcpy(std::forward<TViewDstFwd>(viewDst), viewSrc, getExtents(viewDst)));
    }
    template<typename TExtent, typename TViewFwd, typename TQueue>
    void memset(TQueue& queue, TViewFwd&& view, std::uint8_t const& byte, TExtent const& extent)
    {
        enqueue(queue, createTaskMemset(std::forward<TViewFwd>(view), byte, extent));
    }
    template<typename TViewFwd, typename TQueue>
    void memset(TQueue& queue, TViewFwd&& view, std::uint8_t const& byte)
    {
        enqueue(queue, createTaskMemset(std::forward<TViewFwd>(view), byte, getExtents(view)));
    }
} // namespace alpaka
