// include/alpaka/b200/Global.hpp -- variables in static device memory (`__device__` / `__constant__`) and the copies
// that initialise and read them.
//
// API parity with the reference's mem/global/Traits.hpp:13-46 (DevGlobal<TAcc, T>, get(), operator&) and
// mem/global/DeviceGlobalUniformCudaHipBuiltIn.hpp:43-200 (the four memcpy overloads), declared with the
// ALPAKA_STATIC_ACC_MEM_GLOBAL / ALPAKA_STATIC_ACC_MEM_CONSTANT macros of Config.hpp:
//
//     ALPAKA_STATIC_ACC_MEM_GLOBAL alpaka::DevGlobal<TAcc, float[16]> g_table;   // TAcc is supplied by the macro
//     ... in a kernel:  g_table<TAcc>.get()[i]
//     ... on the host:  alpaka::memcpy(queue, g_table<Acc>, hostView);           // and the reverse
//
// The host side never dereferences the variable: its device address comes from b200_symbol_address (the user's TU and
// libalpaka_b200.so share one cudart instance, see Kernel.hpp) and the copy itself is an ordinary TaskCopy.
#pragma once

#include "Mem.hpp"

namespace alpaka
{
    namespace detail
    {
        //! storage of a device global; T may be const-qualified (constant memory) and may be an array type
        template<typename TTag, typename T>
        struct DevGlobalImplGeneric
        {
            using Type = std::remove_const_t<T>;
            Type value;

            ALPAKA_FN_HOST_ACC auto operator&() -> T*
            {
                return &value;
            }
            ALPAKA_FN_HOST_ACC auto get() -> T&
            {
                return value;
            }
        };

        template<typename TTag, typename T>
        struct DevGlobalTrait;
        template<typename T>
        struct DevGlobalTrait<TagGpuB200, T>
        {
            using Type = DevGlobalImplGeneric<TagGpuB200, T>;
        };
    } // namespace detail

    template<typename TAcc, typename T>
    using DevGlobal = typename detail::DevGlobalTrait<typename trait::AccToTag<TAcc>::type, T>::Type;

    namespace b200
    {
        //! a plain-pointer device view over the storage of a device global, shaped by `extent`
        template<typename TProperty, typename T, typename TExtent>
        [[nodiscard]] auto viewOfGlobal(
            QueueB200<TProperty> const& queue,
            alpaka::detail::DevGlobalImplGeneric<TagGpuB200, T>& global,
            TExtent const& extent)
        {
            using E = std::remove_const_t<std::remove_all_extents_t<T>>;
            DevB200 const dev = getDev(queue);
            void* p = nullptr;
            check(b200_symbol_address(dev.getNativeHandle(), static_cast<void const*>(std::addressof(global.value)), &p));
            return ViewPlainPtr<DevB200, E, Dim<TExtent>, Idx<TExtent>>(static_cast<E*>(p), dev, getExtents(extent));
        }
    } // namespace b200

    //! device global -> view
    template<typename TProperty, typename TViewDst, typename T>
    void memcpy(QueueB200<TProperty>& queue, TViewDst& viewDst, alpaka::detail::DevGlobalImplGeneric<TagGpuB200, T>& src)
    {
        auto const extent = getExtents(viewDst);
        auto const view = b200::viewOfGlobal(queue, src, extent);
        enqueue(queue, createTaskMemcpy(viewDst, view, extent));
    }
    //! view -> device global
    template<typename TProperty, typename T, typename TViewSrc>
    void memcpy(QueueB200<TProperty>& queue, alpaka::detail::DevGlobalImplGeneric<TagGpuB200, T>& dst, TViewSrc const& viewSrc)
    {
        auto const extent = getExtents(viewSrc);
        auto view = b200::viewOfGlobal(queue, dst, extent);
        enqueue(queue, createTaskMemcpy(view, viewSrc, extent));
    }
    //! device global -> view, explicit extent
    template<typename TProperty, typename TViewDst, typename T, typename TExtent>
    void memcpy(
        QueueB200<TProperty>& queue,
        TViewDst& viewDst,
        alpaka::detail::DevGlobalImplGeneric<TagGpuB200, T>& src,
        TExtent const& extent)
    {
        auto const view = b200::viewOfGlobal(queue, src, extent);
        enqueue(queue, createTaskMemcpy(viewDst, view, extent));
    }
    //! view -> device global, explicit extent
    template<typename TProperty, typename T, typename TViewSrc, typename TExtent>
    void memcpy(
        QueueB200<TProperty>& queue,
        alpaka::detail::DevGlobalImplGeneric<TagGpuB200, T>& dst,
        TViewSrc const& viewSrc,
        TExtent const& extent)
    {
        auto view = b200::viewOfGlobal(queue, dst, extent);
        enqueue(queue, createTaskMemcpy(view, viewSrc, extent));
    }
} // namespace alpaka
