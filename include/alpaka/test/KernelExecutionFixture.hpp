// include/alpaka/test/KernelExecutionFixture.hpp -- runs a test kernel `kernel(acc, bool* success, args...)` on device 0
// of the accelerator's platform and returns the flag (reference: include/alpaka/test/KernelExecutionFixture.hpp:25-103).
#pragma once

#include <alpaka/alpaka.hpp>
#include <alpaka/test/Check.hpp>
#include <alpaka/test/queue/Queue.hpp>

#include <utility>

namespace alpaka::test
{
    //! The fixture for executing a kernel on a given accelerator.
    template<typename TAcc>
    class KernelExecutionFixture
    {
    public:
        using Acc = TAcc;
        using Dim = alpaka::Dim<Acc>;
        using Idx = alpaka::Idx<Acc>;
        using Platform = alpaka::Platform<Acc>;
        using Device = Dev<Acc>;
        using Queue = test::DefaultQueue<Device>;
        using WorkDiv = WorkDivMembers<Dim, Idx>;

        //! run with exactly this work division
        explicit KernelExecutionFixture(WorkDiv workDiv) : m_queue{m_device}, m_workDiv{std::move(workDiv)}, m_haveWorkDiv{true}
        {
        }
        //! run over this many grid elements, one element per thread, with a valid work division chosen per kernel
        template<typename TExtent, typename = std::enable_if_t<!std::is_same_v<std::decay_t<TExtent>, WorkDiv>>>
        explicit KernelExecutionFixture(TExtent const& extent) : m_queue{m_device}
                                                               , m_extent{castVec<Idx>(getExtents(extent))}
        {
        }
        KernelExecutionFixture(Queue queue, WorkDiv workDiv)
            : m_device{alpaka::getDev(queue)}
            , m_queue{std::move(queue)}
            , m_workDiv{std::move(workDiv)}
            , m_haveWorkDiv{true}
        {
        }
        template<typename TExtent, typename = std::enable_if_t<!std::is_same_v<std::decay_t<TExtent>, WorkDiv>>>
        KernelExecutionFixture(Queue queue, TExtent const& extent)
            : m_device{alpaka::getDev(queue)}
            , m_queue{std::move(queue)}
            , m_extent{castVec<Idx>(getExtents(extent))}
        {
        }

        template<typename TKernelFnObj, typename... TArgs>
        auto operator()(TKernelFnObj kernelFnObj, TArgs&&... args) -> bool
        {
            // the success flag lives on the device, initialised to true
            auto bufAccResult = allocBuf<bool, Idx>(m_device, static_cast<Idx>(1u));
            memset(m_queue, bufAccResult, static_cast<std::uint8_t>(true));

            if(!m_haveWorkDiv)
            {
                alpaka::KernelCfg<Acc> const kernelCfg = {m_extent, Vec<Dim, Idx>::ones()};
                m_workDiv = alpaka::getValidWorkDiv(
                    kernelCfg,
                    m_device,
                    kernelFnObj,
                    getPtrNative(bufAccResult),
                    std::forward<TArgs>(args)...);
                m_haveWorkDiv = true;
            }
            exec<Acc>(m_queue, m_workDiv, kernelFnObj, getPtrNative(bufAccResult), std::forward<TArgs>(args)...);

            auto bufHostResult = allocBuf<bool, Idx>(m_devHost, static_cast<Idx>(1u));
            memcpy(m_queue, bufHostResult, bufAccResult);
            wait(m_queue);
            return *getPtrNative(bufHostResult);
        }

    private:
        PlatformCpu m_platformHost{};
        DevCpu m_devHost{getDevByIdx(m_platformHost, 0)};
        Platform m_platform{};
        Device m_device{getDevByIdx(m_platform, 0)};
        Queue m_queue;
        WorkDiv m_workDiv{Vec<Dim, Idx>::all(0), Vec<Dim, Idx>::all(0), Vec<Dim, Idx>::all(0)};
        bool m_haveWorkDiv = false;
        Vec<Dim, Idx> m_extent{};
    };
} // namespace alpaka::test
