// include/alpaka/test/queue/Queue.hpp -- queue selection helpers of the test suite
// (reference: include/alpaka/test/queue/Queue.hpp:16-146): DefaultQueue<TDev>, isBlockingQueue, TestQueues.
#pragma once

#include <alpaka/alpaka.hpp>

#include <tuple>

namespace alpaka::test
{
    namespace trait
    {
        template<typename TDev, typename TSfinae = void>
        struct DefaultQueueType;
        template<>
        struct DefaultQueueType<DevCpu>
        {
#if(ALPAKA_DEBUG >= ALPAKA_DEBUG_FULL)
            using type = QueueCpuBlocking;
#else
            using type = QueueCpuNonBlocking;
#endif
        };
        template<>
        struct DefaultQueueType<DevB200>
        {
#if(ALPAKA_DEBUG >= ALPAKA_DEBUG_FULL)
            using type = QueueB200Blocking;
#else
            using type = QueueB200NonBlocking;
#endif
        };

        template<typename TQueue, typename TSfinae = void>
        struct IsBlockingQueue;
        template<typename TProperty>
        struct IsBlockingQueue<QueueCpu<TProperty>> : std::is_same<TProperty, Blocking>
        {
        };
        template<typename TProperty>
        struct IsBlockingQueue<QueueB200<TProperty>> : std::is_same<TProperty, Blocking>
        {
        };
    } // namespace trait

    //! The queue type that should be used for the given device.
    template<typename TDev>
    using DefaultQueue = typename trait::DefaultQueueType<TDev>::type;

    //! The blocking queue trait.
    template<typename TQueue>
    using IsBlockingQueue = trait::IsBlockingQueue<TQueue>;
    template<typename TQueue>
    inline constexpr bool isBlockingQueue = trait::IsBlockingQueue<TQueue>::value;

    //! A std::tuple holding tuples of devices and corresponding queue types.
    using TestQueues = std::tuple<
        std::tuple<DevCpu, QueueCpuBlocking>,
        std::tuple<DevCpu, QueueCpuNonBlocking>,
        std::tuple<DevB200, QueueB200Blocking>,
        std::tuple<DevB200, QueueB200NonBlocking>>;
} // namespace alpaka::test
