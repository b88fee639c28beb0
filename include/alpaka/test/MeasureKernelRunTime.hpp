// include/alpaka/test/MeasureKernelRunTime.hpp -- host-clock time of one task incl. the wait for its completion
// (reference: include/alpaka/test/MeasureKernelRunTime.hpp:17-46).
#pragma once

#include <alpaka/alpaka.hpp>

#include <chrono>

namespace alpaka::test::integ
{
    template<typename TCallable>
    auto measureRunTimeMs(TCallable&& callable) -> std::chrono::milliseconds::rep
    {
        auto const start = std::chrono::high_resolution_clock::now();
        std::forward<TCallable>(callable)();
        auto const end = std::chrono::high_resolution_clock::now();
        return std::chrono::duration_cast<std::chrono::milliseconds>(end - start).count();
    }

    //! enqueues the task, waits for it, returns the elapsed milliseconds
    template<typename TQueue, typename TTask>
    auto measureTaskRunTimeMs(TQueue& queue, TTask&& task) -> std::chrono::milliseconds::rep
    {
        alpaka::wait(queue);
        return measureRunTimeMs(
            [&]
            {
                alpaka::enqueue(queue, std::forward<TTask>(task));
                alpaka::wait(queue);
            });
    }
} // namespace alpaka::test::integ
