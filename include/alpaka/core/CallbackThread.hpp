// alpaka::core::CallbackThread (reference: include/alpaka/core/CallbackThread.hpp; pinned by
// test/unit/core/src/CallbackThread.cpp): one worker thread running tasks in submission order; submit() returns a future
// that rethrows what the task threw. A thin front of the worker the B200 queues use for host tasks
// (alpaka::b200::CallbackThread, include/alpaka/b200/Dev.hpp), extended to move-only callables.
#pragma once
#include <alpaka/alpaka.hpp>

#include <future>
#include <memory>
#include <type_traits>
#include <utility>

namespace alpaka::core
{
    class CallbackThread
    {
    public:
        template<typename TFn>
        auto submit(TFn&& fn) -> std::shared_future<void>
        {
            // the worker stores std::function (copyable): keep move-only callables behind a shared pointer
            auto held = std::make_shared<std::decay_t<TFn>>(std::forward<TFn>(fn));
            return m_worker.submit([held] { (*held)(); });
        }

    private:
        b200::CallbackThread m_worker;
    };
} // namespace alpaka::core
