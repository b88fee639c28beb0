// alpaka::core::detail::ThreadPool (reference: include/alpaka/core/ThreadPool.hpp; pinned by
// test/unit/core/src/ThreadPool.cpp): N worker threads, enqueueTask() returns a future that rethrows what the task
// threw. The reference runs the blocks/threads of its CPU accelerators on such a pool; the B200 layer has no CPU
// accelerator, the class exists for user code that borrowed it.
#pragma once
#include <alpaka/alpaka.hpp>

#include <condition_variable>
#include <deque>
#include <functional>
#include <future>
#include <memory>
#include <mutex>
#include <thread>
#include <type_traits>
#include <vector>

namespace alpaka::core::detail
{
    class ThreadPool
    {
    public:
        explicit ThreadPool(std::size_t threadCount)
        {
            if(threadCount < 1)
                throw std::invalid_argument("The argument 'threadCount' has to be greate or equal to one!");
            m_threads.reserve(threadCount);
            for(std::size_t i = 0; i < threadCount; ++i)
                m_threads.emplace_back([this] { work(); });
        }
        ThreadPool(ThreadPool const&) = delete;
        auto operator=(ThreadPool const&) -> ThreadPool& = delete;
        ~ThreadPool()
        {
            {
                std::lock_guard<std::mutex> l(m_mutex);
                m_stop = true;
            }
            m_cv.notify_all();
            for(auto& t : m_threads)
                t.join(); // drains the backlog first: work() only leaves on an empty queue
        }

        template<typename TFn, typename... TArgs>
        auto enqueueTask(TFn&& fn, TArgs&&... args) -> std::future<void>
        {
            auto task = std::make_shared<std::packaged_task<void()>>(
                [f = std::forward<TFn>(fn), ... a = std::forward<TArgs>(args)]() mutable { (void) f(a...); });
            auto fut = task->get_future();
            {
                std::lock_guard<std::mutex> l(m_mutex);
                m_tasks.emplace_back([task] { (*task)(); });
            }
            m_cv.notify_one();
            return fut;
        }

    private:
        void work()
        {
            for(;;)
            {
                std::function<void()> job;
                {
                    std::unique_lock<std::mutex> l(m_mutex);
                    m_cv.wait(l, [this] { return m_stop || !m_tasks.empty(); });
                    if(m_tasks.empty())
                        return;
                    job = std::move(m_tasks.front());
                    m_tasks.pop_front();
                }
                job(); // exceptions end up in the task's future
            }
        }

        std::mutex m_mutex;
        std::condition_variable m_cv;
        std::deque<std::function<void()>> m_tasks;
        std::vector<std::thread> m_threads;
        bool m_stop = false;
    };
} // namespace alpaka::core::detail
