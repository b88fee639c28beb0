// include/alpaka/b200/Dev.hpp -- platforms, devices, queues and events of the B200 back-end, plus the host
// "device" (DevCpu) that owns host buffers.
//
// API parity with the reference's platform/Traits.hpp:54-81, dev/Traits.hpp:56-126, queue/Traits.hpp:46-70,
// wait/Traits.hpp:33-49, event/Traits.hpp and their CUDA implementations (platform/PlatformUniformCudaHipRt.hpp:28-136,
// dev/DevUniformCudaHipRt.hpp:55-266, queue/cuda_hip/QueueUniformCudaHipRt.hpp:40-242,
// event/EventUniformCudaHipRt.hpp:26-260). Where the reference calls cudart through ApiCudaRt, this layer calls the
// C ABI of libalpaka_b200.so (include/b200/b200.h) and converts its error codes into std::runtime_error carrying the
// same message text (reference: core/UniformCudaHip.hpp:23-112).
//
// Ownership, as in the reference: Platform is an empty value type; Dev is a cheap copyable handle; Queue and Event
// are shared_ptr handles with value equality; a queue's destructor waits for its work, then destroys the stream.
#pragma once

#include "Tags.hpp"
#include "b200/b200.h"

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <deque>
#include <exception>
#include <fstream>
#include <functional>
#include <future>
#include <memory>
#include <mutex>
#include <sstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <unistd.h>
#include <vector>

namespace alpaka
{
    namespace b200
    {
        //! error policy: a non-zero C-ABI code becomes std::runtime_error(b200_last_error_string())
        inline void check(int const rc)
        {
            if(rc != 0)
                throw std::runtime_error(std::string(b200_last_error_string()));
        }

        //! for destructors: report, never throw (reference: *_NOEXCEPT check variants)
        inline void checkNoexcept(int const rc) noexcept
        {
            if(rc != 0)
                std::cerr << "[alpaka-b200] " << b200_last_error_string() << std::endl;
        }
    } // namespace b200

    // -----------------------------------------------------------------------------------------------------------
    // trait declarations shared by host and device objects
    namespace trait
    {
        template<typename T, typename TSfinae = void>
        struct DevType;
        template<typename T, typename TSfinae = void>
        struct PlatformType;
        template<typename TDev, typename TProperty, typename TSfinae = void>
        struct QueueType;
        template<typename T, typename TSfinae = void>
        struct EventType;
        template<typename T, typename TSfinae = void>
        struct GetDev;
        template<typename T, typename TSfinae = void>
        struct NativeHandle;
        template<typename TQueue, typename TTask, typename TSfinae = void>
        struct Enqueue;
        template<typename TAwaited, typename TSfinae = void>
        struct CurrentThreadWaitFor;
        template<typename TWaiter, typename TAwaited, typename TSfinae = void>
        struct WaiterWaitFor;
        template<typename TQueue, typename TSfinae = void>
        struct Empty;
        template<typename TEvent, typename TSfinae = void>
        struct IsComplete;
    } // namespace trait

    template<typename T>
    using Dev = typename trait::DevType<std::remove_cv_t<std::remove_reference_t<T>>>::type;
    template<typename T>
    using Platform = typename trait::PlatformType<std::remove_cv_t<std::remove_reference_t<T>>>::type;
    template<typename TEnv, typename TProperty>
    using Queue = typename trait::QueueType<Dev<TEnv>, TProperty>::type;
    template<typename T>
    using Event = typename trait::EventType<std::remove_cv_t<std::remove_reference_t<T>>>::type;

    // -----------------------------------------------------------------------------------------------------------
    // host side: PlatformCpu / DevCpu / QueueCpu. They exist to own host buffers and to be copy end points; there is
    // no CPU accelerator (no kernel can be launched on DevCpu).
    class DevCpu
    {
    public:
        auto operator==(DevCpu const&) const -> bool
        {
            return true;
        }
        auto operator!=(DevCpu const&) const -> bool
        {
            return false;
        }
        [[nodiscard]] auto getNativeHandle() const noexcept -> int
        {
            return 0;
        }
    };

    struct PlatformCpu
    {
    };

    namespace b200
    {
        //! A single worker thread that runs `void()` tasks in submission order. Used by non-blocking host queues and by
        //! device queues for host tasks (a CUDA host-function callback must not call the runtime; reference:
        //! core/CallbackThread.hpp + queue/cuda_hip/QueueUniformCudaHipRt.hpp:194-230). A task object is destroyed
        //! BEFORE its completion is published, so "wait(queue) returned" implies "the task's captures are gone"
        //! (reference test: QueueTest.cpp "taskIsDestroyedAfterExecution"). The destructor drains the backlog.
        class CallbackThread
        {
            struct Item
            {
                std::function<void()> fn;
                std::promise<void> done;
            };

        public:
            CallbackThread() = default;
            CallbackThread(CallbackThread const&) = delete;
            auto operator=(CallbackThread const&) -> CallbackThread& = delete;
            ~CallbackThread()
            {
                {
                    std::lock_guard<std::mutex> l(m_mutex);
                    m_stop = true;
                }
                m_cv.notify_all();
                if(m_thread.joinable())
                {
                    if(m_thread.get_id() == std::this_thread::get_id())
                        m_thread.detach(); // the last handle died inside one of our own tasks
                    else
                        m_thread.join();
                }
            }

            auto submit(std::function<void()> fn) -> std::shared_future<void>
            {
                Item item{std::move(fn), {}};
                std::shared_future<void> fut = item.done.get_future().share();
                {
                    std::lock_guard<std::mutex> l(m_mutex);
                    ++m_pending;
                    m_items.emplace_back(std::move(item));
                    if(!m_thread.joinable())
                        m_thread = std::thread([this] { run(); });
                }
                m_cv.notify_one();
                return fut;
            }

            //! number of tasks submitted and not yet finished (the running one included)
            [[nodiscard]] auto pending() const -> std::size_t
            {
                std::lock_guard<std::mutex> l(m_mutex);
                return m_pending;
            }

            //! true when called from inside one of this thread's tasks
            [[nodiscard]] auto onWorker() const -> bool
            {
                std::lock_guard<std::mutex> l(m_mutex);
                return m_thread.joinable() && m_thread.get_id() == std::this_thread::get_id();
            }

        private:
            void run()
            {
                for(;;)
                {
                    Item item;
                    {
                        std::unique_lock<std::mutex> l(m_mutex);
                        m_cv.wait(l, [this] { return m_stop || !m_items.empty(); });
                        if(m_items.empty())
                            return;
                        item = std::move(m_items.front());
                        m_items.pop_front();
                    }
                    std::exception_ptr error;
                    try
                    {
                        item.fn();
                    }
                    catch(...)
                    {
                        error = std::current_exception();
                    }
                    item.fn = nullptr; // destroy the task (and what it captured) before anybody is told it finished
                    {
                        std::lock_guard<std::mutex> l(m_mutex);
                        --m_pending;
                    }
                    if(error)
                        item.done.set_exception(error);
                    else
                        item.done.set_value();
                }
            }

            mutable std::mutex m_mutex;
            std::condition_variable m_cv;
            std::deque<Item> m_items;
            std::thread m_thread;
            std::size_t m_pending = 0;
            bool m_stop = false;
        };

        //! One recording of a host event: signalled once, when the queue it was enqueued into reaches it.
        class HostMarker
        {
        public:
            void signal()
            {
                {
                    std::lock_guard<std::mutex> l(m_mutex);
                    m_done = true;
                }
                m_cv.notify_all();
            }
            void wait()
            {
                std::unique_lock<std::mutex> l(m_mutex);
                m_cv.wait(l, [this] { return m_done; });
            }
            [[nodiscard]] auto done() const -> bool
            {
                std::lock_guard<std::mutex> l(m_mutex);
                return m_done;
            }

        private:
            mutable std::mutex m_mutex;
            std::condition_variable m_cv;
            bool m_done = false;
        };

        //! Host task queue. Blocking: the task runs on the calling thread, serialised by a mutex
        //! (reference: queue/QueueGenericThreadsBlocking.hpp:127-163). Non-blocking: the task is handed to the queue's
        //! worker thread and the call returns (reference: queue/QueueGenericThreadsNonBlocking.hpp:100-140).
        class QueueCpuImpl
        {
        public:
            explicit QueueCpuImpl(bool blocking) : m_blocking(blocking)
            {
            }

            void run(std::function<void()> fn)
            {
                if(m_blocking)
                {
                    std::lock_guard<std::mutex> l(m_mutex);
                    m_busy.store(true);
                    struct Clear
                    {
                        std::atomic<bool>& b;
                        ~Clear()
                        {
                            b.store(false);
                        }
                    } clear{m_busy};
                    fn();
                    fn = nullptr;
                }
                else
                    (void) m_worker.submit(std::move(fn));
            }

            [[nodiscard]] auto empty() const -> bool
            {
                return m_blocking ? !m_busy.load() : m_worker.pending() == 0u;
            }

            //! returns once everything submitted before the call has finished
            void drain()
            {
                if(m_blocking)
                {
                    std::lock_guard<std::mutex> l(m_mutex);
                }
                else if(!m_worker.onWorker())
                    m_worker.submit([] {}).wait();
            }

            bool const m_blocking;
            std::mutex m_mutex;
            std::atomic<bool> m_busy{false};
            CallbackThread m_worker;
        };

        //! every live host queue, so that wait(DevCpu) can drain them (reference: dev/DevCpu.hpp queue registry)
        class HostQueueRegistry
        {
        public:
            static auto instance() -> HostQueueRegistry&
            {
                static HostQueueRegistry r;
                return r;
            }
            void add(std::shared_ptr<QueueCpuImpl> const& q)
            {
                std::lock_guard<std::mutex> l(m_mutex);
                std::erase_if(m_queues, [](auto const& w) { return w.expired(); });
                m_queues.emplace_back(q);
            }
            [[nodiscard]] auto snapshot() -> std::vector<std::shared_ptr<QueueCpuImpl>>
            {
                std::lock_guard<std::mutex> l(m_mutex);
                std::vector<std::shared_ptr<QueueCpuImpl>> live;
                for(auto const& w : m_queues)
                    if(auto sp = w.lock())
                        live.emplace_back(std::move(sp));
                return live;
            }

        private:
            std::mutex m_mutex;
            std::vector<std::weak_ptr<QueueCpuImpl>> m_queues;
        };
    } // namespace b200

    template<typename TProperty>
    class QueueCpu
    {
    public:
        explicit QueueCpu(DevCpu const& dev)
            : m_dev(dev)
            , m_impl(std::make_shared<b200::QueueCpuImpl>(std::is_same_v<TProperty, Blocking>))
        {
            b200::HostQueueRegistry::instance().add(m_impl);
        }
        auto operator==(QueueCpu const& rhs) const -> bool
        {
            return m_impl == rhs.m_impl;
        }
        auto operator!=(QueueCpu const& rhs) const -> bool
        {
            return !(*this == rhs);
        }
        DevCpu m_dev;
        std::shared_ptr<b200::QueueCpuImpl> m_impl;
    };
    using QueueCpuBlocking = QueueCpu<Blocking>;
    using QueueCpuNonBlocking = QueueCpu<NonBlocking>;

    // -----------------------------------------------------------------------------------------------------------
    // B200 side
    class DevB200;
    template<typename TProperty>
    class QueueB200;
    class EventB200;

    struct PlatformB200
    {
    };
    //! reference names
    using PlatformCudaRt = PlatformB200;
    using DevCudaRt = DevB200;

    class DevB200
    {
        friend struct trait::GetDev<DevB200>;

    public:
        DevB200() = default;
        explicit DevB200(int ordinal) : m_ordinal(ordinal)
        {
        }
        auto operator==(DevB200 const& rhs) const -> bool
        {
            return m_ordinal == rhs.m_ordinal;
        }
        auto operator!=(DevB200 const& rhs) const -> bool
        {
            return !(*this == rhs);
        }
        [[nodiscard]] auto getNativeHandle() const noexcept -> int
        {
            return m_ordinal;
        }

    private:
        int m_ordinal = 0;
    };

    namespace b200
    {
        //! the stream behind a queue, plus the per-queue scratch used by the native single-pass reductions
        class QueueB200Impl
        {
        public:
            explicit QueueB200Impl(DevB200 const& dev) : m_dev(dev)
            {
                check(b200_stream_create(dev.getNativeHandle(), &m_stream));
            }
            QueueB200Impl(QueueB200Impl const&) = delete;
            auto operator=(QueueB200Impl const&) -> QueueB200Impl& = delete;
            ~QueueB200Impl()
            {
                // the reference's queue destructor waits for outstanding work, then destroys the stream
                // (queue/cuda_hip/QueueUniformCudaHipRt.hpp:67-76)
                checkNoexcept(b200_stream_sync(m_stream));
                if(m_reduceScratch != nullptr)
                    checkNoexcept(b200_free_async(m_dev.getNativeHandle(), m_stream, m_reduceScratch));
                checkNoexcept(b200_stream_destroy(m_dev.getNativeHandle(), m_stream));
            }

            //! zero-initialised B200_REDUCE_SCRATCH_BYTES, allocated on first use (b200.h, "Reductions")
            auto reduceScratch() -> void*
            {
                std::lock_guard<std::mutex> l(m_mutex);
                if(m_reduceScratch == nullptr)
                {
                    int const d = m_dev.getNativeHandle();
                    check(b200_malloc_async(d, m_stream, B200_REDUCE_SCRATCH_BYTES, &m_reduceScratch));
                    check(b200_memset_async(d, m_reduceScratch, 0, B200_REDUCE_SCRATCH_BYTES, m_stream));
                }
                return m_reduceScratch;
            }

            DevB200 m_dev;
            b200_stream_t m_stream = nullptr;
            CallbackThread m_callbackThread;
            std::mutex m_mutex;
            void* m_reduceScratch = nullptr;
        };
    } // namespace b200

    template<typename TProperty>
    class QueueB200
    {
    public:
        explicit QueueB200(DevB200 const& dev) : m_impl(std::make_shared<b200::QueueB200Impl>(dev))
        {
        }
        auto operator==(QueueB200 const& rhs) const -> bool
        {
            return m_impl == rhs.m_impl;
        }
        auto operator!=(QueueB200 const& rhs) const -> bool
        {
            return !(*this == rhs);
        }
        //! the cudaStream_t as an opaque pointer
        [[nodiscard]] auto getNativeHandle() const noexcept -> b200_stream_t
        {
            return m_impl->m_stream;
        }
        //! blocking queues synchronise after every enqueue
        void afterEnqueue() const
        {
            if constexpr(std::is_same_v<TProperty, Blocking>)
                b200::check(b200_stream_sync(m_impl->m_stream));
        }
        std::shared_ptr<b200::QueueB200Impl> m_impl;
    };
    using QueueB200Blocking = QueueB200<Blocking>;
    using QueueB200NonBlocking = QueueB200<NonBlocking>;
    using QueueCudaRtBlocking = QueueB200Blocking;
    using QueueCudaRtNonBlocking = QueueB200NonBlocking;

    namespace b200
    {
        class EventB200Impl
        {
        public:
            EventB200Impl(DevB200 const& dev, bool timing) : m_dev(dev)
            {
                check(b200_event_create(dev.getNativeHandle(), timing ? 1 : 0, &m_event));
            }
            EventB200Impl(EventB200Impl const&) = delete;
            auto operator=(EventB200Impl const&) -> EventB200Impl& = delete;
            ~EventB200Impl()
            {
                checkNoexcept(b200_event_destroy(m_event));
            }
            DevB200 m_dev;
            b200_event_t m_event = nullptr;
        };
    } // namespace b200

    //! Device event. Like the reference's it is created with timing disabled (event/EventUniformCudaHipRt.hpp:50-52);
    //! `timing = true` is an extension used by the benchmark drivers (b200::elapsedMs).
    class EventB200
    {
    public:
        explicit EventB200(DevB200 const& dev, bool busyWait = true, bool timing = false)
            : m_impl(std::make_shared<b200::EventB200Impl>(dev, timing))
        {
            (void) busyWait;
        }
        auto operator==(EventB200 const& rhs) const -> bool
        {
            return m_impl == rhs.m_impl;
        }
        auto operator!=(EventB200 const& rhs) const -> bool
        {
            return !(*this == rhs);
        }
        [[nodiscard]] auto getNativeHandle() const noexcept -> b200_event_t
        {
            return m_impl->m_event;
        }
        std::shared_ptr<b200::EventB200Impl> m_impl;
    };
    using EventCudaRt = EventB200;

    namespace b200
    {
        class EventCpuImpl
        {
        public:
            explicit EventCpuImpl(DevCpu const& dev) : m_dev(dev)
            {
            }
            //! the most recent recording (null: never enqueued)
            [[nodiscard]] auto last() const -> std::shared_ptr<HostMarker>
            {
                std::lock_guard<std::mutex> l(m_mutex);
                return m_last;
            }
            //! starts a new recording; earlier ones stay alive for whoever already waits on them
            [[nodiscard]] auto record() -> std::shared_ptr<HostMarker>
            {
                auto marker = std::make_shared<HostMarker>();
                std::lock_guard<std::mutex> l(m_mutex);
                m_last = marker;
                return marker;
            }
            DevCpu m_dev;

        private:
            mutable std::mutex m_mutex;
            std::shared_ptr<HostMarker> m_last;
        };
    } // namespace b200

    //! Host event with the semantics of a CUDA event: every enqueue is a new recording; isComplete / wait(event) look at
    //! the latest recording, wait(queue, event) captures the recording that is current at the time of the call
    //! (reference: event/EventGenericThreads.hpp:27-330, pinned by test/unit/event/src/EventTest.cpp).
    class EventCpu
    {
    public:
        explicit EventCpu(DevCpu const& dev, bool = true) : m_impl(std::make_shared<b200::EventCpuImpl>(dev))
        {
        }
        auto operator==(EventCpu const& rhs) const -> bool
        {
            return m_impl == rhs.m_impl;
        }
        auto operator!=(EventCpu const& rhs) const -> bool
        {
            return !(*this == rhs);
        }
        std::shared_ptr<b200::EventCpuImpl> m_impl;
    };

    namespace b200
    {
        //! milliseconds between two recorded timing events (extension; the reference's events cannot time)
        inline auto elapsedMs(EventB200 const& start, EventB200 const& stop) -> float
        {
            float ms = 0.f;
            check(b200_event_elapsed_ms(start.getNativeHandle(), stop.getNativeHandle(), &ms));
            return ms;
        }
    } // namespace b200

    // -----------------------------------------------------------------------------------------------------------
    // trait specialisations
    namespace trait
    {
        template<>
        struct DevType<DevCpu>
        {
            using type = DevCpu;
        };
        template<>
        struct DevType<PlatformCpu>
        {
            using type = DevCpu;
        };
        template<>
        struct PlatformType<DevCpu>
        {
            using type = PlatformCpu;
        };
        template<>
        struct PlatformType<PlatformCpu>
        {
            using type = PlatformCpu;
        };
        template<typename TProperty>
        struct QueueType<DevCpu, TProperty>
        {
            using type = QueueCpu<TProperty>;
        };
        template<typename TProperty>
        struct DevType<QueueCpu<TProperty>>
        {
            using type = DevCpu;
        };
        template<typename TProperty>
        struct EventType<QueueCpu<TProperty>>
        {
            using type = EventCpu;
        };
        template<>
        struct EventType<DevCpu>
        {
            using type = EventCpu;
        };
        template<>
        struct DevType<EventCpu>
        {
            using type = DevCpu;
        };
        template<>
        struct GetDev<DevCpu>
        {
            static auto getDev(DevCpu const& d) -> DevCpu
            {
                return d;
            }
        };
        template<typename TProperty>
        struct GetDev<QueueCpu<TProperty>>
        {
            static auto getDev(QueueCpu<TProperty> const& q) -> DevCpu
            {
                return q.m_dev;
            }
        };
        template<>
        struct GetDev<EventCpu>
        {
            static auto getDev(EventCpu const& e) -> DevCpu
            {
                return e.m_impl->m_dev;
            }
        };

        template<>
        struct DevType<DevB200>
        {
            using type = DevB200;
        };
        template<>
        struct DevType<PlatformB200>
        {
            using type = DevB200;
        };
        template<>
        struct PlatformType<DevB200>
        {
            using type = PlatformB200;
        };
        template<>
        struct PlatformType<PlatformB200>
        {
            using type = PlatformB200;
        };
        template<typename TProperty>
        struct QueueType<DevB200, TProperty>
        {
            using type = QueueB200<TProperty>;
        };
        template<typename TProperty>
        struct DevType<QueueB200<TProperty>>
        {
            using type = DevB200;
        };
        template<typename TProperty>
        struct EventType<QueueB200<TProperty>>
        {
            using type = EventB200;
        };
        template<>
        struct EventType<DevB200>
        {
            using type = EventB200;
        };
        template<>
        struct DevType<EventB200>
        {
            using type = DevB200;
        };
        template<>
        struct GetDev<DevB200>
        {
            static auto getDev(DevB200 const& d) -> DevB200
            {
                return d;
            }
        };
        template<typename TProperty>
        struct GetDev<QueueB200<TProperty>>
        {
            static auto getDev(QueueB200<TProperty> const& q) -> DevB200
            {
                return q.m_impl->m_dev;
            }
        };
        template<>
        struct GetDev<EventB200>
        {
            static auto getDev(EventB200 const& e) -> DevB200
            {
                return e.m_impl->m_dev;
            }
        };
    } // namespace trait

    template<typename T>
    [[nodiscard]] auto getDev(T const& t)
    {
        return trait::GetDev<T>::getDev(t);
    }

    template<typename T>
    [[nodiscard]] auto getNativeHandle(T const& t)
    {
        return t.getNativeHandle();
    }

    //! the type getNativeHandle(T) returns (reference: traits/Traits.hpp:37-38)
    template<typename T>
    using NativeHandle = decltype(getNativeHandle(std::declval<T>()));

    // ---- platform
    [[nodiscard]] inline auto getDevCount(PlatformCpu const&) -> std::size_t
    {
        return 1u;
    }
    [[nodiscard]] inline auto getDevByIdx(PlatformCpu const&, std::size_t const& idx) -> DevCpu
    {
        if(idx >= 1u)
        {
            std::stringstream ss;
            ss << "Unable to return device handle for CPU device with index " << idx << " because there is only 1 device!";
            throw std::runtime_error(ss.str());
        }
        return DevCpu{};
    }
    [[nodiscard]] inline auto getDevCount(PlatformB200 const&) -> std::size_t
    {
        int n = 0;
        b200::check(b200_device_count(&n));
        return static_cast<std::size_t>(n);
    }
    [[nodiscard]] inline auto getDevByIdx(PlatformB200 const& platform, std::size_t const& idx) -> DevB200
    {
        std::size_t const n = getDevCount(platform);
        if(idx >= n)
        {
            std::stringstream ss;
            ss << "Unable to return device handle for device " << idx << ". There are only " << n << " devices!";
            throw std::runtime_error(ss.str());
        }
        // touch the device once so that a broken device is reported here, as the reference does
        // (platform/PlatformUniformCudaHipRt.hpp:66-100)
        b200_device_props props;
        b200::check(b200_device_props_get(static_cast<int>(idx), &props));
        return DevB200{static_cast<int>(idx)};
    }
    template<typename TPlatform>
    [[nodiscard]] auto getDevs(TPlatform const& platform) -> std::vector<Dev<TPlatform>>
    {
        std::vector<Dev<TPlatform>> devs;
        std::size_t const n = getDevCount(platform);
        devs.reserve(n);
        for(std::size_t i = 0; i < n; ++i)
            devs.push_back(getDevByIdx(platform, i));
        return devs;
    }

    // ---- device properties
    [[nodiscard]] inline auto getName(DevCpu const&) -> std::string
    {
        std::ifstream f("/proc/cpuinfo");
        std::string line;
        while(std::getline(f, line))
        {
            if(line.rfind("model name", 0) == 0)
            {
                auto const pos = line.find(':');
                if(pos != std::string::npos)
                    return line.substr(pos + 2);
            }
        }
        return "<unknown CPU>";
    }
    [[nodiscard]] inline auto getMemBytes(DevCpu const&) -> std::size_t
    {
        return static_cast<std::size_t>(sysconf(_SC_PHYS_PAGES)) * static_cast<std::size_t>(sysconf(_SC_PAGE_SIZE));
    }
    [[nodiscard]] inline auto getFreeMemBytes(DevCpu const&) -> std::size_t
    {
        return static_cast<std::size_t>(sysconf(_SC_AVPHYS_PAGES)) * static_cast<std::size_t>(sysconf(_SC_PAGE_SIZE));
    }
    [[nodiscard]] inline auto getWarpSizes(DevCpu const&) -> std::vector<std::size_t>
    {
        return {1u};
    }
    [[nodiscard]] inline auto getPreferredWarpSize(DevCpu const&) -> std::size_t
    {
        return 1u;
    }
    inline void reset(DevCpu const&)
    {
    }

    [[nodiscard]] inline auto getName(DevB200 const& dev) -> std::string
    {
        b200_device_props props;
        b200::check(b200_device_props_get(dev.getNativeHandle(), &props));
        return std::string(props.name);
    }
    [[nodiscard]] inline auto getMemBytes(DevB200 const& dev) -> std::size_t
    {
        uint64_t freeB = 0, totalB = 0;
        b200::check(b200_device_mem_info(dev.getNativeHandle(), &freeB, &totalB));
        return static_cast<std::size_t>(totalB);
    }
    [[nodiscard]] inline auto getFreeMemBytes(DevB200 const& dev) -> std::size_t
    {
        uint64_t freeB = 0, totalB = 0;
        b200::check(b200_device_mem_info(dev.getNativeHandle(), &freeB, &totalB));
        return static_cast<std::size_t>(freeB);
    }
    [[nodiscard]] inline auto getWarpSizes(DevB200 const& dev) -> std::vector<std::size_t>
    {
        b200_device_props props;
        b200::check(b200_device_props_get(dev.getNativeHandle(), &props));
        return {static_cast<std::size_t>(props.warp_size)};
    }
    [[nodiscard]] inline auto getPreferredWarpSize(DevB200 const& dev) -> std::size_t
    {
        return getWarpSizes(dev).front();
    }
    inline void reset(DevB200 const& dev)
    {
        b200::check(b200_device_reset(dev.getNativeHandle()));
    }

    // ---- waiting
    namespace trait
    {
        template<>
        struct CurrentThreadWaitFor<DevCpu>
        {
            static void currentThreadWaitFor(DevCpu const&)
            {
                for(auto const& q : b200::HostQueueRegistry::instance().snapshot())
                    q->drain();
            }
        };
        template<typename TProperty>
        struct CurrentThreadWaitFor<QueueCpu<TProperty>>
        {
            static void currentThreadWaitFor(QueueCpu<TProperty> const& q)
            {
                q.m_impl->drain();
            }
        };
        template<>
        struct CurrentThreadWaitFor<EventCpu>
        {
            static void currentThreadWaitFor(EventCpu const& e)
            {
                if(auto const marker = e.m_impl->last())
                    marker->wait();
            }
        };
        template<>
        struct CurrentThreadWaitFor<DevB200>
        {
            static void currentThreadWaitFor(DevB200 const& dev)
            {
                b200::check(b200_device_sync(dev.getNativeHandle()));
            }
        };
        template<typename TProperty>
        struct CurrentThreadWaitFor<QueueB200<TProperty>>
        {
            static void currentThreadWaitFor(QueueB200<TProperty> const& q)
            {
                b200::check(b200_stream_sync(q.getNativeHandle()));
            }
        };
        template<>
        struct CurrentThreadWaitFor<EventB200>
        {
            static void currentThreadWaitFor(EventB200 const& e)
            {
                b200::check(b200_event_sync(e.getNativeHandle()));
            }
        };
        template<typename TProperty>
        struct WaiterWaitFor<QueueB200<TProperty>, EventB200>
        {
            static void waiterWaitFor(QueueB200<TProperty>& q, EventB200 const& e)
            {
                b200::check(b200_stream_wait_event(q.getNativeHandle(), e.getNativeHandle()));
            }
        };
        template<>
        struct WaiterWaitFor<DevB200, EventB200>
        {
            static void waiterWaitFor(DevB200& dev, EventB200 const& e)
            {
                b200::check(b200_device_wait_event(dev.getNativeHandle(), e.getNativeHandle()));
            }
        };
        template<typename TProperty>
        struct WaiterWaitFor<QueueCpu<TProperty>, EventCpu>
        {
            static void waiterWaitFor(QueueCpu<TProperty>& q, EventCpu const& e)
            {
                // the recording current NOW is the one to wait for, even if the event is re-enqueued later
                auto const marker = e.m_impl->last();
                if(marker && !marker->done())
                    q.m_impl->run([marker] { marker->wait(); });
            }
        };
        template<>
        struct WaiterWaitFor<DevCpu, EventCpu>
        {
            static void waiterWaitFor(DevCpu&, EventCpu const& e)
            {
                auto const marker = e.m_impl->last();
                if(marker && !marker->done())
                    for(auto const& q : b200::HostQueueRegistry::instance().snapshot())
                        q->run([marker] { marker->wait(); });
            }
        };
        //! a host queue waiting for a device event: the wait happens in queue order
        template<typename TProperty>
        struct WaiterWaitFor<QueueCpu<TProperty>, EventB200>
        {
            static void waiterWaitFor(QueueCpu<TProperty>& q, EventB200 const& e)
            {
                q.m_impl->run([e] { b200::check(b200_event_sync(e.getNativeHandle())); });
            }
        };

        template<typename TProperty>
        struct Empty<QueueCpu<TProperty>>
        {
            static auto empty(QueueCpu<TProperty> const& q) -> bool
            {
                return q.m_impl->empty();
            }
        };
        template<typename TProperty>
        struct Empty<QueueB200<TProperty>>
        {
            static auto empty(QueueB200<TProperty> const& q) -> bool
            {
                int isEmpty = 0;
                b200::check(b200_stream_query(q.getNativeHandle(), &isEmpty));
                return isEmpty != 0;
            }
        };
        template<>
        struct IsComplete<EventCpu>
        {
            static auto isComplete(EventCpu const& e) -> bool
            {
                auto const marker = e.m_impl->last();
                return !marker || marker->done();
            }
        };
        template<>
        struct IsComplete<EventB200>
        {
            static auto isComplete(EventB200 const& e) -> bool
            {
                int done = 0;
                b200::check(b200_event_query(e.getNativeHandle(), &done));
                return done != 0;
            }
        };

        // ---- enqueue: events
        template<typename TProperty>
        struct Enqueue<QueueB200<TProperty>, EventB200>
        {
            static void enqueue(QueueB200<TProperty>& q, EventB200& e)
            {
                b200::check(b200_event_record(e.getNativeHandle(), q.getNativeHandle()));
                q.afterEnqueue();
            }
        };
        template<typename TProperty>
        struct Enqueue<QueueCpu<TProperty>, EventCpu>
        {
            static void enqueue(QueueCpu<TProperty>& q, EventCpu& e)
            {
                auto marker = e.m_impl->record();
                q.m_impl->run([marker] { marker->signal(); });
            }
        };

        // ---- enqueue: any host callable (`void()`), the generic case
        template<typename TProperty, typename TTask>
        struct Enqueue<QueueCpu<TProperty>, TTask, std::enable_if_t<std::is_invocable_v<TTask&>>>
        {
            static void enqueue(QueueCpu<TProperty>& q, TTask const& task)
            {
                q.m_impl->run(std::function<void()>(task));
            }
        };
        template<typename TProperty, typename TTask>
        struct Enqueue<QueueB200<TProperty>, TTask, std::enable_if_t<std::is_invocable_v<TTask&>>>
        {
            //! The queue is referenced, not owned: its destructor drains the stream (and with it this host function)
            //! before anything is torn down, whereas owning it here could make the CUDA callback thread run that
            //! destructor, where stream synchronisation is not permitted.
            struct Payload
            {
                b200::QueueB200Impl* impl;
                std::function<void()> fn;
            };
            static void trampoline(void* user)
            {
                std::unique_ptr<Payload> p(static_cast<Payload*>(user));
                // run on the queue's callback thread (it may call the runtime), block the stream until done
                p->impl->m_callbackThread.submit(std::move(p->fn)).wait();
            }
            static void enqueue(QueueB200<TProperty>& q, TTask const& task)
            {
                auto* p = new Payload{q.m_impl.get(), std::function<void()>(task)};
                int const rc = b200_launch_host_func(q.getNativeHandle(), &trampoline, p);
                if(rc != 0)
                {
                    delete p;
                    b200::check(rc);
                }
                q.afterEnqueue();
            }
        };
    } // namespace trait

    //! Waits the calling thread for the completion of the given awaited action (device, queue or event).
    template<typename TAwaited>
    void wait(TAwaited const& awaited)
    {
        trait::CurrentThreadWaitFor<TAwaited>::currentThreadWaitFor(awaited);
    }
    //! Makes `waiter` (a queue or a device) wait for `awaited` (an event) without blocking the caller.
    template<typename TWaiter, typename TAwaited>
    void wait(TWaiter& waiter, TAwaited const& awaited)
    {
        trait::WaiterWaitFor<TWaiter, TAwaited>::waiterWaitFor(waiter, awaited);
    }
    template<typename TQueue, typename TTask>
    void enqueue(TQueue& queue, TTask&& task)
    {
        trait::Enqueue<TQueue, std::decay_t<TTask>>::enqueue(queue, task);
    }
    template<typename TQueue>
    [[nodiscard]] auto empty(TQueue const& queue) -> bool
    {
        return trait::Empty<TQueue>::empty(queue);
    }
    template<typename TEvent>
    [[nodiscard]] auto isComplete(TEvent const& event) -> bool
    {
        return trait::IsComplete<TEvent>::isComplete(event);
    }

    namespace concepts
    {
        template<typename T>
        concept Queue = requires { typename trait::DevType<T>::type; } && requires(T const& q) { q.m_impl; };
    }
    template<typename T>
    inline constexpr bool isQueue = concepts::Queue<std::decay_t<T>>;
    template<typename T>
    inline constexpr bool isDevice = std::is_same_v<std::decay_t<T>, DevB200> || std::is_same_v<std::decay_t<T>, DevCpu>;
    template<typename T>
    inline constexpr bool isPlatform
        = std::is_same_v<std::decay_t<T>, PlatformB200> || std::is_same_v<std::decay_t<T>, PlatformCpu>;
} // namespace alpaka
