// include/alpaka/test/event/EventHostManualTrigger.hpp -- a queue entry that blocks the queue until the HOST calls
// trigger(); the event tests use it to hold a queue in a known state.
//
// Interface parity with the reference's helper (include/alpaka/test/event/EventHostManualTrigger.hpp:30-779):
// EventHostManualTrigger<TDev>, isEventHostManualTriggerSupported(dev), enqueue(queue, trigger), isComplete(trigger),
// trigger.trigger(). What is behind it here:
//   * DevCpu: a HostMarker (Dev.hpp) the queue's worker blocks on;
//   * DevB200: a 32-bit word in pinned, device-mapped host memory and a stream memory operation
//     (b200_stream_wait_value32, i.e. cuStreamWaitValue32 resolved at run time) -- the stream waits until the word
//     becomes non-zero, trigger() is a plain host store. No kernel occupies an SM while the queue is held.
#pragma once

#include <alpaka/alpaka.hpp>

#include <atomic>
#include <chrono>
#include <memory>
#include <mutex>
#include <thread>

namespace alpaka::test
{
    namespace trait
    {
        template<typename TDev>
        struct EventHostManualTriggerType;
        template<typename TDev>
        struct IsEventHostManualTriggerSupported;
    } // namespace trait

    //! The event host manual trigger type of the device.
    template<typename TDev>
    using EventHostManualTrigger = typename trait::EventHostManualTriggerType<TDev>::type;

    template<typename TDev>
    auto isEventHostManualTriggerSupported(TDev const& dev) -> bool
    {
        return trait::IsEventHostManualTriggerSupported<TDev>::isSupported(dev);
    }

    // ---- host
    class EventHostManualTriggerCpu
    {
    public:
        struct Impl
        {
            explicit Impl(DevCpu const& dev) : m_dev(dev)
            {
            }
            DevCpu m_dev;
            std::mutex m_mutex;
            std::shared_ptr<b200::HostMarker> m_gate; //!< null: not enqueued or already triggered
        };

        explicit EventHostManualTriggerCpu(DevCpu const& dev) : m_impl(std::make_shared<Impl>(dev))
        {
        }
        auto operator==(EventHostManualTriggerCpu const& rhs) const -> bool
        {
            return m_impl == rhs.m_impl;
        }
        auto operator!=(EventHostManualTriggerCpu const& rhs) const -> bool
        {
            return !(*this == rhs);
        }
        void trigger()
        {
            std::shared_ptr<b200::HostMarker> gate;
            {
                std::lock_guard<std::mutex> l(m_impl->m_mutex);
                gate.swap(m_impl->m_gate);
            }
            if(gate)
                gate->signal();
            // let the released queues make progress before the caller inspects them (as the reference's helper does)
            std::this_thread::sleep_for(std::chrono::milliseconds(200));
        }
        std::shared_ptr<Impl> m_impl;
    };

    // ---- B200
    class EventHostManualTriggerB200
    {
    public:
        struct Impl
        {
            explicit Impl(DevB200 const& dev) : m_dev(dev)
            {
                void* p = nullptr;
                b200::check(b200_host_alloc_pinned(sizeof(std::uint32_t), &p));
                m_word = static_cast<std::uint32_t*>(p);
                *m_word = 0u;
            }
            Impl(Impl const&) = delete;
            auto operator=(Impl const&) -> Impl& = delete;
            ~Impl()
            {
                b200::checkNoexcept(b200_host_free_pinned(m_word));
            }
            DevB200 m_dev;
            std::mutex m_mutex;
            std::uint32_t* m_word = nullptr; //!< pinned + mapped: the same address is valid on the device (UVA)
            bool m_ready = true; //!< not enqueued, or already triggered
        };

        explicit EventHostManualTriggerB200(DevB200 const& dev) : m_impl(std::make_shared<Impl>(dev))
        {
        }
        auto operator==(EventHostManualTriggerB200 const& rhs) const -> bool
        {
            return m_impl == rhs.m_impl;
        }
        auto operator!=(EventHostManualTriggerB200 const& rhs) const -> bool
        {
            return !(*this == rhs);
        }
        void trigger()
        {
            {
                std::lock_guard<std::mutex> l(m_impl->m_mutex);
                m_impl->m_ready = true;
                std::atomic_ref<std::uint32_t>(*m_impl->m_word).store(1u, std::memory_order_release);
            }
            std::this_thread::sleep_for(std::chrono::milliseconds(200));
        }
        std::shared_ptr<Impl> m_impl;
    };
    using EventHostManualTriggerCuda = EventHostManualTriggerB200;

    namespace trait
    {
        template<>
        struct EventHostManualTriggerType<DevCpu>
        {
            using type = EventHostManualTriggerCpu;
        };
        template<>
        struct EventHostManualTriggerType<DevB200>
        {
            using type = EventHostManualTriggerB200;
        };
        template<>
        struct IsEventHostManualTriggerSupported<DevCpu>
        {
            static auto isSupported(DevCpu const&) -> bool
            {
                return true;
            }
        };
        template<>
        struct IsEventHostManualTriggerSupported<DevB200>
        {
            //! probes the stream memory operation once on a scratch stream (the wait is already satisfied)
            static auto isSupported(DevB200 const& dev) -> bool
            {
                void* p = nullptr;
                if(b200_host_alloc_pinned(sizeof(std::uint32_t), &p) != 0)
                    return false;
                *static_cast<std::uint32_t*>(p) = 1u;
                b200_stream_t s = nullptr;
                bool ok = b200_stream_create(dev.getNativeHandle(), &s) == 0;
                if(ok)
                {
                    ok = b200_stream_wait_value32(dev.getNativeHandle(), s, p, 1u) == 0 && b200_stream_sync(s) == 0;
                    (void) b200_stream_destroy(dev.getNativeHandle(), s);
                }
                (void) b200_host_free_pinned(p);
                return ok;
            }
        };
    } // namespace trait
} // namespace alpaka::test

namespace alpaka::trait
{
    template<>
    struct DevType<test::EventHostManualTriggerCpu>
    {
        using type = DevCpu;
    };
    template<>
    struct GetDev<test::EventHostManualTriggerCpu>
    {
        static auto getDev(test::EventHostManualTriggerCpu const& e) -> DevCpu
        {
            return e.m_impl->m_dev;
        }
    };
    template<>
    struct IsComplete<test::EventHostManualTriggerCpu>
    {
        static auto isComplete(test::EventHostManualTriggerCpu const& e) -> bool
        {
            std::lock_guard<std::mutex> l(e.m_impl->m_mutex);
            return !e.m_impl->m_gate;
        }
    };
    template<typename TProperty>
    struct Enqueue<QueueCpu<TProperty>, test::EventHostManualTriggerCpu>
    {
        static void enqueue(QueueCpu<TProperty>& q, test::EventHostManualTriggerCpu& e)
        {
            auto gate = std::make_shared<b200::HostMarker>();
            {
                std::lock_guard<std::mutex> l(e.m_impl->m_mutex);
                ALPAKA_ASSERT(!e.m_impl->m_gate); // must not be enqueued twice without a trigger in between
                e.m_impl->m_gate = gate;
            }
            // a blocking queue blocks its caller here until another thread triggers
            q.m_impl->run([gate] { gate->wait(); });
        }
    };

    template<>
    struct DevType<test::EventHostManualTriggerB200>
    {
        using type = DevB200;
    };
    template<>
    struct GetDev<test::EventHostManualTriggerB200>
    {
        static auto getDev(test::EventHostManualTriggerB200 const& e) -> DevB200
        {
            return e.m_impl->m_dev;
        }
    };
    template<>
    struct IsComplete<test::EventHostManualTriggerB200>
    {
        static auto isComplete(test::EventHostManualTriggerB200 const& e) -> bool
        {
            std::lock_guard<std::mutex> l(e.m_impl->m_mutex);
            return e.m_impl->m_ready;
        }
    };
    template<typename TProperty>
    struct Enqueue<QueueB200<TProperty>, test::EventHostManualTriggerB200>
    {
        static void enqueue(QueueB200<TProperty>& q, test::EventHostManualTriggerB200& e)
        {
            auto const impl = e.m_impl;
            std::lock_guard<std::mutex> l(impl->m_mutex);
            ALPAKA_ASSERT(impl->m_ready);
            impl->m_ready = false;
            std::atomic_ref<std::uint32_t>(*impl->m_word).store(0u, std::memory_order_release);
            // no afterEnqueue(): a blocking queue must not wait here, nobody could trigger (reference :440-468)
            b200::check(b200_stream_wait_value32(impl->m_dev.getNativeHandle(), q.getNativeHandle(), impl->m_word, 1u));
        }
    };
} // namespace alpaka::trait
