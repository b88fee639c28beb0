// include/alpaka/test/queue/QueueTestFixture.hpp -- platform + device 0 + one queue for a (Dev, Queue) tuple of
// test::TestQueues (interface parity with the reference's include/alpaka/test/queue/QueueTestFixture.hpp:11-22).
#pragma once

#include <alpaka/alpaka.hpp>

#include <tuple>

namespace alpaka::test
{
    template<typename TDevQueue>
    struct QueueTestFixture
    {
        using Dev = std::tuple_element_t<0, TDevQueue>;
        using Queue = std::tuple_element_t<1, TDevQueue>;
        using Platform = alpaka::Platform<Dev>;

        Platform m_platform{};
        Dev m_dev{getDevByIdx(m_platform, 0)};
        Queue m_queue{m_dev};
    };
} // namespace alpaka::test
