// tests/cpp/atomic_blocks.cpp -- hierarchy::Blocks / hierarchy::Grids atomics must be atomic BETWEEN blocks (device scope),
// hierarchy::Threads may use CTA scope (reference: include/alpaka/atomic/AtomicUniformCudaHip.hpp:80-130 uses the *_block
// intrinsics for hierarchy::Threads only). One global counter, many blocks, every thread adds 1 under each hierarchy.
// Built by examples/Makefile into build/examples/test_atomic_blocks; tests/test_gpu_cpp_layer.py runs it on the GPU and
// tests/test_sass_properties.py checks the scope of the emitted atomics on CPU.
#include <alpaka/alpaka.hpp>

#include <cstdint>
#include <cstdio>

struct CountBlocksKernel
{
    template<typename TAcc>
    ALPAKA_FN_ACC void operator()(TAcc const& acc, std::uint32_t* counters, double* sums) const
    {
        alpaka::atomicAdd(acc, &counters[0], 1u, alpaka::hierarchy::Blocks{});
        alpaka::atomicAdd(acc, &counters[1], 1u, alpaka::hierarchy::Grids{});
        alpaka::atomicAdd(acc, &counters[2], 1u); // default hierarchy (Grids)
        alpaka::atomicAdd(acc, &sums[0], 1.0, alpaka::hierarchy::Blocks{});
        alpaka::atomicMax(acc, &counters[3], alpaka::getIdx<alpaka::Grid, alpaka::Threads>(acc)[0], alpaka::hierarchy::Blocks{});
        alpaka::atomicCas(acc, &counters[4], 0u, 7u, alpaka::hierarchy::Blocks{});
    }
};

struct CountThreadsKernel
{
    template<typename TAcc>
    ALPAKA_FN_ACC void operator()(TAcc const& acc, std::uint32_t* perBlock) const
    {
        // one counter per block: only the threads of that block touch it
        alpaka::atomicAdd(acc, &perBlock[alpaka::getIdx<alpaka::Grid, alpaka::Blocks>(acc)[0]], 1u, alpaka::hierarchy::Threads{});
    }
};

auto main() -> int
{
    using Dim = alpaka::DimInt<1u>;
    using Idx = std::uint32_t;
    using Acc = alpaka::AccGpuB200<Dim, Idx>;
    auto const devHost = alpaka::getDevByIdx(alpaka::PlatformCpu{}, 0);
    auto const devAcc = alpaka::getDevByIdx(alpaka::Platform<Acc>{}, 0);
    alpaka::Queue<Acc, alpaka::Blocking> queue{devAcc};
    constexpr Idx blocks = 4096, threads = 256;
    auto counters = alpaka::allocBuf<std::uint32_t, Idx>(devAcc, Idx{8});
    auto sums = alpaka::allocBuf<double, Idx>(devAcc, Idx{1});
    auto perBlock = alpaka::allocBuf<std::uint32_t, Idx>(devAcc, blocks);
    alpaka::memset(queue, counters, 0);
    alpaka::memset(queue, sums, 0);
    alpaka::memset(queue, perBlock, 0);
    alpaka::WorkDivMembers<Dim, Idx> const wd{alpaka::Vec<Dim, Idx>{blocks}, alpaka::Vec<Dim, Idx>{threads}, alpaka::Vec<Dim, Idx>{Idx{1}}};
    alpaka::exec<Acc>(queue, wd, CountBlocksKernel{}, counters.data(), sums.data());
    alpaka::exec<Acc>(queue, wd, CountThreadsKernel{}, perBlock.data());
    auto hc = alpaka::allocBuf<std::uint32_t, Idx>(devHost, Idx{8});
    auto hs = alpaka::allocBuf<double, Idx>(devHost, Idx{1});
    auto hb = alpaka::allocBuf<std::uint32_t, Idx>(devHost, blocks);
    alpaka::memcpy(queue, hc, counters);
    alpaka::memcpy(queue, hs, sums);
    alpaka::memcpy(queue, hb, perBlock);
    alpaka::wait(queue);
    std::uint32_t const want = blocks * threads;
    bool ok = hc.data()[0] == want && hc.data()[1] == want && hc.data()[2] == want && hs.data()[0] == double(want)
              && hc.data()[3] == want - 1u && hc.data()[4] == 7u;
    for(Idx b = 0; b < blocks; ++b)
        ok = ok && hb.data()[b] == threads;
    std::printf("atomic_blocks: Blocks %u Grids %u default %u double %.0f max %u cas %u (want %u) -> %s\n", hc.data()[0], hc.data()[1],
                hc.data()[2], hs.data()[0], hc.data()[3], hc.data()[4], want, ok ? "OK" : "FAILED");
    return ok ? 0 : 1;
}
