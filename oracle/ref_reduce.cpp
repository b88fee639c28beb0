/* oracle/_ref harness, part 2/3: example/reduce on the reference's own CPU back-ends.
 *
 * TEST INFRASTRUCTURE ONLY (see ref_babelstream.cpp). ReduceKernel, cheapArray, the iterators and the
 * per-accelerator configs are the reference's files included verbatim from
 * -I/root/reference/example/reduce/src (kernel.hpp:42-132, iterator.hpp, alpakaConfig.hpp). The shipped
 * driver hard-wires `Accelerator = CpuSerial`, n = 1<<28 and T = uint32_t (reduce.cpp:25,112,114), so the
 * two-launch host sequence of reduce.cpp:47-105 is re-issued here, parameterised on accelerator config,
 * element type and n; the launch shapes are computed exactly as reduce.cpp:55-63,75-77.
 */
#include "alpakaConfig.hpp" // reference file
#include "kernel.hpp" // reference file

#include <chrono>
#include <cstdint>
#include <omp.h>

namespace
{
    template<typename T>
    struct AddFn
    {
        ALPAKA_FN_HOST_ACC auto operator()(T a, T b) const -> T
        {
            return a + b;
        }
    };

    template<typename TCfg, typename T>
    T reduceOn(T const* src, std::uint64_t n, double* kernelSeconds)
    {
        using Acc = typename TCfg::Acc;
        using QueueAcc = alpaka::Queue<Acc, alpaka::Blocking>;
        static constexpr std::uint64_t blockSize = getMaxBlockSize<TCfg, 256>();

        auto const devHost = alpaka::getDevByIdx(alpaka::PlatformCpu{}, 0);
        auto const devAcc = alpaka::getDevByIdx(alpaka::Platform<Acc>{}, 0);
        QueueAcc queue(devAcc);

        // reduce.cpp:55-63
        auto blockCount = static_cast<std::uint32_t>(alpaka::getAccDevProps<Acc>(devAcc).m_multiProcessorCount * 8);
        auto const maxBlockCount = static_cast<std::uint32_t>((((n + 1) / 2) - 1) / blockSize + 1);
        if(blockCount > maxBlockCount)
            blockCount = maxBlockCount;

        auto destination = alpaka::allocBuf<T, Idx>(devAcc, static_cast<Extent>(blockCount));

        using Fn = AddFn<T>;
        ReduceKernel<blockSize, T, Fn> kernel1, kernel2;
        WorkDiv workDiv1{static_cast<Extent>(blockCount), static_cast<Extent>(blockSize), static_cast<Extent>(1)};
        WorkDiv workDiv2{static_cast<Extent>(1), static_cast<Extent>(blockSize), static_cast<Extent>(1)};

        // reduce.cpp:79-98; the source is read in place (CPU acc: device memory == host memory).
        auto const task1 = alpaka::createTaskKernel<Acc>(workDiv1, kernel1, src, std::data(destination), n, Fn{});
        auto const task2 = alpaka::createTaskKernel<Acc>(
            workDiv2,
            kernel2,
            static_cast<T const*>(std::data(destination)),
            std::data(destination),
            static_cast<std::uint64_t>(blockCount),
            Fn{});

        auto const t0 = std::chrono::high_resolution_clock::now();
        alpaka::enqueue(queue, task1);
        alpaka::enqueue(queue, task2);
        alpaka::wait(queue);
        auto const t1 = std::chrono::high_resolution_clock::now();
        if(kernelSeconds)
            *kernelSeconds = std::chrono::duration<double>(t1 - t0).count();
        (void) devHost;
        return std::data(destination)[0];
    }

    template<typename T>
    int dispatch(int acc, T const* src, std::uint64_t n, T* out, double* seconds)
    {
        try
        {
            // IteratorCpu casts its begin/end to uint32_t (iterator.hpp:126-138): n >= 2^32 is not
            // representable on the reference CPU path.
            if(n == 0 || n > 0xffffffffull)
                return -1;
            *out = acc == 1 ? reduceOn<CpuOmp2Blocks, T>(src, n, seconds) : reduceOn<CpuSerial, T>(src, n, seconds);
            return 0;
        }
        catch(...)
        {
            return -2;
        }
    }
} // namespace

extern "C"
{
    //! acc: 0 AccCpuSerial, 1 AccCpuOmp2Blocks. seconds may be NULL.
    int ref_reduce_u32(int acc, std::uint32_t const* src, std::uint64_t n, std::uint32_t* out, double* seconds)
    {
        return dispatch<std::uint32_t>(acc, src, n, out, seconds);
    }

    int ref_reduce_i32(int acc, std::int32_t const* src, std::uint64_t n, std::int32_t* out, double* seconds)
    {
        return dispatch<std::int32_t>(acc, src, n, out, seconds);
    }

    int ref_reduce_u64(int acc, std::uint64_t const* src, std::uint64_t n, std::uint64_t* out, double* seconds)
    {
        return dispatch<std::uint64_t>(acc, src, n, out, seconds);
    }

    int ref_reduce_f32(int acc, float const* src, std::uint64_t n, float* out, double* seconds)
    {
        return dispatch<float>(acc, src, n, out, seconds);
    }

    int ref_reduce_f64(int acc, double const* src, std::uint64_t n, double* out, double* seconds)
    {
        return dispatch<double>(acc, src, n, out, seconds);
    }
}
