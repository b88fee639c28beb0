// examples/reduce/reduce_b200.cpp -- example/reduce on the B200 back-end.
//
// The reference's ReduceKernel, its iterators and its accelerator configuration are used UNMODIFIED: this file
// includes example/reduce/src/{alpakaConfig.hpp, kernel.hpp, iterator.hpp} from the reference tree through -I (nothing
// is copied) and only replaces the driver translation unit, because the shipped reduce.cpp hard-codes
// `using Accelerator = CpuSerial;`, n = 2^28 and uint32 (example/reduce/src/reduce.cpp:25, 112-114).
// The host sequence below follows reduce.cpp:47-108 (block count = min(8 * SMs, ceil(ceil(n/2)/blockSize)), two
// createTaskKernel + enqueue, result read back through createView + memcpy of one element).
//
// Two reduction functors are run:
//   --mode=lambda  an extended __host__ __device__ lambda, as reduce.cpp:141 -> the generic trampoline runs the
//                  reference kernel twice (what the reference does on its CUDA back-end);
//   --mode=sum     alpaka::b200::Sum<T> -> recognised (alpaka/b200/Native.hpp), ONE single-pass native reduction.
// Both must agree bit for bit for integers. --input=<file> reads the data (parity tests); default is i+1 as reduce.cpp:137-138
// with the closed-form check of reduce.cpp:148.
#include "../common/cli.hpp"
#include "alpakaConfig.hpp" // reference: example/reduce/src
#include "kernel.hpp" // reference: example/reduce/src

#include <alpaka/alpaka.hpp>

#include <chrono>
#include <cstdlib>
#include <iostream>
#include <sstream>

using Accelerator = GpuCudaRt; // = the B200 accelerator (AccGpuCudaRt is an alias of AccGpuB200)
using Acc = Accelerator::Acc;
using QueueAcc = alpaka::Queue<Acc, alpaka::Blocking>;

namespace
{
    constexpr uint64_t blockSize = getMaxBlockSize<Accelerator, 256>();

    struct Timing
    {
        double bestSeconds = 0;
    };

    //! reduce.cpp:47-108 with the device buffers and the functor passed in
    template<typename T, typename TFunc, typename DevHost, typename DevAcc>
    auto reduce(DevHost const& devHost, DevAcc const& devAcc, QueueAcc& queue, uint64_t n, alpaka::Buf<DevAcc, T, Dim, Extent>& source, TFunc func, int runs, Timing& timing)
        -> T
    {
        auto blockCount = static_cast<uint32_t>(alpaka::getAccDevProps<Acc>(devAcc).m_multiProcessorCount * 8);
        auto const maxBlockCount = static_cast<uint32_t>((((n + 1) / 2) - 1) / blockSize + 1);
        if(blockCount > maxBlockCount)
            blockCount = maxBlockCount;

        alpaka::Buf<DevAcc, T, Dim, Extent> destination = alpaka::allocBuf<T, Idx>(devAcc, static_cast<Extent>(blockCount));

        ReduceKernel<blockSize, T, TFunc> kernel1, kernel2;
        WorkDiv workDiv1{static_cast<Extent>(blockCount), static_cast<Extent>(blockSize), static_cast<Extent>(1)};
        WorkDiv workDiv2{static_cast<Extent>(1), static_cast<Extent>(blockSize), static_cast<Extent>(1)};

        auto const taskMain
            = alpaka::createTaskKernel<Acc>(workDiv1, kernel1, std::data(source), std::data(destination), n, func);
        auto const taskLast = alpaka::createTaskKernel<Acc>(
            workDiv2,
            kernel2,
            std::data(destination),
            std::data(destination),
            blockCount,
            func);

        timing.bestSeconds = 1e30;
        for(int r = 0; r < runs; ++r)
        {
            auto const t0 = std::chrono::high_resolution_clock::now();
            alpaka::enqueue(queue, taskMain);
            alpaka::enqueue(queue, taskLast);
            alpaka::wait(queue);
            auto const t1 = std::chrono::high_resolution_clock::now();
            double const s = std::chrono::duration<double>(t1 - t0).count();
            if(r > 0 || runs == 1)
                timing.bestSeconds = std::min(timing.bestSeconds, s);
        }

        T result;
        auto resultView = alpaka::createView(devHost, &result, static_cast<Extent>(blockSize));
        alpaka::memcpy(queue, resultView, destination, 1);
        alpaka::wait(queue);
        return result;
    }

    template<typename T>
    auto run(cli::Args const& args) -> int
    {
        uint64_t const n = args.u64("n", uint64_t{1} << 28);
        int const runs = static_cast<int>(args.u64("runs", 5));
        std::string const mode = args.str("mode", "both");

        auto const devHost = alpaka::getDevByIdx(alpaka::PlatformCpu{}, 0);
        auto const devAcc = alpaka::getDevByIdx(alpaka::Platform<Acc>{}, 0);
        QueueAcc queue(devAcc);

        auto hostMemory = alpaka::allocBuf<T, Idx>(devHost, n);
        T* h = std::data(hostMemory);
        bool const fromFile = args.has("input");
        if(fromFile)
            cli::readFile(args.str("input"), h, n * sizeof(T));
        else
            for(uint64_t i = 0; i < n; ++i)
                h[i] = static_cast<T>(i + 1); // reduce.cpp:137-138

        alpaka::Buf<alpaka::DevB200, T, Dim, Extent> source = alpaka::allocBuf<T, Idx>(devAcc, n);
        alpaka::memcpy(queue, source, hostMemory, n);

        bool ok = true;
        std::ostringstream json;
        json << "{\"driver\": \"reduce_b200\", \"n\": " << n << ", \"elem_bytes\": " << sizeof(T);
        T results[2] = {};
        int k = 0;
        auto report = [&](char const* name, T result, Timing const& t)
        {
            double const gbs = double(n) * sizeof(T) * 1e-9 / t.bestSeconds;
            std::cout << name << ": result = " << +result << ", best " << t.bestSeconds * 1e3 << " ms, " << gbs << " GB/s" << std::endl;
            json << ", \"" << name << "\": {\"result\": " << +result << ", \"gbs\": " << gbs << "}";
            results[k++] = result;
        };
        if(mode == "lambda" || mode == "both")
        {
            Timing t;
            auto addFn = [] ALPAKA_FN_HOST_ACC(T a, T b) -> T { return a + b; };
            T const r = reduce<T>(devHost, devAcc, queue, n, source, addFn, runs, t);
            report("lambda", r, t);
        }
        if(mode == "sum" || mode == "both")
        {
            Timing t;
            T const r = reduce<T>(devHost, devAcc, queue, n, source, alpaka::b200::Sum<T>{}, runs, t);
            report("sum", r, t);
        }
        if(k == 2 && std::is_integral_v<T> && results[0] != results[1])
        {
            std::cerr << "Results don't match between the two paths" << std::endl;
            ok = false;
        }
        if(!fromFile && std::is_integral_v<T>)
        {
            T const expected = static_cast<T>(n / 2 * (n + 1)); // reduce.cpp:148
            for(int i = 0; i < k; ++i)
                if(results[i] != expected)
                {
                    std::cerr << "Results don't match: " << +results[i] << " != " << +expected << "\n";
                    ok = false;
                }
        }
        if(args.has("output"))
            cli::writeFile(args.str("output"), results, sizeof(results));
        json << ", \"ok\": " << (ok ? "true" : "false") << "}";
        std::cout << json.str() << std::endl;
        std::cout << (ok ? "Results match.\n" : "Results differ.\n");
        return ok ? EXIT_SUCCESS : EXIT_FAILURE;
    }
} // namespace

auto main(int argc, char** argv) -> int
{
    try
    {
        cli::Args const args(argc, argv);
        std::string const dtype = args.str("dtype", "u32");
        if(dtype == "u32")
            return run<uint32_t>(args);
        if(dtype == "f32")
            return run<float>(args);
        if(dtype == "f64")
            return run<double>(args);
        if(dtype == "u64")
            return run<uint64_t>(args);
        std::cerr << "unknown --dtype " << dtype << std::endl;
        return 2;
    }
    catch(std::exception const& e)
    {
        std::cerr << "reduce_b200: " << e.what() << std::endl;
        return 2;
    }
}
