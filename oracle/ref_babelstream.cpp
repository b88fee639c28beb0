/* oracle/_ref harness, part 1/3: BabelStream on the reference's own CPU back-ends.
 *
 * TEST INFRASTRUCTURE ONLY. This file is never part of the product path; it is compiled ONLY into
 * oracle/_ref/libalpaka_ref.so by oracle/Makefile, against the UNMODIFIED reference headers where they
 * lie under /root/reference (-I flags, nothing is copied). The reference kernel functors
 * (InitKernel, CopyKernel, MultKernel, AddKernel, TriadKernel, DotKernel) are taken verbatim by including
 * the reference driver translation unit benchmarks/babelstream/src/babelStreamMainTest.cpp:53-181 with its
 * main() renamed and Catch2 replaced by oracle/catch2_stub. They are launched exactly as the driver does
 * (babelStreamMainTest.cpp:208-339): device 0, Blocking queue, getValidWorkDiv, alpaka::exec, alpaka::wait.
 *
 * Only `tests/`, `__graft_entry__.smoke()` and bench.py's cpu_baseline / --impl reference legs may load the
 * resulting library.
 */
#define main b200_ref_babelstream_driver_main
#include "babelStreamMainTest.cpp" // from -I/root/reference/benchmarks/babelstream/src
#undef main

#include <chrono>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <omp.h>
#include <vector>

namespace
{
    //! Nstream is ABSENT from the reference at this commit: "parity unpinned".
    //! Defined here in the reference's own functor style with upstream BabelStream semantics
    //! a[i] += b[i] + scalar * c[i], so that it runs through the reference back-end machinery.
    struct NstreamKernel
    {
        template<typename TAcc, typename T>
        ALPAKA_FN_ACC void operator()(TAcc const& acc, T* a, T const* b, T const* c) const
        {
            const T scalar = static_cast<T>(scalarVal);
            auto const [i] = alpaka::getIdx<alpaka::Grid, alpaka::Threads>(acc);
            a[i] += b[i] + scalar * c[i];
        }
    };

    using AccOmp = alpaka::AccCpuOmp2Blocks<alpaka::DimInt<1u>, std::uint32_t>;
    using AccSer = alpaka::AccCpuSerial<alpaka::DimInt<1u>, std::uint32_t>;

    template<typename TAcc, typename T>
    struct Launcher
    {
        using Idx = alpaka::Idx<TAcc>;
        using Dim = alpaka::Dim<TAcc>;
        using Vec = alpaka::Vec<Dim, Idx>;
        using Queue = alpaka::Queue<TAcc, alpaka::Blocking>;

        alpaka::Platform<TAcc> platform{};
        alpaka::Dev<TAcc> dev;
        Queue queue;
        Idx n;
        alpaka::KernelCfg<TAcc> cfg;

        explicit Launcher(std::uint64_t n_)
            : dev(alpaka::getDevByIdx(platform, 0))
            , queue(dev)
            , n(static_cast<Idx>(n_))
            , cfg{Vec::all(static_cast<Idx>(n_)), Vec::all(static_cast<Idx>(1))}
        {
        }

        // One launch of kernel `k` the way babelStreamMainTest.cpp:305-339 does it.
        void run(int k, T* a, T* b, T* c, T initA)
        {
            switch(k)
            {
            case 0:
                {
                    auto wd = alpaka::getValidWorkDiv(cfg, dev, InitKernel(), a, b, c, initA);
                    alpaka::exec<TAcc>(queue, wd, InitKernel(), a, b, c, initA);
                    break;
                }
            case 1:
                {
                    auto wd = alpaka::getValidWorkDiv(cfg, dev, CopyKernel(), static_cast<T const*>(a), b);
                    alpaka::exec<TAcc>(queue, wd, CopyKernel(), static_cast<T const*>(a), b);
                    break;
                }
            case 2:
                {
                    auto wd = alpaka::getValidWorkDiv(cfg, dev, MultKernel(), a, b);
                    alpaka::exec<TAcc>(queue, wd, MultKernel(), a, b);
                    break;
                }
            case 3:
                {
                    auto wd = alpaka::getValidWorkDiv(
                        cfg,
                        dev,
                        AddKernel(),
                        static_cast<T const*>(a),
                        static_cast<T const*>(b),
                        c);
                    alpaka::exec<TAcc>(queue, wd, AddKernel(), static_cast<T const*>(a), static_cast<T const*>(b), c);
                    break;
                }
            case 4:
                {
                    auto wd = alpaka::getValidWorkDiv(
                        cfg,
                        dev,
                        TriadKernel(),
                        static_cast<T const*>(a),
                        static_cast<T const*>(b),
                        c);
                    alpaka::exec<TAcc>(
                        queue,
                        wd,
                        TriadKernel(),
                        static_cast<T const*>(a),
                        static_cast<T const*>(b),
                        c);
                    break;
                }
            case 5:
                {
                    auto wd = alpaka::getValidWorkDiv(
                        cfg,
                        dev,
                        NstreamKernel(),
                        a,
                        static_cast<T const*>(b),
                        static_cast<T const*>(c));
                    alpaka::exec<TAcc>(
                        queue,
                        wd,
                        NstreamKernel(),
                        a,
                        static_cast<T const*>(b),
                        static_cast<T const*>(c));
                    break;
                }
            default:
                break;
            }
            alpaka::wait(queue);
        }

        // DotKernel with WorkDiv {gridBlocks, 1, 1}: the only block size a single-thread CPU acc accepts
        // (acc/AccCpuOmp2Blocks.hpp:192-197). The shipped driver gates Dot to GPU tags
        // (babelStreamMainTest.cpp:372); host-side finish is std::reduce as in :402-403.
        T dot(T const* a, T const* b, std::uint32_t gridBlocks, T* partials)
        {
            using WorkDiv = alpaka::WorkDivMembers<Dim, Idx>;
            auto const wd = WorkDiv{Vec{static_cast<Idx>(gridBlocks)}, Vec{static_cast<Idx>(1)}, Vec::all(1)};
            alpaka::exec<TAcc>(queue, wd, DotKernel(), a, b, partials, n);
            alpaka::wait(queue);
            return std::reduce(partials, partials + gridBlocks, T{0});
        }
    };

    template<typename TAcc, typename T>
    int runOne(int k, void* a, void* b, void* c, double initA, std::uint64_t n)
    {
        Launcher<TAcc, T> l(n);
        l.run(k, static_cast<T*>(a), static_cast<T*>(b), static_cast<T*>(c), static_cast<T>(initA));
        return 0;
    }

    template<typename TAcc, typename T>
    double dotOne(void const* a, void const* b, std::uint64_t n, std::uint32_t g, void* partials)
    {
        Launcher<TAcc, T> l(n);
        std::vector<T> local;
        T* p = static_cast<T*>(partials);
        if(p == nullptr)
        {
            local.resize(g);
            p = local.data();
        }
        return static_cast<double>(l.dot(static_cast<T const*>(a), static_cast<T const*>(b), g, p));
    }

    // measureKernelExec semantics (babelStreamMainTest.cpp:279-301): host clock around exec + wait.
    template<typename TAcc, typename T>
    int timeOne(int k, std::uint64_t n, int runs, double* secondsOut, std::uint32_t dotBlocks)
    {
        Launcher<TAcc, T> l(n);
        auto const devHost = alpaka::getDevByIdx(alpaka::PlatformCpu{}, 0);
        using Idx = std::uint32_t;
        auto bufA = alpaka::allocBuf<T, Idx>(devHost, static_cast<Idx>(n));
        auto bufB = alpaka::allocBuf<T, Idx>(devHost, static_cast<Idx>(n));
        auto bufC = alpaka::allocBuf<T, Idx>(devHost, static_cast<Idx>(n));
        T* a = std::data(bufA);
        T* b = std::data(bufB);
        T* c = std::data(bufC);
        std::vector<T> partials(dotBlocks ? dotBlocks : 1);
        // first touch through the Init kernel itself, as the driver does (first timed kernel is Init).
        l.run(0, a, b, c, static_cast<T>(valA));
        for(int r = 0; r < runs; ++r)
        {
            auto const t0 = std::chrono::high_resolution_clock::now();
            if(k == 6)
                (void) l.dot(a, b, dotBlocks, partials.data());
            else
                l.run(k, a, b, c, static_cast<T>(valA));
            auto const t1 = std::chrono::high_resolution_clock::now();
            secondsOut[r] = std::chrono::duration<double>(t1 - t0).count();
        }
        return 0;
    }
} // namespace

extern "C"
{
    //! kernel: 0 Init, 1 Copy, 2 Mul, 3 Add, 4 Triad, 5 Nstream. dtype: 0 float, 1 double.
    //! acc: 0 AccCpuSerial, 1 AccCpuOmp2Blocks. Scalar is the reference's fixed scalarVal = 2.0f
    //! (babelStreamCommon.hpp:31). Returns 0 on success, -1 on bad arguments, -2 on exception.
    int ref_babelstream_run(int acc, int kernel, int dtype, void* a, void* b, void* c, double initA, std::uint64_t n)
    {
        try
        {
            if(kernel < 0 || kernel > 5 || n == 0 || n > 0xffffffffull)
                return -1;
            if(acc == 1)
                return dtype == 0 ? runOne<AccOmp, float>(kernel, a, b, c, initA, n)
                                  : runOne<AccOmp, double>(kernel, a, b, c, initA, n);
            return dtype == 0 ? runOne<AccSer, float>(kernel, a, b, c, initA, n)
                              : runOne<AccSer, double>(kernel, a, b, c, initA, n);
        }
        catch(...)
        {
            return -2;
        }
    }

    //! Reference DotKernel on a CPU acc with WorkDiv {gridBlocks,1,1}; partials may be NULL.
    double ref_babelstream_dot(
        int acc,
        int dtype,
        void const* a,
        void const* b,
        std::uint64_t n,
        std::uint32_t gridBlocks,
        void* partials)
    {
        if(acc == 1)
            return dtype == 0 ? dotOne<AccOmp, float>(a, b, n, gridBlocks, partials)
                              : dotOne<AccOmp, double>(a, b, n, gridBlocks, partials);
        return dtype == 0 ? dotOne<AccSer, float>(a, b, n, gridBlocks, partials)
                          : dotOne<AccSer, double>(a, b, n, gridBlocks, partials);
    }

    //! Times `runs` launches of one kernel (6 = Dot) on AccCpuOmp2Blocks with the reference's own method.
    int ref_babelstream_time(int kernel, int dtype, std::uint64_t n, int runs, double* secondsOut, std::uint32_t dotBlocks)
    {
        try
        {
            return dtype == 0 ? timeOne<AccOmp, float>(kernel, n, runs, secondsOut, dotBlocks)
                              : timeOne<AccOmp, double>(kernel, n, runs, secondsOut, dotBlocks);
        }
        catch(...)
        {
            return -2;
        }
    }

    int ref_omp_max_threads()
    {
        return omp_get_max_threads();
    }

    void ref_omp_set_threads(int n)
    {
        omp_set_num_threads(n);
    }
}
