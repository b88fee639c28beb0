// examples/babelstream/babelstream_b200.cpp -- BabelStream on the B200 back-end through the alpaka API.
//
// The parameterised counterpart of the reference driver (benchmarks/babelstream/src/babelStreamMainTest.cpp:186-460):
// same kernel sequence (Init, Copy, Mult, Add, Triad, Dot), same launch shapes (getValidWorkDiv over N elements, one
// element per thread; Dot with {256 blocks, 1024 threads}), same timing method (host clock around exec + wait, minimum
// excluding the first run, babelStreamCommon.hpp:168-206) and the same verification (A=1, B=2, C=5, Dot=2N). Added:
//   * Nstream (a[i] += b[i] + scalar*c[i]; upstream BabelStream; absent from the reference),
//   * --devices=K: contiguous slabs of --array-size elements on each of K devices of one process, one queue per device,
//     Dot partials combined on the host in device order,
//   * --parity-kernel=<name> --input=<file> --output=<file>: run ONE kernel once on given a,b,c and dump a,b,c (and the
//     Dot value) -- used by tests/test_gpu_cpp_layer.py to compare this path bit for bit with the oracle.
// The functors below are ordinary alpaka kernels. Built with -DALPAKA_B200_RECOGNIZE_REFERENCE_KERNELS their type names
// are recognised (alpaka/b200/Native.hpp) and the launches run the hand-written sm_100a kernels; with
// ALPAKA_B200_NATIVE=0 in the environment they run through the generic trampoline instead.
#include "../common/cli.hpp"

#include <alpaka/alpaka.hpp>

#include <algorithm>
#include <chrono>
#include <iomanip>
#include <iostream>
#include <numeric>
#include <sstream>
#include <string>
#include <vector>

#ifdef BABELSTREAM_RENAMED_FUNCTORS
// The same functors under names the library has never heard of (build/examples/babelstream_b200_renamed): nothing is
// recognised, every launch takes the GENERIC path -- the block-coarsened trampoline where the launch proves it safe.
#    define InitKernel UserInit
#    define CopyKernel UserCopy
#    define MultKernel UserScale
#    define AddKernel UserSum
#    define TriadKernel UserTriad
#    define NstreamKernel UserNstream
#    define DotKernel UserDot
#endif

constexpr double scalarVal = 2.0;
constexpr double valA = 1.0;
constexpr unsigned dotBlockThreads = 1024;
constexpr unsigned dotGridBlocks = 256;

struct InitKernel
{
    template<typename TAcc, typename T>
    ALPAKA_FN_ACC void operator()(TAcc const& acc, T* a, T* b, T* c, T initA) const
    {
        auto const i = alpaka::getIdx<alpaka::Grid, alpaka::Threads>(acc)[0];
        a[i] = initA;
        b[i] = T(0);
        c[i] = T(0);
    }
};

struct CopyKernel
{
    template<typename TAcc, typename T>
    ALPAKA_FN_ACC void operator()(TAcc const& acc, T const* a, T* b) const
    {
        auto const i = alpaka::getIdx<alpaka::Grid, alpaka::Threads>(acc)[0];
        b[i] = a[i];
    }
};

struct MultKernel
{
    template<typename TAcc, typename T>
    ALPAKA_FN_ACC void operator()(TAcc const& acc, T const* a, T* b) const
    {
        auto const i = alpaka::getIdx<alpaka::Grid, alpaka::Threads>(acc)[0];
        b[i] = T(scalarVal) * a[i];
    }
};

struct AddKernel
{
    template<typename TAcc, typename T>
    ALPAKA_FN_ACC void operator()(TAcc const& acc, T const* a, T const* b, T* c) const
    {
        auto const i = alpaka::getIdx<alpaka::Grid, alpaka::Threads>(acc)[0];
        c[i] = a[i] + b[i];
    }
};

struct TriadKernel
{
    template<typename TAcc, typename T>
    ALPAKA_FN_ACC void operator()(TAcc const& acc, T const* a, T const* b, T* c) const
    {
        auto const i = alpaka::getIdx<alpaka::Grid, alpaka::Threads>(acc)[0];
        c[i] = a[i] + T(scalarVal) * b[i];
    }
};

struct NstreamKernel
{
    template<typename TAcc, typename T>
    ALPAKA_FN_ACC void operator()(TAcc const& acc, T* a, T const* b, T const* c) const
    {
        auto const i = alpaka::getIdx<alpaka::Grid, alpaka::Threads>(acc)[0];
        a[i] += b[i] + T(scalarVal) * c[i];
    }
};

struct DotKernel
{
    template<typename TAcc, typename T>
    ALPAKA_FN_ACC void operator()(TAcc const& acc, T const* a, T const* b, T* sum, alpaka::Idx<TAcc> n) const
    {
        auto& partial = alpaka::declareSharedVar<T[dotBlockThreads], __COUNTER__>(acc);
        auto const t = alpaka::getIdx<alpaka::Block, alpaka::Threads>(acc)[0];
        auto const stride = alpaka::getWorkDiv<alpaka::Grid, alpaka::Threads>(acc)[0];
        T s = 0;
        for(auto i = alpaka::getIdx<alpaka::Grid, alpaka::Threads>(acc)[0]; i < n; i += stride)
            s += a[i] * b[i];
        partial[t] = s;
        for(auto half = alpaka::getWorkDiv<alpaka::Block, alpaka::Threads>(acc)[0] / 2; half > 0; half /= 2)
        {
            alpaka::syncBlockThreads(acc);
            if(t < half)
                partial[t] += partial[t + half];
        }
        if(t == 0)
            sum[alpaka::getIdx<alpaka::Grid, alpaka::Blocks>(acc)[0]] = partial[0];
    }
};

namespace
{
    using Dim = alpaka::DimInt<1u>;
    using Idx = std::uint32_t; // the reference driver's index type; array sizes up to 2^32-1 per device
    using Acc = alpaka::AccGpuB200<Dim, Idx>;
    using Vec = alpaka::Vec<Dim, Idx>;
    using WorkDiv = alpaka::WorkDivMembers<Dim, Idx>;
    using Queue = alpaka::Queue<Acc, alpaka::NonBlocking>;

    template<typename T>
    struct Shard
    {
        alpaka::DevB200 dev;
        Queue queue;
        alpaka::BufB200<T, Dim, Idx> a, b, c, sum;
        Shard(alpaka::DevB200 const& d, Idx n)
            : dev(d)
            , queue(d)
            , a(alpaka::allocBuf<T, Idx>(d, n))
            , b(alpaka::allocBuf<T, Idx>(d, n))
            , c(alpaka::allocBuf<T, Idx>(d, n))
            , sum(alpaka::allocBuf<T, Idx>(d, Idx{dotGridBlocks}))
        {
        }
    };

    template<typename T>
    struct Bench
    {
        Idx n;
        std::vector<Shard<T>> shards;
        WorkDiv wd;
        WorkDiv wdDot{Vec{dotGridBlocks}, Vec{dotBlockThreads}, Vec{1u}};

        Bench(Idx n_, unsigned devices) : n(n_), wd(Vec{1u}, Vec{1u}, Vec{1u})
        {
            auto const platform = alpaka::Platform<Acc>{};
            if(alpaka::getDevCount(platform) < devices)
                throw std::runtime_error("not enough B200 devices for --devices");
            for(unsigned d = 0; d < devices; ++d)
                shards.emplace_back(alpaka::getDevByIdx(platform, d), n);
            alpaka::KernelCfg<Acc> const cfg{Vec{n}, Vec{1u}};
            auto& s0 = shards[0];
            wd = alpaka::getValidWorkDiv(cfg, s0.dev, TriadKernel{}, ptrC(s0.a), ptrC(s0.b), std::data(s0.c));
        }

        static auto ptrC(alpaka::BufB200<T, Dim, Idx> const& buf) -> T const*
        {
            return std::data(buf);
        }

        void launch(std::string const& k)
        {
            for(auto& s : shards)
            {
                T* a = std::data(s.a);
                T* b = std::data(s.b);
                T* c = std::data(s.c);
                if(k == "init")
                    alpaka::exec<Acc>(s.queue, wd, InitKernel{}, a, b, c, static_cast<T>(valA));
                else if(k == "copy")
                    alpaka::exec<Acc>(s.queue, wd, CopyKernel{}, static_cast<T const*>(a), b);
                else if(k == "mul")
                    alpaka::exec<Acc>(s.queue, wd, MultKernel{}, static_cast<T const*>(a), b);
                else if(k == "add")
                    alpaka::exec<Acc>(s.queue, wd, AddKernel{}, static_cast<T const*>(a), static_cast<T const*>(b), c);
                else if(k == "triad")
                    alpaka::exec<Acc>(s.queue, wd, TriadKernel{}, static_cast<T const*>(a), static_cast<T const*>(b), c);
                else if(k == "nstream")
                    alpaka::exec<Acc>(s.queue, wd, NstreamKernel{}, a, static_cast<T const*>(b), static_cast<T const*>(c));
                else if(k == "dot")
                    alpaka::exec<Acc>(
                        s.queue,
                        wdDot,
                        DotKernel{},
                        static_cast<T const*>(a),
                        static_cast<T const*>(b),
                        std::data(s.sum),
                        n);
                else
                    throw std::runtime_error("unknown kernel " + k);
            }
        }

        void waitAll()
        {
            for(auto& s : shards)
                alpaka::wait(s.queue);
        }

        //! host finish of Dot: std::reduce over each device's block sums, devices combined in order
        auto dotResult() -> T
        {
            auto const host = alpaka::getDevByIdx(alpaka::PlatformCpu{}, 0);
            T total = 0;
            for(auto& s : shards)
            {
                auto h = alpaka::allocBuf<T, Idx>(host, Idx{dotGridBlocks});
                alpaka::memcpy(s.queue, h, s.sum);
                alpaka::wait(s.queue);
                T const* p = std::data(h);
                total += std::reduce(p, p + dotGridBlocks, T{0});
            }
            return total;
        }
    };

    template<typename T>
    auto parity(cli::Args const& args) -> int
    {
        auto const n = static_cast<Idx>(args.u64("array-size", 1u << 20));
        std::string const k = args.str("parity-kernel");
        Bench<T> bench(n, 1);
        auto& s = bench.shards[0];
        auto const host = alpaka::getDevByIdx(alpaka::PlatformCpu{}, 0);
        auto h = alpaka::allocBuf<T, Idx>(host, Idx{3u * n});
        cli::readFile(args.str("input"), std::data(h), sizeof(T) * 3u * n);
        auto view = [&](Idx which) { return alpaka::createView(host, std::data(h) + std::size_t(which) * n, n); };
        auto va = view(0), vb = view(1), vc = view(2);
        alpaka::memcpy(s.queue, s.a, va);
        alpaka::memcpy(s.queue, s.b, vb);
        alpaka::memcpy(s.queue, s.c, vc);
        if(n % 1024u != 0u && k != "dot")
        {
            // ragged sizes: the functors have no bounds check, so pick a work division that covers n exactly
            bench.wd = WorkDiv{Vec{n}, Vec{1u}, Vec{1u}};
            for(Idx bt = 1024u; bt >= 1u; bt /= 2u)
                if(n % bt == 0u)
                {
                    bench.wd = WorkDiv{Vec{n / bt}, Vec{bt}, Vec{1u}};
                    break;
                }
        }
        bench.launch(k);
        alpaka::memcpy(s.queue, va, s.a);
        alpaka::memcpy(s.queue, vb, s.b);
        alpaka::memcpy(s.queue, vc, s.c);
        alpaka::wait(s.queue);
        std::vector<char> out(sizeof(T) * 3u * n + sizeof(T));
        std::memcpy(out.data(), std::data(h), sizeof(T) * 3u * n);
        T dot = 0;
        if(k == "dot")
            dot = bench.dotResult();
        std::memcpy(out.data() + sizeof(T) * 3u * n, &dot, sizeof(T));
        cli::writeFile(args.str("output"), out.data(), out.size());
        std::cout << "parity " << k << " n=" << n << " workdiv " << bench.wd << " done" << std::endl;
        return 0;
    }

    template<typename T>
    auto benchmark(cli::Args const& args) -> int
    {
        auto const n = static_cast<Idx>(args.u64("array-size", 1u << 25));
        auto const runs = static_cast<int>(args.u64("number-runs", 20));
        auto const devices = static_cast<unsigned>(args.u64("devices", 1));
        Bench<T> bench(n, devices);
        std::cout << "AcceleratorType:" << alpaka::getAccName<Acc>() << "\nDeviceName:" << alpaka::getName(bench.shards[0].dev)
                  << "\nDevices:" << devices << "\nDataSize(items per device):" << n << "\nPrecision:"
                  << (sizeof(T) == 8 ? "double" : "single") << "\nNumberOfRuns:" << runs << "\nWorkDiv:" << bench.wd
                  << "\nNativeKernels:" << (alpaka::b200::nativeKernelsEnabled() ? "on" : "off") << std::endl;

        struct Row
        {
            std::string name;
            double arrays;
            double minS, maxS, avgS;
        };
        std::vector<Row> rows;
        auto measure = [&](std::string const& k, double arrays)
        {
            std::vector<double> t;
            for(int r = 0; r < runs; ++r)
            {
                auto const t0 = std::chrono::high_resolution_clock::now();
                bench.launch(k);
                bench.waitAll();
                auto const t1 = std::chrono::high_resolution_clock::now();
                t.push_back(std::chrono::duration<double>(t1 - t0).count());
            }
            // the first run is excluded (babelStreamCommon.hpp:168-206)
            auto const first = t.size() > 1 ? t.begin() + 1 : t.begin();
            double const mn = *std::min_element(first, t.end());
            double const mx = *std::max_element(first, t.end());
            double const avg = std::accumulate(first, t.end(), 0.0) / double(t.end() - first);
            rows.push_back({k, arrays, mn, mx, avg});
        };
        measure("init", 3); // true traffic: three arrays written (the reference books two, :415)
        measure("copy", 2);
        measure("mul", 2);
        measure("add", 3);
        measure("triad", 3);

        // verification, as the reference: A = 1, B = 2, C = 5 (babelStreamMainTest.cpp:343-368)
        auto const host = alpaka::getDevByIdx(alpaka::PlatformCpu{}, 0);
        bool ok = true;
        for(auto& s : bench.shards)
        {
            auto h = alpaka::allocBuf<T, Idx>(host, n);
            auto check = [&](alpaka::BufB200<T, Dim, Idx>& d, T expected, char const* name)
            {
                alpaka::memcpy(s.queue, h, d);
                alpaka::wait(s.queue);
                for(Idx i = 0; i < n; ++i)
                    if(h[i] != expected)
                    {
                        std::cerr << "verification failed: " << name << "[" << i << "] = " << h[i] << " != " << expected << std::endl;
                        ok = false;
                        return;
                    }
            };
            check(s.c, static_cast<T>(valA + scalarVal * scalarVal * valA), "c");
            check(s.b, static_cast<T>(scalarVal * valA), "b");
            check(s.a, static_cast<T>(valA), "a");
        }
        measure("dot", 2);
        T const dot = bench.dotResult();
        T const expectedDot = static_cast<T>(2.0 * double(n) * devices);
        if(std::abs(double(dot - expectedDot)) > 100.0 * double(std::numeric_limits<T>::epsilon()) * double(expectedDot))
        {
            std::cerr << "verification failed: dot = " << dot << " != " << expectedDot << std::endl;
            ok = false;
        }
        measure("nstream", 4); // last: it is not idempotent, so it runs after the verification

        std::cout << std::left << std::setw(10) << "Kernel" << std::setw(16) << "Bandwidth(GB/s)" << std::setw(14) << "MinTime(s)"
                  << std::setw(14) << "MaxTime(s)" << std::setw(14) << "AvgTime(s)" << "\n";
        std::ostringstream json;
        json << "{\"driver\": \"babelstream_b200\", \"devices\": " << devices << ", \"n_per_device\": " << n << ", \"dtype\": \""
             << (sizeof(T) == 8 ? "f64" : "f32") << "\", \"native\": " << (alpaka::b200::nativeKernelsEnabled() ? "true" : "false")
             << ", \"gbs\": {";
        for(std::size_t i = 0; i < rows.size(); ++i)
        {
            auto const& r = rows[i];
            double const gbs = r.arrays * sizeof(T) * double(n) * devices * 1e-9 / r.minS;
            std::cout << std::left << std::setw(10) << r.name << std::setw(16) << gbs << std::setw(14) << r.minS << std::setw(14)
                      << r.maxS << std::setw(14) << r.avgS << "\n";
            json << (i ? ", " : "") << "\"" << r.name << "\": " << gbs;
        }
        json << "}, \"verified\": " << (ok ? "true" : "false") << "}";
        std::cout << json.str() << std::endl;
        return ok ? 0 : 1;
    }
} // namespace

auto main(int argc, char** argv) -> int
{
    try
    {
        cli::Args const args(argc, argv);
        bool const single = args.str("precision", "double") == "float" || args.str("precision", "double") == "single";
        if(args.has("parity-kernel"))
            return single ? parity<float>(args) : parity<double>(args);
        return single ? benchmark<float>(args) : benchmark<double>(args);
    }
    catch(std::exception const& e)
    {
        std::cerr << "babelstream_b200: " << e.what() << std::endl;
        return 2;
    }
}
