// include/alpaka/b200/Native.hpp -- trait::NativeKernel specialisations that route the reference drivers' kernel
// functors to the hand-written sm_100a kernels of libalpaka_b200.so.
//
// Problem: the reference's contract is "arbitrary user functor, one element per thread";
// the generic trampoline therefore reproduces the reference's scalar LDG.E.64/STG.E.64 code. The babelstream functors
// are global-namespace structs defined in the driver's translation unit (benchmarks/babelstream/src/
// babelStreamMainTest.cpp:53-181), the heat functors likewise (example/heatEquation2D/src/StencilKernel.hpp:30-89,
// BoundaryKernel.hpp:24-86). A library can forward-declare those names and specialise a trait on them; that is what
// this header does, so the UNMODIFIED drivers reach the vectorised streams, the single-pass Dot and the TMA stencil.
//
// Opt-in: define ALPAKA_B200_RECOGNIZE_REFERENCE_KERNELS on the build line (the forward declarations below would clash
// with a user type of the same name but a different kind). ALPAKA_B200_NATIVE=0 at run time falls back to the generic
// trampoline for every launch (A/B measurement, see tests/test_gpu_cpp_layer.py).
//
// Safety net: the first launch of each recognised stream functor is cross-checked -- the user's functor (generic
// trampoline) and the native kernel both run on a small scratch input with exactly representable values, and the
// native path is only kept if the results are bit-identical. A same-named functor with different semantics therefore
// keeps working, just without the fast path. A launch whose argument types do not fit a specialisation's `launch`
// signature is not offered to it at all (Kernel.hpp checks callability); a specialisation also returns false when
// the work division or the alignment is not what it was written for.
#pragma once

#include "Kernel.hpp"

#if defined(ALPAKA_B200_RECOGNIZE_REFERENCE_KERNELS) && defined(__CUDACC__)

#    include <atomic>
#    include <cmath>
#    include <map>
#    include <mutex>
#    include <tuple>
#    include <vector>

// the reference drivers' functor names (global namespace)
struct InitKernel;
struct CopyKernel;
struct MultKernel;
struct AddKernel;
struct TriadKernel;
struct DotKernel;
struct NstreamKernel; // not in the reference; examples/babelstream/babelstream_b200.cpp (upstream BabelStream semantics)
template<std::uint32_t TBlockSize, typename T, typename TFunc>
struct ReduceKernel; // example/reduce/src/kernel.hpp:41-132
template<std::size_t T_SharedMemSize1D>
struct StencilKernel;
struct BoundaryKernel;

namespace alpaka::b200::native
{
    //! scalarVal of the reference driver (benchmarks/babelstream/src/babelStreamCommon.hpp:31); cross-checked at run
    //! time against the functor itself, see verifyStream()
    inline constexpr double babelstreamScalar = 2.0;

    template<typename T>
    inline constexpr bool isStreamElem = std::is_same_v<T, float> || std::is_same_v<T, double>;

    // ---- C-ABI entry per element type
    inline auto streamInit(b200_stream_t s, double* a, double* b, double* c, double v, uint64_t n) -> int
    {
        return b200_stream_init_f64(s, a, b, c, v, n);
    }
    inline auto streamInit(b200_stream_t s, float* a, float* b, float* c, float v, uint64_t n) -> int
    {
        return b200_stream_init_f32(s, a, b, c, v, n);
    }
    inline auto streamCopy(b200_stream_t s, double const* a, double* b, uint64_t n) -> int
    {
        return b200_stream_copy_f64(s, a, b, n);
    }
    inline auto streamCopy(b200_stream_t s, float const* a, float* b, uint64_t n) -> int
    {
        return b200_stream_copy_f32(s, a, b, n);
    }
    inline auto streamMul(b200_stream_t s, double const* a, double* b, uint64_t n) -> int
    {
        return b200_stream_mul_f64(s, a, b, babelstreamScalar, n);
    }
    inline auto streamMul(b200_stream_t s, float const* a, float* b, uint64_t n) -> int
    {
        return b200_stream_mul_f32(s, a, b, static_cast<float>(babelstreamScalar), n);
    }
    inline auto streamAdd(b200_stream_t s, double const* a, double const* b, double* c, uint64_t n) -> int
    {
        return b200_stream_add_f64(s, a, b, c, n);
    }
    inline auto streamAdd(b200_stream_t s, float const* a, float const* b, float* c, uint64_t n) -> int
    {
        return b200_stream_add_f32(s, a, b, c, n);
    }
    inline auto streamTriad(b200_stream_t s, double const* a, double const* b, double* c, uint64_t n) -> int
    {
        return b200_stream_triad_f64(s, a, b, c, babelstreamScalar, n);
    }
    inline auto streamTriad(b200_stream_t s, float const* a, float const* b, float* c, uint64_t n) -> int
    {
        return b200_stream_triad_f32(s, a, b, c, static_cast<float>(babelstreamScalar), n);
    }
    inline auto streamNstream(b200_stream_t s, double* a, double const* b, double const* c, uint64_t n) -> int
    {
        return b200_stream_nstream_f64(s, a, b, c, babelstreamScalar, n);
    }
    inline auto streamNstream(b200_stream_t s, float* a, float const* b, float const* c, uint64_t n) -> int
    {
        return b200_stream_nstream_f32(s, a, b, c, static_cast<float>(babelstreamScalar), n);
    }
    inline auto reduceSum(b200_stream_t s, std::uint32_t const* in, uint64_t n, std::uint32_t* out, void* scratch) -> int
    {
        return b200_reduce_sum_u32(s, in, n, out, scratch);
    }
    inline auto reduceSum(b200_stream_t s, std::int32_t const* in, uint64_t n, std::int32_t* out, void* scratch) -> int
    {
        return b200_reduce_sum_i32(s, in, n, out, scratch);
    }
    inline auto reduceSum(b200_stream_t s, std::uint64_t const* in, uint64_t n, std::uint64_t* out, void* scratch) -> int
    {
        return b200_reduce_sum_u64(s, in, n, out, scratch);
    }
    inline auto reduceSum(b200_stream_t s, float const* in, uint64_t n, float* out, void* scratch) -> int
    {
        return b200_reduce_sum_f32(s, in, n, out, scratch);
    }
    inline auto reduceSum(b200_stream_t s, double const* in, uint64_t n, double* out, void* scratch) -> int
    {
        return b200_reduce_sum_f64(s, in, n, out, scratch);
    }
    template<typename T>
    inline constexpr bool isReduceElem = std::is_same_v<T, std::uint32_t> || std::is_same_v<T, std::int32_t>
                                         || std::is_same_v<T, std::uint64_t> || std::is_same_v<T, float> || std::is_same_v<T, double>;

    inline auto dotPartials(b200_stream_t s, double const* a, double const* b, uint64_t n, double* out, uint32_t k, void* scratch)
        -> int
    {
        return b200_dot_partials_f64(s, a, b, n, out, k, scratch);
    }
    inline auto dotPartials(b200_stream_t s, float const* a, float const* b, uint64_t n, float* out, uint32_t k, void* scratch)
        -> int
    {
        return b200_dot_partials_f32(s, a, b, n, out, k, scratch);
    }

    //! number of elements the reference functor touches: one per grid thread (the functors have no bounds check)
    template<typename TDim, typename TIdx>
    auto gridThreads(WorkDivMembers<TDim, TIdx> const& wd) -> uint64_t
    {
        uint64_t n = 1;
        for(std::size_t d = 0; d < TDim::value; ++d)
            n *= static_cast<uint64_t>(wd.m_gridBlockExtent[d]) * static_cast<uint64_t>(wd.m_blockThreadExtent[d]);
        return n;
    }

    enum class Verdict : int
    {
        Untested = 0,
        Native = 1,
        Generic = 2
    };

    //! One verdict per (functor, accelerator, element type).
    template<typename TKernel, typename TAcc, typename T>
    auto verdict() -> std::atomic<int>&
    {
        static std::atomic<int> v{static_cast<int>(Verdict::Untested)};
        return v;
    }

    //! Runs `generic` (the user's functor through the trampoline) and `native` on separate copies of a small input
    //! (values k mod 7 - 3: every product/sum is exact, so FMA contraction cannot matter) and compares all three
    //! arrays bit for bit. Synchronous; executed once per functor type.
    template<typename TKernel, typename TAcc, typename T, typename TQueue, typename FGeneric, typename FNative>
    auto verifyStream(TQueue& queue, FGeneric&& generic, FNative&& native) -> bool
    {
        auto& v = verdict<TKernel, TAcc, T>();
        int const known = v.load(std::memory_order_acquire);
        if(known != static_cast<int>(Verdict::Untested))
            return known == static_cast<int>(Verdict::Native);

        constexpr std::size_t n = 4096; // 4 blocks of 1024 threads
        int const dev = getDev(queue).getNativeHandle();
        b200_stream_t const s = queue.getNativeHandle();
        std::vector<T> h(3 * n), outG(3 * n), outN(3 * n);
        for(std::size_t k = 0; k < 3 * n; ++k)
            h[k] = static_cast<T>(static_cast<int>((k * 2654435761u >> 7) % 7u) - 3);
        void* dG = nullptr;
        void* dN = nullptr;
        check(b200_malloc_async(dev, s, 3 * n * sizeof(T), &dG));
        check(b200_malloc_async(dev, s, 3 * n * sizeof(T), &dN));
        check(b200_memcpy_async(dev, dG, h.data(), 3 * n * sizeof(T), B200_COPY_H2D, s));
        check(b200_memcpy_async(dev, dN, h.data(), 3 * n * sizeof(T), B200_COPY_H2D, s));
        T* g = static_cast<T*>(dG);
        T* m = static_cast<T*>(dN);
        generic(g, g + n, g + 2 * n, n);
        native(m, m + n, m + 2 * n, n);
        check(b200_memcpy_async(dev, outG.data(), dG, 3 * n * sizeof(T), B200_COPY_D2H, s));
        check(b200_memcpy_async(dev, outN.data(), dN, 3 * n * sizeof(T), B200_COPY_D2H, s));
        check(b200_stream_sync(s));
        check(b200_free_async(dev, s, dG));
        check(b200_free_async(dev, s, dN));
        bool const same = std::memcmp(outG.data(), outN.data(), 3 * n * sizeof(T)) == 0;
        if(!same)
            std::cerr << "[alpaka-b200] a kernel functor named like a reference BabelStream kernel computes something "
                         "else; it keeps running through the generic trampoline"
                      << std::endl;
        v.store(static_cast<int>(same ? Verdict::Native : Verdict::Generic), std::memory_order_release);
        return same;
    }

    template<typename TDim, typename TIdx>
    auto verifyWorkDiv() -> WorkDivMembers<TDim, TIdx>
    {
        using V = Vec<TDim, TIdx>;
        return WorkDivMembers<TDim, TIdx>{V::all(4), V::all(1024), V::all(1)};
    }

    //! Cross-check of a functor claimed as the reference DotKernel (babelStreamMainTest.cpp:145-181): the user's functor
    //! (generic trampoline, the caller's block size, 4 blocks) and the native single-pass kernel both reduce a small input
    //! of small integers (every product and partial sum exact, so no summation order or FMA contraction can matter); the
    //! native path is kept only if the two totals are identical. `generic(a, b, partials, n, blocks)` launches the user's
    //! functor, `native(a, b, partials, n, blocks)` the library's; both fill `blocks` partials whose sum is the result.
    template<typename TKernel, typename TAcc, typename T, typename TQueue, typename FGeneric, typename FNative>
    auto verifyDot(TQueue& queue, FGeneric&& generic, FNative&& native) -> bool
    {
        auto& v = verdict<TKernel, TAcc, T>();
        int const known = v.load(std::memory_order_acquire);
        if(known != static_cast<int>(Verdict::Untested))
            return known == static_cast<int>(Verdict::Native);

        constexpr std::size_t n = 10007; // ragged: not a multiple of any block size
        constexpr std::uint32_t blocks = 4;
        int const dev = getDev(queue).getNativeHandle();
        b200_stream_t const s = queue.getNativeHandle();
        std::vector<T> h(2 * n);
        for(std::size_t k = 0; k < 2 * n; ++k)
            h[k] = static_cast<T>(static_cast<int>((k * 2654435761u >> 7) % 7u) - 3);
        void* dIn = nullptr;
        void* dOut = nullptr;
        check(b200_malloc_async(dev, s, 2 * n * sizeof(T), &dIn));
        check(b200_malloc_async(dev, s, 2 * blocks * sizeof(T), &dOut));
        check(b200_memcpy_async(dev, dIn, h.data(), 2 * n * sizeof(T), B200_COPY_H2D, s));
        check(b200_memset_async(dev, dOut, 0, 2 * blocks * sizeof(T), s));
        T* in = static_cast<T*>(dIn);
        T* out = static_cast<T*>(dOut);
        generic(static_cast<T const*>(in), static_cast<T const*>(in + n), out, n, blocks);
        native(static_cast<T const*>(in), static_cast<T const*>(in + n), out + blocks, n, blocks);
        T res[2 * blocks];
        check(b200_memcpy_async(dev, res, dOut, sizeof(res), B200_COPY_D2H, s));
        check(b200_stream_sync(s));
        check(b200_free_async(dev, s, dIn));
        check(b200_free_async(dev, s, dOut));
        T sumG{0}, sumN{0};
        long long want = 0;
        for(std::uint32_t k = 0; k < blocks; ++k)
        {
            sumG += res[k];
            sumN += res[blocks + k];
        }
        for(std::size_t k = 0; k < n; ++k)
            want += static_cast<long long>(h[k]) * static_cast<long long>(h[n + k]);
        bool const same = sumG == sumN && sumN == static_cast<T>(want);
        if(!same)
            std::cerr << "[alpaka-b200] a kernel functor named DotKernel does not compute the reference's blockwise dot "
                         "product; it keeps running through the generic trampoline"
                      << std::endl;
        v.store(static_cast<int>(same ? Verdict::Native : Verdict::Generic), std::memory_order_release);
        return same;
    }

    //! Cross-check of a functor claimed as the reference ReduceKernel<B, T, Sum<T>> (example/reduce/src/kernel.hpp:42-132):
    //! the user's functor (generic trampoline, the driver's two launches: G blocks, then one block over the partials) and the
    //! native single-pass kernel reduce the same small-integer input; kept only if destination[0] is identical.
    template<typename TKernel, typename TAcc, typename T, typename TQueue, typename FGeneric, typename FNative>
    auto verifyReduce(TQueue& queue, FGeneric&& generic, FNative&& native) -> bool
    {
        auto& v = verdict<TKernel, TAcc, T>();
        int const known = v.load(std::memory_order_acquire);
        if(known != static_cast<int>(Verdict::Untested))
            return known == static_cast<int>(Verdict::Native);

        constexpr std::size_t n = 10007;
        constexpr std::uint32_t blocks = 3;
        int const dev = getDev(queue).getNativeHandle();
        b200_stream_t const s = queue.getNativeHandle();
        std::vector<T> h(n);
        long long want = 0;
        for(std::size_t k = 0; k < n; ++k)
        {
            int const x = static_cast<int>((k * 2654435761u >> 7) % 7u); // 0..6: exact in every element type
            h[k] = static_cast<T>(x);
            want += x;
        }
        void* dIn = nullptr;
        void* dOut = nullptr;
        check(b200_malloc_async(dev, s, n * sizeof(T), &dIn));
        check(b200_malloc_async(dev, s, 2 * blocks * sizeof(T), &dOut));
        check(b200_memcpy_async(dev, dIn, h.data(), n * sizeof(T), B200_COPY_H2D, s));
        check(b200_memset_async(dev, dOut, 0, 2 * blocks * sizeof(T), s));
        T* out = static_cast<T*>(dOut);
        generic(static_cast<T const*>(dIn), out, n, blocks);
        native(static_cast<T const*>(dIn), out + blocks, n);
        T res[2 * blocks];
        check(b200_memcpy_async(dev, res, dOut, sizeof(res), B200_COPY_D2H, s));
        check(b200_stream_sync(s));
        check(b200_free_async(dev, s, dIn));
        check(b200_free_async(dev, s, dOut));
        bool const same = res[0] == res[blocks] && res[blocks] == static_cast<T>(want);
        if(!same)
            std::cerr << "[alpaka-b200] a kernel functor named ReduceKernel<..., Sum<T>> does not compute the reference's "
                         "sum; it keeps running through the generic trampoline"
                      << std::endl;
        v.store(static_cast<int>(same ? Verdict::Native : Verdict::Generic), std::memory_order_release);
        return same;
    }

    // ---- heatEquation2D plans: TMA descriptors + boundary tables per ping-pong buffer pair
    struct HeatPlanKey
    {
        int dev;
        void const* lo;
        void const* hi;
        std::size_t pitch;
        uint32_t ny, nx;
        double dx, dy;
        auto operator<(HeatPlanKey const& o) const -> bool
        {
            return std::tie(dev, lo, hi, pitch, ny, nx, dx, dy) < std::tie(o.dev, o.lo, o.hi, o.pitch, o.ny, o.nx, o.dx, o.dy);
        }
    };
    struct HeatPlan
    {
        b200_heat2d_plan_t plan = nullptr;
        double* u0 = nullptr; //!< buffer index 0 of the plan (the lower address)
    };
    struct HeatPlanCache
    {
        std::mutex mutex;
        std::map<HeatPlanKey, HeatPlan> plans;
        ~HeatPlanCache()
        {
            for(auto& kv : plans)
                (void) b200_heat2d_plan_destroy(kv.second.plan);
        }
    };
    inline auto heatPlans() -> HeatPlanCache&
    {
        static HeatPlanCache cache;
        return cache;
    }

    //! finds or builds the plan for the buffer pair {a, b}; the boundary tables are computed HERE, on the host, with
    //! the C library's sin -- the values the reference's CPU back-end produces
    inline auto heatPlanFor(int dev, double* a, double* b, std::size_t pitch, uint32_t ny, uint32_t nx, double dx, double dy)
        -> HeatPlan
    {
        double* lo = a < b ? a : b;
        double* hi = a < b ? b : a;
        HeatPlanKey const key{dev, lo, hi, pitch, ny, nx, dx, dy};
        auto& cache = heatPlans();
        std::lock_guard<std::mutex> l(cache.mutex);
        auto const it = cache.plans.find(key);
        if(it != cache.plans.end())
            return it->second;
        constexpr double pi = math::constants::pi;
        std::vector<double> sx(static_cast<std::size_t>(nx) + 2u), sy(static_cast<std::size_t>(ny) + 2u);
        for(uint32_t i = 0; i < nx + 2u; ++i)
            sx[i] = std::sin(pi * (i * dx)); // exactSolution(idx2D[1] * dx, ...), analyticalSolution.hpp:17-21
        for(uint32_t j = 0; j < ny + 2u; ++j)
            sy[j] = std::sin(pi * (j * dy));
        HeatPlan p;
        p.u0 = lo;
        check(b200_heat2d_plan_create(dev, lo, hi, pitch, ny, nx, sx.data(), sy.data(), B200_EDGE_ALL, &p.plan));
        cache.plans.emplace(key, p);
        return p;
    }

    //! the plan whose pair contains `u` with matching geometry (BoundaryKernel only sees one buffer)
    inline auto heatPlanContaining(int dev, double const* u, std::size_t pitch, uint32_t ny, uint32_t nx, double dx, double dy, HeatPlan& out, int& index)
        -> bool
    {
        auto& cache = heatPlans();
        std::lock_guard<std::mutex> l(cache.mutex);
        for(auto const& [k, p] : cache.plans)
        {
            if(k.dev == dev && k.pitch == pitch && k.ny == ny && k.nx == nx && k.dx == dx && k.dy == dy
               && (k.lo == u || k.hi == u))
            {
                out = p;
                index = (k.lo == u) ? 0 : 1;
                return true;
            }
        }
        return false;
    }
} // namespace alpaka::b200::native

namespace alpaka::trait
{
    // ---------------------------------------------------------------------------------------------------------
    // BabelStream. Launch shape of the driver: 1-D, one element per thread, grid threads == array size
    // (babelStreamMainTest.cpp:245-268, 305-339).
    template<typename TAcc>
    struct NativeKernel<::InitKernel, TAcc>
    {
        static constexpr bool available = true;
        template<typename TQueue, typename TDim, typename TIdx, typename TK, typename T>
        static auto launch(TQueue& q, WorkDivMembers<TDim, TIdx> const& wd, TK const& k, T* a, T* b, T* c, T initA) -> bool
        {
            namespace nv = b200::native;
            if constexpr(!nv::isStreamElem<T> || TDim::value != 1u)
                return false;
            else
            {
                bool const ok = nv::verifyStream<TK, TAcc, T>(
                    q,
                    [&](T* x, T* y, T* z, std::size_t)
                    { b200::launchGeneric<TAcc>(q, nv::verifyWorkDiv<TDim, TIdx>(), k, x, y, z, static_cast<T>(3)); },
                    [&](T* x, T* y, T* z, std::size_t n)
                    { b200::check(nv::streamInit(q.getNativeHandle(), x, y, z, static_cast<T>(3), n)); });
                if(!ok)
                    return false;
                b200::check(nv::streamInit(q.getNativeHandle(), a, b, c, initA, nv::gridThreads(wd)));
                return true;
            }
        }
    };

#    define ALPAKA_B200_NATIVE_STREAM2(KERNEL, CALL)                                                                  \
        template<typename TAcc>                                                                                       \
        struct NativeKernel<::KERNEL, TAcc>                                                                           \
        {                                                                                                             \
            static constexpr bool available = true;                                                                   \
            template<typename TQueue, typename TDim, typename TIdx, typename TK, typename TA, typename T>             \
            static auto launch(TQueue& q, WorkDivMembers<TDim, TIdx> const& wd, TK const& k, TA* a, T* b)             \
                -> std::enable_if_t<std::is_same_v<std::remove_const_t<TA>, T>, bool>                                 \
            {                                                                                                         \
                namespace nv = b200::native;                                                                          \
                if constexpr(!nv::isStreamElem<T> || TDim::value != 1u)                                               \
                    return false;                                                                                     \
                else                                                                                                  \
                {                                                                                                     \
                    bool const ok = nv::verifyStream<TK, TAcc, T>(                                                    \
                        q,                                                                                            \
                        [&](T* x, T* y, T*, std::size_t)                                                              \
                        { b200::launchGeneric<TAcc>(q, nv::verifyWorkDiv<TDim, TIdx>(), k, static_cast<TA*>(x), y); }, \
                        [&](T* x, T* y, T*, std::size_t n) { b200::check(nv::CALL(q.getNativeHandle(), x, y, n)); }); \
                    if(!ok)                                                                                           \
                        return false;                                                                                 \
                    b200::check(nv::CALL(q.getNativeHandle(), a, b, nv::gridThreads(wd)));                            \
                    return true;                                                                                      \
                }                                                                                                     \
            }                                                                                                         \
        };

#    define ALPAKA_B200_NATIVE_STREAM3(KERNEL, CALL)                                                                  \
        template<typename TAcc>                                                                                       \
        struct NativeKernel<::KERNEL, TAcc>                                                                           \
        {                                                                                                             \
            static constexpr bool available = true;                                                                   \
            template<typename TQueue, typename TDim, typename TIdx, typename TK, typename TA, typename TB, typename T> \
            static auto launch(TQueue& q, WorkDivMembers<TDim, TIdx> const& wd, TK const& k, TA* a, TB* b, T* c)      \
                -> std::enable_if_t<                                                                                  \
                    std::is_same_v<std::remove_const_t<TA>, T> && std::is_same_v<std::remove_const_t<TB>, T>,         \
                    bool>                                                                                             \
            {                                                                                                         \
                namespace nv = b200::native;                                                                          \
                if constexpr(!nv::isStreamElem<T> || TDim::value != 1u)                                               \
                    return false;                                                                                     \
                else                                                                                                  \
                {                                                                                                     \
                    bool const ok = nv::verifyStream<TK, TAcc, T>(                                                    \
                        q,                                                                                            \
                        [&](T* x, T* y, T* z, std::size_t) {                                                          \
                            b200::launchGeneric<TAcc>(                                                                \
                                q,                                                                                    \
                                nv::verifyWorkDiv<TDim, TIdx>(),                                                      \
                                k,                                                                                    \
                                static_cast<TA*>(x),                                                                  \
                                static_cast<TB*>(y),                                                                  \
                                z);                                                                                   \
                        },                                                                                            \
                        [&](T* x, T* y, T* z, std::size_t n)                                                          \
                        { b200::check(nv::CALL(q.getNativeHandle(), x, y, z, n)); });                                 \
                    if(!ok)                                                                                           \
                        return false;                                                                                 \
                    b200::check(nv::CALL(q.getNativeHandle(), a, b, c, nv::gridThreads(wd)));                         \
                    return true;                                                                                      \
                }                                                                                                     \
            }                                                                                                         \
        };

    ALPAKA_B200_NATIVE_STREAM2(CopyKernel, streamCopy)
    ALPAKA_B200_NATIVE_STREAM2(MultKernel, streamMul)
    ALPAKA_B200_NATIVE_STREAM3(AddKernel, streamAdd)
    ALPAKA_B200_NATIVE_STREAM3(TriadKernel, streamTriad)
#    undef ALPAKA_B200_NATIVE_STREAM2
#    undef ALPAKA_B200_NATIVE_STREAM3

    template<typename TAcc>
    struct NativeKernel<::NstreamKernel, TAcc>
    {
        static constexpr bool available = true;
        template<typename TQueue, typename TDim, typename TIdx, typename TK, typename T>
        static auto launch(TQueue& q, WorkDivMembers<TDim, TIdx> const& wd, TK const& k, T* a, T const* b, T const* c) -> bool
        {
            namespace nv = b200::native;
            if constexpr(!nv::isStreamElem<T> || TDim::value != 1u)
                return false;
            else
            {
                bool const ok = nv::verifyStream<TK, TAcc, T>(
                    q,
                    [&](T* x, T* y, T* z, std::size_t)
                    { b200::launchGeneric<TAcc>(q, nv::verifyWorkDiv<TDim, TIdx>(), k, x, static_cast<T const*>(y), static_cast<T const*>(z)); },
                    [&](T* x, T* y, T* z, std::size_t n) { b200::check(nv::streamNstream(q.getNativeHandle(), x, y, z, n)); });
                if(!ok)
                    return false;
                b200::check(nv::streamNstream(q.getNativeHandle(), a, b, c, nv::gridThreads(wd)));
                return true;
            }
        }
    };

    //! DotKernel(a, b, sum, arraySize) with WorkDiv {G, B, 1}: sum[0..G) are per-block partials that the driver folds
    //! with std::reduce on the host (babelStreamMainTest.cpp:378-405). The native single-pass kernel fills the same G
    //! slots (b200_dot_partials_*); their sum is the dot product, so the driver's host fold and check are unchanged.
    template<typename TAcc>
    struct NativeKernel<::DotKernel, TAcc>
    {
        static constexpr bool available = true;
        template<typename TQueue, typename TDim, typename TIdx, typename TK, typename TA, typename TB, typename T, typename TN>
        static auto launch(TQueue& q, WorkDivMembers<TDim, TIdx> const& wd, TK const& k, TA* a, TB* b, T* sum, TN arraySize)
            -> std::enable_if_t<
                std::is_same_v<std::remove_const_t<TA>, T> && std::is_same_v<std::remove_const_t<TB>, T> && std::is_integral_v<TN>,
                bool>
        {
            namespace nv = b200::native;
            if constexpr(!nv::isStreamElem<T> || TDim::value != 1u)
                return false;
            else
            {
                auto const partials = static_cast<uint64_t>(wd.m_gridBlockExtent[0]);
                if(partials < 2u || partials > 0xffffffffull)
                    return false;
                // claimed by NAME: cross-check once against the user's own functor (same block size, 4 blocks)
                bool const ok = nv::verifyDot<TK, TAcc, T>(
                    q,
                    [&](T const* x, T const* y, T* out, std::size_t n, std::uint32_t blocks)
                    {
                        using V = Vec<TDim, TIdx>;
                        WorkDivMembers<TDim, TIdx> const small{V::all(static_cast<TIdx>(blocks)), wd.m_blockThreadExtent, V::all(1)};
                        b200::launchGeneric<TAcc>(q, small, k, static_cast<TA*>(const_cast<T*>(x)), static_cast<TB*>(const_cast<T*>(y)), out, static_cast<TN>(n));
                    },
                    [&](T const* x, T const* y, T* out, std::size_t n, std::uint32_t blocks)
                    { b200::check(nv::dotPartials(q.getNativeHandle(), x, y, n, out, blocks, q.m_impl->reduceScratch())); });
                if(!ok)
                    return false;
                b200::check(nv::dotPartials(
                    q.getNativeHandle(),
                    a,
                    b,
                    static_cast<uint64_t>(arraySize),
                    sum,
                    static_cast<uint32_t>(partials),
                    q.m_impl->reduceScratch()));
                return true;
            }
        }
    };

    // ---------------------------------------------------------------------------------------------------------
    // example/reduce. The driver launches ReduceKernel twice (reduce.cpp:79-98): main grid  source -> destination[0..G),
    // then one block  destination -> destination[0]. The reduction functor is an opaque type, so only
    // alpaka::b200::Sum<T> (the functor this library provides for "+") is claimed: the main launch becomes ONE
    // single-pass native reduction that leaves the total in destination[0] and zeros in destination[1..G); the
    // second launch (source == destination) runs the same native kernel in place over those G values.
    template<std::uint32_t TBlockSize, typename T, typename TAcc>
    struct NativeKernel<::ReduceKernel<TBlockSize, T, b200::Sum<T>>, TAcc>
    {
        static constexpr bool available = true;
        template<typename TQueue, typename TDim, typename TIdx, typename TK, typename TSrc, typename TN>
        static auto launch(
            TQueue& q,
            WorkDivMembers<TDim, TIdx> const& wd,
            TK const& k,
            TSrc* source,
            T* destination,
            TN const& n,
            b200::Sum<T> const& fn) -> std::enable_if_t<std::is_integral_v<TN> && std::is_same_v<std::remove_const_t<TSrc>, T>, bool>
        {
            namespace nv = b200::native;
            if constexpr(!nv::isReduceElem<T> || TDim::value != 1u)
                return false;
            else
            {
                b200_stream_t const s = q.getNativeHandle();
                // claimed by NAME: cross-check once against the user's own functor (the driver's two launches)
                bool const ok = nv::verifyReduce<TK, TAcc, T>(
                    q,
                    [&](T const* in, T* out, std::size_t m, std::uint32_t blocks)
                    {
                        using V = Vec<TDim, TIdx>;
                        WorkDivMembers<TDim, TIdx> const wd1{V::all(static_cast<TIdx>(blocks)), V::all(static_cast<TIdx>(TBlockSize)), V::all(1)};
                        WorkDivMembers<TDim, TIdx> const wd2{V::all(1), V::all(static_cast<TIdx>(TBlockSize)), V::all(1)};
                        // exactly the argument types of THIS launch: the trampoline instantiation that is its generic
                        // fall-back anyway (nvcc does not emit a kernel first instantiated inside a lambda of a template)
                        b200::launchGeneric<TAcc>(q, wd1, k, static_cast<TSrc*>(const_cast<T*>(in)), out, static_cast<TN>(m), fn);
                        b200::launchGeneric<TAcc>(q, wd2, k, static_cast<TSrc*>(out), out, static_cast<TN>(blocks), fn);
                    },
                    [&](T const* in, T* out, std::size_t m) { b200::check(nv::reduceSum(s, in, m, out, q.m_impl->reduceScratch())); });
                if(!ok)
                    return false;
                if(source == destination)
                {
                    // In-place reduction of destination[0..n) into destination[0]: the second launch of the driver's pair
                    // (after the native first launch it sums the total and G-1 zeros, x + 0 == x), or a stand-alone
                    // in-place call -- never skipped, the kernel cannot know which. Safe in place: out[0] is written by the
                    // last block after every block has finished reading (b200_reduce.cu, single-pass ticket).
                    b200::check(nv::reduceSum(s, source, static_cast<uint64_t>(n), destination, q.m_impl->reduceScratch()));
                    return true;
                }
                auto const blocks = static_cast<std::size_t>(wd.m_gridBlockExtent[0]);
                int const dev = getDev(q).getNativeHandle();
                if(blocks > 1u)
                    b200::check(b200_memset_async(dev, destination + 1, 0, (blocks - 1u) * sizeof(T), s));
                b200::check(nv::reduceSum(s, source, static_cast<uint64_t>(n), destination, q.m_impl->reduceScratch()));
                return true;
            }
        }
    };

    // ---------------------------------------------------------------------------------------------------------
    // heatEquation2D. Launch shape of the driver: grid = (ny/chunk, nx/chunk) blocks, field (ny+2) x (nx+2) with byte
    // pitches, Stencil then Boundary per step (heatEquation2D.cpp:141-168).
    template<std::size_t N, typename TAcc>
    struct NativeKernel<::StencilKernel<N>, TAcc>
    {
        static constexpr bool available = true;
        template<typename TQueue, typename TDim, typename TIdx, typename TK, typename TV>
        static auto launch(
            TQueue& q,
            WorkDivMembers<TDim, TIdx> const& wd,
            TK const&,
            double const* uCurr,
            double* uNext,
            Vec<TDim, TV> const& chunk,
            Vec<TDim, TV> const& pitchCurr,
            Vec<TDim, TV> const& pitchNext,
            double dx,
            double dy,
            double dt) -> bool
        {
            namespace nv = b200::native;
            if constexpr(TDim::value != 2u)
                return false;
            else
            {
                if(pitchCurr != pitchNext || pitchCurr[1] != sizeof(double) || pitchCurr[0] % 16u != 0u
                   || reinterpret_cast<std::uintptr_t>(uCurr) % 16u != 0u || reinterpret_cast<std::uintptr_t>(uNext) % 16u != 0u
                   || uCurr == uNext)
                    return false;
                auto const ny = static_cast<uint64_t>(wd.m_gridBlockExtent[0]) * static_cast<uint64_t>(chunk[0]);
                auto const nx = static_cast<uint64_t>(wd.m_gridBlockExtent[1]) * static_cast<uint64_t>(chunk[1]);
                if(ny == 0u || nx == 0u || ny > 0x7ffffff0ull || nx > 0x7ffffff0ull)
                    return false;
                // the functor's shared tile must be the chunk plus its 1-cell halo, or it is not the reference kernel
                if(static_cast<std::size_t>((chunk[0] + 2) * (chunk[1] + 2)) != N)
                    return false;
                int const dev = getDev(q).getNativeHandle();
                auto const plan = nv::heatPlanFor(
                    dev,
                    const_cast<double*>(uCurr),
                    uNext,
                    static_cast<std::size_t>(pitchCurr[0]),
                    static_cast<uint32_t>(ny),
                    static_cast<uint32_t>(nx),
                    dx,
                    dy);
                int const srcIndex = (plan.u0 == uCurr) ? 0 : 1;
                double const rX = dt / (dx * dx); // StencilKernel.hpp:70-71
                double const rY = dt / (dy * dy);
                // core cells only: rows 1..ny, columns 1..nx; the ring belongs to BoundaryKernel
                b200::check(b200_heat2d_step_window_f64(
                    plan.plan,
                    q.getNativeHandle(),
                    srcIndex,
                    rX,
                    rY,
                    0.0,
                    1u,
                    static_cast<uint32_t>(ny) + 1u,
                    1u,
                    static_cast<uint32_t>(nx) + 1u));
                return true;
            }
        }
    };

    template<typename TAcc>
    struct NativeKernel<::BoundaryKernel, TAcc>
    {
        static constexpr bool available = true;
        template<typename TQueue, typename TDim, typename TIdx, typename TK, typename TV>
        static auto launch(
            TQueue& q,
            WorkDivMembers<TDim, TIdx> const& wd,
            TK const&,
            double* uBuf,
            Vec<TDim, TV> const& chunk,
            Vec<TDim, TV> const& pitch,
            uint32_t step,
            double dx,
            double dy,
            double dt) -> bool
        {
            namespace nv = b200::native;
            if constexpr(TDim::value != 2u)
                return false;
            else
            {
                auto const ny = static_cast<uint64_t>(wd.m_gridBlockExtent[0]) * static_cast<uint64_t>(chunk[0]);
                auto const nx = static_cast<uint64_t>(wd.m_gridBlockExtent[1]) * static_cast<uint64_t>(chunk[1]);
                nv::HeatPlan plan;
                int index = 0;
                if(!nv::heatPlanContaining(
                       getDev(q).getNativeHandle(),
                       uBuf,
                       static_cast<std::size_t>(pitch[0]),
                       static_cast<uint32_t>(ny),
                       static_cast<uint32_t>(nx),
                       dx,
                       dy,
                       plan,
                       index))
                    return false; // no stencil launch has described this field yet
                constexpr double pi = math::constants::pi;
                double const tf = std::exp(-pi * pi * (step * dt)); // exactSolution(..., step * dt), host libm
                b200::check(b200_heat2d_boundary_f64(plan.plan, q.getNativeHandle(), index, tf));
                return true;
            }
        }
    };
} // namespace alpaka::trait

#endif // ALPAKA_B200_RECOGNIZE_REFERENCE_KERNELS && __CUDACC__
