// include/alpaka/b200/Kernel.hpp -- kernel launch path of the B200 back-end.
//
// API parity with the reference's kernel/Traits.hpp:29-382 (createTaskKernel, exec, getFunctionAttributes,
// trait::BlockSharedMemDynSizeBytes, trait::WarpSize, the trivially-copyable checks),
// kernel/KernelFunctionAttributes.hpp:14-24, kernel/TaskKernelGpuUniformCudaHipRt.hpp:61-366 (trampoline, task,
// Enqueue, FunctionAttributes) and workdiv/WorkDivHelpers.hpp:315-396 (KernelCfg, getValidWorkDiv).
//
// Two ways a task reaches the GPU:
//   1. generic: the trampoline alpaka::b200k::run<...> (the framework's only generic __global__) builds the
//      accelerator object and calls the user functor -- any alpaka kernel runs unchanged. The launch goes through
//      the C ABI (b200_launch: cudaLaunchKernel by host function address on the queue's stream).
//   2. native: trait::NativeKernel<TKernelFnObj, TAcc> can claim a functor type and forward the launch to one of the
//      hand-written sm_100a kernels of libalpaka_b200.so (b200_stream_*, b200_dot_*, b200_heat2d_*). The library
//      ships specialisations for the reference drivers' functors in alpaka/b200/Native.hpp (opt-in); a specialisation
//      returns false to decline a particular launch, which then takes path 1. ALPAKA_B200_NATIVE=0 in the
//      environment disables path 2 at run time (A/B measurement).
#pragma once

#include "Acc.hpp"

#include <cstdlib>
#include <map>
#include <mutex>
#include <tuple>
#include <type_traits>
#include <utility>

namespace alpaka
{
    //! Kernel function attributes struct. Attributes are filled by calling the API of the accelerator using the kernel
    //! function as an argument.
    struct KernelFunctionAttributes
    {
        std::size_t constSizeBytes{0};
        std::size_t localSizeBytes{0};
        std::size_t sharedSizeBytes{0};
        int maxDynamicSharedSizeBytes{0};
        int numRegs{0};
        int asmVersion{0}; //!< PTX version the kernel was compiled for
        int maxThreadsPerBlock{0};
    };

    namespace trait
    {
        //! The trait for getting the size of the block shared dynamic memory of a kernel.
        //! The default implementation returns 0; specialise per (kernel functor, accelerator).
        template<typename TKernelFnObj, typename TAcc, typename TSfinae = void>
        struct BlockSharedMemDynSizeBytes
        {
            template<typename TDim, typename... TArgs>
            ALPAKA_FN_HOST_ACC static auto getBlockSharedMemDynSizeBytes(
                TKernelFnObj const& /*kernelFnObj*/,
                Vec<TDim, Idx<TAcc>> const& /*blockThreadExtent*/,
                Vec<TDim, Idx<TAcc>> const& /*threadElemExtent*/,
                TArgs const&... /*args*/) -> std::size_t
            {
                return 0u;
            }
        };

        //! The trait for getting the warp size required by a kernel (0 = any). The B200 only has 32-wide warps.
        template<typename TKernelFnObj, typename TAcc, typename TSfinae = void>
        struct WarpSize : std::integral_constant<std::uint32_t, 0>
        {
        };

        //! Hook: lets a library claim a kernel functor type for a hand-written implementation.
        //! A specialisation defines `available = true` and
        //!   template<typename TQueue, typename TWorkDiv, typename... TArgs>
        //!   static bool launch(TQueue&, TWorkDiv const&, TKernelFnObj const&, TArgs const&...);
        //! returning false if it declines this particular launch.
        template<typename TKernelFnObj, typename TAcc, typename TSfinae = void>
        struct NativeKernel
        {
            static constexpr bool available = false;
        };

        template<typename TAcc, typename TDev, typename TKernelFnObj, typename... TArgs>
        struct FunctionAttributes;

        template<typename TAcc, typename TWorkDiv, typename TKernelFnObj, typename... TArgs>
        struct CreateTaskKernel;
    } // namespace trait

    template<typename TKernelFnObj, typename TAcc>
    inline constexpr std::uint32_t warpSize = trait::WarpSize<TKernelFnObj, TAcc>::value;

    //! \return The size of the shared memory allocated for a block in bytes.
    template<typename TAcc, typename TKernelFnObj, typename TDim, typename... TArgs>
    ALPAKA_FN_HOST_ACC auto getBlockSharedMemDynSizeBytes(
        TKernelFnObj const& kernelFnObj,
        Vec<TDim, Idx<TAcc>> const& blockThreadExtent,
        Vec<TDim, Idx<TAcc>> const& threadElemExtent,
        TArgs const&... args) -> std::size_t
    {
        return trait::BlockSharedMemDynSizeBytes<TKernelFnObj, TAcc>::getBlockSharedMemDynSizeBytes(
            kernelFnObj,
            blockThreadExtent,
            threadElemExtent,
            args...);
    }

    namespace b200
    {
        //! The "+" reduction functor. A plain struct instead of a lambda so that libraries can recognise it
        //! (alpaka/b200/Native.hpp routes ReduceKernel<.., Sum<T>> to the native single-pass reduction).
        template<typename T>
        struct Sum
        {
            ALPAKA_FN_HOST_ACC constexpr auto operator()(T const& a, T const& b) const -> T
            {
                return a + b;
            }
        };

        //! ALPAKA_B200_NATIVE=0 routes every launch through the generic trampoline
        inline auto nativeKernelsEnabled() -> bool
        {
            static bool const enabled = []
            {
                char const* e = std::getenv("ALPAKA_B200_NATIVE");
                return !(e != nullptr && e[0] == '0');
            }();
            return enabled;
        }
    } // namespace b200

#if defined(__CUDACC__)
    // short namespace: keeps the mangled kernel name in profiler output readable
    // (same consideration as the reference, kernel/TaskKernelGpuUniformCudaHipRt.hpp:58-60)
    namespace b200k
    {
        //! The generic kernel trampoline: builds the accelerator and calls the user's functor.
        template<typename TKernelFnObj, typename TAcc, typename TDim, typename TIdx, typename... TArgs>
        __global__ void run(Vec<TDim, TIdx> const threadElemExtent, TKernelFnObj const kernelFnObj, TArgs... args)
        {
            TAcc const acc(threadElemExtent);
            kernelFnObj(acc, args...);
        }

        //! pointer parameters of the coarsened trampoline are restrict-qualified (the launch proved they cannot alias)
        template<typename T>
        struct NoAlias
        {
            using type = T;
        };
        template<typename T>
        struct NoAlias<T*>
        {
            using type = T* __restrict__;
        };

        //! The BLOCK-COARSENED trampoline. One-element-per-thread functors (the reference's contract,
        //! kernel/TaskKernelGpuUniformCudaHipRt.hpp:61-76) keep 8 bytes per load in flight per thread: 16 KB per SM, about
        //! 60 % of the HBM bandwidth of a B200 whatever the block size. Here a physical block runs V consecutive VIRTUAL
        //! blocks of the user's grid (fastest dimension) one after the other, unrolled; the accelerator object answers
        //! getIdx / getWorkDiv with the user's grid. Blocks are independent by the programming model, so this is the
        //! same program; because the launch site has PROVEN that the pointer arguments lie in pairwise distinct
        //! allocations (b200_mem_range) they are restrict-qualified, and the compiler hoists the loads of all V virtual
        //! blocks above the first store: V times the bytes in flight with the user's functor unchanged.
        //! Only launched for functors without shared memory (nothing to protect between virtual blocks).
        //! TArgs arrive ALREADY restrict-qualified (the launch site instantiates with NoAlias<T>::type...): nvcc honours
        //! restrict on the parameters of a __global__ function only -- measured, tools/README.md -- and a pack of dependent
        //! parameter types does not survive its host stub, so the qualifier travels inside the template arguments.
        template<int V, typename TKernelFnObj, typename TAcc, typename TDim, typename TIdx, typename... TArgs>
        __global__ void runCoarse(
            Vec<TDim, TIdx> const threadElemExtent,
            Vec<TDim, TIdx> const gridBlockExtent,
            TKernelFnObj const kernelFnObj,
            TArgs... args)
        {
            constexpr std::size_t x = TDim::value - 1u; // alpaka's last component is CUDA's x
            Vec<TDim, TIdx> block = alpaka::b200::fromBuiltin<TDim, TIdx>(blockIdx);
            TIdx const first = block[x] * static_cast<TIdx>(V);
            if(first + static_cast<TIdx>(V) <= gridBlockExtent[x])
            {
#    pragma unroll
                for(int v = 0; v < V; ++v)
                {
                    block[x] = first + static_cast<TIdx>(v);
                    TAcc const acc(threadElemExtent, gridBlockExtent, block);
                    kernelFnObj(acc, args...);
                }
            }
            else
            {
                for(TIdx b = first; b < gridBlockExtent[x]; ++b)
                {
                    block[x] = b;
                    TAcc const acc(threadElemExtent, gridBlockExtent, block);
                    kernelFnObj(acc, args...);
                }
            }
        }
    } // namespace b200k
#endif

#if defined(__CUDACC__)
    namespace b200::detail
    {
        //! Arguments through which no hidden pointer can reach the kernel: raw pointers (checked at the launch),
        //! arithmetic / enum values and alpaka vectors of them. Anything else (a view object, a struct with pointer
        //! members) keeps the plain trampoline. The functor itself must be stateless for the same reason.
        template<typename T>
        inline constexpr bool plainValueArg = std::is_arithmetic_v<T> || std::is_enum_v<T>;
        template<typename TDim, typename TVal>
        inline constexpr bool plainValueArg<Vec<TDim, TVal>> = std::is_arithmetic_v<TVal>;
        template<typename T>
        inline constexpr bool coarsenableArg
            = plainValueArg<std::remove_cv_t<T>>
              || (std::is_pointer_v<T> && (std::is_arithmetic_v<std::remove_cv_t<std::remove_pointer_t<T>>>) );
        template<typename TKernelFnObj, typename... TArgs>
        inline constexpr bool coarsenable = std::is_empty_v<TKernelFnObj> && (coarsenableArg<TArgs> && ...)
                                            && (std::is_pointer_v<TArgs> || ...);

        //! tunable `generic.coarsen` (B200_TUNE / b200_tune_set): 4 (default) or 0 = always the plain trampoline
        inline auto coarsenFactor() -> int
        {
            int64_t v = 4;
            (void) b200_tune_get("generic.coarsen", &v);
            return static_cast<int>(v);
        }

        //! coarsen only grids that still fill the device afterwards: >= 32 virtual blocks per SM along x
        template<typename TDev>
        auto coarsenMinBlocks(TDev const& dev) -> std::uint64_t
        {
            static std::uint64_t const n = [&]
            {
                b200_device_props p{};
                return b200_device_props_get(dev.getNativeHandle(), &p) == 0 ? static_cast<std::uint64_t>(p.multi_processor_count) * 32u
                                                                            : std::uint64_t{148u * 32u};
            }();
            return n;
        }

        struct PtrArg
        {
            void const* p;
            bool writable;
        };

        //! true iff every pointer argument lies in an allocation of this library and no WRITABLE pointer shares its
        //! allocation with any other pointer argument (two read-only pointers may: nothing is modified through them)
        template<typename... TArgs>
        auto argsCannotAlias(TArgs const&... args) -> bool
        {
            PtrArg ptrs[sizeof...(TArgs) + 1u];
            std::size_t n = 0;
            auto const collect = [&](auto const& a)
            {
                using A = std::decay_t<decltype(a)>;
                if constexpr(std::is_pointer_v<A>)
                    ptrs[n++] = PtrArg{static_cast<void const*>(a), !std::is_const_v<std::remove_pointer_t<A>>};
            };
            (collect(args), ...);
            void* base[sizeof...(TArgs) + 1u];
            for(std::size_t i = 0; i < n; ++i)
            {
                std::size_t bytes = 0;
                if(b200_mem_range(ptrs[i].p, &base[i], &bytes) != 0 || base[i] == nullptr)
                    return false; // foreign memory: nothing is known about it
            }
            for(std::size_t i = 0; i < n; ++i)
                for(std::size_t j = i + 1; j < n; ++j)
                    if(base[i] == base[j] && (ptrs[i].writable || ptrs[j].writable))
                        return false;
            return true;
        }

        //! the coarsened instantiation holds no static shared memory (cached per kernel and device)
        inline auto usesNoSharedMemory(int dev, void const* kernel) -> bool
        {
            static std::mutex mutex;
            static std::map<std::pair<int, void const*>, bool> cache;
            std::lock_guard<std::mutex> l(mutex);
            auto const key = std::make_pair(dev, kernel);
            auto const it = cache.find(key);
            if(it != cache.end())
                return it->second;
            b200_func_attributes a{};
            bool const ok = b200_func_attributes_get(dev, kernel, &a) == 0 && a.shared_size_bytes == 0u;
            cache.emplace(key, ok);
            return ok;
        }
    } // namespace b200::detail

    namespace b200
    {
        //! Launches `kernelFnObj(acc, args...)` through the generic trampoline on the queue's stream (no
        //! blocking-queue synchronisation; the caller does that). Also what a NativeKernel specialisation calls to
        //! cross-check itself against the user's functor.
        template<typename TAcc, typename TQueue, typename TDim, typename TIdx, typename TKernelFnObj, typename... TArgs>
        void launchGeneric(
            TQueue& queue,
            WorkDivMembers<TDim, TIdx> const& workDiv,
            TKernelFnObj const& kernelFnObj,
            TArgs const&... args)
        {
            auto const gridBlockExtent = workDiv.m_gridBlockExtent;
            auto const blockThreadExtent = workDiv.m_blockThreadExtent;
            auto threadElemExtent = workDiv.m_threadElemExtent;

            uint32_t grid[3] = {1u, 1u, 1u};
            uint32_t block[3] = {1u, 1u, 1u};
            for(std::size_t d = 0; d < TDim::value; ++d)
            {
                // alpaka's fastest (last) dimension is CUDA x
                grid[d] = static_cast<uint32_t>(gridBlockExtent[TDim::value - 1u - d]);
                block[d] = static_cast<uint32_t>(blockThreadExtent[TDim::value - 1u - d]);
            }

#    if ALPAKA_DEBUG >= ALPAKA_DEBUG_MINIMAL
            if(!isValidWorkDiv(workDiv, getAccDevProps<TAcc>(getDev(queue))))
                throw std::runtime_error(
                    "The given work division is not valid or not supported by the device of type " + getAccName<TAcc>()
                    + "!");
#    endif
            std::size_t const dynSmemBytes
                = getBlockSharedMemDynSizeBytes<TAcc>(kernelFnObj, blockThreadExtent, threadElemExtent, args...);

            // block-coarsened, restrict-qualified trampoline when it is provably the same program (see b200k::runCoarse)
            if constexpr(TDim::value >= 1u && detail::coarsenable<TKernelFnObj, TArgs...>)
            {
                constexpr int V = 4;
                if(dynSmemBytes == 0u && detail::coarsenFactor() == V
                   && static_cast<std::uint64_t>(gridBlockExtent[TDim::value - 1u]) >= detail::coarsenMinBlocks(getDev(queue))
                   && detail::argsCannotAlias(args...))
                {
                    void const* const coarse = (void const*) b200k::runCoarse<V, TKernelFnObj, TAcc, TDim, TIdx, typename b200k::NoAlias<TArgs>::type...>;
                    if(detail::usesNoSharedMemory(getDev(queue).getNativeHandle(), coarse))
                    {
                        grid[0] = (grid[0] + static_cast<uint32_t>(V) - 1u) / static_cast<uint32_t>(V);
                        void* argvCoarse[3u + sizeof...(TArgs)]
                            = {const_cast<void*>(static_cast<void const*>(&threadElemExtent)),
                               const_cast<void*>(static_cast<void const*>(&gridBlockExtent)),
                               const_cast<void*>(static_cast<void const*>(&kernelFnObj)),
                               const_cast<void*>(static_cast<void const*>(&args))...};
                        check(b200_launch(
                            getDev(queue).getNativeHandle(),
                            coarse,
                            grid,
                            block,
                            0u,
                            queue.getNativeHandle(),
                            argvCoarse));
#    if ALPAKA_DEBUG >= ALPAKA_DEBUG_MINIMAL
                        check(b200_stream_sync(queue.getNativeHandle()));
#    endif
                        return;
                    }
                }
            }

            auto const kernel = b200k::run<TKernelFnObj, TAcc, TDim, TIdx, TArgs...>;

            // cudaLaunchKernel wants one pointer per kernel parameter, in order
            void* argv[2u + sizeof...(TArgs)]
                = {const_cast<void*>(static_cast<void const*>(&threadElemExtent)),
                   const_cast<void*>(static_cast<void const*>(&kernelFnObj)),
                   const_cast<void*>(static_cast<void const*>(&args))...};

            check(b200_launch(
                getDev(queue).getNativeHandle(),
                reinterpret_cast<void const*>(kernel),
                grid,
                block,
                dynSmemBytes,
                queue.getNativeHandle(),
                argv));
#    if ALPAKA_DEBUG >= ALPAKA_DEBUG_MINIMAL
            // debug builds surface asynchronous launch failures at the launch site, like the reference
            check(b200_stream_sync(queue.getNativeHandle()));
#    endif
        }
    } // namespace b200
#endif

    namespace detail
    {
        template<typename T>
        struct IsKernelArgumentTriviallyCopyable
        {
#if defined(__NVCC__) && defined(__CUDACC_EXTENDED_LAMBDA__)
            // extended lambdas are seen by the host pass through a placeholder closure type
            static constexpr bool value = std::is_trivially_copyable_v<T> || __nv_is_extended_device_lambda_closure_type(T)
                                          || __nv_is_extended_host_device_lambda_closure_type(T);
#else
            static constexpr bool value = std::is_trivially_copyable_v<T>;
#endif
        };
    } // namespace detail
    //! customisation point: declare a kernel argument type copyable by memcpy although the language cannot prove it
    template<typename T, typename = void>
    struct IsKernelArgumentTriviallyCopyable : detail::IsKernelArgumentTriviallyCopyable<T>
    {
    };
    template<typename T>
    inline constexpr bool isKernelArgumentTriviallyCopyable = IsKernelArgumentTriviallyCopyable<T>::value;
    template<typename T>
    inline constexpr bool isKernelTriviallyCopyable = IsKernelArgumentTriviallyCopyable<T>::value;

    //! The kernel execution task of the B200 back-end: work division + functor + arguments, all held by value.
    template<typename TAcc, typename TDim, typename TIdx, typename TKernelFnObj, typename... TArgs>
    class TaskKernelB200 final : public WorkDivMembers<TDim, TIdx>
    {
    public:
        template<typename TWorkDiv>
        TaskKernelB200(TWorkDiv&& workDiv, TKernelFnObj const& kernelFnObj, TArgs&&... args)
            : WorkDivMembers<TDim, TIdx>(std::forward<TWorkDiv>(workDiv))
            , m_kernelFnObj(kernelFnObj)
            , m_args(std::forward<TArgs>(args)...)
        {
            static_assert(
                Dim<std::decay_t<TWorkDiv>>::value == TDim::value,
                "The work division and the execution task have to be of the same dimensionality!");
        }

        TKernelFnObj m_kernelFnObj;
        std::tuple<std::decay_t<TArgs>...> m_args;
    };

    namespace trait
    {
        template<typename TAcc, typename TDim, typename TIdx, typename TKernelFnObj, typename... TArgs>
        struct AccType<TaskKernelB200<TAcc, TDim, TIdx, TKernelFnObj, TArgs...>>
        {
            using type = TAcc;
        };
        template<typename TAcc, typename TDim, typename TIdx, typename TKernelFnObj, typename... TArgs>
        struct DevType<TaskKernelB200<TAcc, TDim, TIdx, TKernelFnObj, TArgs...>>
        {
            using type = DevB200;
        };
        template<typename TAcc, typename TDim, typename TIdx, typename TKernelFnObj, typename... TArgs>
        struct DimType<TaskKernelB200<TAcc, TDim, TIdx, TKernelFnObj, TArgs...>>
        {
            using type = TDim;
        };
        template<typename TAcc, typename TDim, typename TIdx, typename TKernelFnObj, typename... TArgs>
        struct IdxType<TaskKernelB200<TAcc, TDim, TIdx, TKernelFnObj, TArgs...>>
        {
            using type = TIdx;
        };
        template<typename TAcc, typename TDim, typename TIdx, typename TKernelFnObj, typename... TArgs>
        struct GetWorkDiv<TaskKernelB200<TAcc, TDim, TIdx, TKernelFnObj, TArgs...>, origin::Grid, unit::Blocks>
        {
            static auto getWorkDiv(WorkDivMembers<TDim, TIdx> const& w) -> Vec<TDim, TIdx>
            {
                return w.m_gridBlockExtent;
            }
        };
        template<typename TAcc, typename TDim, typename TIdx, typename TKernelFnObj, typename... TArgs>
        struct GetWorkDiv<TaskKernelB200<TAcc, TDim, TIdx, TKernelFnObj, TArgs...>, origin::Block, unit::Threads>
        {
            static auto getWorkDiv(WorkDivMembers<TDim, TIdx> const& w) -> Vec<TDim, TIdx>
            {
                return w.m_blockThreadExtent;
            }
        };
        template<typename TAcc, typename TDim, typename TIdx, typename TKernelFnObj, typename... TArgs>
        struct GetWorkDiv<TaskKernelB200<TAcc, TDim, TIdx, TKernelFnObj, TArgs...>, origin::Thread, unit::Elems>
        {
            static auto getWorkDiv(WorkDivMembers<TDim, TIdx> const& w) -> Vec<TDim, TIdx>
            {
                return w.m_threadElemExtent;
            }
        };

        template<typename TApi, typename TDim, typename TIdx, typename TWorkDiv, typename TKernelFnObj, typename... TArgs>
        struct CreateTaskKernel<AccGpuUniformCudaHipRt<TApi, TDim, TIdx>, TWorkDiv, TKernelFnObj, TArgs...>
        {
            static auto createTaskKernel(TWorkDiv const& workDiv, TKernelFnObj const& kernelFnObj, TArgs&&... args)
            {
                return TaskKernelB200<AccGpuUniformCudaHipRt<TApi, TDim, TIdx>, TDim, TIdx, TKernelFnObj, TArgs...>(
                    workDiv,
                    kernelFnObj,
                    std::forward<TArgs>(args)...);
            }
        };

#if defined(__CUDACC__)
        template<typename TApi, typename TDim, typename TIdx, typename TKernelFnObj, typename... TArgs>
        struct FunctionAttributes<AccGpuUniformCudaHipRt<TApi, TDim, TIdx>, DevB200, TKernelFnObj, TArgs...>
        {
            static auto getFunctionAttributes(DevB200 const& dev, TKernelFnObj const&, TArgs&&...) -> KernelFunctionAttributes
            {
                using TAcc = AccGpuUniformCudaHipRt<TApi, TDim, TIdx>;
                auto const kernel = b200k::run<TKernelFnObj, TAcc, TDim, TIdx, std::decay_t<TArgs>...>;
                b200_func_attributes a{};
                b200::check(b200_func_attributes_get(dev.getNativeHandle(), reinterpret_cast<void const*>(kernel), &a));
                KernelFunctionAttributes r;
                r.constSizeBytes = a.const_size_bytes;
                r.localSizeBytes = a.local_size_bytes;
                r.sharedSizeBytes = a.shared_size_bytes;
                r.maxDynamicSharedSizeBytes = a.max_dynamic_shared_size_bytes;
                r.numRegs = a.num_regs;
                r.asmVersion = a.ptx_version;
                r.maxThreadsPerBlock = a.max_threads_per_block;
                return r;
            }
        };

        template<typename TProperty, typename TAcc, typename TDim, typename TIdx, typename TKernelFnObj, typename... TArgs>
        struct Enqueue<QueueB200<TProperty>, TaskKernelB200<TAcc, TDim, TIdx, TKernelFnObj, TArgs...>>
        {
            using Task = TaskKernelB200<TAcc, TDim, TIdx, TKernelFnObj, TArgs...>;

            static void enqueue(QueueB200<TProperty>& queue, Task const& task)
            {
                // 1. a hand-written kernel may claim this functor
                if constexpr(NativeKernel<TKernelFnObj, TAcc>::available)
                {
                    if(b200::nativeKernelsEnabled())
                    {
                        bool const handled = std::apply(
                            [&](auto const&... args) -> bool
                            {
                                auto const& workDiv = static_cast<WorkDivMembers<TDim, TIdx> const&>(task);
                                // only offered if the specialisation's signature accepts these argument types
                                if constexpr(requires {
                                                 NativeKernel<TKernelFnObj, TAcc>::launch(queue, workDiv, task.m_kernelFnObj, args...);
                                             })
                                    return NativeKernel<TKernelFnObj, TAcc>::launch(queue, workDiv, task.m_kernelFnObj, args...);
                                else
                                    return false;
                            },
                            task.m_args);
                        if(handled)
                        {
                            queue.afterEnqueue();
                            return;
                        }
                    }
                }

                // 2. generic trampoline
                std::apply(
                    [&](auto const&... args)
                    {
                        b200::launchGeneric<TAcc>(
                            queue,
                            static_cast<WorkDivMembers<TDim, TIdx> const&>(task),
                            task.m_kernelFnObj,
                            args...);
                    },
                    task.m_args);
                queue.afterEnqueue();
            }
        };
#endif // __CUDACC__
    } // namespace trait

    //! Creates a kernel execution task.
    //! \tparam TAcc The accelerator type.
    //! \param workDiv The index domain work division.
    //! \param kernelFnObj The kernel function object which should be executed.
    //! \param args The kernel invocation arguments.
    template<typename TAcc, typename TWorkDiv, typename TKernelFnObj, typename... TArgs>
    [[nodiscard]] auto createTaskKernel(TWorkDiv const& workDiv, TKernelFnObj const& kernelFnObj, TArgs&&... args)
    {
        static_assert(
            isKernelTriviallyCopyable<TKernelFnObj>,
            "Kernels must be trivially copyable or specialize trait::IsKernelTriviallyCopyable<>!");
        static_assert(
            (isKernelArgumentTriviallyCopyable<std::decay_t<TArgs>> && ...),
            "The kernel arguments must be trivially copyable or specialize trait::IsKernelArgumentTriviallyCopyable<>!");
        static_assert(
            Dim<std::decay_t<TWorkDiv>>::value == Dim<TAcc>::value,
            "The dimensions of TAcc and TWorkDiv have to be identical!");
        static_assert(
            std::is_same_v<Idx<std::decay_t<TWorkDiv>>, Idx<TAcc>>,
            "The idx type of TAcc and the idx type of TWorkDiv have to be identical!");
        return trait::CreateTaskKernel<TAcc, TWorkDiv, TKernelFnObj, TArgs...>::createTaskKernel(
            workDiv,
            kernelFnObj,
            std::forward<TArgs>(args)...);
    }

    //! Executes the given kernel in the given queue.
    template<typename TAcc, typename TQueue, typename TWorkDiv, typename TKernelFnObj, typename... TArgs>
    auto exec(TQueue& queue, TWorkDiv const& workDiv, TKernelFnObj const& kernelFnObj, TArgs&&... args) -> void
    {
        enqueue(queue, createTaskKernel<TAcc>(workDiv, kernelFnObj, std::forward<TArgs>(args)...));
    }

    //! \return The kernel function attributes (registers, shared memory, maximum threads per block) of the generic
    //! launch of `kernelFnObj(acc, args...)` on the given device.
    template<typename TAcc, typename TDev, typename TKernelFnObj, typename... TArgs>
    [[nodiscard]] auto getFunctionAttributes(TDev const& dev, TKernelFnObj const& kernelFnObj, TArgs&&... args)
        -> KernelFunctionAttributes
    {
        return trait::FunctionAttributes<TAcc, TDev, TKernelFnObj, TArgs...>::getFunctionAttributes(
            dev,
            kernelFnObj,
            std::forward<TArgs>(args)...);
    }

    //! Kernel start configuration to determine a valid work division.
    template<
        typename TAcc,
        typename TGridElemExtent = Vec<Dim<TAcc>, Idx<TAcc>>,
        typename TThreadElemExtent = Vec<Dim<TAcc>, Idx<TAcc>>>
    struct KernelCfg
    {
        //! The full extent of elements in the grid.
        TGridElemExtent const gridElemExtent = Vec<Dim<TAcc>, Idx<TAcc>>::ones();
        //! The number of elements computed per thread.
        TThreadElemExtent const threadElemExtent = Vec<Dim<TAcc>, Idx<TAcc>>::ones();
        //! If true, the grid thread extent is a multiple of the block thread extent in every dimension.
        bool blockThreadMustDivideGridThreadExtent = true;
        //! The grid block extent subdivision restrictions.
        GridBlockExtentSubDivRestrictions gridBlockExtentSubDivRestrictions = GridBlockExtentSubDivRestrictions::Unrestricted;

        static_assert(
            Dim<TGridElemExtent>::value == Dim<TAcc>::value && Dim<TThreadElemExtent>::value == Dim<TAcc>::value,
            "The dimension of Acc and of the extents have to be identical!");
        static_assert(
            std::is_same_v<Idx<TGridElemExtent>, Idx<TAcc>> && std::is_same_v<Idx<TThreadElemExtent>, Idx<TAcc>>,
            "The idx type of Acc and of the extents have to be identical!");
    };

    //! \return A work division valid for the accelerator, the device and the kernel
    //! (block size limited by the kernel's registers/shared memory).
    template<typename TAcc, typename TDev, typename TGridElemExtent, typename TThreadElemExtent, typename TKernelFnObj, typename... TArgs>
    [[nodiscard]] auto getValidWorkDiv(
        KernelCfg<TAcc, TGridElemExtent, TThreadElemExtent> const& kernelCfg,
        TDev const& dev,
        TKernelFnObj const& kernelFnObj,
        TArgs&&... args) -> WorkDivMembers<Dim<TAcc>, Idx<TAcc>>
    {
        using I = Idx<TAcc>;
        if constexpr(Dim<TAcc>::value == 0u)
        {
            auto const zero = Vec<DimInt<0u>, I>{};
            return WorkDivMembers<DimInt<0u>, I>{zero, zero, zero};
        }
        else
        {
            auto const attrs = getFunctionAttributes<TAcc>(dev, kernelFnObj, std::forward<TArgs>(args)...);
            return subDivideGridElems(
                getExtents(kernelCfg.gridElemExtent),
                getExtents(kernelCfg.threadElemExtent),
                getAccDevProps<TAcc>(dev),
                static_cast<I>(attrs.maxThreadsPerBlock),
                kernelCfg.blockThreadMustDivideGridThreadExtent,
                kernelCfg.gridBlockExtentSubDivRestrictions);
        }
    }

    //! Checks if the work division is supported by the accelerator on the device for the given kernel.
    template<typename TAcc, typename TWorkDiv, typename TDev, typename TKernelFnObj, typename... TArgs>
    [[nodiscard]] auto isValidWorkDiv(TWorkDiv const& workDiv, TDev const& dev, TKernelFnObj const& kernelFnObj, TArgs&&... args)
        -> bool
    {
        auto const attrs = getFunctionAttributes<TAcc>(dev, kernelFnObj, std::forward<TArgs>(args)...);
        return isValidWorkDiv(workDiv, getAccDevProps<TAcc>(dev), static_cast<std::size_t>(attrs.maxThreadsPerBlock));
    }
    //! Checks if the work division is supported by the accelerator on the device.
    template<typename TAcc, typename TWorkDiv, typename TDev>
    [[nodiscard]] auto isValidWorkDiv(TWorkDiv const& workDiv, TDev const& dev) -> bool
    {
        return isValidWorkDiv(workDiv, getAccDevProps<TAcc>(dev));
    }
} // namespace alpaka
