// include/alpaka/b200/Acc.hpp -- the B200 accelerator: its type, host-side traits and everything a kernel can ask
// of `acc` on the device (indices, work division, shared memory, barriers, warp collectives, atomics, fences,
// math, bit intrinsics).
//
// API parity with the reference's acc/AccGpuUniformCudaHipRt.hpp:49-304 + acc/AccGpuCudaRt.hpp:15-31 (type, traits),
// acc/Traits.hpp:77-103 (getAccDevProps/getAccName), idx/Accessors.hpp:21-83 + idx/{bt,gb}/*UniformCudaHipBuiltIn.hpp,
// workdiv/WorkDivUniformCudaHipBuiltIn.hpp:20-111, block/shared/{st,dyn}/*, block/sync/Traits.hpp:35-106,
// warp/Traits.hpp:64-316 + warp/WarpUniformCudaHipBuiltIn.hpp:34-183, atomic/Traits.hpp:72-304 + atomic/Op.hpp,
// mem/fence/Traits.hpp, math/Traits.hpp (the subset the drivers and conformance tests use), intrinsic/Traits.hpp:38-79.
//
// The class is named AccGpuUniformCudaHipRt<TApi,TDim,TIdx> with TApi = ApiB200Rt because user code partially
// specialises on that template name (example/reduce/src/alpakaConfig.hpp:110-114). AccGpuB200<TDim,TIdx> is the
// name of the new accelerator; AccGpuCudaRt<TDim,TIdx> is an alias so reference drivers select it unchanged.
// Device code is only visible to nvcc (__CUDACC__); host-only translation units still see every type and trait.
#pragma once

#include "WorkDiv.hpp"

#include <cxxabi.h>
#include <string>
#include <typeinfo>

namespace alpaka
{
    //! Size of the dynamic shared-memory arena of the reference's CPU accelerators
    //! (block/shared/dyn/BlockSharedDynMemberAllocKiB.hpp). There is no such arena here; the name exists because code
    //! written against the reference mentions it inside `if constexpr` branches for CPU tags.
    inline constexpr std::uint32_t BlockSharedDynMemberAllocKiB = 47u;

    //! The vendor-API tag of this back-end (the reference's slot for ApiCudaRt / ApiHipRt).
    struct ApiB200Rt
    {
        static constexpr char name[] = "B200";
    };
    using ApiCudaRt = ApiB200Rt;

#if defined(__CUDACC__)
    namespace b200
    {
        //! dim3-like built-in -> alpaka vector (defined with the device-side traits below)
        template<typename TDim, typename TIdx, typename TBuiltin>
        __device__ __forceinline__ auto fromBuiltin(TBuiltin const& v) -> Vec<TDim, TIdx>;
    } // namespace b200
#endif

    //! The B200 accelerator. Constructed on the device by the kernel trampoline; never copied.
    template<typename TApi, typename TDim, typename TIdx>
    class AccGpuUniformCudaHipRt final
    {
        static_assert(sizeof(TIdx) >= sizeof(int), "Index type is not supported, consider using int or a larger type.");
        static_assert(TDim::value <= 3u, "The B200 accelerator supports 0 to 3 dimensions (CUDA grids are 3-D).");

    public:
        //! the launch as the hardware sees it: grid extent and block index are CUDA's built-ins
        ALPAKA_FN_HOST_ACC explicit AccGpuUniformCudaHipRt(Vec<TDim, TIdx> const& threadElemExtent)
            : m_threadElemExtent(threadElemExtent)
#if defined(__CUDA_ARCH__)
            , m_gridBlockExtent(b200::fromBuiltin<TDim, TIdx>(gridDim))
            , m_blockIdx(b200::fromBuiltin<TDim, TIdx>(blockIdx))
#endif
        {
        }
        //! a VIRTUAL block of the user's grid executed by a physical block of a coarsened launch (b200k::runCoarse):
        //! getWorkDiv<Grid, Blocks> and getIdx<Grid, Blocks> answer with the user's grid, not the hardware's
        ALPAKA_FN_HOST_ACC AccGpuUniformCudaHipRt(
            Vec<TDim, TIdx> const& threadElemExtent,
            Vec<TDim, TIdx> const& gridBlockExtent,
            Vec<TDim, TIdx> const& blockIdxInGrid)
            : m_threadElemExtent(threadElemExtent)
            , m_gridBlockExtent(gridBlockExtent)
            , m_blockIdx(blockIdxInGrid)
        {
        }
        AccGpuUniformCudaHipRt(AccGpuUniformCudaHipRt const&) = delete;
        AccGpuUniformCudaHipRt(AccGpuUniformCudaHipRt&&) = delete;
        auto operator=(AccGpuUniformCudaHipRt const&) -> AccGpuUniformCudaHipRt& = delete;
        auto operator=(AccGpuUniformCudaHipRt&&) -> AccGpuUniformCudaHipRt& = delete;

        Vec<TDim, TIdx> const& m_threadElemExtent;
        Vec<TDim, TIdx> const m_gridBlockExtent{};
        Vec<TDim, TIdx> const m_blockIdx{};
    };

    template<typename TDim, typename TIdx>
    using AccGpuB200 = AccGpuUniformCudaHipRt<ApiB200Rt, TDim, TIdx>;
    template<typename TDim, typename TIdx>
    using AccGpuCudaRt = AccGpuB200<TDim, TIdx>;

    //! Names of the reference's CPU accelerators. They are host placeholders only: user configuration headers name
    //! them (example/reduce/src/alpakaConfig.hpp:69-72, 102); nothing can be launched on them here.
    template<typename TDim, typename TIdx>
    class AccCpuSerial;
    template<typename TDim, typename TIdx>
    class AccCpuOmp2Blocks;
    template<typename TDim, typename TIdx>
    class AccCpuThreads;

    namespace trait
    {
        template<typename TAcc, typename TSfinae = void>
        struct AccType;
        template<typename TAcc, typename TSfinae = void>
        struct IsSingleThreadAcc : std::false_type
        {
        };
        template<typename TAcc, typename TSfinae = void>
        struct IsMultiThreadAcc : std::false_type
        {
        };
        template<typename TAcc, typename TSfinae = void>
        struct GetAccDevProps;
        template<typename TAcc, typename TSfinae = void>
        struct GetAccName;

        template<typename TApi, typename TDim, typename TIdx>
        struct AccType<AccGpuUniformCudaHipRt<TApi, TDim, TIdx>>
        {
            using type = AccGpuUniformCudaHipRt<TApi, TDim, TIdx>;
        };
        template<typename TApi, typename TDim, typename TIdx>
        struct IsMultiThreadAcc<AccGpuUniformCudaHipRt<TApi, TDim, TIdx>> : std::true_type
        {
        };
        template<typename TApi, typename TDim, typename TIdx>
        struct DevType<AccGpuUniformCudaHipRt<TApi, TDim, TIdx>>
        {
            using type = DevB200;
        };
        template<typename TApi, typename TDim, typename TIdx>
        struct PlatformType<AccGpuUniformCudaHipRt<TApi, TDim, TIdx>>
        {
            using type = PlatformB200;
        };
        template<typename TApi, typename TDim, typename TIdx>
        struct DimType<AccGpuUniformCudaHipRt<TApi, TDim, TIdx>>
        {
            using type = TDim;
        };
        template<typename TApi, typename TDim, typename TIdx>
        struct IdxType<AccGpuUniformCudaHipRt<TApi, TDim, TIdx>>
        {
            using type = TIdx;
        };
        template<typename TDim, typename TIdx>
        struct AccToTag<AccGpuUniformCudaHipRt<ApiB200Rt, TDim, TIdx>>
        {
            using type = TagGpuB200;
        };
        template<typename TDim, typename TIdx>
        struct TagToAcc<TagGpuB200, TDim, TIdx>
        {
            using type = AccGpuB200<TDim, TIdx>;
        };

        // host placeholders: enough for `Dev<AccCpuSerial<...>>` style aliases in user configuration code
        template<typename TDim, typename TIdx>
        struct DevType<AccCpuSerial<TDim, TIdx>>
        {
            using type = DevCpu;
        };
        template<typename TDim, typename TIdx>
        struct PlatformType<AccCpuSerial<TDim, TIdx>>
        {
            using type = PlatformCpu;
        };
        template<typename TDim, typename TIdx>
        struct DimType<AccCpuSerial<TDim, TIdx>>
        {
            using type = TDim;
        };
        template<typename TDim, typename TIdx>
        struct IdxType<AccCpuSerial<TDim, TIdx>>
        {
            using type = TIdx;
        };

        template<typename TApi, typename TDim, typename TIdx>
        struct GetAccDevProps<AccGpuUniformCudaHipRt<TApi, TDim, TIdx>>
        {
            static auto getAccDevProps(DevB200 const& dev) -> AccDevProps<TDim, TIdx>
            {
                b200_acc_dev_props p{};
                // the C ABI reports limits in alpaka order (index 0 = slowest dimension) for `dim` dimensions
                b200::check(b200_acc_dev_props_get(dev.getNativeHandle(), static_cast<int>(TDim::value == 0u ? 1u : TDim::value), &p));
                AccDevProps<TDim, TIdx> r{};
                r.m_multiProcessorCount = b200::clampIdx<TIdx>(p.multi_processor_count);
                r.m_gridBlockCountMax = b200::clampIdx<TIdx>(p.grid_block_count_max);
                r.m_blockThreadCountMax = b200::clampIdx<TIdx>(p.block_thread_count_max);
                r.m_threadElemCountMax = b200::clampIdx<TIdx>(p.thread_elem_count_max);
                for(std::size_t d = 0; d < TDim::value; ++d)
                {
                    r.m_gridBlockExtentMax[d] = b200::clampIdx<TIdx>(p.grid_block_extent_max[d]);
                    r.m_blockThreadExtentMax[d] = b200::clampIdx<TIdx>(p.block_thread_extent_max[d]);
                    r.m_threadElemExtentMax[d] = b200::clampIdx<TIdx>(p.thread_elem_extent_max[d]);
                }
                r.m_sharedMemSizeBytes = static_cast<std::size_t>(p.shared_mem_size_bytes);
                r.m_globalMemSizeBytes = static_cast<std::size_t>(p.global_mem_size_bytes);
                return r;
            }
        };

        template<typename TApi, typename TDim, typename TIdx>
        struct GetAccName<AccGpuUniformCudaHipRt<TApi, TDim, TIdx>>
        {
            static auto getAccName() -> std::string
            {
                int status = 0;
                char* dm = abi::__cxa_demangle(typeid(TIdx).name(), nullptr, nullptr, &status);
                std::string const idx = (status == 0 && dm != nullptr) ? dm : typeid(TIdx).name();
                std::free(dm);
                return std::string("AccGpuB200<") + std::to_string(TDim::value) + "," + idx + ">";
            }
        };
    } // namespace trait

    namespace core
    {
        namespace detail
        {
            template<typename T>
            inline auto demangle() -> std::string
            {
                int status = 0;
                char* dm = abi::__cxa_demangle(typeid(T).name(), nullptr, nullptr, &status);
                std::string const r = (status == 0 && dm != nullptr) ? dm : typeid(T).name();
                std::free(dm);
                return r;
            }
        } // namespace detail
        //! human-readable name of T (reference: core/DemangleTypeNames.hpp)
        template<typename T>
        inline std::string const demangled = detail::demangle<T>();
    } // namespace core

    template<typename TAcc>
    using Acc = typename trait::AccType<TAcc>::type;

    template<typename TAcc>
    inline constexpr bool isAccelerator = requires { typename trait::AccType<std::decay_t<TAcc>>::type; };
    template<typename TAcc>
    inline constexpr bool isSingleThreadAcc = trait::IsSingleThreadAcc<TAcc>::value;
    template<typename TAcc>
    inline constexpr bool isMultiThreadAcc = trait::IsMultiThreadAcc<TAcc>::value;

    //! \return The acceleration properties on the given device.
    template<typename TAcc, typename TDev>
    [[nodiscard]] auto getAccDevProps(TDev const& dev) -> AccDevProps<Dim<TAcc>, Idx<TAcc>>
    {
        return trait::GetAccDevProps<TAcc>::getAccDevProps(dev);
    }
    //! \return The accelerator name
    template<typename TAcc>
    [[nodiscard]] auto getAccName() -> std::string
    {
        return trait::GetAccName<TAcc>::getAccName();
    }

    namespace detail
    {
        //! gives every read-modify-write op tag the reference's call operator: apply NON-atomically, return the old value
        template<typename TOp>
        struct AtomicOpCall
        {
            template<typename T>
            ALPAKA_FN_HOST_ACC auto operator()(T* addr, T const& value) const -> T
            {
                T const old = *addr;
                *addr = TOp::next(old, value);
                return old;
            }
        };
    } // namespace detail

    // ---- atomic operation tags. operator() applies the operation NON-atomically and returns the old value, which
    // is how the reference's tests compute expected results on the host (atomic/Op.hpp).
    struct AtomicAdd : detail::AtomicOpCall<AtomicAdd>
    {
        template<typename T>
        ALPAKA_FN_HOST_ACC static auto next(T const& old, T const& value) -> T
        {
            return static_cast<T>(old + value);
        }
    };
    struct AtomicSub : detail::AtomicOpCall<AtomicSub>
    {
        template<typename T>
        ALPAKA_FN_HOST_ACC static auto next(T const& old, T const& value) -> T
        {
            return static_cast<T>(old - value);
        }
    };
    struct AtomicMin : detail::AtomicOpCall<AtomicMin>
    {
        template<typename T>
        ALPAKA_FN_HOST_ACC static auto next(T const& old, T const& value) -> T
        {
            return value < old ? value : old;
        }
    };
    struct AtomicMax : detail::AtomicOpCall<AtomicMax>
    {
        template<typename T>
        ALPAKA_FN_HOST_ACC static auto next(T const& old, T const& value) -> T
        {
            return value > old ? value : old;
        }
    };
    struct AtomicExch : detail::AtomicOpCall<AtomicExch>
    {
        template<typename T>
        ALPAKA_FN_HOST_ACC static auto next(T const&, T const& value) -> T
        {
            return value;
        }
    };
    //! old >= value ? 0 : old + 1
    struct AtomicInc : detail::AtomicOpCall<AtomicInc>
    {
        template<typename T>
        ALPAKA_FN_HOST_ACC static auto next(T const& old, T const& value) -> T
        {
            return old >= value ? static_cast<T>(0) : static_cast<T>(old + 1);
        }
    };
    //! (old == 0 || old > value) ? value : old - 1
    struct AtomicDec : detail::AtomicOpCall<AtomicDec>
    {
        template<typename T>
        ALPAKA_FN_HOST_ACC static auto next(T const& old, T const& value) -> T
        {
            return (old == static_cast<T>(0) || old > value) ? value : static_cast<T>(old - 1);
        }
    };
    struct AtomicAnd : detail::AtomicOpCall<AtomicAnd>
    {
        template<typename T>
        ALPAKA_FN_HOST_ACC static auto next(T const& old, T const& value) -> T
        {
            return static_cast<T>(old & value);
        }
    };
    struct AtomicOr : detail::AtomicOpCall<AtomicOr>
    {
        template<typename T>
        ALPAKA_FN_HOST_ACC static auto next(T const& old, T const& value) -> T
        {
            return static_cast<T>(old | value);
        }
    };
    struct AtomicXor : detail::AtomicOpCall<AtomicXor>
    {
        template<typename T>
        ALPAKA_FN_HOST_ACC static auto next(T const& old, T const& value) -> T
        {
            return static_cast<T>(old ^ value);
        }
    };
    //! old == compare ? value : old
    struct AtomicCas
    {
        template<typename T>
        ALPAKA_FN_HOST_ACC static auto next(T const& old, T const& compare, T const& value) -> T
        {
            return old == compare ? value : old;
        }
        template<typename T>
        ALPAKA_FN_HOST_ACC auto operator()(T* addr, T const& compare, T const& value) const -> T
        {
            T const old = *addr;
            *addr = next(old, compare, value);
            return old;
        }
    };

    // ---- block synchronisation predicate operations
    struct BlockCount
    {
        enum
        {
            InitialValue = 0u
        };
        template<typename T>
        ALPAKA_FN_HOST_ACC auto operator()(T const& currentResult, T const& value) const -> T
        {
            return currentResult + static_cast<T>(value != static_cast<T>(0));
        }
    };
    struct BlockAnd
    {
        enum
        {
            InitialValue = 1u
        };
        template<typename T>
        ALPAKA_FN_HOST_ACC auto operator()(T const& currentResult, T const& value) const -> T
        {
            return static_cast<T>(currentResult && (value != static_cast<T>(0)));
        }
    };
    struct BlockOr
    {
        enum
        {
            InitialValue = 0u
        };
        template<typename T>
        ALPAKA_FN_HOST_ACC auto operator()(T const& currentResult, T const& value) const -> T
        {
            return static_cast<T>(currentResult || (value != static_cast<T>(0)));
        }
    };

    namespace math::constants
    {
        inline constexpr double e = 2.718281828459045235360287471352662498;
        inline constexpr double log2e = 1.442695040888963407359924681001892137;
        inline constexpr double log10e = 0.434294481903251827651128918916605082;
        inline constexpr double pi = 3.141592653589793238462643383279502884;
        inline constexpr double inv_pi = 0.318309886183790671537767526745028724;
        inline constexpr double ln2 = 0.693147180559945309417232121458176568;
        inline constexpr double ln10 = 2.302585092994045684017991454684364208;
        inline constexpr double sqrt2 = 1.414213562373095048801688724209698079;
        inline constexpr double sqrt3 = 1.732050807568877293527446341505872367;
    } // namespace math::constants

#if defined(__CUDACC__)
    // =============================================================================================================
    // device side
    namespace b200
    {
        //! dim3-like built-in -> alpaka vector: x is the LAST component
        template<typename TDim, typename TIdx, typename TBuiltin>
        __device__ __forceinline__ auto fromBuiltin(TBuiltin const& v) -> Vec<TDim, TIdx>
        {
            if constexpr(TDim::value == 0u)
                return Vec<TDim, TIdx>{};
            else if constexpr(TDim::value == 1u)
                return Vec<TDim, TIdx>{static_cast<TIdx>(v.x)};
            else if constexpr(TDim::value == 2u)
                return Vec<TDim, TIdx>{static_cast<TIdx>(v.y), static_cast<TIdx>(v.x)};
            else
                return Vec<TDim, TIdx>{static_cast<TIdx>(v.z), static_cast<TIdx>(v.y), static_cast<TIdx>(v.x)};
        }
    } // namespace b200

    namespace trait
    {
        template<typename TApi, typename TDim, typename TIdx>
        struct GetWorkDiv<AccGpuUniformCudaHipRt<TApi, TDim, TIdx>, origin::Grid, unit::Blocks>
        {
            __device__ static auto getWorkDiv(AccGpuUniformCudaHipRt<TApi, TDim, TIdx> const& acc) -> Vec<TDim, TIdx>
            {
                return acc.m_gridBlockExtent;
            }
        };
        template<typename TApi, typename TDim, typename TIdx>
        struct GetWorkDiv<AccGpuUniformCudaHipRt<TApi, TDim, TIdx>, origin::Block, unit::Threads>
        {
            __device__ static auto getWorkDiv(AccGpuUniformCudaHipRt<TApi, TDim, TIdx> const&) -> Vec<TDim, TIdx>
            {
                return b200::fromBuiltin<TDim, TIdx>(blockDim);
            }
        };
        template<typename TApi, typename TDim, typename TIdx>
        struct GetWorkDiv<AccGpuUniformCudaHipRt<TApi, TDim, TIdx>, origin::Thread, unit::Elems>
        {
            __device__ static auto getWorkDiv(AccGpuUniformCudaHipRt<TApi, TDim, TIdx> const& acc) -> Vec<TDim, TIdx>
            {
                return acc.m_threadElemExtent;
            }
        };

        template<typename TIdxProvider, typename TOrigin, typename TUnit, typename TSfinae = void>
        struct GetIdx;

        template<typename TApi, typename TDim, typename TIdx>
        struct GetIdx<AccGpuUniformCudaHipRt<TApi, TDim, TIdx>, origin::Grid, unit::Blocks>
        {
            template<typename TWorkDiv>
            __device__ static auto getIdx(AccGpuUniformCudaHipRt<TApi, TDim, TIdx> const& acc, TWorkDiv const&) -> Vec<TDim, TIdx>
            {
                return acc.m_blockIdx;
            }
        };
        template<typename TApi, typename TDim, typename TIdx>
        struct GetIdx<AccGpuUniformCudaHipRt<TApi, TDim, TIdx>, origin::Block, unit::Threads>
        {
            template<typename TWorkDiv>
            __device__ static auto getIdx(AccGpuUniformCudaHipRt<TApi, TDim, TIdx> const&, TWorkDiv const&) -> Vec<TDim, TIdx>
            {
                return b200::fromBuiltin<TDim, TIdx>(threadIdx);
            }
        };
    } // namespace trait

    //! Index of the calling thread/block measured from TOrigin in TUnit:
    //! Grid/Blocks, Block/Threads, Grid/Threads, Grid/Elems (first element of the thread), Block/Elems.
    template<typename TOrigin, typename TUnit, typename TIdxProvider, typename TWorkDiv>
    [[nodiscard]] __device__ auto getIdx(TIdxProvider const& idx, TWorkDiv const& workDiv) -> Vec<Dim<TWorkDiv>, Idx<TWorkDiv>>
    {
        if constexpr(std::is_same_v<TOrigin, origin::Grid> && std::is_same_v<TUnit, unit::Threads>)
            return getIdx<origin::Grid, unit::Blocks>(idx, workDiv) * getWorkDiv<origin::Block, unit::Threads>(workDiv)
                   + getIdx<origin::Block, unit::Threads>(idx, workDiv);
        else if constexpr(std::is_same_v<TOrigin, origin::Grid> && std::is_same_v<TUnit, unit::Elems>)
            return getIdx<origin::Grid, unit::Threads>(idx, workDiv) * getWorkDiv<origin::Thread, unit::Elems>(workDiv);
        else if constexpr(std::is_same_v<TOrigin, origin::Block> && std::is_same_v<TUnit, unit::Elems>)
            return getIdx<origin::Block, unit::Threads>(idx, workDiv) * getWorkDiv<origin::Thread, unit::Elems>(workDiv);
        else
            return trait::GetIdx<TIdxProvider, TOrigin, TUnit>::getIdx(idx, workDiv);
    }
    template<typename TOrigin, typename TUnit, typename TAcc>
    [[nodiscard]] __device__ auto getIdx(TAcc const& acc) -> Vec<Dim<TAcc>, Idx<TAcc>>
    {
        return getIdx<TOrigin, TUnit>(acc, acc);
    }

    // ---- block shared memory
    //! A block-shared variable of type T. The same (T, id) names the same storage on every call within a block;
    //! different ids name distinct storage; the memory is uninitialised.
    template<typename T, std::size_t TuniqueId, typename TAcc>
    __device__ auto declareSharedVar(TAcc const&) -> T&
    {
        __shared__ uint8_t shMem alignas(alignof(T))[sizeof(T)];
        return *reinterpret_cast<T*>(shMem);
    }
    template<typename TAcc>
    __device__ void freeSharedVars(TAcc&)
    {
    }
    //! The dynamic shared memory of the block (size = trait::BlockSharedMemDynSizeBytes at launch), same base
    //! pointer for every T.
    template<typename T, typename TAcc>
    __device__ auto getDynSharedMem(TAcc const&) -> T*
    {
        extern __shared__ std::byte alpakaB200DynSharedMem alignas(std::max_align_t)[];
        return reinterpret_cast<T*>(alpakaB200DynSharedMem);
    }

    // ---- block synchronisation
    template<typename TAcc>
    __device__ void syncBlockThreads(TAcc const&)
    {
        __syncthreads();
    }
    //! Barrier that also combines `predicate` over the block: BlockCount / BlockAnd / BlockOr.
    template<typename TOp, typename TAcc>
    __device__ auto syncBlockThreadsPredicate(TAcc const&, int predicate) -> int
    {
        if constexpr(std::is_same_v<TOp, BlockCount>)
            return __syncthreads_count(predicate);
        else if constexpr(std::is_same_v<TOp, BlockAnd>)
            return __syncthreads_and(predicate);
        else
        {
            static_assert(std::is_same_v<TOp, BlockOr>, "syncBlockThreadsPredicate: BlockCount, BlockAnd or BlockOr");
            return __syncthreads_or(predicate);
        }
    }

    // ---- memory fences
    template<typename TAcc, typename TMemScope>
    __device__ void mem_fence(TAcc const&, TMemScope const&)
    {
        if constexpr(std::is_same_v<TMemScope, memory_scope::Block>)
            __threadfence_block();
        else
            __threadfence();
    }

    // ---- warp collectives. All use the full member mask, like the reference: lanes that have exited do not
    // take part (warp/WarpUniformCudaHipBuiltIn.hpp:61-183).
    namespace warp
    {
        template<typename TAcc>
        [[nodiscard]] __device__ auto getSize(TAcc const&) -> std::int32_t
        {
            return warpSize;
        }
        template<typename TAcc>
        [[nodiscard]] __device__ auto activemask(TAcc const&) -> std::uint32_t
        {
            return __activemask();
        }
        template<typename TAcc>
        [[nodiscard]] __device__ auto all(TAcc const&, std::int32_t predicate) -> std::int32_t
        {
            return __all_sync(0xffff'ffffu, predicate);
        }
        template<typename TAcc>
        [[nodiscard]] __device__ auto any(TAcc const&, std::int32_t predicate) -> std::int32_t
        {
            return __any_sync(0xffff'ffffu, predicate);
        }
        template<typename TAcc>
        [[nodiscard]] __device__ auto ballot(TAcc const&, std::int32_t predicate) -> std::uint32_t
        {
            return __ballot_sync(0xffff'ffffu, predicate);
        }
        //! value of `srcLane` (modulo width); width = 0 means the warp size
        template<typename TAcc, typename T>
        [[nodiscard]] __device__ auto shfl(TAcc const&, T value, std::int32_t srcLane, std::int32_t width = 0) -> T
        {
            return __shfl_sync(0xffff'ffffu, value, srcLane, width ? width : warpSize);
        }
        template<typename TAcc, typename T>
        [[nodiscard]] __device__ auto shfl_up(TAcc const&, T value, std::uint32_t offset, std::int32_t width = 0) -> T
        {
            return __shfl_up_sync(0xffff'ffffu, value, offset, width ? width : warpSize);
        }
        template<typename TAcc, typename T>
        [[nodiscard]] __device__ auto shfl_down(TAcc const&, T value, std::uint32_t offset, std::int32_t width = 0) -> T
        {
            return __shfl_down_sync(0xffff'ffffu, value, offset, width ? width : warpSize);
        }
        template<typename TAcc, typename T>
        [[nodiscard]] __device__ auto shfl_xor(TAcc const&, T value, std::int32_t mask, std::int32_t width = 0) -> T
        {
            return __shfl_xor_sync(0xffff'ffffu, value, mask, width ? width : warpSize);
        }
    } // namespace warp

    // ---- atomics
    namespace b200
    {
        template<std::size_t N>
        struct UIntOfSize;
        template<>
        struct UIntOfSize<4u>
        {
            using type = unsigned int;
        };
        template<>
        struct UIntOfSize<8u>
        {
            using type = unsigned long long;
        };

        template<typename TTo, typename TFrom>
        __device__ __forceinline__ auto bitCast(TFrom const& v) -> TTo
        {
            static_assert(sizeof(TTo) == sizeof(TFrom));
            TTo r;
            ::memcpy(&r, &v, sizeof(TTo));
            return r;
        }

        //! CTA scope only for hierarchy::Threads (atomic between the threads of ONE block); hierarchy::Blocks means
        //! "atomic between all blocks of a grid" and hierarchy::Grids "between grids", so both need device scope
        //! (reference: atomic/AtomicUniformCudaHip.hpp:80-130 uses the *_block intrinsics for Threads only).
        template<typename THierarchy>
        inline constexpr bool blockScope = std::is_same_v<THierarchy, hierarchy::Threads>;

        template<typename THierarchy, typename U>
        __device__ __forceinline__ auto cas(U* addr, U compare, U value) -> U
        {
            if constexpr(blockScope<THierarchy>)
                return atomicCAS_block(addr, compare, value);
            else
                return atomicCAS(addr, compare, value);
        }

        //! any read-modify-write as a compare-and-swap loop on the 4/8-byte word
        template<typename TOp, typename THierarchy, typename T>
        __device__ auto rmwLoop(T* addr, T const& value) -> T
        {
            static_assert(sizeof(T) == 4u || sizeof(T) == 8u, "atomics are defined for 4- and 8-byte types");
            using U = typename UIntOfSize<sizeof(T)>::type;
            U* const a = reinterpret_cast<U*>(addr);
            U old = *a;
            U assumed;
            do
            {
                assumed = old;
                T const next = TOp::next(bitCast<T>(assumed), value);
                old = cas<THierarchy>(a, assumed, bitCast<U>(next));
            } while(assumed != old);
            return bitCast<T>(old);
        }

#    define ALPAKA_B200_NATIVE_ATOMIC(fn, addr, value)                                                                \
        (blockScope<THierarchy> ? fn##_block((addr), (value)) : fn((addr), (value)))

        template<typename TOp, typename THierarchy, typename T>
        __device__ auto atomicDispatch(T* addr, T const& value) -> T
        {
            using U = typename UIntOfSize<sizeof(T)>::type;
            constexpr bool isInt = std::is_integral_v<T>;
            if constexpr(std::is_same_v<TOp, AtomicAdd>)
            {
                if constexpr(isInt) // two's complement: unsigned add of the same width
                    return static_cast<T>(ALPAKA_B200_NATIVE_ATOMIC(atomicAdd, reinterpret_cast<U*>(addr), static_cast<U>(value)));
                else if constexpr(std::is_same_v<T, float> || std::is_same_v<T, double>)
                    return ALPAKA_B200_NATIVE_ATOMIC(atomicAdd, addr, value);
                else
                    return rmwLoop<TOp, THierarchy>(addr, value);
            }
            else if constexpr(std::is_same_v<TOp, AtomicSub> && isInt)
            {
                return static_cast<T>(
                    ALPAKA_B200_NATIVE_ATOMIC(atomicAdd, reinterpret_cast<U*>(addr), static_cast<U>(0) - static_cast<U>(value)));
            }
            else if constexpr((std::is_same_v<TOp, AtomicMin> || std::is_same_v<TOp, AtomicMax>) &&isInt)
            {
                using N = std::conditional_t<
                    sizeof(T) == 4u,
                    std::conditional_t<std::is_signed_v<T>, int, unsigned int>,
                    std::conditional_t<std::is_signed_v<T>, long long, unsigned long long>>;
                if constexpr(std::is_same_v<TOp, AtomicMin>)
                    return static_cast<T>(ALPAKA_B200_NATIVE_ATOMIC(atomicMin, reinterpret_cast<N*>(addr), static_cast<N>(value)));
                else
                    return static_cast<T>(ALPAKA_B200_NATIVE_ATOMIC(atomicMax, reinterpret_cast<N*>(addr), static_cast<N>(value)));
            }
            else if constexpr(std::is_same_v<TOp, AtomicExch>)
            {
                return bitCast<T>(ALPAKA_B200_NATIVE_ATOMIC(atomicExch, reinterpret_cast<U*>(addr), bitCast<U>(value)));
            }
            else if constexpr(std::is_same_v<TOp, AtomicAnd> && isInt)
            {
                return static_cast<T>(ALPAKA_B200_NATIVE_ATOMIC(atomicAnd, reinterpret_cast<U*>(addr), static_cast<U>(value)));
            }
            else if constexpr(std::is_same_v<TOp, AtomicOr> && isInt)
            {
                return static_cast<T>(ALPAKA_B200_NATIVE_ATOMIC(atomicOr, reinterpret_cast<U*>(addr), static_cast<U>(value)));
            }
            else if constexpr(std::is_same_v<TOp, AtomicXor> && isInt)
            {
                return static_cast<T>(ALPAKA_B200_NATIVE_ATOMIC(atomicXor, reinterpret_cast<U*>(addr), static_cast<U>(value)));
            }
            else if constexpr((std::is_same_v<TOp, AtomicInc> || std::is_same_v<TOp, AtomicDec>) &&std::is_same_v<T, unsigned int>)
            {
                if constexpr(std::is_same_v<TOp, AtomicInc>)
                    return ALPAKA_B200_NATIVE_ATOMIC(atomicInc, addr, value);
                else
                    return ALPAKA_B200_NATIVE_ATOMIC(atomicDec, addr, value);
            }
            else
            {
                return rmwLoop<TOp, THierarchy>(addr, value);
            }
        }
#    undef ALPAKA_B200_NATIVE_ATOMIC
    } // namespace b200

    //! Executes the given operation atomically. \return The old value of *addr.
    template<typename TOp, typename TAcc, typename T, typename THierarchy = hierarchy::Grids>
    __device__ auto atomicOp(TAcc const&, T* const addr, T const& value, THierarchy const& = THierarchy()) -> T
    {
        return b200::atomicDispatch<TOp, THierarchy>(addr, value);
    }
    //! Compare-and-swap form.
    template<typename TOp, typename TAcc, typename T, typename THierarchy = hierarchy::Grids>
    __device__ auto atomicOp(TAcc const&, T* const addr, T const& compare, T const& value, THierarchy const& = THierarchy()) -> T
    {
        static_assert(std::is_same_v<TOp, AtomicCas>, "the 4-argument atomicOp is the compare-and-swap");
        static_assert(sizeof(T) == 4u || sizeof(T) == 8u, "atomics are defined for 4- and 8-byte types");
        using U = typename b200::UIntOfSize<sizeof(T)>::type;
        if constexpr(std::is_integral_v<T>)
        {
            return static_cast<T>(b200::cas<THierarchy>(reinterpret_cast<U*>(addr), static_cast<U>(compare), static_cast<U>(value)));
        }
        else
        {
            // floating point: value comparison (so that -0.0 == +0.0, like the reference's CPU semantics)
            U* const a = reinterpret_cast<U*>(addr);
            U old = *a;
            U assumed;
            do
            {
                assumed = old;
                T const cur = b200::bitCast<T>(assumed);
                if(!(cur == compare))
                    return cur;
                old = b200::cas<THierarchy>(a, assumed, b200::bitCast<U>(value));
            } while(assumed != old);
            return b200::bitCast<T>(old);
        }
    }

#    define ALPAKA_B200_NAMED_ATOMIC(name, op)                                                                        \
        template<typename TAcc, typename T, typename THierarchy = hierarchy::Grids>                                   \
        __device__ auto name(TAcc const& acc, T* const addr, T const& value, THierarchy const& hier = THierarchy()) -> T \
        {                                                                                                             \
            return atomicOp<op>(acc, addr, value, hier);                                                              \
        }
    ALPAKA_B200_NAMED_ATOMIC(atomicAdd, AtomicAdd)
    ALPAKA_B200_NAMED_ATOMIC(atomicSub, AtomicSub)
    ALPAKA_B200_NAMED_ATOMIC(atomicMin, AtomicMin)
    ALPAKA_B200_NAMED_ATOMIC(atomicMax, AtomicMax)
    ALPAKA_B200_NAMED_ATOMIC(atomicExch, AtomicExch)
    ALPAKA_B200_NAMED_ATOMIC(atomicInc, AtomicInc)
    ALPAKA_B200_NAMED_ATOMIC(atomicDec, AtomicDec)
    ALPAKA_B200_NAMED_ATOMIC(atomicAnd, AtomicAnd)
    ALPAKA_B200_NAMED_ATOMIC(atomicOr, AtomicOr)
    ALPAKA_B200_NAMED_ATOMIC(atomicXor, AtomicXor)
#    undef ALPAKA_B200_NAMED_ATOMIC
    template<typename TAcc, typename T, typename THierarchy = hierarchy::Grids>
    __device__ auto atomicCas(TAcc const& acc, T* const addr, T const& compare, T const& value, THierarchy const& hier = THierarchy()) -> T
    {
        return atomicOp<AtomicCas>(acc, addr, compare, value, hier);
    }

    // ---- bit intrinsics
    template<typename TAcc, typename T>
    [[nodiscard]] __device__ auto popcount(TAcc const&, T value) -> std::int32_t
    {
        static_assert(std::is_integral_v<T> && (sizeof(T) == 4u || sizeof(T) == 8u));
        if constexpr(sizeof(T) == 4u)
            return __popc(static_cast<unsigned int>(value));
        else
            return __popcll(static_cast<unsigned long long>(value));
    }
    //! 1-based position of the least significant set bit, 0 if none
    template<typename TAcc, typename T>
    [[nodiscard]] __device__ auto ffs(TAcc const&, T value) -> std::int32_t
    {
        static_assert(std::is_integral_v<T> && (sizeof(T) == 4u || sizeof(T) == 8u));
        if constexpr(sizeof(T) == 4u)
            return __ffs(static_cast<int>(value));
        else
            return __ffsll(static_cast<long long>(value));
    }

    // ---- math: alpaka::math::f(acc, x) -> the CUDA device overload of f
    namespace math
    {
#    define ALPAKA_B200_MATH_1(name)                                                                                  \
        template<typename TAcc, typename T>                                                                           \
        __device__ auto name(TAcc const&, T const& x)                                                                 \
        {                                                                                                             \
            return ::name(x);                                                                                         \
        }
#    define ALPAKA_B200_MATH_2(name)                                                                                  \
        template<typename TAcc, typename T, typename U>                                                               \
        __device__ auto name(TAcc const&, T const& x, U const& y)                                                     \
        {                                                                                                             \
            using C = std::common_type_t<T, U>;                                                                       \
            return ::name(static_cast<C>(x), static_cast<C>(y));                                                      \
        }
        ALPAKA_B200_MATH_1(sqrt)
        ALPAKA_B200_MATH_1(cbrt)
        ALPAKA_B200_MATH_1(exp)
        ALPAKA_B200_MATH_1(log)
        ALPAKA_B200_MATH_1(log2)
        ALPAKA_B200_MATH_1(log10)
        ALPAKA_B200_MATH_1(sin)
        ALPAKA_B200_MATH_1(cos)
        ALPAKA_B200_MATH_1(tan)
        ALPAKA_B200_MATH_1(asin)
        ALPAKA_B200_MATH_1(acos)
        ALPAKA_B200_MATH_1(atan)
        ALPAKA_B200_MATH_1(sinh)
        ALPAKA_B200_MATH_1(cosh)
        ALPAKA_B200_MATH_1(tanh)
        ALPAKA_B200_MATH_1(asinh)
        ALPAKA_B200_MATH_1(acosh)
        ALPAKA_B200_MATH_1(atanh)
        ALPAKA_B200_MATH_1(erf)
        ALPAKA_B200_MATH_1(floor)
        ALPAKA_B200_MATH_1(ceil)
        ALPAKA_B200_MATH_1(trunc)
        ALPAKA_B200_MATH_1(round)
        ALPAKA_B200_MATH_1(lround)
        ALPAKA_B200_MATH_1(llround)
        ALPAKA_B200_MATH_1(isnan)
        ALPAKA_B200_MATH_1(isinf)
        ALPAKA_B200_MATH_1(isfinite)
        ALPAKA_B200_MATH_2(atan2)
        ALPAKA_B200_MATH_2(pow)
        ALPAKA_B200_MATH_2(fmod)
        ALPAKA_B200_MATH_2(remainder)
        ALPAKA_B200_MATH_2(copysign)
#    undef ALPAKA_B200_MATH_1
#    undef ALPAKA_B200_MATH_2

        template<typename TAcc, typename T>
        [[nodiscard]] __device__ auto abs(TAcc const&, T const& x) -> T
        {
            if constexpr(std::is_floating_point_v<T>)
                return ::fabs(x);
            else if constexpr(std::is_signed_v<T>)
                return x < 0 ? static_cast<T>(-x) : x;
            else
                return x;
        }
        template<typename TAcc, typename T>
        [[nodiscard]] __device__ auto rsqrt(TAcc const&, T const& x) -> T
        {
            if constexpr(std::is_same_v<T, float>)
                return ::rsqrtf(x);
            else
                return ::rsqrt(static_cast<double>(x));
        }
        template<typename TAcc, typename T>
        __device__ void sincos(TAcc const&, T const& x, T& s, T& c)
        {
            if constexpr(std::is_same_v<T, float>)
                ::sincosf(x, &s, &c);
            else
                ::sincos(x, &s, &c);
        }
        template<typename TAcc, typename T>
        [[nodiscard]] __device__ auto fma(TAcc const&, T const& x, T const& y, T const& z) -> T
        {
            return ::fma(x, y, z);
        }
        //! min/max: integers compare, floating point follows fmin/fmax (a NaN operand yields the other one)
        template<typename TAcc, typename T, typename U>
        [[nodiscard]] __device__ auto min(TAcc const&, T const& x, U const& y) -> std::common_type_t<T, U>
        {
            using C = std::common_type_t<T, U>;
            if constexpr(std::is_floating_point_v<C>)
                return ::fmin(static_cast<C>(x), static_cast<C>(y));
            else
                return static_cast<C>(y) < static_cast<C>(x) ? static_cast<C>(y) : static_cast<C>(x);
        }
        template<typename TAcc, typename T, typename U>
        [[nodiscard]] __device__ auto max(TAcc const&, T const& x, U const& y) -> std::common_type_t<T, U>
        {
            using C = std::common_type_t<T, U>;
            if constexpr(std::is_floating_point_v<C>)
                return ::fmax(static_cast<C>(x), static_cast<C>(y));
            else
                return static_cast<C>(x) < static_cast<C>(y) ? static_cast<C>(y) : static_cast<C>(x);
        }
    } // namespace math
#endif // __CUDACC__
} // namespace alpaka
