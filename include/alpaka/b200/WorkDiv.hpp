// include/alpaka/b200/WorkDiv.hpp -- work division: WorkDivMembers, AccDevProps, subDivideGridElems, isValidWorkDiv.
//
// API parity with the reference's workdiv/WorkDivMembers.hpp:18-97, acc/AccDevProps.hpp:16-33 and
// workdiv/WorkDivHelpers.hpp:30-309, 406-549. The selection algorithm itself lives behind the C ABI
// (b200_subdivide_grid_elems / b200_is_valid_work_div, alpaka_b200/csrc/b200_workdiv.cpp) so that every host
// language binds the same logic; it is pinned against the reference's golden vectors
// (test/unit/workDiv/src/WorkDivHelpersTest.cpp:34-180) in tests/test_workdiv.py.
#pragma once

#include "Dev.hpp"

#include <ostream>

namespace alpaka
{
    //! A basic class holding the work division as grid block extent, block thread extent and thread element extent.
    template<typename TDim, typename TIdx>
    class WorkDivMembers
    {
    public:
        ALPAKA_FN_HOST_ACC WorkDivMembers() = delete;

        //! Accepts anything with extents -- Vecs, user vector types with a trait::GetExtents, plain scalars for 1-D --
        //! of TDim or MORE dimensions; extra (slowest) dimensions are dropped
        //! (reference: workdiv/WorkDivMembers.hpp:26-35, test/unit/workDiv/src/WorkDivHelpersTest.cpp:223-249).
        template<typename TGridBlockExtent, typename TBlockThreadExtent, typename TThreadElemExtent>
        ALPAKA_FN_HOST_ACC explicit WorkDivMembers(
            TGridBlockExtent const& gridBlockExtent = TGridBlockExtent(),
            TBlockThreadExtent const& blockThreadExtent = TBlockThreadExtent(),
            TThreadElemExtent const& threadElemExtent = TThreadElemExtent())
            : m_gridBlockExtent(subVecEnd<TDim>(castVec<TIdx>(getExtents(gridBlockExtent))))
            , m_blockThreadExtent(subVecEnd<TDim>(castVec<TIdx>(getExtents(blockThreadExtent))))
            , m_threadElemExtent(subVecEnd<TDim>(castVec<TIdx>(getExtents(threadElemExtent))))
        {
        }

        //! braced lists: WorkDivMembers<Dim,Idx>{{2,2}, {16,16}, {1,1}}
        ALPAKA_FN_HOST_ACC WorkDivMembers(
            Vec<TDim, TIdx> const& gridBlockExtent,
            Vec<TDim, TIdx> const& blockThreadExtent,
            Vec<TDim, TIdx> const& threadElemExtent)
            : m_gridBlockExtent(gridBlockExtent)
            , m_blockThreadExtent(blockThreadExtent)
            , m_threadElemExtent(threadElemExtent)
        {
        }

        //! copy from any other work-division-like object (e.g. an accelerator on the device)
        template<typename TWorkDiv, typename = decltype(std::declval<TWorkDiv const&>().m_gridBlockExtent)>
        ALPAKA_FN_HOST_ACC explicit WorkDivMembers(TWorkDiv const& other)
            : m_gridBlockExtent(castVec<TIdx>(other.m_gridBlockExtent))
            , m_blockThreadExtent(castVec<TIdx>(other.m_blockThreadExtent))
            , m_threadElemExtent(castVec<TIdx>(other.m_threadElemExtent))
        {
        }

        ALPAKA_FN_HOST_ACC friend constexpr auto operator==(WorkDivMembers const& a, WorkDivMembers const& b) -> bool
        {
            return a.m_gridBlockExtent == b.m_gridBlockExtent && a.m_blockThreadExtent == b.m_blockThreadExtent
                   && a.m_threadElemExtent == b.m_threadElemExtent;
        }
        ALPAKA_FN_HOST_ACC friend constexpr auto operator!=(WorkDivMembers const& a, WorkDivMembers const& b) -> bool
        {
            return !(a == b);
        }
        friend auto operator<<(std::ostream& os, WorkDivMembers const& w) -> std::ostream&
        {
            return os << "{gridBlockExtent: " << w.m_gridBlockExtent << ", blockThreadExtent: " << w.m_blockThreadExtent
                      << ", threadElemExtent: " << w.m_threadElemExtent << "}";
        }

        Vec<TDim, TIdx> m_gridBlockExtent;
        Vec<TDim, TIdx> m_blockThreadExtent;
        Vec<TDim, TIdx> m_threadElemExtent;
    };

    // deduction guide: WorkDivMembers{Vec, Vec, Vec}
    template<typename TDim, typename TIdx>
    ALPAKA_FN_HOST_ACC WorkDivMembers(Vec<TDim, TIdx> const&, Vec<TDim, TIdx> const&, Vec<TDim, TIdx> const&)
        -> WorkDivMembers<TDim, TIdx>;

    namespace trait
    {
        template<typename TDim, typename TIdx>
        struct DimType<WorkDivMembers<TDim, TIdx>>
        {
            using type = TDim;
        };
        template<typename TDim, typename TIdx>
        struct IdxType<WorkDivMembers<TDim, TIdx>>
        {
            using type = TIdx;
        };

        //! GetWorkDiv<TWorkDiv, TOrigin, TUnit>: the three basic extents; derived ones are composed below
        template<typename TWorkDiv, typename TOrigin, typename TUnit, typename TSfinae = void>
        struct GetWorkDiv;

        template<typename TDim, typename TIdx>
        struct GetWorkDiv<WorkDivMembers<TDim, TIdx>, origin::Grid, unit::Blocks>
        {
            ALPAKA_FN_HOST_ACC static auto getWorkDiv(WorkDivMembers<TDim, TIdx> const& w) -> Vec<TDim, TIdx>
            {
                return w.m_gridBlockExtent;
            }
        };
        template<typename TDim, typename TIdx>
        struct GetWorkDiv<WorkDivMembers<TDim, TIdx>, origin::Block, unit::Threads>
        {
            ALPAKA_FN_HOST_ACC static auto getWorkDiv(WorkDivMembers<TDim, TIdx> const& w) -> Vec<TDim, TIdx>
            {
                return w.m_blockThreadExtent;
            }
        };
        template<typename TDim, typename TIdx>
        struct GetWorkDiv<WorkDivMembers<TDim, TIdx>, origin::Thread, unit::Elems>
        {
            ALPAKA_FN_HOST_ACC static auto getWorkDiv(WorkDivMembers<TDim, TIdx> const& w) -> Vec<TDim, TIdx>
            {
                return w.m_threadElemExtent;
            }
        };
    } // namespace trait

    //! Extent of the work division measured from TOrigin in TUnit
    //! (Grid/Blocks, Block/Threads, Thread/Elems, Grid/Threads, Grid/Elems, Block/Elems).
    ALPAKA_NO_HOST_ACC_WARNING
    template<typename TOrigin, typename TUnit, typename TWorkDiv>
    [[nodiscard]] ALPAKA_FN_HOST_ACC auto getWorkDiv(TWorkDiv const& workDiv) -> Vec<Dim<TWorkDiv>, Idx<TWorkDiv>>
    {
        if constexpr(std::is_same_v<TOrigin, origin::Grid> && std::is_same_v<TUnit, unit::Threads>)
            return getWorkDiv<origin::Grid, unit::Blocks>(workDiv) * getWorkDiv<origin::Block, unit::Threads>(workDiv);
        else if constexpr(std::is_same_v<TOrigin, origin::Grid> && std::is_same_v<TUnit, unit::Elems>)
            return getWorkDiv<origin::Grid, unit::Threads>(workDiv) * getWorkDiv<origin::Thread, unit::Elems>(workDiv);
        else if constexpr(std::is_same_v<TOrigin, origin::Block> && std::is_same_v<TUnit, unit::Elems>)
            return getWorkDiv<origin::Block, unit::Threads>(workDiv) * getWorkDiv<origin::Thread, unit::Elems>(workDiv);
        else
            return trait::GetWorkDiv<TWorkDiv, TOrigin, TUnit>::getWorkDiv(workDiv);
    }

    //! The acceleration properties on a device.
    template<typename TDim, typename TIdx>
    struct AccDevProps
    {
        static_assert(sizeof(TIdx) >= sizeof(int), "Index type is not supported, consider using int or a larger type.");

        // member order is part of the API (aggregate initialisation in user code and the reference's tests)
        TIdx m_multiProcessorCount; //!< The number of multiprocessors.
        Vec<TDim, TIdx> m_gridBlockExtentMax; //!< The maximum number of blocks in each dimension of the grid.
        TIdx m_gridBlockCountMax; //!< The maximum number of blocks in a grid.
        Vec<TDim, TIdx> m_blockThreadExtentMax; //!< The maximum number of threads in each dimension of a block.
        TIdx m_blockThreadCountMax; //!< The maximum number of threads in a block.
        Vec<TDim, TIdx> m_threadElemExtentMax; //!< The maximum number of elements in each dimension of a thread.
        TIdx m_threadElemCountMax; //!< The maximum number of elements in a threads.
        std::size_t m_sharedMemSizeBytes; //!< The size of shared memory per block
        std::size_t m_globalMemSizeBytes; //!< The size of global memory
    };

    //! The grid block extent subdivision restrictions.
    enum class GridBlockExtentSubDivRestrictions
    {
        EqualExtent, //!< The block thread extent will be equal in all dimensions.
        CloseToEqualExtent, //!< The block thread extent will be as close to equal as possible in all dimensions.
        Unrestricted, //!< The block thread extent will not have any restrictions.
    };

    namespace b200
    {
        template<typename TDim, typename TIdx>
        inline auto toAbiProps(AccDevProps<TDim, TIdx> const& p) -> b200_acc_dev_props
        {
            b200_acc_dev_props a{};
            a.multi_processor_count = static_cast<uint64_t>(p.m_multiProcessorCount);
            a.grid_block_count_max = static_cast<uint64_t>(p.m_gridBlockCountMax);
            a.block_thread_count_max = static_cast<uint64_t>(p.m_blockThreadCountMax);
            a.thread_elem_count_max = static_cast<uint64_t>(p.m_threadElemCountMax);
            for(std::size_t d = 0; d < TDim::value; ++d)
            {
                a.grid_block_extent_max[d] = static_cast<uint64_t>(p.m_gridBlockExtentMax[d]);
                a.block_thread_extent_max[d] = static_cast<uint64_t>(p.m_blockThreadExtentMax[d]);
                a.thread_elem_extent_max[d] = static_cast<uint64_t>(p.m_threadElemExtentMax[d]);
            }
            a.shared_mem_size_bytes = p.m_sharedMemSizeBytes;
            a.global_mem_size_bytes = p.m_globalMemSizeBytes;
            return a;
        }

        //! clamp a 64-bit limit reported by the C ABI into TIdx
        template<typename TIdx>
        constexpr auto clampIdx(uint64_t v) -> TIdx
        {
            constexpr uint64_t m = static_cast<uint64_t>(std::numeric_limits<TIdx>::max());
            return static_cast<TIdx>(v > m ? m : v);
        }
    } // namespace b200

    //! Subdivides the given grid element extent into blocks, restricted by the device properties.
    //! \param gridElemExtent The full extent of elements in the grid.
    //! \param threadElemExtent the number of elements computed per thread.
    //! \param accDevProps The maxima for the work division.
    //! \param kernelBlockThreadCountMax The maximum number of threads per block for the kernel (0 = device limit).
    //! \param blockThreadMustDivideGridThreadExtent If true, the grid thread extent will be a multiple of the block
    //!     thread extent in every dimension.
    template<typename TDim, typename TIdx>
    [[nodiscard]] ALPAKA_FN_HOST auto subDivideGridElems(
        Vec<TDim, TIdx> const& gridElemExtent,
        Vec<TDim, TIdx> const& threadElemExtent,
        AccDevProps<TDim, TIdx> const& accDevProps,
        TIdx kernelBlockThreadCountMax = static_cast<TIdx>(0u),
        bool blockThreadMustDivideGridThreadExtent = true,
        GridBlockExtentSubDivRestrictions gridBlockExtentSubDivRestrictions = GridBlockExtentSubDivRestrictions::Unrestricted)
        -> WorkDivMembers<TDim, TIdx>
    {
        using V = Vec<TDim, TIdx>;
        if constexpr(TDim::value == 0u)
        {
            return WorkDivMembers<TDim, TIdx>{V{}, V{}, V{}};
        }
        else
        {
            uint64_t ge[8], te[8], gb[8], bt[8], teOut[8];
            for(std::size_t d = 0; d < TDim::value; ++d)
            {
                ge[d] = static_cast<uint64_t>(gridElemExtent[d]);
                te[d] = static_cast<uint64_t>(threadElemExtent[d]);
            }
            auto const props = b200::toAbiProps(accDevProps);
            b200::check(b200_subdivide_grid_elems(
                static_cast<int>(TDim::value),
                ge,
                te,
                &props,
                static_cast<uint64_t>(kernelBlockThreadCountMax),
                blockThreadMustDivideGridThreadExtent ? 1 : 0,
                static_cast<int>(gridBlockExtentSubDivRestrictions),
                gb,
                bt,
                teOut));
            V g, b, t;
            for(std::size_t d = 0; d < TDim::value; ++d)
            {
                g[d] = static_cast<TIdx>(gb[d]);
                b[d] = static_cast<TIdx>(bt[d]);
                t[d] = static_cast<TIdx>(teOut[d]);
            }
            return WorkDivMembers<TDim, TIdx>{g, b, t};
        }
    }

    //! Checks a work division against explicit device properties (and optionally a kernel's block-size limit).
    template<typename TDim, typename TIdx, typename TWorkDiv>
    [[nodiscard]] ALPAKA_FN_HOST auto isValidWorkDiv(
        TWorkDiv const& workDiv,
        AccDevProps<TDim, TIdx> const& accDevProps,
        std::size_t kernelBlockThreadCountMax = 0u) -> bool
    {
        if constexpr(TDim::value == 0u)
        {
            return true;
        }
        else
        {
            uint64_t gb[8], bt[8], te[8];
            auto const g = getWorkDiv<Grid, Blocks>(workDiv);
            auto const b = getWorkDiv<Block, Threads>(workDiv);
            auto const t = getWorkDiv<Thread, Elems>(workDiv);
            for(std::size_t d = 0; d < TDim::value; ++d)
            {
                gb[d] = static_cast<uint64_t>(g[d]);
                bt[d] = static_cast<uint64_t>(b[d]);
                te[d] = static_cast<uint64_t>(t[d]);
            }
            auto const props = b200::toAbiProps(accDevProps);
            int valid = 0;
            b200::check(b200_is_valid_work_div(
                static_cast<int>(TDim::value),
                gb,
                bt,
                te,
                &props,
                static_cast<uint64_t>(kernelBlockThreadCountMax),
                &valid));
            return valid != 0;
        }
    }
} // namespace alpaka
