// include/alpaka/b200/Exec.hpp -- element / group index ranges for kernels.
//
// API parity with the reference's include/alpaka/exec/{ElementIndex,UniformElements,IndependentElements,Once}.hpp:
//   uniformElements[Along<D>|AlongX|Y|Z](acc [, first], extent)      UniformElements.hpp:258-319
//   uniformElementsND(acc [, extentVec])                             UniformElements.hpp:592-608
//   uniformGroups[Along<D>|AlongX|Y|Z](acc [, elements])             UniformElements.hpp:815-876
//   uniformGroupElements[Along<D>|AlongX|Y|Z](acc, group [, extent]) UniformElements.hpp:1082-1143
//   independentGroups[Along...](acc [, groups])                      IndependentElements.hpp:139-199
//   independentGroupElements[Along...](acc [, first], extent)        IndependentElements.hpp:372-433
//   oncePerGrid(acc), oncePerBlock(acc)                              Once.hpp:27-54
//   ElementIndex<TIdx>{global, local}                                ElementIndex.hpp:12-17
// and the class names in alpaka::detail (UniformElementsAlong<TAcc,D>, UniformElementsND<TAcc>, ...) that user code
// may spell out.
//
// Design. Every 1-D range above is one of three index walks, so three small value types carry all of them:
//   RunHopRange   runs of `run` consecutive indices whose starts are `pitch` apart, clipped to `extent`
//                 (grid-strided elements, block-strided elements);
//   HopRange      start, start+pitch, ... clipped to `extent` (grid-strided groups);
//   GroupRange    consecutive local indices of one group, yielding {global, local} pairs.
// Iterators compare by position only and saturate at `extent`, which makes `it != end()` a single integer compare in
// SASS and lets nvcc turn the common one-element-per-thread case into a plain grid-stride loop: this is the form
// through which a user kernel gets the multi-element-per-thread shape of the hand-written stream kernels
// (b200_stream.cu) without writing index arithmetic.
// The N-dimensional range walks the thread's element box with an odometer and then hops the box by the grid pitch,
// also as an odometer (last dimension fastest), as UniformElements.hpp:545-583 specifies.
#pragma once

#include "Acc.hpp"

#include <cstddef>
#include <type_traits>

namespace alpaka
{
    //! The index of an element along one dimension: within the whole problem and within its group.
    template<typename TIdx>
    struct ElementIndex
    {
        TIdx global;
        TIdx local;
    };

    namespace b200x
    {
        template<typename TIdx>
        ALPAKA_FN_HOST_ACC constexpr TIdx clipTo(TIdx v, TIdx hi)
        {
            return v < hi ? v : hi;
        }

        //! Runs of `run` consecutive indices; run starts are `pitch` apart; nothing at or beyond `extent`.
        template<typename TIdx>
        class RunHopRange
        {
        public:
            class const_iterator
            {
            public:
                ALPAKA_FN_HOST_ACC constexpr const_iterator(TIdx pos, TIdx run, TIdx pitch, TIdx extent)
                    : m_pos{clipTo(pos, extent)}
                    , m_left{run}
                    , m_run{run}
                    , m_gap{pitch - run}
                    , m_extent{extent}
                {
                }

                ALPAKA_FN_HOST_ACC constexpr TIdx operator*() const
                {
                    return m_pos;
                }

                ALPAKA_FN_HOST_ACC constexpr const_iterator& operator++()
                {
                    ++m_pos;
                    if(--m_left == 0)
                    {
                        m_left = m_run;
                        m_pos += m_gap;
                    }
                    m_pos = clipTo(m_pos, m_extent);
                    return *this;
                }

                ALPAKA_FN_HOST_ACC constexpr const_iterator operator++(int)
                {
                    auto const before = *this;
                    ++*this;
                    return before;
                }

                ALPAKA_FN_HOST_ACC friend constexpr bool operator==(const_iterator const& a, const_iterator const& b)
                {
                    return a.m_pos == b.m_pos;
                }

                ALPAKA_FN_HOST_ACC friend constexpr bool operator!=(const_iterator const& a, const_iterator const& b)
                {
                    return a.m_pos != b.m_pos;
                }

            private:
                TIdx m_pos;
                TIdx m_left; // indices left in the current run, including the current one
                TIdx m_run;
                TIdx m_gap;
                TIdx m_extent;
            };

            using iterator = const_iterator;

            ALPAKA_FN_HOST_ACC constexpr RunHopRange(TIdx start, TIdx run, TIdx pitch, TIdx extent)
                : m_start{start}
                , m_run{run}
                , m_pitch{pitch}
                , m_extent{extent}
            {
            }

            ALPAKA_FN_HOST_ACC constexpr const_iterator begin() const
            {
                return {m_start, m_run, m_pitch, m_extent};
            }

            ALPAKA_FN_HOST_ACC constexpr const_iterator end() const
            {
                return {m_extent, m_run, m_pitch, m_extent};
            }

        private:
            TIdx m_start, m_run, m_pitch, m_extent;
        };

        //! start, start + pitch, ... below `extent`.
        template<typename TIdx>
        class HopRange
        {
        public:
            class const_iterator
            {
            public:
                ALPAKA_FN_HOST_ACC constexpr const_iterator(TIdx pos, TIdx pitch, TIdx extent)
                    : m_pos{clipTo(pos, extent)}
                    , m_pitch{pitch}
                    , m_extent{extent}
                {
                }

                ALPAKA_FN_HOST_ACC constexpr TIdx operator*() const
                {
                    return m_pos;
                }

                ALPAKA_FN_HOST_ACC constexpr const_iterator& operator++()
                {
                    m_pos = clipTo(static_cast<TIdx>(m_pos + m_pitch), m_extent);
                    return *this;
                }

                ALPAKA_FN_HOST_ACC constexpr const_iterator operator++(int)
                {
                    auto const before = *this;
                    ++*this;
                    return before;
                }

                ALPAKA_FN_HOST_ACC friend constexpr bool operator==(const_iterator const& a, const_iterator const& b)
                {
                    return a.m_pos == b.m_pos;
                }

                ALPAKA_FN_HOST_ACC friend constexpr bool operator!=(const_iterator const& a, const_iterator const& b)
                {
                    return a.m_pos != b.m_pos;
                }

            private:
                TIdx m_pos, m_pitch, m_extent;
            };

            using iterator = const_iterator;

            ALPAKA_FN_HOST_ACC constexpr HopRange(TIdx start, TIdx pitch, TIdx extent)
                : m_start{start}
                , m_pitch{pitch}
                , m_extent{extent}
            {
            }

            ALPAKA_FN_HOST_ACC constexpr const_iterator begin() const
            {
                return {m_start, m_pitch, m_extent};
            }

            ALPAKA_FN_HOST_ACC constexpr const_iterator end() const
            {
                return {m_extent, m_pitch, m_extent};
            }

        private:
            TIdx m_start, m_pitch, m_extent;
        };

        //! Local indices [lo, hi) of the group whose first element is `origin`; yields {origin + i, i}.
        template<typename TIdx>
        class GroupRange
        {
        public:
            class const_iterator
            {
            public:
                ALPAKA_FN_HOST_ACC constexpr const_iterator(TIdx local, TIdx origin, TIdx hi)
                    : m_local{local}
                    , m_origin{origin}
                    , m_hi{hi}
                {
                }

                ALPAKA_FN_HOST_ACC constexpr ElementIndex<TIdx> operator*() const
                {
                    return {static_cast<TIdx>(m_origin + m_local), m_local};
                }

                ALPAKA_FN_HOST_ACC constexpr const_iterator& operator++()
                {
                    m_local = clipTo(static_cast<TIdx>(m_local + 1), m_hi);
                    return *this;
                }

                ALPAKA_FN_HOST_ACC constexpr const_iterator operator++(int)
                {
                    auto const before = *this;
                    ++*this;
                    return before;
                }

                ALPAKA_FN_HOST_ACC friend constexpr bool operator==(const_iterator const& a, const_iterator const& b)
                {
                    return a.m_local == b.m_local;
                }

                ALPAKA_FN_HOST_ACC friend constexpr bool operator!=(const_iterator const& a, const_iterator const& b)
                {
                    return a.m_local != b.m_local;
                }

            private:
                TIdx m_local, m_origin, m_hi;
            };

            using iterator = const_iterator;

            ALPAKA_FN_HOST_ACC constexpr GroupRange(TIdx origin, TIdx lo, TIdx hi) : m_origin{origin}, m_lo{lo}, m_hi{hi}
            {
            }

            ALPAKA_FN_HOST_ACC constexpr const_iterator begin() const
            {
                return {m_lo, m_origin, m_hi};
            }

            ALPAKA_FN_HOST_ACC constexpr const_iterator end() const
            {
                return {m_hi, m_origin, m_hi};
            }

        private:
            TIdx m_origin, m_lo, m_hi;
        };

        template<typename TAcc, std::size_t D>
        inline constexpr bool accHasDim = isAccelerator<TAcc> && (alpaka::Dim<TAcc>::value >= D);
    } // namespace b200x

    namespace detail
    {
        //! Elements [first, extent) along dimension D, shared uniformly by all threads of the grid:
        //! thread t visits [t*e, t*e+e), then the same run one grid pitch further on, ...
        template<typename TAcc, std::size_t D, typename = std::enable_if_t<b200x::accHasDim<TAcc, D>>>
        class UniformElementsAlong : public b200x::RunHopRange<alpaka::Idx<TAcc>>
        {
            using Base = b200x::RunHopRange<alpaka::Idx<TAcc>>;

            ALPAKA_FN_ACC static auto perThread(TAcc const& acc)
            {
                return getWorkDiv<Thread, Elems>(acc)[D];
            }

            ALPAKA_FN_ACC static auto pitch(TAcc const& acc)
            {
                return getWorkDiv<Grid, Threads>(acc)[D] * perThread(acc);
            }

            ALPAKA_FN_ACC static auto origin(TAcc const& acc)
            {
                return getIdx<Grid, Threads>(acc)[D] * perThread(acc);
            }

        public:
            using Idx = alpaka::Idx<TAcc>;

            ALPAKA_FN_ACC explicit UniformElementsAlong(TAcc const& acc)
                : Base{origin(acc), perThread(acc), pitch(acc), pitch(acc)}
            {
            }

            ALPAKA_FN_ACC UniformElementsAlong(TAcc const& acc, Idx extent)
                : Base{origin(acc), perThread(acc), pitch(acc), extent}
            {
            }

            ALPAKA_FN_ACC UniformElementsAlong(TAcc const& acc, Idx first, Idx extent)
                : Base{static_cast<Idx>(origin(acc) + first), perThread(acc), pitch(acc), extent}
            {
            }
        };

        //! Groups (blocks' worth of elements) needed to cover `elements` elements along D, grid-strided over blocks.
        template<typename TAcc, std::size_t D, typename = std::enable_if_t<b200x::accHasDim<TAcc, D>>>
        class UniformGroupsAlong : public b200x::HopRange<alpaka::Idx<TAcc>>
        {
            using Base = b200x::HopRange<alpaka::Idx<TAcc>>;

        public:
            using Idx = alpaka::Idx<TAcc>;

            ALPAKA_FN_ACC explicit UniformGroupsAlong(TAcc const& acc)
                : Base{getIdx<Grid, Blocks>(acc)[D], getWorkDiv<Grid, Blocks>(acc)[D], getWorkDiv<Grid, Blocks>(acc)[D]}
            {
            }

            ALPAKA_FN_ACC UniformGroupsAlong(TAcc const& acc, Idx elements)
                : Base{
                      getIdx<Grid, Blocks>(acc)[D],
                      getWorkDiv<Grid, Blocks>(acc)[D],
                      core::divCeil(elements, getWorkDiv<Block, Elems>(acc)[D])}
            {
            }
        };

        //! The elements of group `group` that belong to the calling thread, as {global, local} pairs.
        template<typename TAcc, std::size_t D, typename = std::enable_if_t<b200x::accHasDim<TAcc, D>>>
        class UniformGroupElementsAlong : public b200x::GroupRange<alpaka::Idx<TAcc>>
        {
            using Base = b200x::GroupRange<alpaka::Idx<TAcc>>;

            ALPAKA_FN_ACC static auto groupOrigin(TAcc const& acc, alpaka::Idx<TAcc> group)
            {
                return static_cast<alpaka::Idx<TAcc>>(group * getWorkDiv<Block, Elems>(acc)[D]);
            }

            ALPAKA_FN_ACC static auto localLo(TAcc const& acc)
            {
                return static_cast<alpaka::Idx<TAcc>>(getIdx<Block, Threads>(acc)[D] * getWorkDiv<Thread, Elems>(acc)[D]);
            }

            ALPAKA_FN_ACC static auto localHi(TAcc const& acc)
            {
                return static_cast<alpaka::Idx<TAcc>>(localLo(acc) + getWorkDiv<Thread, Elems>(acc)[D]);
            }

        public:
            using Idx = alpaka::Idx<TAcc>;

            ALPAKA_FN_ACC UniformGroupElementsAlong(TAcc const& acc, Idx group)
                : Base{groupOrigin(acc, group), localLo(acc), localHi(acc)}
            {
            }

            ALPAKA_FN_ACC UniformGroupElementsAlong(TAcc const& acc, Idx group, Idx extent)
                : Base{
                      groupOrigin(acc, group),
                      b200x::clipTo(localLo(acc), static_cast<Idx>(extent - groupOrigin(acc, group))),
                      b200x::clipTo(localHi(acc), static_cast<Idx>(extent - groupOrigin(acc, group)))}
            {
            }
        };

        //! Groups [0, groups) along D that the blocks of the grid process independently of each other.
        template<typename TAcc, std::size_t D, typename = std::enable_if_t<b200x::accHasDim<TAcc, D>>>
        class IndependentGroupsAlong : public b200x::HopRange<alpaka::Idx<TAcc>>
        {
            using Base = b200x::HopRange<alpaka::Idx<TAcc>>;

        public:
            using Idx = alpaka::Idx<TAcc>;

            ALPAKA_FN_ACC explicit IndependentGroupsAlong(TAcc const& acc)
                : Base{getIdx<Grid, Blocks>(acc)[D], getWorkDiv<Grid, Blocks>(acc)[D], getWorkDiv<Grid, Blocks>(acc)[D]}
            {
            }

            ALPAKA_FN_ACC IndependentGroupsAlong(TAcc const& acc, Idx groups)
                : Base{getIdx<Grid, Blocks>(acc)[D], getWorkDiv<Grid, Blocks>(acc)[D], groups}
            {
            }
        };

        //! Elements [first, extent) along D shared by the threads of ONE block (block-strided).
        template<typename TAcc, std::size_t D, typename = std::enable_if_t<b200x::accHasDim<TAcc, D>>>
        class IndependentGroupElementsAlong : public b200x::RunHopRange<alpaka::Idx<TAcc>>
        {
            using Base = b200x::RunHopRange<alpaka::Idx<TAcc>>;

            ALPAKA_FN_ACC static auto perThread(TAcc const& acc)
            {
                return getWorkDiv<Thread, Elems>(acc)[D];
            }

            ALPAKA_FN_ACC static auto pitch(TAcc const& acc)
            {
                return getWorkDiv<Block, Threads>(acc)[D] * perThread(acc);
            }

            ALPAKA_FN_ACC static auto origin(TAcc const& acc)
            {
                return getIdx<Block, Threads>(acc)[D] * perThread(acc);
            }

        public:
            using Idx = alpaka::Idx<TAcc>;

            ALPAKA_FN_ACC explicit IndependentGroupElementsAlong(TAcc const& acc)
                : Base{origin(acc), perThread(acc), pitch(acc), pitch(acc)}
            {
            }

            ALPAKA_FN_ACC IndependentGroupElementsAlong(TAcc const& acc, Idx extent)
                : Base{origin(acc), perThread(acc), pitch(acc), extent}
            {
            }

            ALPAKA_FN_ACC IndependentGroupElementsAlong(TAcc const& acc, Idx first, Idx extent)
                : Base{static_cast<Idx>(origin(acc) + first), perThread(acc), pitch(acc), extent}
            {
            }
        };

        //! All N-dimensional element indices below `extent`, shared uniformly by the threads of the grid.
        template<typename TAcc, typename = std::enable_if_t<isAccelerator<TAcc> && (alpaka::Dim<TAcc>::value > 0)>>
        class UniformElementsND
        {
        public:
            using Dim = alpaka::Dim<TAcc>;
            using Idx = alpaka::Idx<TAcc>;
            using Vec = alpaka::Vec<Dim, Idx>;

            ALPAKA_FN_ACC explicit UniformElementsND(TAcc const& acc)
                : m_box{getWorkDiv<Thread, Elems>(acc)}
                , m_home{getIdx<Grid, Threads>(acc) * m_box}
                , m_pitch{getWorkDiv<Grid, Threads>(acc) * m_box}
                , m_extent{m_pitch}
            {
            }

            ALPAKA_FN_ACC UniformElementsND(TAcc const& acc, Vec extent)
                : m_box{getWorkDiv<Thread, Elems>(acc)}
                , m_home{getIdx<Grid, Threads>(acc) * m_box}
                , m_pitch{getWorkDiv<Grid, Threads>(acc) * m_box}
                , m_extent{extent}
            {
            }

            class const_iterator
            {
                friend class UniformElementsND;
                static constexpr std::size_t N = Dim::value;

                //! Positioned at the first element of the box whose corner is `corner` (all corner[d] < extent[d]).
                ALPAKA_FN_ACC const_iterator(UniformElementsND const* range, Vec corner)
                    : m_range{range}
                    , m_lo{corner}
                    , m_hi{corner}
                    , m_at{corner}
                {
                    for(std::size_t d = 0; d < N; ++d)
                        m_hi[d] = upper(d);
                }

                //! One past the last element: the extent itself.
                ALPAKA_FN_ACC explicit const_iterator(UniformElementsND const* range)
                    : m_range{range}
                    , m_lo{range->m_extent}
                    , m_hi{range->m_extent}
                    , m_at{range->m_extent}
                {
                }

                ALPAKA_FN_ACC Idx upper(std::size_t d) const
                {
                    return b200x::clipTo(static_cast<Idx>(m_lo[d] + m_range->m_box[d]), m_range->m_extent[d]);
                }

                ALPAKA_FN_ACC void advance()
                {
                    // odometer over the thread's current box, last dimension fastest
                    for(std::size_t d = N; d-- > 0;)
                    {
                        if(++m_at[d] < m_hi[d])
                            return;
                        m_at[d] = m_lo[d];
                    }
                    // box exhausted: hop it by the grid pitch, again as an odometer
                    for(std::size_t d = N; d-- > 0;)
                    {
                        m_lo[d] += m_range->m_pitch[d];
                        bool const wrapped = !(m_lo[d] < m_range->m_extent[d]);
                        if(wrapped)
                            m_lo[d] = m_range->m_home[d];
                        m_at[d] = m_lo[d];
                        m_hi[d] = upper(d);
                        if(!wrapped)
                            return;
                    }
                    // every dimension wrapped: the walk is over
                    m_lo = m_hi = m_at = m_range->m_extent;
                }

            public:
                ALPAKA_FN_ACC Vec operator*() const
                {
                    return m_at;
                }

                ALPAKA_FN_ACC const_iterator& operator++()
                {
                    advance();
                    return *this;
                }

                ALPAKA_FN_ACC const_iterator operator++(int)
                {
                    auto const before = *this;
                    advance();
                    return before;
                }

                ALPAKA_FN_ACC friend bool operator==(const_iterator const& a, const_iterator const& b)
                {
                    return a.m_at == b.m_at;
                }

                ALPAKA_FN_ACC friend bool operator!=(const_iterator const& a, const_iterator const& b)
                {
                    return !(a.m_at == b.m_at);
                }

            private:
                UniformElementsND const* m_range;
                Vec m_lo; // corner of the current box
                Vec m_hi; // one past its last element, clipped to the extent
                Vec m_at; // current element
            };

            using iterator = const_iterator;

            ALPAKA_FN_ACC const_iterator begin() const
            {
                for(std::size_t d = 0; d < Dim::value; ++d)
                    if(!(m_home[d] < m_extent[d]))
                        return const_iterator{this};
                return const_iterator{this, m_home};
            }

            ALPAKA_FN_ACC const_iterator end() const
            {
                return const_iterator{this};
            }

        private:
            Vec m_box; // elements per thread
            Vec m_home; // first element of this thread's first box
            Vec m_pitch; // elements covered by the whole grid in one pass
            Vec m_extent;
        };
    } // namespace detail

    // ---- factory functions: the spelling user kernels use ----------------------------------------------------------

#define ALPAKA_B200_RANGE_FACTORIES(fn, Cls)                                                                            \
    template<typename TAcc, typename... TArgs, typename = std::enable_if_t<isAccelerator<TAcc> && alpaka::Dim<TAcc>::value == 1>> \
    ALPAKA_FN_ACC inline auto fn(TAcc const& acc, TArgs... args)                                                        \
    {                                                                                                                   \
        return detail::Cls<TAcc, 0>(acc, static_cast<alpaka::Idx<TAcc>>(args)...);                                      \
    }                                                                                                                   \
    template<std::size_t D, typename TAcc, typename... TArgs, typename = std::enable_if_t<b200x::accHasDim<TAcc, D>>>   \
    ALPAKA_FN_ACC inline auto fn##Along(TAcc const& acc, TArgs... args)                                                 \
    {                                                                                                                   \
        return detail::Cls<TAcc, D>(acc, static_cast<alpaka::Idx<TAcc>>(args)...);                                      \
    }                                                                                                                   \
    template<typename TAcc, typename... TArgs, typename = std::enable_if_t<isAccelerator<TAcc> && (alpaka::Dim<TAcc>::value > 0)>> \
    ALPAKA_FN_ACC inline auto fn##AlongX(TAcc const& acc, TArgs... args)                                                \
    {                                                                                                                   \
        return detail::Cls<TAcc, alpaka::Dim<TAcc>::value - 1>(acc, static_cast<alpaka::Idx<TAcc>>(args)...);           \
    }                                                                                                                   \
    template<typename TAcc, typename... TArgs, typename = std::enable_if_t<isAccelerator<TAcc> && (alpaka::Dim<TAcc>::value > 1)>> \
    ALPAKA_FN_ACC inline auto fn##AlongY(TAcc const& acc, TArgs... args)                                                \
    {                                                                                                                   \
        return detail::Cls<TAcc, alpaka::Dim<TAcc>::value - 2>(acc, static_cast<alpaka::Idx<TAcc>>(args)...);           \
    }                                                                                                                   \
    template<typename TAcc, typename... TArgs, typename = std::enable_if_t<isAccelerator<TAcc> && (alpaka::Dim<TAcc>::value > 2)>> \
    ALPAKA_FN_ACC inline auto fn##AlongZ(TAcc const& acc, TArgs... args)                                                \
    {                                                                                                                   \
        return detail::Cls<TAcc, alpaka::Dim<TAcc>::value - 3>(acc, static_cast<alpaka::Idx<TAcc>>(args)...);           \
    }

    ALPAKA_B200_RANGE_FACTORIES(uniformElements, UniformElementsAlong)
    ALPAKA_B200_RANGE_FACTORIES(uniformGroups, UniformGroupsAlong)
    ALPAKA_B200_RANGE_FACTORIES(uniformGroupElements, UniformGroupElementsAlong)
    ALPAKA_B200_RANGE_FACTORIES(independentGroups, IndependentGroupsAlong)
    ALPAKA_B200_RANGE_FACTORIES(independentGroupElements, IndependentGroupElementsAlong)
#undef ALPAKA_B200_RANGE_FACTORIES

    template<typename TAcc, typename = std::enable_if_t<isAccelerator<TAcc> && (alpaka::Dim<TAcc>::value > 0)>>
    ALPAKA_FN_ACC inline auto uniformElementsND(TAcc const& acc)
    {
        return detail::UniformElementsND<TAcc>(acc);
    }

    template<typename TAcc, typename = std::enable_if_t<isAccelerator<TAcc> && (alpaka::Dim<TAcc>::value > 0)>>
    ALPAKA_FN_ACC inline auto uniformElementsND(TAcc const& acc, alpaka::Vec<alpaka::Dim<TAcc>, alpaka::Idx<TAcc>> extent)
    {
        return detail::UniformElementsND<TAcc>(acc, extent);
    }

    //! True in exactly one thread of the grid.
    template<typename TAcc, typename = std::enable_if_t<isAccelerator<TAcc>>>
    ALPAKA_FN_ACC inline bool oncePerGrid(TAcc const& acc)
    {
        return getIdx<Grid, Threads>(acc) == alpaka::Vec<alpaka::Dim<TAcc>, alpaka::Idx<TAcc>>::zeros();
    }

    //! True in exactly one thread of every block.
    template<typename TAcc, typename = std::enable_if_t<isAccelerator<TAcc>>>
    ALPAKA_FN_ACC inline bool oncePerBlock(TAcc const& acc)
    {
        return getIdx<Block, Threads>(acc) == alpaka::Vec<alpaka::Dim<TAcc>, alpaka::Idx<TAcc>>::zeros();
    }
} // namespace alpaka
