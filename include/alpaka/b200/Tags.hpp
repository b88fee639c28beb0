// include/alpaka/b200/Tags.hpp -- accelerator tags, origin/unit/hierarchy/scope tags, queue properties.
//
// API parity with the reference's acc/Tag.hpp:20-90 (InterfaceTag, AccToTag, TagToAcc, accMatchesTags, AccTags),
// acc/TagAccIsEnabled.hpp:24-35 (EnabledAccTags), core/Positioning.hpp (origin::/unit:: + the un-namespaced aliases),
// atomic/AtomicHierarchy.hpp:26-33 (hierarchy::), mem/fence/Traits.hpp (memory_scope::) and queue/Properties.hpp
// (Blocking / NonBlocking).
//
// There is ONE accelerator in this framework: TagGpuB200. The reference's other tag NAMES are declared so that
// user code that lists them (e.g. `accMatchesTags<Acc, TagGpuCudaRt, TagGpuHipRt, TagGpuSyclIntel>`,
// benchmarks/babelstream/src/babelStreamMainTest.cpp:372) compiles; none of them has an accelerator behind it.
// TagGpuCudaRt is an ALIAS of TagGpuB200: a driver written against the reference's CUDA back-end selects the B200
// back-end without source changes.
#pragma once

#include "Vec.hpp"

#include <concepts>
#include <iostream>
#include <string>
#include <tuple>
#include <type_traits>

namespace alpaka
{
    // ---- where an index/extent is measured from, and in which unit
    namespace origin
    {
        struct Grid
        {
        };
        struct Block
        {
        };
        struct Thread
        {
        };
    } // namespace origin

    namespace unit
    {
        struct Blocks
        {
        };
        struct Threads
        {
        };
        struct Elems
        {
        };
    } // namespace unit

    using namespace origin;
    using namespace unit;

    // ---- atomic hierarchy levels
    namespace hierarchy
    {
        struct Grids
        {
        };
        struct Blocks
        {
        };
        struct Threads
        {
        };
    } // namespace hierarchy

    // ---- memory fence scopes
    namespace memory_scope
    {
        struct Block
        {
        };
        struct Grid
        {
        };
        struct Device
        {
        };
    } // namespace memory_scope

    // ---- queue behaviour
    namespace property
    {
        struct Blocking
        {
        };
        struct NonBlocking
        {
        };
    } // namespace property
    using namespace property;

    // ---- accelerator tags
    struct InterfaceTag
    {
    };

#define ALPAKA_B200_DECLARE_TAG(name)                                                                                 \
    struct name : InterfaceTag                                                                                        \
    {                                                                                                                 \
        static auto get_name() -> std::string                                                                         \
        {                                                                                                             \
            return #name;                                                                                             \
        }                                                                                                             \
    }

    ALPAKA_B200_DECLARE_TAG(TagGpuB200);
    // names of the reference's other back-ends: declared, never enabled
    ALPAKA_B200_DECLARE_TAG(TagCpuSerial);
    ALPAKA_B200_DECLARE_TAG(TagCpuThreads);
    ALPAKA_B200_DECLARE_TAG(TagCpuTbbBlocks);
    ALPAKA_B200_DECLARE_TAG(TagCpuOmp2Blocks);
    ALPAKA_B200_DECLARE_TAG(TagCpuOmp2Threads);
    ALPAKA_B200_DECLARE_TAG(TagCpuSycl);
    ALPAKA_B200_DECLARE_TAG(TagFpgaSyclIntel);
    ALPAKA_B200_DECLARE_TAG(TagGenericSycl);
    ALPAKA_B200_DECLARE_TAG(TagGpuSyclIntel);
    ALPAKA_B200_DECLARE_TAG(TagGpuHipRt);
#undef ALPAKA_B200_DECLARE_TAG

    //! The reference's CUDA tag name resolves to the B200 back-end.
    using TagGpuCudaRt = TagGpuB200;

    namespace concepts
    {
        //! a tag: an InterfaceTag that can be default-constructed and names itself (reference: acc/Tag.hpp:43-53 --
        //! deriving from InterfaceTag alone is not enough, nor is a get_name() on an unrelated type)
        template<typename T>
        concept Tag = std::derived_from<T, InterfaceTag> && std::default_initializable<T> && requires {
            { T::get_name() } -> std::same_as<std::string>;
        };
    } // namespace concepts

    template<typename T>
    inline constexpr bool isTag = concepts::Tag<T>;

    namespace trait
    {
        template<typename TAcc>
        struct AccToTag;

        template<typename TTag, typename TDim, typename TIdx>
        struct TagToAcc;
    } // namespace trait

    template<typename TAcc>
    using AccToTag = typename trait::AccToTag<TAcc>::type;

    template<concepts::Tag TTag, typename TDim, typename TIdx>
    using TagToAcc = typename trait::TagToAcc<TTag, TDim, TIdx>::type;

    template<typename TAcc, concepts::Tag... TTag>
    inline constexpr bool accMatchesTags = (std::is_same_v<AccToTag<TAcc>, TTag> || ...);

    //! every tag name that exists
    using AccTags = std::tuple<
        TagCpuSerial,
        TagCpuThreads,
        TagCpuTbbBlocks,
        TagCpuOmp2Blocks,
        TagCpuOmp2Threads,
        TagGpuB200,
        TagGpuHipRt,
        TagCpuSycl,
        TagFpgaSyclIntel,
        TagGpuSyclIntel>;

    //! the tags that have an accelerator in this build: exactly one
    using EnabledAccTags = std::tuple<TagGpuB200>;

    //! true iff an accelerator stands behind the tag, i.e. trait::TagToAcc maps it (reference: acc/TagAccIsEnabled.hpp:24-35)
    template<concepts::Tag TTag>
    struct AccIsEnabled : std::bool_constant<requires { typename trait::TagToAcc<TTag, DimInt<1u>, int>::type; }>
    {
    };

    namespace detail
    {
        template<typename TTuple>
        struct PrintTagNames;
        template<typename... TTags>
        struct PrintTagNames<std::tuple<TTags...>>
        {
            static void print()
            {
                bool first = true;
                ((std::cout << (first ? "" : ", ") << TTags::get_name(), first = false), ...);
                std::cout << std::endl;
            }
        };
    } // namespace detail

    //! prints "TagA, TagB" for a std::tuple of tags (reference: acc/TagAccIsEnabled.hpp / example helpers)
    template<typename TTuple>
    void printTagNames()
    {
        detail::PrintTagNames<TTuple>::print();
    }

    //! calls `callable(tag)` once per enabled tag and returns the disjunction of the results -- a driver's
    //! `return executeForEachAccTag(...)` is therefore EXIT_SUCCESS only if every run returned 0
    //! (reference: example/ExecuteForEachAccTag.hpp:19-26)
    template<typename TCallable>
    inline auto executeForEachAccTag(TCallable&& callable)
    {
        return std::apply([&](auto const&... tags) { return (callable(tags) || ...); }, EnabledAccTags{});
    }
} // namespace alpaka
