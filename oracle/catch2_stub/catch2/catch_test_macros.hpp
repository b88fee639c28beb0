/* Stub for the few Catch2 names the reference babelstream driver TU mentions
 * (benchmarks/babelstream/src/babelStreamMainTest.cpp:2,7-9,45,366-368,405,465-478).
 * TEST INFRASTRUCTURE ONLY: lets oracle/ref_babelstream.cpp include that TU verbatim so that the
 * reference kernel functors (InitKernel ... DotKernel) are compiled unmodified into oracle/_ref.
 * The Catch2 test cases themselves become never-instantiated function templates. */
#pragma once
#include <cstdio>
#include <cstdlib>

#define B200_CATCH_STUB_CAT2(a, b) a##b
#define B200_CATCH_STUB_CAT(a, b) B200_CATCH_STUB_CAT2(a, b)

#define REQUIRE(...)                                                                                                  \
    do                                                                                                                \
    {                                                                                                                 \
        if(!(__VA_ARGS__))                                                                                            \
        {                                                                                                             \
            std::fprintf(stderr, "REQUIRE failed: %s (%s:%d)\n", #__VA_ARGS__, __FILE__, __LINE__);                   \
            std::abort();                                                                                             \
        }                                                                                                             \
    } while(0)

#define TEMPLATE_LIST_TEST_CASE(name, tags, list)                                                                     \
    template<typename TestType>                                                                                       \
    [[maybe_unused]] static void B200_CATCH_STUB_CAT(b200_catch_stub_case_, __LINE__)()
