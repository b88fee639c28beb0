// include/alpaka/b200/Config.hpp -- compile-time configuration and function-attribute macros of the B200 back-end.
//
// API parity with the reference's include/alpaka/core/Common.hpp:29-221 (ALPAKA_FN_*), core/Unroll.hpp:16-24,
// core/Assert.hpp, core/Debug.hpp:13-60 and version.hpp:9-11 -- same macro names and meaning, written fresh for a
// single target: nvcc >= 12.8 compiling for sm_100a. There is exactly one accelerator here (AccGpuB200, which also
// answers to the reference's AccGpuCudaRt / TagGpuCudaRt names); no other back-end and no CPU fallback exist, so
// the reference's per-back-end ALPAKA_ACC_*_ENABLED matrix collapses to the two macros defined below.
#pragma once

// standard headers the reference's umbrella header brings in transitively and user code relies on (e.g. std::iota in
// example/convolution2D/src/convolution2D.cpp:253 with no <numeric> of its own)
#include <algorithm>
#include <cassert>
#include <functional>
#include <numeric>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <stdexcept>

#define ALPAKA_VERSION_MAJOR 2
#define ALPAKA_VERSION_MINOR 0
#define ALPAKA_VERSION_PATCH 0
#define ALPAKA_B200 1

// The one enabled accelerator. Reference drivers key GPU-only code on ALPAKA_ACC_GPU_CUDA_ENABLED
// (e.g. example/reduce/src/alpakaConfig.hpp:96) so that name is defined as well.
#ifndef ALPAKA_ACC_GPU_B200_ENABLED
#    define ALPAKA_ACC_GPU_B200_ENABLED
#endif
#ifndef ALPAKA_ACC_GPU_CUDA_ENABLED
#    define ALPAKA_ACC_GPU_CUDA_ENABLED
#endif

#ifndef ALPAKA_DEBUG
#    define ALPAKA_DEBUG 0
#endif
#define ALPAKA_DEBUG_DISABLED 0
#define ALPAKA_DEBUG_MINIMAL 1
#define ALPAKA_DEBUG_FULL 2

#if defined(__CUDACC__)
#    define ALPAKA_FN_ACC __device__
#    define ALPAKA_FN_HOST_ACC __host__ __device__
#    define ALPAKA_FN_HOST __host__
#    define ALPAKA_FN_INLINE __forceinline__
#    if defined(__NVCC__)
#        define ALPAKA_NO_HOST_ACC_WARNING _Pragma("nv_exec_check_disable")
#    else
#        define ALPAKA_NO_HOST_ACC_WARNING
#    endif
// Device-global variable templates (see Global.hpp). The template parameter is spelled TAcc because user code names
// it inside the declaration: `ALPAKA_STATIC_ACC_MEM_GLOBAL alpaka::DevGlobal<TAcc, int> g;` (reference:
// core/Common.hpp:136-206). Without relocatable device code CUDA requires internal linkage for such variables.
#    if defined(__CUDACC_RDC__)
#        define ALPAKA_STATIC_ACC_MEM_GLOBAL                                                                          \
            template<typename TAcc>                                                                                   \
            __device__ inline
#        define ALPAKA_STATIC_ACC_MEM_CONSTANT                                                                        \
            template<typename TAcc>                                                                                   \
            __constant__ inline
#    else
#        define ALPAKA_STATIC_ACC_MEM_GLOBAL                                                                          \
            template<typename TAcc>                                                                                   \
            __device__ static
#        define ALPAKA_STATIC_ACC_MEM_CONSTANT                                                                        \
            template<typename TAcc>                                                                                   \
            __constant__ static
#    endif
#else
#    define ALPAKA_FN_ACC
#    define ALPAKA_FN_HOST_ACC
#    define ALPAKA_FN_HOST
#    define ALPAKA_FN_INLINE inline __attribute__((always_inline))
#    define ALPAKA_NO_HOST_ACC_WARNING
#endif

#define ALPAKA_FN_EXTERN extern

// Compiler/language identification in the Boost.Predef vocabulary the reference (and code written against it) tests
// with `#if BOOST_COMP_NVCC` etc. (core/BoostPredef.hpp). Boost itself is not a dependency here: only these names,
// with Boost's value encoding (major * 10'000'000 + minor * 100'000 + patch, 0 = not this compiler).
#ifndef BOOST_VERSION_NUMBER
#    define BOOST_VERSION_NUMBER(major, minor, patch)                                                                \
        ((((major) % 100) * 10000000) + (((minor) % 100) * 100000) + ((patch) % 100000))
#    define BOOST_VERSION_NUMBER_NOT_AVAILABLE 0
#    define BOOST_VERSION_NUMBER_AVAILABLE BOOST_VERSION_NUMBER(0, 0, 1)
#endif
#ifndef BOOST_LANG_CUDA
#    if defined(__CUDACC__)
#        define BOOST_LANG_CUDA BOOST_VERSION_NUMBER(__CUDACC_VER_MAJOR__, __CUDACC_VER_MINOR__, __CUDACC_VER_BUILD__)
#    else
#        define BOOST_LANG_CUDA BOOST_VERSION_NUMBER_NOT_AVAILABLE
#    endif
#endif
#ifndef BOOST_LANG_HIP
#    define BOOST_LANG_HIP BOOST_VERSION_NUMBER_NOT_AVAILABLE
#endif
#ifndef BOOST_COMP_NVCC
#    if defined(__NVCC__)
#        define BOOST_COMP_NVCC BOOST_VERSION_NUMBER(__CUDACC_VER_MAJOR__, __CUDACC_VER_MINOR__, __CUDACC_VER_BUILD__)
#    else
#        define BOOST_COMP_NVCC BOOST_VERSION_NUMBER_NOT_AVAILABLE
#    endif
#endif
#ifndef BOOST_COMP_GNUC
#    if defined(__GNUC__) && !defined(__clang__)
#        define BOOST_COMP_GNUC BOOST_VERSION_NUMBER(__GNUC__, __GNUC_MINOR__, __GNUC_PATCHLEVEL__)
#    else
#        define BOOST_COMP_GNUC BOOST_VERSION_NUMBER_NOT_AVAILABLE
#    endif
#endif
#ifndef BOOST_COMP_CLANG
#    if defined(__clang__)
#        define BOOST_COMP_CLANG BOOST_VERSION_NUMBER(__clang_major__, __clang_minor__, __clang_patchlevel__)
#    else
#        define BOOST_COMP_CLANG BOOST_VERSION_NUMBER_NOT_AVAILABLE
#    endif
#endif
#ifndef BOOST_COMP_CLANG_CUDA
#    define BOOST_COMP_CLANG_CUDA BOOST_VERSION_NUMBER_NOT_AVAILABLE
#endif
#ifndef BOOST_COMP_MSVC
#    define BOOST_COMP_MSVC BOOST_VERSION_NUMBER_NOT_AVAILABLE
#endif
#ifndef BOOST_COMP_MSVC_EMULATED
#    define BOOST_COMP_MSVC_EMULATED BOOST_VERSION_NUMBER_NOT_AVAILABLE
#endif
#ifndef BOOST_COMP_PGI
#    define BOOST_COMP_PGI BOOST_VERSION_NUMBER_NOT_AVAILABLE
#endif
#ifndef BOOST_COMP_HIP
#    define BOOST_COMP_HIP BOOST_VERSION_NUMBER_NOT_AVAILABLE
#endif
#ifndef BOOST_ARCH_PTX
#    if defined(__CUDA_ARCH__)
#        define BOOST_ARCH_PTX BOOST_VERSION_NUMBER(__CUDA_ARCH__ / 100, (__CUDA_ARCH__ % 100) / 10, 0)
#    else
#        define BOOST_ARCH_PTX BOOST_VERSION_NUMBER_NOT_AVAILABLE
#    endif
#endif
#ifndef BOOST_ARCH_HSA
#    define BOOST_ARCH_HSA BOOST_VERSION_NUMBER_NOT_AVAILABLE
#endif

// ALPAKA_UNROLL(n) / ALPAKA_UNROLL(): loop unrolling hint placed in front of a loop.
#define ALPAKA_B200_PRAGMA(x) _Pragma(#x)
#if defined(__CUDACC__)
#    define ALPAKA_UNROLL(...) ALPAKA_B200_PRAGMA(unroll __VA_ARGS__)
#else
#    define ALPAKA_UNROLL(...) ALPAKA_B200_PRAGMA(GCC unroll 8)
#endif

#define ALPAKA_ASSERT(...) assert((__VA_ARGS__))
#if defined(__CUDA_ARCH__)
#    define ALPAKA_ASSERT_ACC(...) assert((__VA_ARGS__))
#else
#    define ALPAKA_ASSERT_ACC(...) assert((__VA_ARGS__))
#endif
#define ALPAKA_ASSERT_OFFLOAD(...) ALPAKA_ASSERT_ACC(__VA_ARGS__)

#if defined(__CUDA_ARCH__)
#    define ALPAKA_UNREACHABLE(...) __builtin_unreachable()
#else
#    define ALPAKA_UNREACHABLE(...) __builtin_unreachable()
#endif

#define ALPAKA_DEVICE_VOLATILE volatile

// ALPAKA_THROW_ACC(msg): user-defined fatal error inside a kernel (reference: core/RuntimeMacros.hpp:21-51). On the
// device it prints the message and traps; the host sees cudaErrorLaunchFailure as std::runtime_error at the next
// wait(). In host code it throws std::runtime_error directly.
#if defined(__CUDA_ARCH__)
#    define ALPAKA_THROW_ACC(MSG)                                                                                     \
        {                                                                                                             \
            printf("alpaka encountered a user-defined error condition while running on the B200 back-end:\n%s", (MSG)); \
            __trap();                                                                                                 \
        }
#else
#    define ALPAKA_THROW_ACC(MSG)                                                                                     \
        {                                                                                                             \
            printf("alpaka encountered a user-defined error condition:\n%s", (MSG));                                  \
            throw std::runtime_error(MSG);                                                                            \
        }
#endif

// scope logging of the reference (core/Debug.hpp) is a no-op unless ALPAKA_DEBUG >= 2
#if ALPAKA_DEBUG >= ALPAKA_DEBUG_FULL
#    include <iostream>
namespace alpaka::core::detail
{
    struct ScopeLog
    {
        char const* m_name;
        explicit ScopeLog(char const* n) : m_name(n)
        {
            std::cout << "[+] " << m_name << std::endl;
        }
        ~ScopeLog()
        {
            std::cout << "[-] " << m_name << std::endl;
        }
    };
} // namespace alpaka::core::detail
#    define ALPAKA_DEBUG_FULL_LOG_SCOPE ::alpaka::core::detail::ScopeLog const alpakaScopeLog_(__func__)
#else
#    define ALPAKA_DEBUG_FULL_LOG_SCOPE
#endif
#define ALPAKA_DEBUG_MINIMAL_LOG_SCOPE ALPAKA_DEBUG_FULL_LOG_SCOPE
