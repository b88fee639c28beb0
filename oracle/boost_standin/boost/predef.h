/* Minimal stand-in for <boost/predef.h>: TEST INFRASTRUCTURE ONLY (see predef/version_number.h).
 * Defines exactly the BOOST_COMP_ / BOOST_LANG_ / BOOST_ARCH_ / BOOST_OS_ macros the reference tests
 * (census: grep -o 'BOOST_[A-Z_0-9]*' over /root/reference/include/alpaka). Supported host
 * compilers: g++ / clang++ on Linux x86-64, optionally under nvcc. */
#ifndef B200_ORACLE_BOOST_PREDEF_H
#define B200_ORACLE_BOOST_PREDEF_H
#include <boost/predef/version_number.h>

/* ---- compilers ---- */
#if defined(__clang__)
#    define BOOST_COMP_CLANG BOOST_VERSION_NUMBER(__clang_major__, __clang_minor__, __clang_patchlevel__)
#    define BOOST_COMP_CLANG_AVAILABLE
#    define BOOST_COMP_GNUC BOOST_VERSION_NUMBER_NOT_AVAILABLE
#elif defined(__GNUC__)
#    define BOOST_COMP_CLANG BOOST_VERSION_NUMBER_NOT_AVAILABLE
#    define BOOST_COMP_GNUC BOOST_VERSION_NUMBER(__GNUC__, __GNUC_MINOR__, __GNUC_PATCHLEVEL__)
#else
#    define BOOST_COMP_CLANG BOOST_VERSION_NUMBER_NOT_AVAILABLE
#    define BOOST_COMP_GNUC BOOST_VERSION_NUMBER_NOT_AVAILABLE
#endif
#if defined(__NVCC__)
#    define BOOST_COMP_NVCC BOOST_VERSION_NUMBER(__CUDACC_VER_MAJOR__, __CUDACC_VER_MINOR__, __CUDACC_VER_BUILD__)
#else
#    define BOOST_COMP_NVCC BOOST_VERSION_NUMBER_NOT_AVAILABLE
#endif
/* the *_EMULATED macros stay UNDEFINED, as in real Boost.Predef when nothing is emulated */
#define BOOST_COMP_MSVC BOOST_VERSION_NUMBER_NOT_AVAILABLE
#define BOOST_COMP_PGI BOOST_VERSION_NUMBER_NOT_AVAILABLE
#define BOOST_COMP_HPACC BOOST_VERSION_NUMBER_NOT_AVAILABLE
#define BOOST_COMP_SUNPRO BOOST_VERSION_NUMBER_NOT_AVAILABLE
#define BOOST_COMP_IBM BOOST_VERSION_NUMBER_NOT_AVAILABLE
#define BOOST_COMP_INTEL BOOST_VERSION_NUMBER_NOT_AVAILABLE

/* ---- languages ---- */
#if defined(__CUDACC__) || defined(__CUDA__)
#    include <cuda.h>
#    define BOOST_LANG_CUDA BOOST_PREDEF_MAKE_10_VVRRP(CUDA_VERSION)
#else
#    define BOOST_LANG_CUDA BOOST_VERSION_NUMBER_NOT_AVAILABLE
#endif

/* ---- architectures ---- */
#if defined(__CUDA_ARCH__)
#    define BOOST_ARCH_PTX BOOST_PREDEF_MAKE_10_VVRRP(__CUDA_ARCH__)
#else
#    define BOOST_ARCH_PTX BOOST_VERSION_NUMBER_NOT_AVAILABLE
#endif
#if defined(__x86_64__) || defined(__i386__)
#    define BOOST_ARCH_X86 BOOST_VERSION_NUMBER_AVAILABLE
#else
#    define BOOST_ARCH_X86 BOOST_VERSION_NUMBER_NOT_AVAILABLE
#endif

/* ---- operating systems ---- */
#if defined(__linux__)
#    define BOOST_OS_LINUX BOOST_VERSION_NUMBER_AVAILABLE
#    define BOOST_OS_UNIX BOOST_VERSION_NUMBER_AVAILABLE
#else
#    define BOOST_OS_LINUX BOOST_VERSION_NUMBER_NOT_AVAILABLE
#    define BOOST_OS_UNIX BOOST_VERSION_NUMBER_NOT_AVAILABLE
#endif
#define BOOST_OS_WINDOWS BOOST_VERSION_NUMBER_NOT_AVAILABLE
#define BOOST_OS_MACOS BOOST_VERSION_NUMBER_NOT_AVAILABLE
#define BOOST_OS_CYGWIN BOOST_VERSION_NUMBER_NOT_AVAILABLE
#define BOOST_OS_BSD BOOST_VERSION_NUMBER_NOT_AVAILABLE
#define BOOST_OS_IOS BOOST_VERSION_NUMBER_NOT_AVAILABLE

#endif
