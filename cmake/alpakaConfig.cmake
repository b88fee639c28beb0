# cmake/alpakaConfig.cmake -- `find_package(alpaka)` for the B200 back-end of this repository.
#
# The CMake boundary existing alpaka projects use (reference: cmake/alpakaCommon.cmake:61-80,479-483 builds the INTERFACE
# target `alpaka` / alias `alpaka::alpaka`; cmake/addExecutable.cmake:1-19 provides `alpaka_add_executable`). A project's
# own CMakeLists.txt stays as it is:
#
#     cmake -S <reference>/example/heatEquation2D -B build -Dalpaka_DIR=<this repo>/cmake
#
# What the target carries: include/ of this repository (alpaka/alpaka.hpp -> the B200 accelerator), C++20, nvcc's
# --extended-lambda --expt-relaxed-constexpr, libalpaka_b200.so + its rpath. What a target cannot inherit from an INTERFACE
# library -- CUDA as the language of .cpp sources, compute_100a, the SHARED CUDA runtime (the generic launch entry
# b200_launch and the user's translation unit must share one cudart instance) -- is set by alpaka_add_executable, exactly
# where the reference sets its per-target CUDA properties.
#
# Options (cache variables):
#   alpaka_B200_RECOGNIZE_REFERENCE_KERNELS  ON   route kernel functors named like the reference drivers' to the native kernels
#                                                 (-DALPAKA_B200_RECOGNIZE_REFERENCE_KERNELS; each claim is cross-checked at run time)
#   alpaka_B200_STRICT_FP                    OFF  -fmad=false: user functors bit-identical to a CPU build with -ffp-contract=off
#   alpaka_DEBUG                             0    ALPAKA_DEBUG level (reference: cmake/alpakaCommon.cmake alpaka_DEBUG)
# There is one accelerator and no CPU back-end: alpaka_ACC_GPU_CUDA_ENABLE is always ON, every other alpaka_ACC_* is ignored.
cmake_minimum_required(VERSION 3.25)

get_filename_component(_alpaka_b200_ROOT "${CMAKE_CURRENT_LIST_DIR}/.." ABSOLUTE)
set(_alpaka_b200_LIBDIR "${_alpaka_b200_ROOT}/alpaka_b200/lib")
if(NOT EXISTS "${_alpaka_b200_LIBDIR}/libalpaka_b200.so")
    message(FATAL_ERROR "alpaka (B200): ${_alpaka_b200_LIBDIR}/libalpaka_b200.so is missing -- build it with "
                        "`make -C ${_alpaka_b200_ROOT}/alpaka_b200/csrc` (there is no header-only or CPU fallback).")
endif()

option(alpaka_B200_RECOGNIZE_REFERENCE_KERNELS "Route the reference drivers' kernel functors to the hand-written sm_100a kernels" ON)
option(alpaka_B200_STRICT_FP "Compile user kernels with -fmad=false (no FMA contraction)" OFF)
set(alpaka_DEBUG "0" CACHE STRING "Debug level")
set(alpaka_ACC_GPU_CUDA_ENABLE ON)
set(alpaka_CXX_STANDARD 20)
set(alpaka_VERSION 2.0.0)

if(NOT CMAKE_CUDA_COMPILER AND EXISTS "/usr/local/cuda/bin/nvcc")
    set(CMAKE_CUDA_COMPILER "/usr/local/cuda/bin/nvcc")
endif()
if(NOT CMAKE_CUDA_HOST_COMPILER AND EXISTS "/usr/bin/g++")
    set(CMAKE_CUDA_HOST_COMPILER "/usr/bin/g++")
endif()
if(NOT DEFINED CMAKE_CUDA_ARCHITECTURES)
    set(CMAKE_CUDA_ARCHITECTURES "100a")
endif()

if(NOT TARGET alpaka::alpaka)
    add_library(alpaka::alpaka INTERFACE IMPORTED)
    target_include_directories(alpaka::alpaka INTERFACE "${_alpaka_b200_ROOT}/include")
    target_compile_features(alpaka::alpaka INTERFACE cxx_std_20)
    target_compile_options(alpaka::alpaka INTERFACE
        "$<$<COMPILE_LANGUAGE:CUDA>:--extended-lambda>"
        "$<$<COMPILE_LANGUAGE:CUDA>:--expt-relaxed-constexpr>"
        "$<$<COMPILE_LANGUAGE:CUDA>:-lineinfo>")
    if(alpaka_B200_STRICT_FP)
        target_compile_options(alpaka::alpaka INTERFACE "$<$<COMPILE_LANGUAGE:CUDA>:-fmad=false>"
                                                        "$<$<COMPILE_LANGUAGE:CUDA>:-Xcompiler=-ffp-contract=off>")
    endif()
    if(alpaka_B200_RECOGNIZE_REFERENCE_KERNELS)
        target_compile_definitions(alpaka::alpaka INTERFACE ALPAKA_B200_RECOGNIZE_REFERENCE_KERNELS)
    endif()
    if(NOT alpaka_DEBUG STREQUAL "0")
        target_compile_definitions(alpaka::alpaka INTERFACE "ALPAKA_DEBUG=${alpaka_DEBUG}")
    endif()
    target_link_libraries(alpaka::alpaka INTERFACE "${_alpaka_b200_LIBDIR}/libalpaka_b200.so")
    target_link_options(alpaka::alpaka INTERFACE "LINKER:-rpath,${_alpaka_b200_LIBDIR}")
endif()

#------------------------------------------------------------------------------
# alpaka_add_executable(<name> [WIN32] [MACOSX_BUNDLE] [EXCLUDE_FROM_ALL] [<source>...])
# A macro, like the reference's, so that enable_language(CUDA) happens in the caller's scope.
macro(alpaka_add_executable In_Name)
    enable_language(CUDA)
    add_executable(${In_Name} ${ARGN})
    foreach(_alpaka_b200_file ${ARGN})
        if(("${_alpaka_b200_file}" MATCHES "\\.cpp$") OR ("${_alpaka_b200_file}" MATCHES "\\.cxx$") OR ("${_alpaka_b200_file}" MATCHES "\\.cu$"))
            set_source_files_properties(${_alpaka_b200_file} PROPERTIES LANGUAGE CUDA)
        endif()
    endforeach()
    set_target_properties(${In_Name} PROPERTIES
        CUDA_STANDARD 20
        CUDA_STANDARD_REQUIRED ON
        CUDA_ARCHITECTURES "${CMAKE_CUDA_ARCHITECTURES}"
        CUDA_RUNTIME_LIBRARY Shared
        CUDA_SEPARABLE_COMPILATION OFF
        LINKER_LANGUAGE CUDA)
endmacro()

set(alpaka_FOUND TRUE)
