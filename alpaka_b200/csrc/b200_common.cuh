// alpaka_b200/csrc/b200_common.cuh -- shared helpers of libalpaka_b200.so (error policy, tuning registry,
// PTX load/store wrappers). sm_100a only.
#pragma once
#include "b200/b200.h"

#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>

namespace b200
{
    // ---- error policy: thread-local last-error text, mirroring the message the reference throws
    // (core/UniformCudaHip.hpp:62-82: "file(line) 'cmd' returned error : 'name': 'string'!").
    std::string& lastError();
    int fail(int code, char const* what, char const* file, int line);
    int cudaFail(cudaError_t e, char const* cmd, char const* file, int line);
    extern std::atomic<uint64_t> g_launchCount;
    int64_t tune(char const* key, int64_t dflt);
    int smCount(int dev);
    int currentDevice();
    int useDeviceOf(cudaStream_t s); // makes the stream's device current (multi-device processes)

    inline void countLaunch()
    {
        g_launchCount.fetch_add(1, std::memory_order_relaxed);
    }
} // namespace b200

#define B200_CUDA(cmd)                                                                                                \
    do                                                                                                                \
    {                                                                                                                 \
        cudaError_t const b200_e_ = (cmd);                                                                            \
        if(b200_e_ != cudaSuccess)                                                                                    \
            return ::b200::cudaFail(b200_e_, #cmd, __FILE__, __LINE__);                                               \
    } while(0)

#define B200_REQUIRE(cond, code)                                                                                      \
    do                                                                                                                \
    {                                                                                                                 \
        if(!(cond))                                                                                                   \
            return ::b200::fail((code), #cond, __FILE__, __LINE__);                                                   \
    } while(0)

// launch check: catches configuration errors of the launch just issued without synchronising
#define B200_LAUNCH_CHECK()                                                                                           \
    do                                                                                                                \
    {                                                                                                                 \
        ::b200::countLaunch();                                                                                        \
        B200_CUDA(cudaPeekAtLastError());                                                                             \
    } while(0)

#ifdef __CUDACC__
namespace b200
{
    // ---- bounded flag waits of the fused exchanges (Dot / reduce scalar exchange, heat halo exchange).
    // A peer that died must not hang this GPU for ever, but a peer that is merely late (first-launch module load, a
    // paused rank) must not produce a wrong result either: the bound is generous (tunable `exchange.timeout_ms`,
    // default 60 s), a timeout raises the status word the host checks AND poisons the result (NaN for floating point).
    inline uint64_t waitLimitNs()
    {
        return uint64_t(tune("exchange.timeout_ms", 60000)) * 1000000ull;
    }

    __device__ __forceinline__ uint64_t globalTimerNs()
    {
        uint64_t t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        return t;
    }

    //! spins until *flag (acquire, system scope) satisfies `flag + slack >= want`; false on timeout
    __device__ __forceinline__ bool waitFlagAtLeast(uint32_t const* flag, uint32_t want, uint32_t slack, uint64_t limitNs)
    {
        uint64_t t0 = 0;
        for(uint32_t spins = 0;; ++spins)
        {
            uint32_t seen;
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(flag) : "memory");
            if(seen + slack >= want)
                return true;
            if(spins == 64u)
                t0 = globalTimerNs();
            else if(spins > 64u && (spins & 63u) == 0u && globalTimerNs() - t0 > limitNs)
                return false;
            if(spins > 16u)
                __nanosleep(spins > 1024u ? 1000 : 100);
        }
    }

    // ---- 128-bit and 256-bit global accesses with streaming cache policy.
    // HINT: 0 = default, 1 = ld.nc / L1::no_allocate + st .cs (evict-first streaming),
    //       2 (256-bit loads only) = ld.nc / L1::no_allocate / L2::evict_last: the lines just read stay behind the dirty
    //         lines of the stores in the L2's replacement order -- Triad/Add 7153 vs 7131 GB/s (profiles/r02/tune_stream_fine.log)
    template<int HINT>
    __device__ __forceinline__ void ldg128(void const* p, uint32_t (&r)[4])
    {
        if constexpr(HINT == 0)
            asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                         : "l"(p));
        else
            asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                         : "l"(p));
    }

    template<int HINT>
    __device__ __forceinline__ void stg128(void* p, uint32_t const (&r)[4])
    {
        if constexpr(HINT == 0)
            asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3])
                         : "memory");
        else
            asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p),
                         "r"(r[0]),
                         "r"(r[1]),
                         "r"(r[2]),
                         "r"(r[3])
                         : "memory");
    }

    // 256-bit accesses: new with sm_100 (PTX ISA 8.8, ld/st.global.v4.b64 / .v8.b32).
    template<int HINT>
    __device__ __forceinline__ void ldg256(void const* p, uint64_t (&r)[4])
    {
        if constexpr(HINT == 0)
            asm volatile("ld.global.v4.u64 {%0,%1,%2,%3}, [%4];"
                         : "=l"(r[0]), "=l"(r[1]), "=l"(r[2]), "=l"(r[3])
                         : "l"(p));
        else if constexpr(HINT == 2)
            asm volatile("ld.global.nc.L1::no_allocate.L2::evict_last.v4.u64 {%0,%1,%2,%3}, [%4];"
                         : "=l"(r[0]), "=l"(r[1]), "=l"(r[2]), "=l"(r[3])
                         : "l"(p));
        else
            asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];"
                         : "=l"(r[0]), "=l"(r[1]), "=l"(r[2]), "=l"(r[3])
                         : "l"(p));
    }

    template<int HINT>
    __device__ __forceinline__ void stg256(void* p, uint64_t const (&r)[4])
    {
        if constexpr(HINT == 0)
            asm volatile("st.global.v4.u64 [%0], {%1,%2,%3,%4};" ::"l"(p), "l"(r[0]), "l"(r[1]), "l"(r[2]), "l"(r[3])
                         : "memory");
        else
            asm volatile("st.global.cs.v4.u64 [%0], {%1,%2,%3,%4};" ::"l"(p),
                         "l"(r[0]),
                         "l"(r[1]),
                         "l"(r[2]),
                         "l"(r[3])
                         : "memory");
    }
} // namespace b200
#endif
