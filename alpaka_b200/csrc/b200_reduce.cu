// alpaka_b200/csrc/b200_reduce.cu -- Dot and sum-reduce as single-pass grid reductions for sm_100a.
//
// Replaces (a) the reference DotKernel + host std::reduce (benchmarks/babelstream/src/babelStreamMainTest.cpp:
// 145-181, 399-405: fixed 256 blocks x 1024 threads, shared-memory halving tree with a barrier per level, D2H
// copy of 256 partials and a host fold) and (b) example/reduce's ReduceKernel launched twice
// (example/reduce/src/kernel.hpp:42-132, reduce.cpp:79-98).
//
// Design (HBM-bound: 16 B/element for Dot, sizeof(T) for reduce):
//   level 1  per-thread grid-stride accumulation over 32-byte vectors (ld.global.nc.v4.b64), UNROLL independent
//            accumulators so UNROLL*(1|2) loads are in flight per thread;
//   level 2  warp tree with __shfl_down_sync (no barriers), one shared-memory slot per warp, first warp finishes;
//   level 3  block partial -> scratch[blockIdx]; __threadfence(); atomic ticket; the LAST block to arrive folds the
//            partials in a fixed order (thread t takes t, t+B, ...; then the same block tree) and writes the scalar.
//   level 4  (several GPUs, optional) the same last block stores the device's scalar straight into every rank's slot
//            array through peer pointers (NVLink P2P / CUDA-IPC mappings), publishes the call number in their flag words
//            (st.release.sys), waits for the other ranks' flags (ld.acquire.sys) and folds the slots in RANK ORDER: the
//            all-ranks result comes out of the SAME launch on every rank, bit-identical everywhere -- no NCCL, no host.
// One launch, no host fold, no second kernel. The order of additions depends only on (n, grid, block), which depend
// only on the device -> bit-reproducible run to run. Products use __dmul_rn + __dadd_rn (no FMA), matching the
// oracle's pinned contraction; integer sums wrap (order-free, bit-exact vs the reference).
#include "b200_common.cuh"

#include <cmath>
#include <type_traits>

namespace
{
    using b200::ldg256;

    constexpr int kBlock = 512;
    constexpr int kMaxPartials = 4096; // 4096 * 8 B = 32 KiB of the 64 KiB scratch
    constexpr size_t kTicketOffset = size_t(kMaxPartials) * 8;
    static_assert(kTicketOffset + 64 <= B200_REDUCE_SCRATCH_BYTES);

    __device__ __forceinline__ double addv(double a, double b)
    {
        return __dadd_rn(a, b);
    }

    __device__ __forceinline__ float addv(float a, float b)
    {
        return __fadd_rn(a, b);
    }

    __device__ __forceinline__ uint32_t addv(uint32_t a, uint32_t b)
    {
        return a + b;
    }

    __device__ __forceinline__ int32_t addv(int32_t a, int32_t b)
    {
        return int32_t(uint32_t(a) + uint32_t(b)); // two's complement wrap, like the reference on x86
    }

    __device__ __forceinline__ uint64_t addv(uint64_t a, uint64_t b)
    {
        return a + b;
    }

    __device__ __forceinline__ double mulv(double a, double b)
    {
        return __dmul_rn(a, b);
    }

    __device__ __forceinline__ float mulv(float a, float b)
    {
        return __fmul_rn(a, b);
    }

    template<typename T>
    __device__ __forceinline__ T mulv(T a, T b)
    {
        return a * b;
    }

    template<typename T>
    __device__ __forceinline__ T shflDown(T v, int delta)
    {
        if constexpr(sizeof(T) == 8)
        {
            uint64_t u;
            memcpy(&u, &v, 8);
            uint32_t lo = uint32_t(u), hi = uint32_t(u >> 32);
            lo = __shfl_down_sync(0xffffffffu, lo, delta);
            hi = __shfl_down_sync(0xffffffffu, hi, delta);
            u = (uint64_t(hi) << 32) | lo;
            T r;
            memcpy(&r, &u, 8);
            return r;
        }
        else
        {
            return __shfl_down_sync(0xffffffffu, v, delta);
        }
    }

    // block-wide sum, result valid in thread 0. Fixed order: lane tree (16,8,4,2,1) then warp tree.
    template<typename T>
    __device__ __forceinline__ T blockSum(T v, T* warpSlots)
    {
        int const lane = threadIdx.x & 31;
        int const warp = threadIdx.x >> 5;
#pragma unroll
        for(int d = 16; d > 0; d >>= 1)
            v = addv(v, shflDown(v, d));
        if(lane == 0)
            warpSlots[warp] = v;
        __syncthreads();
        int const nWarps = blockDim.x >> 5;
        if(warp == 0)
        {
            v = lane < nWarps ? warpSlots[lane] : T(0);
#pragma unroll
            for(int d = 16; d > 0; d >>= 1)
                v = addv(v, shflDown(v, d));
        }
        return v;
    }

    template<typename T>
    __device__ __forceinline__ T ldVolatile(T const* p)
    {
        return *reinterpret_cast<T const volatile*>(p);
    }

    // Exchange buffer of ONE rank (B200_EXCHANGE_BYTES, zeroed once): slots[2][kMaxRanks] 8-byte containers (two parities of
    // the call number), then flags[kMaxRanks] (u32, the call number each rank last published here), then a status word.
    constexpr int kMaxRanks = B200_EXCHANGE_MAX_RANKS;
    constexpr size_t kFlagsOffset = 2 * size_t(kMaxRanks) * 8;
    constexpr size_t kStatusOffset = kFlagsOffset + size_t(kMaxRanks) * 4;
    static_assert(kStatusOffset + 4 <= B200_EXCHANGE_BYTES);

    struct Exchange
    {
        char* base[kMaxRanks]; // base[r]: rank r's exchange buffer as seen from this device; base[rank] is the own one
        uint32_t world, rank, step; // world <= 1: no exchange
        uint64_t waitNs; // bound of the flag wait (b200::waitLimitNs)
    };

    // Called by the threads of the last block after thread 0 produced this device's scalar `mine`. Returns (in thread 0)
    // the sum over all ranks, folded left to right in rank order.
    template<typename T>
    __device__ __forceinline__ T exchangeAllRanks(Exchange const& X, T mine, uint64_t* smemSlots)
    {
        __shared__ uint64_t mineBits;
        __shared__ int timedOut;
        if(threadIdx.x == 0)
        {
            timedOut = 0;
            uint64_t u = 0;
            memcpy(&u, &mine, sizeof(T));
            mineBits = u;
        }
        __syncthreads();
        uint32_t const parity = X.step & 1u;
        if(threadIdx.x < X.world)
        {
            // thread p: my scalar into rank p's slot for me, then the call number into its flag for me
            uint32_t const p = threadIdx.x;
            auto* slot = reinterpret_cast<uint64_t*>(X.base[p]) + parity * kMaxRanks + X.rank;
            auto* flag = reinterpret_cast<uint32_t*>(X.base[p] + kFlagsOffset) + X.rank;
            asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(slot), "l"(mineBits) : "memory");
            asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(X.step) : "memory");
            // ... and rank p's scalar out of MY buffer once its flag says it has arrived (bounded wait, b200::waitLimitNs)
            char* const my = X.base[X.rank];
            auto const* myFlag = reinterpret_cast<uint32_t const*>(my + kFlagsOffset) + p;
            if(!b200::waitFlagAtLeast(myFlag, X.step, 0u, X.waitNs))
            {
                atomicExch(reinterpret_cast<uint32_t*>(my + kStatusOffset), 1u + p);
                timedOut = 1;
            }
            uint64_t v;
            asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(reinterpret_cast<uint64_t const*>(my) + parity * kMaxRanks + p) : "memory");
            smemSlots[p] = v;
        }
        __syncthreads();
        T total = T(0);
        if(threadIdx.x == 0)
        {
            memcpy(&total, &smemSlots[0], sizeof(T));
            for(uint32_t r = 1; r < X.world; ++r)
            {
                T v;
                memcpy(&v, &smemSlots[r], sizeof(T));
                total = addv(total, v);
            }
            // a peer never arrived: the status word is set; floating-point results are poisoned as well
            if constexpr(std::is_floating_point_v<T>)
                if(timedOut)
                    total = T(NAN);
        }
        return total;
    }

    // DOT: sum a[i]*b[i]; otherwise sum a[i].
    template<typename T, bool DOT, int UNROLL>
    __global__ void __launch_bounds__(kBlock) reduceKernel(
        T const* __restrict__ a,
        T const* __restrict__ b,
        uint64_t const n,
        uint64_t const nVec, // full 32-byte vectors (0 when the inputs are not 32-byte aligned)
        T* __restrict__ partials,
        unsigned int* __restrict__ ticket,
        T* __restrict__ out,
        uint32_t const nOut,
        Exchange const X)
    {
        constexpr int N = 32 / int(sizeof(T));
        __shared__ T warpSlots[32];
        __shared__ uint64_t rankSlots[kMaxRanks];
        __shared__ bool isLast;

        T acc[UNROLL];
#pragma unroll
        for(int u = 0; u < UNROLL; ++u)
            acc[u] = T(0);

        // ---- level 1: vector body. A block owns chunks of blockDim*UNROLL vectors, grid-strided.
        uint64_t const chunk = uint64_t(blockDim.x) * UNROLL;
        uint64_t const nFull = nVec / chunk;
        for(uint64_t c = blockIdx.x; c < nFull; c += gridDim.x)
        {
            uint64_t const base = c * chunk + threadIdx.x;
            uint64_t ra[UNROLL][4];
            uint64_t rb[UNROLL][4];
#pragma unroll
            for(int u = 0; u < UNROLL; ++u)
            {
                ldg256<1>(a + (base + uint64_t(u) * blockDim.x) * N, ra[u]);
                if constexpr(DOT)
                    ldg256<1>(b + (base + uint64_t(u) * blockDim.x) * N, rb[u]);
            }
#pragma unroll
            for(int u = 0; u < UNROLL; ++u)
            {
                T va[N], vb[N];
                memcpy(va, ra[u], 32);
                if constexpr(DOT)
                    memcpy(vb, rb[u], 32);
#pragma unroll
                for(int k = 0; k < N; ++k)
                    acc[u] = addv(acc[u], DOT ? mulv(va[k], vb[k]) : va[k]);
            }
        }
        // remainder vectors and scalar tail: plain grid-stride over elements, folded into acc[0]
        {
            uint64_t const start = nFull * chunk * N;
            uint64_t const tid = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
            uint64_t const stride = uint64_t(gridDim.x) * blockDim.x;
            for(uint64_t i = start + tid; i < n; i += stride)
                acc[0] = addv(acc[0], DOT ? mulv(a[i], b[i]) : a[i]);
        }
        T v = acc[0];
#pragma unroll
        for(int u = 1; u < UNROLL; ++u)
            v = addv(v, acc[u]);

        // ---- level 2
        v = blockSum(v, warpSlots);

        // ---- level 3: single-pass grid reduction
        if(threadIdx.x == 0)
        {
            partials[blockIdx.x] = v;
            __threadfence();
            unsigned int const t = atomicAdd(ticket, 1u);
            isLast = (t == gridDim.x - 1);
        }
        __syncthreads();
        if(!isLast)
            return;
        __threadfence();

        if(nOut == 1)
        {
            T s = T(0);
            for(uint32_t i = threadIdx.x; i < gridDim.x; i += blockDim.x)
                s = addv(s, ldVolatile(partials + i));
            __syncthreads(); // warpSlots reuse
            s = blockSum(s, warpSlots);
            if(X.world > 1)
                s = exchangeAllRanks(X, s, rankSlots); // level 4: all ranks, rank order, same launch
            if(threadIdx.x == 0)
                out[0] = s;
        }
        else
        {
            // reference-shaped output: out[k] = sum of partials k, k+nOut, ...  (std::reduce(out) == total)
            for(uint32_t k = threadIdx.x; k < nOut; k += blockDim.x)
            {
                T s = T(0);
                for(uint32_t i = k; i < gridDim.x; i += nOut)
                    s = addv(s, ldVolatile(partials + i));
                out[k] = s;
            }
        }
        if(threadIdx.x == 0)
            *ticket = 0u; // ready for the next launch on this scratch
    }

    template<typename T, bool DOT>
    int launchReduce(b200_stream_t stream, T const* a, T const* b, uint64_t n, T* out, uint32_t nOut, void* scratch, b200_exchange const* ex = nullptr, uint32_t step = 0)
    {
        B200_REQUIRE(out && scratch && nOut >= 1, B200_EINVAL);
        Exchange X{};
        if(ex != nullptr)
        {
            B200_REQUIRE(nOut == 1 && step >= 1 && ex->world >= 1 && ex->world <= uint32_t(kMaxRanks) && ex->rank < ex->world, B200_EINVAL);
            for(uint32_t r = 0; r < ex->world; ++r)
            {
                B200_REQUIRE(ex->base[r] != nullptr && reinterpret_cast<uintptr_t>(ex->base[r]) % 8 == 0, B200_EINVAL);
                X.base[r] = static_cast<char*>(ex->base[r]);
            }
            X.world = ex->world;
            X.rank = ex->rank;
            X.step = step;
            X.waitNs = b200::waitLimitNs();
        }
        B200_REQUIRE(n == 0 || (a && (!DOT || b)), B200_EINVAL);
        B200_REQUIRE(reinterpret_cast<uintptr_t>(scratch) % 16 == 0, B200_EALIGN);
        auto const s = reinterpret_cast<cudaStream_t>(stream);
        if(int const rc = b200::useDeviceOf(s))
            return rc;
        constexpr int N = 32 / int(sizeof(T));
        bool const aligned
            = reinterpret_cast<uintptr_t>(a) % 32 == 0 && (!DOT || reinterpret_cast<uintptr_t>(b) % 32 == 0);
        uint64_t const nVec = aligned ? n / N : 0;
        int const unroll = int(b200::tune(DOT ? "dot.unroll" : "reduce.unroll", DOT ? 2 : 4));
        int const ctasPerSm = int(b200::tune(DOT ? "dot.ctas_per_sm" : "reduce.ctas_per_sm", DOT ? 2 : 3));
        uint64_t const chunk = uint64_t(kBlock) * unroll;
        uint64_t grid = uint64_t(b200::smCount(b200::currentDevice())) * (ctasPerSm > 0 ? ctasPerSm : 4);
        uint64_t const need = (n / N + chunk - 1) / chunk;
        if(grid > need)
            grid = need ? need : 1;
        if(grid > kMaxPartials)
            grid = kMaxPartials;
        T* partials = static_cast<T*>(scratch);
        auto* ticket = reinterpret_cast<unsigned int*>(static_cast<char*>(scratch) + kTicketOffset);
        switch(unroll)
        {
        case 1:
            reduceKernel<T, DOT, 1><<<unsigned(grid), kBlock, 0, s>>>(a, b, n, nVec, partials, ticket, out, nOut, X);
            break;
        case 2:
            reduceKernel<T, DOT, 2><<<unsigned(grid), kBlock, 0, s>>>(a, b, n, nVec, partials, ticket, out, nOut, X);
            break;
        case 4:
            reduceKernel<T, DOT, 4><<<unsigned(grid), kBlock, 0, s>>>(a, b, n, nVec, partials, ticket, out, nOut, X);
            break;
        default:
            return b200::fail(B200_EINVAL, "unroll must be 1, 2 or 4", __FILE__, __LINE__);
        }
        B200_LAUNCH_CHECK();
        return 0;
    }
} // namespace

extern "C"
{
    int b200_dot_f64(b200_stream_t s, double const* a, double const* b, uint64_t n, double* out_dev, void* scratch)
    {
        return launchReduce<double, true>(s, a, b, n, out_dev, 1, scratch);
    }

    int b200_dot_f32(b200_stream_t s, float const* a, float const* b, uint64_t n, float* out_dev, void* scratch)
    {
        return launchReduce<float, true>(s, a, b, n, out_dev, 1, scratch);
    }

    int b200_dot_partials_f64(b200_stream_t s, double const* a, double const* b, uint64_t n, double* partials_dev, uint32_t n_partials, void* scratch)
    {
        B200_REQUIRE(n_partials >= 2, B200_EINVAL);
        return launchReduce<double, true>(s, a, b, n, partials_dev, n_partials, scratch);
    }

    int b200_dot_partials_f32(b200_stream_t s, float const* a, float const* b, uint64_t n, float* partials_dev, uint32_t n_partials, void* scratch)
    {
        B200_REQUIRE(n_partials >= 2, B200_EINVAL);
        return launchReduce<float, true>(s, a, b, n, partials_dev, n_partials, scratch);
    }

    int b200_reduce_sum_u32(b200_stream_t s, uint32_t const* in, uint64_t n, uint32_t* out_dev, void* scratch)
    {
        return launchReduce<uint32_t, false>(s, in, nullptr, n, out_dev, 1, scratch);
    }

    int b200_reduce_sum_i32(b200_stream_t s, int32_t const* in, uint64_t n, int32_t* out_dev, void* scratch)
    {
        return launchReduce<int32_t, false>(s, in, nullptr, n, out_dev, 1, scratch);
    }

    int b200_reduce_sum_u64(b200_stream_t s, uint64_t const* in, uint64_t n, uint64_t* out_dev, void* scratch)
    {
        return launchReduce<uint64_t, false>(s, in, nullptr, n, out_dev, 1, scratch);
    }

    int b200_reduce_sum_f32(b200_stream_t s, float const* in, uint64_t n, float* out_dev, void* scratch)
    {
        return launchReduce<float, false>(s, in, nullptr, n, out_dev, 1, scratch);
    }

    int b200_reduce_sum_f64(b200_stream_t s, double const* in, uint64_t n, double* out_dev, void* scratch)
    {
        return launchReduce<double, false>(s, in, nullptr, n, out_dev, 1, scratch);
    }

    // ---- the same reductions with the all-ranks exchange fused into the launch
    int b200_dot_allranks_f64(b200_stream_t s, double const* a, double const* b, uint64_t n, double* out_dev, void* scratch, b200_exchange const* ex, uint32_t step)
    {
        B200_REQUIRE(ex, B200_EINVAL);
        return launchReduce<double, true>(s, a, b, n, out_dev, 1, scratch, ex, step);
    }

    int b200_dot_allranks_f32(b200_stream_t s, float const* a, float const* b, uint64_t n, float* out_dev, void* scratch, b200_exchange const* ex, uint32_t step)
    {
        B200_REQUIRE(ex, B200_EINVAL);
        return launchReduce<float, true>(s, a, b, n, out_dev, 1, scratch, ex, step);
    }

    int b200_reduce_sum_allranks_u32(b200_stream_t s, uint32_t const* in, uint64_t n, uint32_t* out_dev, void* scratch, b200_exchange const* ex, uint32_t step)
    {
        B200_REQUIRE(ex, B200_EINVAL);
        return launchReduce<uint32_t, false>(s, in, nullptr, n, out_dev, 1, scratch, ex, step);
    }

    int b200_reduce_sum_allranks_f32(b200_stream_t s, float const* in, uint64_t n, float* out_dev, void* scratch, b200_exchange const* ex, uint32_t step)
    {
        B200_REQUIRE(ex, B200_EINVAL);
        return launchReduce<float, false>(s, in, nullptr, n, out_dev, 1, scratch, ex, step);
    }

    int b200_reduce_sum_allranks_f64(b200_stream_t s, double const* in, uint64_t n, double* out_dev, void* scratch, b200_exchange const* ex, uint32_t step)
    {
        B200_REQUIRE(ex, B200_EINVAL);
        return launchReduce<double, false>(s, in, nullptr, n, out_dev, 1, scratch, ex, step);
    }

    int b200_exchange_status(int dev, void const* own_exchange_buffer, uint32_t* status)
    {
        B200_REQUIRE(own_exchange_buffer && status, B200_EINVAL);
        B200_CUDA(cudaSetDevice(dev));
        B200_CUDA(cudaMemcpy(status, static_cast<char const*>(own_exchange_buffer) + kStatusOffset, sizeof(uint32_t), cudaMemcpyDeviceToHost));
        return 0;
    }
}
