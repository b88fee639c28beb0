// tools/stream_lab.cu -- measurement bench for the BabelStream Copy / Triad kernels on one B200 (development tool).
//
// Every way of moving the same bytes that VERDICT r01 (weak #3) asked to try, timed with CUDA events on 2^30 doubles
// per array; the winner is what alpaka_b200/csrc/b200_stream.cu ships. One line per configuration:
//     <op> <variant> <params> : <ms> ms <GB/s>
// Variants
//   ldg      per-thread 32-byte LDG/STG (the r01 kernel): launch shape x cache policy x array placement
//            shape  0 = one 512-vector chunk per block
//                   1 = persistent, lock-step grid stride
//                   2 = persistent, every CTA owns ONE CONTIGUOUS range
//            policy lp: 0 default | 1 nc.L1::no_allocate | 2 nc + L2::evict_first | 3 nc + L2::evict_last
//                       4 nc + L2::256B prefetch | 5 nc + L2::evict_first + L2::256B
//                   sp: 0 default | 1 .cs | 2 L2::evict_first | 3 L2::evict_last | 4 .wt | 5 L1::no_allocate + L2::evict_first
//   bulk     shared-memory staging with 1-D bulk copies (cp.async.bulk, SASS UBLKCP): a producer thread streams
//            multi-KB tiles in, the consumers compute in shared memory, one thread streams the result tile out
//            (long same-direction bursts per CTA). tile bytes x stages x CTAs/SM x shape (0 strided / 1 contiguous)
//   pad      array placement: distance between the arrays = 8 GiB + pad (DRAM bank/row aliasing probe)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 tools/stream_lab.cu -o build/tools/stream_lab
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#define CK(x)                                                                                                         \
    do                                                                                                                \
    {                                                                                                                 \
        cudaError_t e_ = (x);                                                                                         \
        if(e_ != cudaSuccess)                                                                                         \
        {                                                                                                             \
            std::fprintf(stderr, "%s:%d %s -> %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_));                 \
            std::exit(2);                                                                                             \
        }                                                                                                             \
    } while(0)

namespace
{
    struct V4
    {
        double v[4];
    };

    __device__ __forceinline__ uint64_t policyEvictFirst()
    {
        uint64_t p;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
        return p;
    }

    template<int LP>
    __device__ __forceinline__ V4 ld32(double const* p)
    {
        V4 r;
        if constexpr(LP == 0)
            asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.v[0]), "=d"(r.v[1]), "=d"(r.v[2]), "=d"(r.v[3]) : "l"(p));
        else if constexpr(LP == 1)
            asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];"
                         : "=d"(r.v[0]), "=d"(r.v[1]), "=d"(r.v[2]), "=d"(r.v[3])
                         : "l"(p));
        else if constexpr(LP == 2)
            asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v4.f64 {%0,%1,%2,%3}, [%4];"
                         : "=d"(r.v[0]), "=d"(r.v[1]), "=d"(r.v[2]), "=d"(r.v[3])
                         : "l"(p));
        else if constexpr(LP == 3)
            asm volatile("ld.global.nc.L1::no_allocate.L2::evict_last.v4.f64 {%0,%1,%2,%3}, [%4];"
                         : "=d"(r.v[0]), "=d"(r.v[1]), "=d"(r.v[2]), "=d"(r.v[3])
                         : "l"(p));
        else if constexpr(LP == 4)
            asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v4.f64 {%0,%1,%2,%3}, [%4];"
                         : "=d"(r.v[0]), "=d"(r.v[1]), "=d"(r.v[2]), "=d"(r.v[3])
                         : "l"(p));
        else
            asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.L2::256B.v4.f64 {%0,%1,%2,%3}, [%4];"
                         : "=d"(r.v[0]), "=d"(r.v[1]), "=d"(r.v[2]), "=d"(r.v[3])
                         : "l"(p));
        return r;
    }

    template<int SP>
    __device__ __forceinline__ void st32(double* p, V4 const& r)
    {
        if constexpr(SP == 0)
            asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(r.v[0]), "d"(r.v[1]), "d"(r.v[2]), "d"(r.v[3]) : "memory");
        else if constexpr(SP == 1)
            asm volatile("st.global.cs.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(r.v[0]), "d"(r.v[1]), "d"(r.v[2]), "d"(r.v[3]) : "memory");
        else if constexpr(SP == 2)
            asm volatile("st.global.L2::evict_first.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(r.v[0]), "d"(r.v[1]), "d"(r.v[2]), "d"(r.v[3]) : "memory");
        else if constexpr(SP == 3)
            asm volatile("st.global.L2::evict_last.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(r.v[0]), "d"(r.v[1]), "d"(r.v[2]), "d"(r.v[3]) : "memory");
        else if constexpr(SP == 4)
            asm volatile("st.global.wt.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(r.v[0]), "d"(r.v[1]), "d"(r.v[2]), "d"(r.v[3]) : "memory");
        else
            asm volatile("st.global.L1::no_allocate.L2::evict_first.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(r.v[0]), "d"(r.v[1]), "d"(r.v[2]), "d"(r.v[3]) : "memory");
    }

    // OP: 0 copy (c = a), 1 triad (c = a + s*b), 2 read-only (a, b -> one value per thread), 3 write-only (c = s)
    template<int OP, int LP, int SP, int U>
    __device__ __forceinline__ void chunkStep(double const* a, double const* b, double* c, double s, uint64_t base, uint32_t stride, double& sink)
    {
        V4 x[U], y[U];
#pragma unroll
        for(int u = 0; u < U; ++u)
        {
            if constexpr(OP != 3)
                x[u] = ld32<LP>(a + 4 * (base + uint64_t(u) * stride));
            if constexpr(OP == 1 || OP == 2)
                y[u] = ld32<LP>(b + 4 * (base + uint64_t(u) * stride));
        }
#pragma unroll
        for(int u = 0; u < U; ++u)
        {
            V4 o;
#pragma unroll
            for(int k = 0; k < 4; ++k)
            {
                if constexpr(OP == 0)
                    o.v[k] = x[u].v[k];
                else if constexpr(OP == 1)
                    o.v[k] = __dadd_rn(x[u].v[k], __dmul_rn(s, y[u].v[k]));
                else if constexpr(OP == 2)
                    sink += x[u].v[k] * y[u].v[k];
                else
                    o.v[k] = s;
            }
            if constexpr(OP != 2)
                st32<SP>(c + 4 * (base + uint64_t(u) * stride), o);
        }
    }

    // SHAPE 0/1: chunk index = blockIdx.x + k*gridDim.x (grid = #chunks for shape 0); SHAPE 2: contiguous range per CTA
    template<int OP, int LP, int SP, int U, int SHAPE>
    __global__ void __launch_bounds__(U >= 8 ? 256 : (U >= 4 ? 512 : 1024)) ldgKernel(double const* a, double const* b, double* c, double s, uint64_t nVec, double* sinkOut)
    {
        uint64_t const chunk = uint64_t(blockDim.x) * U;
        uint64_t const nChunks = nVec / chunk; // lab sizes are multiples of the chunk
        double sink = 0.0;
        if constexpr(SHAPE == 2)
        {
            uint64_t const per = (nChunks + gridDim.x - 1) / gridDim.x;
            uint64_t const c0 = per * blockIdx.x;
            uint64_t const c1 = c0 + per < nChunks ? c0 + per : nChunks;
            for(uint64_t ch = c0; ch < c1; ++ch)
                chunkStep<OP, LP, SP, U>(a, b, c, s, ch * chunk + threadIdx.x, blockDim.x, sink);
        }
        else
        {
            for(uint64_t ch = blockIdx.x; ch < nChunks; ch += gridDim.x)
                chunkStep<OP, LP, SP, U>(a, b, c, s, ch * chunk + threadIdx.x, blockDim.x, sink);
        }
        if constexpr(OP == 2)
            if(sink == 123.456)
                *sinkOut = sink;
    }

    // ---------------------------------------------------------------------------------------------- bulk staging
    __device__ __forceinline__ uint32_t sAddr(void const* p)
    {
        return uint32_t(__cvta_generic_to_shared(p));
    }

    __device__ __forceinline__ void mbarInit(uint64_t* bar, uint32_t count)
    {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sAddr(bar)), "r"(count));
    }

    __device__ __forceinline__ void mbarExpectTx(uint64_t* bar, uint32_t bytes)
    {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sAddr(bar)), "r"(bytes) : "memory");
    }

    __device__ __forceinline__ void mbarArrive(uint64_t* bar)
    {
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(sAddr(bar)) : "memory");
    }

    __device__ __forceinline__ void mbarWait(uint64_t* bar, uint32_t parity)
    {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "WAIT_%=:\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
            "@p bra DONE_%=;\n"
            "bra WAIT_%=;\n"
            "DONE_%=:\n"
            "}\n" ::"r"(sAddr(bar)),
            "r"(parity)
            : "memory");
    }

    template<bool HINTED>
    __device__ __forceinline__ void bulkIn(void* smemDst, void const* gsrc, uint32_t bytes, uint64_t* bar, uint64_t pol)
    {
        if constexpr(HINTED)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(sAddr(smemDst)),
                         "l"(gsrc),
                         "r"(bytes),
                         "r"(sAddr(bar)),
                         "l"(pol)
                         : "memory");
        else
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sAddr(smemDst)),
                         "l"(gsrc),
                         "r"(bytes),
                         "r"(sAddr(bar))
                         : "memory");
    }

    template<bool HINTED>
    __device__ __forceinline__ void bulkOut(void* gdst, void const* smemSrc, uint32_t bytes, uint64_t pol)
    {
        if constexpr(HINTED)
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(gdst), "r"(sAddr(smemSrc)), "r"(bytes), "l"(pol)
                         : "memory");
        else
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(sAddr(smemSrc)), "r"(bytes) : "memory");
    }

    // One CTA = warp 0 (producer: lane 0 issues the loads AND the stores) + NC consumer warps (triad only).
    // Ring of STAGES slots, each [a tile | b tile]; the result overwrites the a tile in place and leaves by a bulk store.
    //   full[s]  : loads of slot s landed (tx bytes)
    //   done[s]  : consumers finished computing slot s (count = consumer threads)  [triad only]
    // The producer re-fills slot s after the store that read it has finished READING shared memory (wait_group.read).
    // OP 0 copy: no consumers at all -- tiles go global -> shared -> global by the copy engine alone.
    template<int OP, bool HINTED, int CONTIG>
    __global__ void __launch_bounds__(288) bulkKernel(double const* a, double const* b, double* c, double s, uint64_t nTiles, uint32_t tileBytes, int stages)
    {
        extern __shared__ __align__(128) unsigned char smem[];
        constexpr int kArrays = OP == 1 ? 2 : 1;
        uint32_t const slotBytes = tileBytes * kArrays;
        uint64_t* full = reinterpret_cast<uint64_t*>(smem + size_t(stages) * slotBytes);
        uint64_t* done = full + stages;
        int const tid = threadIdx.x;
        int const nCons = blockDim.x - 32;

        uint64_t t0, t1, tstep;
        if constexpr(CONTIG)
        {
            uint64_t const per = (nTiles + gridDim.x - 1) / gridDim.x;
            t0 = per * blockIdx.x;
            t1 = t0 + per < nTiles ? t0 + per : nTiles;
            tstep = 1;
        }
        else
        {
            t0 = blockIdx.x;
            t1 = nTiles;
            tstep = gridDim.x;
        }
        uint64_t const myTiles = t0 < t1 ? (t1 - t0 + tstep - 1) / tstep : 0;
        uint32_t const tileElems = tileBytes / 8;

        if(tid == 0)
        {
            for(int i = 0; i < stages; ++i)
            {
                mbarInit(&full[i], 1);
                mbarInit(&done[i], OP == 1 ? nCons : 1);
            }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();

        if(tid < 32)
        {
            if(tid == 0)
            {
                uint64_t const pol = HINTED ? policyEvictFirst() : 0;
                // prologue: fill the ring
                uint64_t issued = 0;
                for(; issued < myTiles && issued < uint64_t(stages); ++issued)
                {
                    uint64_t const t = t0 + issued * tstep;
                    unsigned char* slot = smem + size_t(issued) * slotBytes;
                    mbarExpectTx(&full[issued], slotBytes);
                    bulkIn<HINTED>(slot, a + t * tileElems, tileBytes, &full[issued], pol);
                    if constexpr(OP == 1)
                        bulkIn<HINTED>(slot + tileBytes, b + t * tileElems, tileBytes, &full[issued], pol);
                }
                for(uint64_t k = 0; k < myTiles; ++k)
                {
                    int const sl = int(k % stages);
                    uint32_t const par = uint32_t(k / stages) & 1u;
                    uint64_t const t = t0 + k * tstep;
                    unsigned char* slot = smem + size_t(sl) * slotBytes;
                    if constexpr(OP == 1)
                        mbarWait(&done[sl], par); // consumers wrote the result (and fenced the async proxy)
                    else
                        mbarWait(&full[sl], par);
                    bulkOut<HINTED>(c + t * tileElems, slot, tileBytes, pol);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    // re-fill the slot whose store was committed one iteration ago: allow 1 store group still reading
                    if(k >= 1 && issued < myTiles)
                    {
                        asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                        int const rs = int((k - 1) % stages);
                        uint64_t const tn = t0 + issued * tstep;
                        unsigned char* rslot = smem + size_t(rs) * slotBytes;
                        mbarExpectTx(&full[rs], slotBytes);
                        bulkIn<HINTED>(rslot, a + tn * tileElems, tileBytes, &full[rs], pol);
                        if constexpr(OP == 1)
                            bulkIn<HINTED>(rslot + tileBytes, b + tn * tileElems, tileBytes, &full[rs], pol);
                        ++issued;
                    }
                }
                asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
            }
        }
        else if constexpr(OP == 1)
        {
            int const ct = tid - 32;
            for(uint64_t k = 0; k < myTiles; ++k)
            {
                int const sl = int(k % stages);
                uint32_t const par = uint32_t(k / stages) & 1u;
                double2* ta = reinterpret_cast<double2*>(smem + size_t(sl) * slotBytes);
                double2 const* tb = reinterpret_cast<double2 const*>(smem + size_t(sl) * slotBytes + tileBytes);
                mbarWait(&full[sl], par);
                for(uint32_t i = ct; i < tileElems / 2; i += nCons)
                {
                    double2 x = ta[i];
                    double2 const y = tb[i];
                    x.x = __dadd_rn(x.x, __dmul_rn(s, y.x));
                    x.y = __dadd_rn(x.y, __dmul_rn(s, y.y));
                    ta[i] = x;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbarArrive(&done[sl]);
            }
        }
    }

    // ---------------------------------------------------------------------------------------------- host side
    struct Lab
    {
        unsigned char* pool = nullptr;
        uint64_t n = 0;
        double *a = nullptr, *b = nullptr, *c = nullptr, *sink = nullptr;
        cudaEvent_t e0, e1;
        int sms = 148;
        int steps = 8, warmup = 2;
    };

    template<typename F>
    float timeIt(Lab& L, F&& launch)
    {
        for(int i = 0; i < L.warmup; ++i)
            launch();
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(L.e0));
        for(int i = 0; i < L.steps; ++i)
            launch();
        CK(cudaEventRecord(L.e1));
        CK(cudaEventSynchronize(L.e1));
        CK(cudaGetLastError());
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, L.e0, L.e1));
        return ms / L.steps;
    }

    double bytesOf(int op, uint64_t n)
    {
        return (op == 1 ? 24.0 : op == 3 ? 8.0 : 16.0) * double(n);
    }

    char const* opName(int op)
    {
        return op == 0 ? "copy" : op == 1 ? "triad" : op == 2 ? "read2" : "write1";
    }

    void report(char const* variant, int op, std::string const& params, float ms, uint64_t n)
    {
        std::printf("%-6s %-5s %-58s : %.4f ms %8.1f GB/s\n", opName(op), variant, params.c_str(), ms, bytesOf(op, n) * 1e-6 / ms);
        std::fflush(stdout);
    }

    template<int OP, int LP, int SP, int U, int SHAPE>
    void runLdg(Lab& L, int block, int ctasPerSm, char const* tag = "")
    {
        if((U >= 4 && block > 512) || (U >= 8 && block > 256))
            return;
        uint64_t const nVec = L.n / 4;
        uint64_t const chunk = uint64_t(block) * U;
        unsigned grid = SHAPE == 0 ? unsigned(nVec / chunk) : unsigned(L.sms * ctasPerSm);
        float ms = timeIt(L, [&] { ldgKernel<OP, LP, SP, U, SHAPE><<<grid, block>>>(L.a, L.b, L.c, 2.0, nVec, L.sink); });
        char buf[160];
        std::snprintf(buf, sizeof buf, "shape=%d lp=%d sp=%d U=%d block=%d ctas=%d %s", SHAPE, LP, SP, U, block, SHAPE == 0 ? 0 : ctasPerSm, tag);
        report("ldg", OP, buf, ms, L.n);
    }

    template<int OP, bool HINTED, int CONTIG>
    void runBulk(Lab& L, uint32_t tileBytes, int stages, int ctasPerSm, int consumers)
    {
        int const arrays = OP == 1 ? 2 : 1;
        size_t const smem = size_t(stages) * tileBytes * arrays + size_t(stages) * 16;
        if(smem * ctasPerSm > 227u * 1024u || smem > 227u * 1024u)
            return;
        uint64_t const nTiles = L.n * 8 / tileBytes;
        auto kern = bulkKernel<OP, HINTED, CONTIG>;
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        int const block = 32 + (OP == 1 ? consumers : 0);
        unsigned const grid = unsigned(L.sms * ctasPerSm);
        float ms = timeIt(L, [&] { kern<<<grid, block, smem>>>(L.a, L.b, L.c, 2.0, nTiles, tileBytes, stages); });
        char buf[160];
        std::snprintf(buf, sizeof buf, "tile=%uB stages=%d ctas=%d cons=%d hint=%d contig=%d", tileBytes, stages, ctasPerSm, OP == 1 ? consumers : 0, int(HINTED), CONTIG);
        report("bulk", OP, buf, ms, L.n);
    }

    void place(Lab& L, uint64_t pad)
    {
        uint64_t const bytes = L.n * 8;
        L.a = reinterpret_cast<double*>(L.pool);
        L.b = reinterpret_cast<double*>(L.pool + bytes + pad);
        L.c = reinterpret_cast<double*>(L.pool + 2 * (bytes + pad));
    }

    template<int OP>
    void sweepPolicies(Lab& L)
    {
        // loads x stores, one chunk per block (the shipped shape)
        runLdg<OP, 0, 0, 1, 0>(L, 512, 0);
        runLdg<OP, 1, 1, 1, 0>(L, 512, 0, "(r01 default)");
        runLdg<OP, 1, 0, 1, 0>(L, 512, 0);
        runLdg<OP, 1, 2, 1, 0>(L, 512, 0);
        runLdg<OP, 1, 3, 1, 0>(L, 512, 0);
        runLdg<OP, 1, 4, 1, 0>(L, 512, 0);
        runLdg<OP, 1, 5, 1, 0>(L, 512, 0);
        runLdg<OP, 2, 1, 1, 0>(L, 512, 0);
        runLdg<OP, 2, 2, 1, 0>(L, 512, 0);
        runLdg<OP, 2, 3, 1, 0>(L, 512, 0);
        runLdg<OP, 2, 5, 1, 0>(L, 512, 0);
        runLdg<OP, 3, 1, 1, 0>(L, 512, 0);
        runLdg<OP, 3, 2, 1, 0>(L, 512, 0);
        runLdg<OP, 4, 1, 1, 0>(L, 512, 0);
        runLdg<OP, 4, 2, 1, 0>(L, 512, 0);
        runLdg<OP, 5, 2, 1, 0>(L, 512, 0);
        runLdg<OP, 5, 5, 1, 0>(L, 512, 0);
    }

    template<int OP>
    void sweepShapes(Lab& L)
    {
        for(int block : {256, 512, 1024})
        {
            runLdg<OP, 1, 1, 1, 0>(L, block, 0);
            runLdg<OP, 1, 1, 2, 0>(L, block, 0);
            runLdg<OP, 1, 1, 4, 0>(L, block, 0);
        }
        for(int ctas : {1, 2, 4})
        {
            for(int block : {256, 512, 1024})
            {
                if(block * ctas > 2048)
                    continue;
                runLdg<OP, 1, 1, 1, 1>(L, block, ctas);
                runLdg<OP, 1, 1, 1, 2>(L, block, ctas);
                runLdg<OP, 1, 1, 2, 2>(L, block, ctas);
                runLdg<OP, 1, 1, 4, 2>(L, block, ctas);
                runLdg<OP, 2, 2, 4, 2>(L, block, ctas);
            }
        }
    }

    // finer look at the one-chunk-per-block shape: block size x vectors per thread, streaming vs evict_last loads
    template<int OP>
    void sweepFine(Lab& L)
    {
        for(int rep = 0; rep < 2; ++rep) // twice: run-to-run noise is about 0.3 %
            for(int block : {128, 256, 512, 1024})
            {
                runLdg<OP, 1, 1, 1, 0>(L, block, 0);
                runLdg<OP, 1, 1, 2, 0>(L, block, 0);
                runLdg<OP, 1, 1, 4, 0>(L, block, 0);
                runLdg<OP, 1, 1, 8, 0>(L, block, 0);
                runLdg<OP, 3, 1, 1, 0>(L, block, 0);
                runLdg<OP, 3, 1, 4, 0>(L, block, 0);
                runLdg<OP, 3, 2, 1, 0>(L, block, 0);
            }
    }

    template<int OP>
    void sweepBulk(Lab& L)
    {
        for(uint32_t tile : {4096u, 8192u, 16384u, 32768u})
            for(int stages : {2, 3, 4, 6})
                for(int ctas : {1, 2, 4})
                {
                    runBulk<OP, false, 0>(L, tile, stages, ctas, 256);
                    runBulk<OP, false, 1>(L, tile, stages, ctas, 256);
                }
        // cache-hinted and consumer-count variants of a middle configuration
        for(int ctas : {1, 2, 4})
        {
            runBulk<OP, true, 0>(L, 8192, 4, ctas, 256);
            runBulk<OP, true, 1>(L, 16384, 3, ctas, 256);
            if(OP == 1)
            {
                runBulk<OP, false, 0>(L, 8192, 4, ctas, 128);
                runBulk<OP, false, 0>(L, 16384, 3, ctas, 128);
            }
        }
    }
} // namespace

int main(int argc, char** argv)
{
    Lab L;
    uint64_t n = 1ull << 30;
    std::string what = "all";
    for(int i = 1; i < argc; ++i)
    {
        if(!std::strncmp(argv[i], "--n=", 4))
            n = std::strtoull(argv[i] + 4, nullptr, 10);
        else if(!std::strncmp(argv[i], "--steps=", 8))
            L.steps = std::atoi(argv[i] + 8);
        else
            what = argv[i];
    }
    L.n = n;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    L.sms = prop.multiProcessorCount;
    uint64_t const maxPad = 64ull << 20;
    CK(cudaMalloc(&L.pool, 3 * (n * 8 + maxPad)));
    CK(cudaMalloc(&L.sink, 8));
    CK(cudaMemset(L.pool, 0, 3 * (n * 8 + maxPad)));
    CK(cudaEventCreate(&L.e0));
    CK(cudaEventCreate(&L.e1));
    std::printf("# stream_lab on %s, %d SMs, n = %llu doubles per array, %d timed launches per line\n", prop.name, L.sms, (unsigned long long) n, L.steps);
    place(L, 0);

    if(what == "all" || what == "ref")
    {
        runLdg<2, 1, 1, 1, 0>(L, 512, 0, "(two read streams)");
        runLdg<2, 1, 1, 2, 2>(L, 512, 2, "(two read streams)");
        runLdg<3, 1, 1, 1, 0>(L, 512, 0, "(one write stream)");
        runLdg<3, 1, 0, 1, 0>(L, 512, 0, "(one write stream)");
    }
    if(what == "all" || what == "pad")
    {
        for(uint64_t pad : {0ull, 256ull, 4096ull, 1ull << 16, (1ull << 16) + 4096, 1ull << 20, (1ull << 20) + (1ull << 12), 3ull << 20, (5ull << 20) + (1ull << 13), 1ull << 25})
        {
            place(L, pad);
            char tag[64];
            std::snprintf(tag, sizeof tag, "pad=%llu", (unsigned long long) pad);
            runLdg<0, 1, 1, 1, 0>(L, 512, 0, tag);
            runLdg<1, 1, 1, 1, 0>(L, 512, 0, tag);
        }
        place(L, 0);
    }
    if(what == "all" || what == "policy")
    {
        sweepPolicies<1>(L);
        sweepPolicies<0>(L);
    }
    if(what == "all" || what == "shape")
    {
        sweepShapes<1>(L);
        sweepShapes<0>(L);
    }
    if(what == "fine")
    {
        sweepFine<1>(L);
        sweepFine<0>(L);
    }
    if(what == "all" || what == "bulk")
    {
        sweepBulk<0>(L);
        sweepBulk<1>(L);
    }
    return 0;
}
