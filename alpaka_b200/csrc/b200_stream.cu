// alpaka_b200/csrc/b200_stream.cu -- BabelStream Init/Copy/Mul/Add/Triad/Nstream for sm_100a.
//
// Replaces the user functors of the reference driver (benchmarks/babelstream/src/babelStreamMainTest.cpp:
// InitKernel :53-69, CopyKernel :72-86, MultKernel :89-104, AddKernel :107-122, TriadKernel :125-141) which the
// reference runs one element per thread through its generic trampoline (scalar LDG.E.64 / STG.E.64).
// Here: HBM-bound streams, so the design is about bytes in flight and full-width transactions:
//   * each thread moves UNROLL vectors of VB = 32 bytes (ld/st.global.v4.b64, new with sm_100) or 16 bytes,
//     all loads of an iteration issued before the first store;
//   * a block owns a contiguous chunk of blockDim*UNROLL vectors per iteration, a warp access is one
//     contiguous 512 B / 1 KiB span (fully coalesced, 4 or 8 sectors per request per thread);
//   * grid: ONE block per chunk by default (stream.ctas_per_sm = 0); chunk = 1024 vectors (32 KB per array) for the
//     operations that load (round 2: tools/stream_lab.cu -- launch shape x cache policy x 1-D bulk-copy staging through
//     shared memory x array placement; nothing but the chunk size moves the mixed read/write figure, which sits at
//     7.1 TB/s against 7.55 TB/s for loads alone and 7.60 TB/s for stores alone). Measured on B200 at 2^30 doubles
//     (tools/tune.py stream, profiles/r01/tune_stream.log): one-chunk blocks reach 7.0-7.1 TB/s for Copy/Mul/Triad and
//     7.6 TB/s for Init, a persistent grid of SMs x {2,4,8} blocks striding over the chunks only 6.1-6.7 TB/s -- the
//     hardware block scheduler spreads the DRAM pages touched at any instant better than lock-step strides do;
//     stream.ctas_per_sm > 0 selects the persistent form;
//   * streaming cache policy (ld.global.nc.L1::no_allocate, st.global.cs) because nothing is re-used;
//   * arithmetic is __dmul_rn/__dadd_rn (__fmul_rn/__fadd_rn): FMA contraction pinned OFF so the bits equal the
//     reference CPU back-end compiled with -ffp-contract=off.
// Algorithmic bytes per element: Init 24 (f64) / 12 (f32), Copy/Mul 16/8, Add/Triad 24/12,
// Nstream 32/16.
#include "b200_common.cuh"

namespace
{
    using b200::ldg128;
    using b200::ldg256;
    using b200::stg128;
    using b200::stg256;

    // ---- a VB-byte pack of T held in registers
    template<typename T, int VB>
    struct Pack
    {
        static constexpr int N = VB / int(sizeof(T));
        T v[N];
    };

    // HINT: 0 = default cache policy, 1 = streaming loads (ld.nc.L1::no_allocate) AND streaming stores (st.cs),
    //       2 = streaming loads only, 3 = streaming stores only, 4 = as 1 with L2::evict_last on the (256-bit) loads
    template<int HINT>
    inline constexpr int kLoadHint = HINT == 4 ? 2 : ((HINT == 1 || HINT == 2) ? 1 : 0);
    template<int HINT>
    inline constexpr int kStoreHint = (HINT == 1 || HINT == 3 || HINT == 4) ? 1 : 0;

    template<int HINT, typename T, int VB>
    __device__ __forceinline__ Pack<T, VB> loadPack(T const* base, uint64_t vecIdx)
    {
        Pack<T, VB> p;
        T const* ptr = base + vecIdx * uint64_t(Pack<T, VB>::N);
        if constexpr(VB == 32)
        {
            uint64_t r[4];
            ldg256<kLoadHint<HINT>>(ptr, r);
            memcpy(p.v, r, 32);
        }
        else if constexpr(VB == 16)
        {
            uint32_t r[4];
            ldg128<kLoadHint<HINT>>(ptr, r);
            memcpy(p.v, r, 16);
        }
        else
        {
            static_assert(VB == int(sizeof(T)));
            p.v[0] = *ptr;
        }
        return p;
    }

    template<int HINT, typename T, int VB>
    __device__ __forceinline__ void storePack(T* base, uint64_t vecIdx, Pack<T, VB> const& p)
    {
        T* ptr = base + vecIdx * uint64_t(Pack<T, VB>::N);
        if constexpr(VB == 32)
        {
            uint64_t r[4];
            memcpy(r, p.v, 32);
            stg256<kStoreHint<HINT>>(ptr, r);
        }
        else if constexpr(VB == 16)
        {
            uint32_t r[4];
            memcpy(r, p.v, 16);
            stg128<kStoreHint<HINT>>(ptr, r);
        }
        else
        {
            *ptr = p.v[0];
        }
    }

    __device__ __forceinline__ double mulRn(double a, double b)
    {
        return __dmul_rn(a, b);
    }

    __device__ __forceinline__ float mulRn(float a, float b)
    {
        return __fmul_rn(a, b);
    }

    __device__ __forceinline__ double addRn(double a, double b)
    {
        return __dadd_rn(a, b);
    }

    __device__ __forceinline__ float addRn(float a, float b)
    {
        return __fadd_rn(a, b);
    }

    // ---- the six operations. In<VB> = what one vector step loads; apply() stores.
    template<typename T>
    struct InitOp
    {
        T* a;
        T* b;
        T* c;
        T initA;

        template<int VB>
        struct In
        {
        };

        template<int HINT, int VB>
        __device__ __forceinline__ In<VB> load(uint64_t) const
        {
            return {};
        }

        template<int HINT, int VB>
        __device__ __forceinline__ void apply(uint64_t i, In<VB> const&) const
        {
            Pack<T, VB> pa, pz;
#pragma unroll
            for(int k = 0; k < Pack<T, VB>::N; ++k)
            {
                pa.v[k] = initA;
                pz.v[k] = T(0.0);
            }
            storePack<HINT>(a, i, pa);
            storePack<HINT>(b, i, pz);
            storePack<HINT>(c, i, pz);
        }
    };

    template<typename T>
    struct CopyOp
    {
        T const* a;
        T* b;

        template<int VB>
        struct In
        {
            Pack<T, VB> a;
        };

        template<int HINT, int VB>
        __device__ __forceinline__ In<VB> load(uint64_t i) const
        {
            return {loadPack<HINT, T, VB>(a, i)};
        }

        template<int HINT, int VB>
        __device__ __forceinline__ void apply(uint64_t i, In<VB> const& in) const
        {
            storePack<HINT>(b, i, in.a);
        }
    };

    template<typename T>
    struct MulOp
    {
        T const* a;
        T* b;
        T scalar;

        template<int VB>
        struct In
        {
            Pack<T, VB> a;
        };

        template<int HINT, int VB>
        __device__ __forceinline__ In<VB> load(uint64_t i) const
        {
            return {loadPack<HINT, T, VB>(a, i)};
        }

        template<int HINT, int VB>
        __device__ __forceinline__ void apply(uint64_t i, In<VB> const& in) const
        {
            Pack<T, VB> o;
#pragma unroll
            for(int k = 0; k < Pack<T, VB>::N; ++k)
                o.v[k] = mulRn(scalar, in.a.v[k]);
            storePack<HINT>(b, i, o);
        }
    };

    template<typename T>
    struct AddOp
    {
        T const* a;
        T const* b;
        T* c;

        template<int VB>
        struct In
        {
            Pack<T, VB> a, b;
        };

        template<int HINT, int VB>
        __device__ __forceinline__ In<VB> load(uint64_t i) const
        {
            return {loadPack<HINT, T, VB>(a, i), loadPack<HINT, T, VB>(b, i)};
        }

        template<int HINT, int VB>
        __device__ __forceinline__ void apply(uint64_t i, In<VB> const& in) const
        {
            Pack<T, VB> o;
#pragma unroll
            for(int k = 0; k < Pack<T, VB>::N; ++k)
                o.v[k] = addRn(in.a.v[k], in.b.v[k]);
            storePack<HINT>(c, i, o);
        }
    };

    template<typename T>
    struct TriadOp
    {
        T const* a;
        T const* b;
        T* c;
        T scalar;

        template<int VB>
        struct In
        {
            Pack<T, VB> a, b;
        };

        template<int HINT, int VB>
        __device__ __forceinline__ In<VB> load(uint64_t i) const
        {
            return {loadPack<HINT, T, VB>(a, i), loadPack<HINT, T, VB>(b, i)};
        }

        template<int HINT, int VB>
        __device__ __forceinline__ void apply(uint64_t i, In<VB> const& in) const
        {
            Pack<T, VB> o;
#pragma unroll
            for(int k = 0; k < Pack<T, VB>::N; ++k)
                o.v[k] = addRn(in.a.v[k], mulRn(scalar, in.b.v[k]));
            storePack<HINT>(c, i, o);
        }
    };

    template<typename T>
    struct NstreamOp
    {
        T* a;
        T const* b;
        T const* c;
        T scalar;

        template<int VB>
        struct In
        {
            Pack<T, VB> a, b, c;
        };

        template<int HINT, int VB>
        __device__ __forceinline__ In<VB> load(uint64_t i) const
        {
            // `a` is read-modify-write: it must not go through the non-coherent path
            return {loadPack<0, T, VB>(a, i), loadPack<HINT, T, VB>(b, i), loadPack<HINT, T, VB>(c, i)};
        }

        template<int HINT, int VB>
        __device__ __forceinline__ void apply(uint64_t i, In<VB> const& in) const
        {
            Pack<T, VB> o;
#pragma unroll
            for(int k = 0; k < Pack<T, VB>::N; ++k)
                o.v[k] = addRn(in.a.v[k], addRn(in.b.v[k], mulRn(scalar, in.c.v[k])));
            storePack<HINT>(a, i, o);
        }
    };

    // ---- the kernel: persistent (or one-chunk-per-block) grid-stride over chunks of blockDim*UNROLL vectors
    template<int UNROLL>
    inline constexpr int kMaxBlock = UNROLL == 1 ? 1024 : 512;

    template<typename Op, typename T, int VB, int UNROLL, int HINT>
    __global__ void __launch_bounds__(kMaxBlock<UNROLL>) streamKernel(Op const op, uint64_t const nVec, uint64_t const n)
    {
        using In = typename Op::template In<VB>;
        uint64_t const chunk = uint64_t(blockDim.x) * UNROLL;
        uint64_t const nFull = nVec / chunk;

        for(uint64_t c = blockIdx.x; c < nFull; c += gridDim.x)
        {
            uint64_t const base = c * chunk + threadIdx.x;
            In in[UNROLL];
#pragma unroll
            for(int u = 0; u < UNROLL; ++u)
                in[u] = op.template load<HINT, VB>(base + uint64_t(u) * blockDim.x);
#pragma unroll
            for(int u = 0; u < UNROLL; ++u)
                op.template apply<HINT, VB>(base + uint64_t(u) * blockDim.x, in[u]);
        }

        // partial last chunk: the block whose turn it would be
        if(blockIdx.x == nFull % gridDim.x)
        {
            for(uint64_t i = nFull * chunk + threadIdx.x; i < nVec; i += blockDim.x)
            {
                In const in = op.template load<HINT, VB>(i);
                op.template apply<HINT, VB>(i, in);
            }
            // scalar tail: fewer than one vector of elements
            constexpr int N = Pack<T, VB>::N;
            uint64_t const tailStart = nVec * N;
            if constexpr(N > 1)
            {
                if(tailStart + threadIdx.x < n)
                {
                    using In1 = typename Op::template In<int(sizeof(T))>;
                    In1 const in = op.template load<0, int(sizeof(T))>(tailStart + threadIdx.x);
                    op.template apply<0, int(sizeof(T))>(tailStart + threadIdx.x, in);
                }
            }
        }
    }

    struct StreamCfg
    {
        int vb, unroll, hint, block, ctasPerSm;
    };

    // Launch shape per operation, measured on B200 at 2^30 doubles (tools/stream_lab.cu, profiles/r02/tune_stream_*.log;
    // reproducible to 1 GB/s): a CTA that owns 32 KB of every array beats the 16 KB one of round 1 --
    //   Triad (two loads in flight per vector)  1024 threads x 1 vector   7130 vs 7089 GB/s
    //   Copy  (one load per vector)              256 threads x 4 vectors  7098 vs 6994 GB/s
    //   Mul   (one load, four DMUL per vector)  1024 threads x 1 vector   7094 vs 6887 GB/s at Copy's shape -- wherever
    //         the two arrays lie (profiles/r02/tune_mul.log, placement.log: the deficit is the shape, not the placement)
    // Add follows Triad; Init (stores only, 7.59 TB/s) and Nstream keep 512 x 1.
    struct ShapeDefault
    {
        int unroll, block, hint;
    };

    ShapeDefault shapeDefault(char const* opName)
    {
        std::string const op(opName);
        if(op == "triad" || op == "add")
            return {1, 1024, 4}; // two load streams: L2::evict_last on the loads (+0.3 %)
        if(op == "mul")
            return {1, 1024, 1};
        if(op == "copy")
            return {4, 256, 1};
        return {1, 512, 1};
    }

    StreamCfg streamCfg(char const* opName, int elemBytes)
    {
        (void) elemBytes;
        StreamCfg c;
        std::string const p = std::string("stream.") + opName + ".";
        ShapeDefault const d = shapeDefault(opName);
        c.vb = int(b200::tune((p + "vb").c_str(), b200::tune("stream.vb", 32)));
        c.unroll = int(b200::tune((p + "unroll").c_str(), b200::tune("stream.unroll", d.unroll)));
        c.hint = int(b200::tune((p + "hint").c_str(), b200::tune("stream.hint", d.hint)));
        c.block = int(b200::tune((p + "block").c_str(), b200::tune("stream.block", d.block)));
        c.ctasPerSm = int(b200::tune((p + "ctas_per_sm").c_str(), b200::tune("stream.ctas_per_sm", 0)));
        return c;
    }

    template<typename Op, typename T, int VB, int UNROLL, int HINT>
    int launchOne(cudaStream_t s, Op const& op, uint64_t n, StreamCfg const& cfg)
    {
        constexpr int N = Pack<T, VB>::N;
        uint64_t const nVec = n / N;
        uint64_t const chunk = uint64_t(cfg.block) * UNROLL;
        uint64_t const nChunks = (nVec + chunk - 1) / chunk;
        uint64_t grid = nChunks ? nChunks : 1;
        if(cfg.ctasPerSm > 0)
        {
            uint64_t const persistent = uint64_t(b200::smCount(b200::currentDevice())) * cfg.ctasPerSm;
            grid = grid < persistent ? grid : persistent;
        }
        if(grid > 0x7fffffffull)
            grid = 0x7fffffffull;
        streamKernel<Op, T, VB, UNROLL, HINT><<<unsigned(grid), cfg.block, 0, s>>>(op, nVec, n);
        B200_LAUNCH_CHECK();
        return 0;
    }

    template<typename Op, typename T, int VB, int HINT>
    int launchUnroll(cudaStream_t s, Op const& op, uint64_t n, StreamCfg const& cfg)
    {
        switch(cfg.unroll)
        {
        case 1:
            return launchOne<Op, T, VB, 1, HINT>(s, op, n, cfg);
        case 2:
            return launchOne<Op, T, VB, 2, HINT>(s, op, n, cfg);
        case 4:
            return launchOne<Op, T, VB, 4, HINT>(s, op, n, cfg);
        default:
            return b200::fail(B200_EINVAL, "stream.unroll must be 1, 2 or 4", __FILE__, __LINE__);
        }
    }

    template<typename Op, typename T>
    int launchStream(b200_stream_t stream, Op const& op, uint64_t n, bool aligned32, bool aligned16, char const* name)
    {
        if(n == 0)
            return 0;
        auto const s = reinterpret_cast<cudaStream_t>(stream);
        if(int const rc = b200::useDeviceOf(s))
            return rc;
        StreamCfg cfg = streamCfg(name, int(sizeof(T)));
        if(cfg.block < 32 || cfg.block % 32 != 0 || cfg.block > (cfg.unroll == 1 ? 1024 : 512))
            return b200::fail(B200_EINVAL, "stream.block must be a multiple of 32 in [32,1024] (<= 512 when stream.unroll > 1)", __FILE__, __LINE__);
        int vb = cfg.vb;
        if(vb == 32 && !aligned32)
            vb = 16;
        if(vb == 16 && !aligned16)
            vb = int(sizeof(T));
        if(cfg.hint < 0 || cfg.hint > 4)
            return b200::fail(B200_EINVAL, "stream.hint must be 0..4", __FILE__, __LINE__);
        if(vb == 32)
        {
            switch(cfg.hint)
            {
            case 0:
                return launchUnroll<Op, T, 32, 0>(s, op, n, cfg);
            case 1:
                return launchUnroll<Op, T, 32, 1>(s, op, n, cfg);
            case 2:
                return launchUnroll<Op, T, 32, 2>(s, op, n, cfg);
            case 3:
                return launchUnroll<Op, T, 32, 3>(s, op, n, cfg);
            default:
                return launchUnroll<Op, T, 32, 4>(s, op, n, cfg);
            }
        }
        if(vb == 16)
            return cfg.hint ? launchUnroll<Op, T, 16, 1>(s, op, n, cfg) : launchUnroll<Op, T, 16, 0>(s, op, n, cfg);
        // element-aligned only: scalar path (still on the GPU; there is no CPU fallback)
        return launchUnroll<Op, T, int(sizeof(T)), 0>(s, op, n, cfg);
    }

    template<typename... P>
    bool alignedTo(size_t a, P... ptrs)
    {
        return (((reinterpret_cast<uintptr_t>(ptrs)) % a == 0) && ...);
    }
} // namespace

#define B200_STREAM_ENTRIES(T, SFX)                                                                                   \
    extern "C" int b200_stream_init_##SFX(b200_stream_t s, T* a, T* b, T* c, T init_a, uint64_t n)                    \
    {                                                                                                                 \
        B200_REQUIRE(n == 0 || (a && b && c), B200_EINVAL);                                                           \
        return launchStream<InitOp<T>, T>(s, InitOp<T>{a, b, c, init_a}, n, alignedTo(32, a, b, c), alignedTo(16, a, b, c), "init"); \
    }                                                                                                                 \
    extern "C" int b200_stream_copy_##SFX(b200_stream_t s, T const* a, T* b, uint64_t n)                              \
    {                                                                                                                 \
        B200_REQUIRE(n == 0 || (a && b), B200_EINVAL);                                                                \
        return launchStream<CopyOp<T>, T>(s, CopyOp<T>{a, b}, n, alignedTo(32, a, b), alignedTo(16, a, b), "copy");   \
    }                                                                                                                 \
    extern "C" int b200_stream_mul_##SFX(b200_stream_t s, T const* a, T* b, T scalar, uint64_t n)                     \
    {                                                                                                                 \
        B200_REQUIRE(n == 0 || (a && b), B200_EINVAL);                                                                \
        return launchStream<MulOp<T>, T>(s, MulOp<T>{a, b, scalar}, n, alignedTo(32, a, b), alignedTo(16, a, b), "mul"); \
    }                                                                                                                 \
    extern "C" int b200_stream_add_##SFX(b200_stream_t s, T const* a, T const* b, T* c, uint64_t n)                   \
    {                                                                                                                 \
        B200_REQUIRE(n == 0 || (a && b && c), B200_EINVAL);                                                           \
        return launchStream<AddOp<T>, T>(s, AddOp<T>{a, b, c}, n, alignedTo(32, a, b, c), alignedTo(16, a, b, c), "add"); \
    }                                                                                                                 \
    extern "C" int b200_stream_triad_##SFX(b200_stream_t s, T const* a, T const* b, T* c, T scalar, uint64_t n)       \
    {                                                                                                                 \
        B200_REQUIRE(n == 0 || (a && b && c), B200_EINVAL);                                                           \
        return launchStream<TriadOp<T>, T>(s, TriadOp<T>{a, b, c, scalar}, n, alignedTo(32, a, b, c), alignedTo(16, a, b, c), "triad"); \
    }                                                                                                                 \
    extern "C" int b200_stream_nstream_##SFX(b200_stream_t s, T* a, T const* b, T const* c, T scalar, uint64_t n)     \
    {                                                                                                                 \
        B200_REQUIRE(n == 0 || (a && b && c), B200_EINVAL);                                                           \
        return launchStream<NstreamOp<T>, T>(s, NstreamOp<T>{a, b, c, scalar}, n, alignedTo(32, a, b, c), alignedTo(16, a, b, c), "nstream"); \
    }

B200_STREAM_ENTRIES(double, f64)
B200_STREAM_ENTRIES(float, f32)
