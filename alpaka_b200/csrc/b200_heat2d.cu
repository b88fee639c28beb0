// alpaka_b200/csrc/b200_heat2d.cu -- fused FTCS step of heatEquation2D for sm_100a.
//
// Replaces the two launches per step of the reference (example/heatEquation2D/src/heatEquation2D.cpp:141-168):
//   StencilKernel<(16+2)*(16+2)> (StencilKernel.hpp:31-89): 16x16 tile + halo staged into shared memory by a
//     per-thread linear-index loop with a div/mod per element (mapIdx, :59-65), one output cell per thread;
//   BoundaryKernel (BoundaryKernel.hpp:24-86): a full-grid launch in which only edge blocks work, calling
//     exp/sin per ring cell.
// Here ONE persistent kernel per step:
//   * the (TY+2) x (TX+4) input box of each 32 x 128 output tile is fetched by TMA (cp.async.bulk.tensor.2d,
//     SASS UTMALDG) into a ring of shared-memory stages; a dedicated producer warp issues the copies; a `full`
//     mbarrier per stage carries the transaction count and an `empty` mbarrier (one arrival per consumer warp)
//     hands the stage back, so STAGES-1 tiles (~36 KB each) stay in flight per CTA and no CTA-wide barrier sits in
//     the tile loop. Out-of-range box parts are zero-filled by the TMA unit,
//     so edge tiles need no address arithmetic.
//   * each consumer thread owns a column PAIR and walks RPT (4) rows, keeping the 3x3 neighbourhood in registers: three
//     16-byte LDS per row (conflict-free: a warp reads 512 contiguous bytes), one 16-byte coalesced global
//     store per row. Tiles start on even columns so every store is 16-byte aligned although core cells start at
//     column 1.
//   * the boundary ring is written by the same kernel from separable host tables:
//     exactSolution(x,y,t) = exp(-pi^2 t) * (sin(pi x) + sin(pi y)) == tf * (sx[i] + sy[j]) with sx, sy, tf computed
//     on the host by glibc, which makes the ring (and therefore the whole field) bit-identical to the reference
//     CPU back-end. Corners are never written, as in the reference.
//   * arithmetic order is the reference's, ((((c*k + l*rX) + r*rX) + u*rY) + d*rY), with explicit
//     __dmul_rn/__dadd_rn (no FMA contraction).
// Roofline: HBM, 16 B per core cell per step (one read + one write); halo columns/rows re-read by neighbouring
// tiles (1.096x at 32x128+halo) are served by the 126 MB L2 because neighbouring tiles are processed within the
// same wave / the next tile-row (4 MB apart).
#include "b200_common.cuh"

#include <cuda.h> // CUtensorMap types only; the encoder is fetched through cudaGetDriverEntryPoint

#include <mutex>

namespace
{
    constexpr int TX = 128; // output tile width  (cells)
    constexpr int TY = 32; // output tile height (cells)
    constexpr int BOX_X = TX + 4; // input box: 2 halo columns on each side (keeps 16-byte alignment of pairs)
    constexpr int BOX_Y = TY + 2;
    constexpr int BOX_BYTES = BOX_X * BOX_Y * 8; // 35904
    constexpr int STAGE_BYTES = (BOX_BYTES + 127) / 128 * 128; // TMA destination must be 128-byte aligned
    constexpr int kMaxStages = 6;
    // Consumer threads: one per column pair x row group; RPT rows per thread (template parameter).
    // + one producer warp that only issues TMA copies (warp-specialised, no CTA-wide barrier in the tile loop).
    __host__ __device__ constexpr int consumerThreads(int rpt)
    {
        return (TX / 2) * (TY / rpt);
    }

    constexpr int kMaxWindows = 5;

    // One rectangular window of OUTPUT cells [j0,j1) x [i0,i1) in padded coordinates, cut into TY x TX tiles whose grid
    // starts at (j0, iw0). A launch processes up to kMaxWindows windows, tiles numbered window by window.
    struct HeatWindow
    {
        uint32_t j0, j1, i0, i1;
        uint32_t iw0; // i0 rounded down to even: tile grid origin
        uint32_t tilesX;
        uint32_t tileBegin; // index of this window's first tile
        uint32_t strip; // 1: edge strip of a decomposed tile (its border cells are also stored into the neighbours' ghosts)
    };

    struct HeatArgs
    {
        double* dst;
        size_t pitchElems;
        uint32_t ny, nx;
        HeatWindow win[kMaxWindows];
        int nWin;
        uint32_t totalTiles;
        double k, rX, rY, tf;
        double const* sx;
        double const* sy;
        int edges;
        int stages;
        // ---- fused halo exchange (2-D decomposition); all null / 0 for a stand-alone field
        double* peerDst[4]; // [top, bottom, left, right]: the neighbour's DESTINATION buffer of this step, or null
        uint32_t* peerFlag[4]; // slot in the neighbour's flag array that this rank sets to `step`
        uint32_t* myFlags; // [4] slots set by the neighbours: "my step-s border cells are in your ghosts"
        uint32_t* stripCounter; // strip tiles finished in this launch (reset by the last one)
        uint32_t* status; // != 0: a flag wait timed out
        uint64_t waitNs; // bound of a flag wait (b200::waitLimitNs)
        uint32_t stripTiles;
        uint32_t step; // 1-based time level this launch produces
    };

    __device__ __forceinline__ uint32_t smemAddr(void const* p)
    {
        return static_cast<uint32_t>(__cvta_generic_to_shared(p));
    }

    __device__ __forceinline__ void mbarInit(uint64_t* bar, uint32_t count)
    {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(bar)), "r"(count));
    }

    __device__ __forceinline__ void mbarExpectTx(uint64_t* bar, uint32_t bytes)
    {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(bar)), "r"(bytes)
                     : "memory");
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
            "}\n" ::"r"(smemAddr(bar)),
            "r"(parity)
            : "memory");
    }

    __device__ __forceinline__ void tmaLoad2d(void* smemDst, CUtensorMap const* map, int32_t cx, int32_t cy, uint64_t* bar)
    {
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::
                "r"(smemAddr(smemDst)),
            "l"(reinterpret_cast<uint64_t>(map)),
            "r"(cx),
            "r"(cy),
            "r"(smemAddr(bar))
            : "memory");
    }

    __device__ __forceinline__ double2 lds128(double const* p)
    {
        double2 v;
        asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(smemAddr(p)));
        return v;
    }

    template<int HINT>
    __device__ __forceinline__ void stg2(double* p, double a, double b)
    {
        if constexpr(HINT)
            asm volatile("st.global.cs.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(a), "d"(b) : "memory");
        else
            asm volatile("st.global.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(a), "d"(b) : "memory");
    }

    // StencilKernel.hpp:84-86, left to right, no contraction
    __device__ __forceinline__ double ftcs(double c, double l, double r, double u, double d, double k, double rX, double rY)
    {
        double t = __dmul_rn(c, k);
        t = __dadd_rn(t, __dmul_rn(l, rX));
        t = __dadd_rn(t, __dmul_rn(r, rX));
        t = __dadd_rn(t, __dmul_rn(u, rY));
        t = __dadd_rn(t, __dmul_rn(d, rY));
        return t;
    }

    // Value of a non-core cell, or "no write". BoundaryKernel.hpp:63-84: top/bottom rows for i in 1..nx, left/right
    // columns for j in 1..ny; corners untouched. Sides not flagged in `edges` are ghost cells of a sub-domain.
    __device__ __forceinline__ bool ringValue(HeatArgs const& A, uint32_t j, uint32_t i, double& v)
    {
        bool const rowRing = (j == 0 && (A.edges & B200_EDGE_TOP)) || (j == A.ny + 1 && (A.edges & B200_EDGE_BOTTOM));
        bool const colRing = (i == 0 && (A.edges & B200_EDGE_LEFT)) || (i == A.nx + 1 && (A.edges & B200_EDGE_RIGHT));
        bool const iCore = i >= 1 && i <= A.nx;
        bool const jCore = j >= 1 && j <= A.ny;
        if((rowRing && iCore) || (colRing && jCore))
        {
            v = __dmul_rn(A.tf, __dadd_rn(__ldg(A.sx + i), __ldg(A.sy + j)));
            return true;
        }
        return false;
    }

    __device__ __forceinline__ void mbarArrive(uint64_t* bar)
    {
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smemAddr(bar)) : "memory");
    }


    // ---- the flag protocol of the fused halo exchange, shared by the three kernel families ---------------------------
    // waitForNeighbours: called by ONE thread of a CTA that is about to read ghost cells. The ghosts hold the neighbours'
    // border cells of the previous launch once their flag words say so (which also means the neighbours are done READING
    // the ghost cells this launch overwrites in their other buffer). Bounded wait (b200::waitLimitNs, 60 s by default): a peer
    // that died must not hang this GPU; a timeout raises the status word, which the host layers turn into an error. Ends with the proxy fence TMA needs: ghosts are written through the generic proxy (peer stores),
    // the TMA unit reads them through the async proxy.
    template<int SIDES>
    __device__ __forceinline__ void waitForNeighbours(double* const (&peerDst)[SIDES], uint32_t const* myFlags, uint32_t step, uint32_t* status, uint64_t waitNs)
    {
        for(int side = 0; side < SIDES; ++side)
        {
            if(peerDst[side] == nullptr)
                continue;
            if(!b200::waitFlagAtLeast(myFlags + side, step, 1u, waitNs)) // seen >= step - 1 without underflow
                atomicExch(status, 1u + uint32_t(side));
        }
        asm volatile("fence.proxy.async;" ::: "memory");
    }

    // publishWhenLastStrip: called by ONE thread of a strip CTA after a CTA-wide barrier (every thread has issued the
    // tile's peer stores). Counts the tile; whoever finishes the LAST strip tile of the launch publishes `step` in the
    // neighbours' flag words.
    template<int SIDES>
    __device__ __forceinline__ void publishWhenLastStrip(uint32_t* stripCounter, uint32_t stripTiles, uint32_t* const (&peerFlag)[SIDES], uint32_t step)
    {
        __threadfence_system();
        uint32_t const done = atomicAdd(stripCounter, 1u);
        if(done == stripTiles - 1u)
        {
            __threadfence_system();
            *stripCounter = 0u; // ready for the next launch
            for(int side = 0; side < SIDES; ++side)
                if(peerFlag[side] != nullptr)
                    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(peerFlag[side]), "r"(step) : "memory");
        }
    }

    template<int HINT, int RPT>
    __global__ void __launch_bounds__(consumerThreads(RPT) + 32) heatStepKernel(const __grid_constant__ CUtensorMap mapSrc, HeatArgs const A)
    {
        constexpr int kConsumers = consumerThreads(RPT);
        constexpr int kConsumerWarps = kConsumers / 32;
        extern __shared__ __align__(128) unsigned char smem[];
        __shared__ uint64_t full[kMaxStages];
        __shared__ uint64_t empty[kMaxStages];

        uint32_t const totalTiles = A.totalTiles;
        int const tid = threadIdx.x;
        int const stages = A.stages;

        // tile index -> window and tile origin
        auto tileOrigin = [&](uint32_t t, uint32_t& y0, uint32_t& x0) -> int
        {
            int w = A.nWin - 1;
            while(w > 0 && t < A.win[w].tileBegin)
                --w;
            uint32_t const local = t - A.win[w].tileBegin;
            uint32_t const ty = local / A.win[w].tilesX;
            uint32_t const tx = local - ty * A.win[w].tilesX;
            y0 = A.win[w].j0 + ty * TY;
            x0 = A.win[w].iw0 + tx * TX;
            return w;
        };

        if(tid == 0)
        {
            for(int s = 0; s < stages; ++s)
            {
                mbarInit(&full[s], 1);
                mbarInit(&empty[s], kConsumerWarps);
            }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            // Only strip tiles read ghost cells, and they come first in the tile order: CTAs without a strip tile skip the wait
            if(A.myFlags != nullptr && blockIdx.x < A.stripTiles)
                waitForNeighbours(A.peerDst, A.myFlags, A.step, A.status, A.waitNs);
        }
        __syncthreads();

        if(tid >= kConsumers)
        {
            // ---- producer warp: one lane streams the CTA's tiles through the stage ring
            if(tid == kConsumers)
            {
                int s = 0;
                uint32_t parity = 1; // first pass over the ring: "empty" completes immediately (preceding phase)
                for(uint32_t t = blockIdx.x; t < totalTiles; t += gridDim.x)
                {
                    mbarWait(&empty[s], parity);
                    uint32_t y0, x0;
                    (void) tileOrigin(t, y0, x0);
                    mbarExpectTx(&full[s], BOX_BYTES);
                    tmaLoad2d(smem + size_t(s) * STAGE_BYTES, &mapSrc, int32_t(x0) - 2, int32_t(y0) - 1, &full[s]);
                    if(++s == stages)
                    {
                        s = 0;
                        parity ^= 1u;
                    }
                }
            }
            return;
        }

        // ---- consumers
        int const cp = tid % (TX / 2); // column pair within the tile
        int const rg = tid / (TX / 2); // row group
        int s = 0;
        uint32_t parity = 0;
        for(uint32_t t = blockIdx.x; t < totalTiles; t += gridDim.x)
        {
            uint32_t y0, x0;
            HeatWindow const& W = A.win[tileOrigin(t, y0, x0)];
            mbarWait(&full[s], parity);

            double const* box = reinterpret_cast<double const*>(smem + size_t(s) * STAGE_BYTES);
            // box(r, c): r = output row offset + 1, c = output col offset + 2
            int const r0 = rg * RPT;
            double const* p = box + size_t(r0) * BOX_X + 2 * cp; // row above the first output row, left pair
            double2 up = lds128(p + 2);
            double2 cl = lds128(p + BOX_X), cc = lds128(p + BOX_X + 2), cr = lds128(p + BOX_X + 4);

            uint32_t const gi = x0 + 2 * cp; // global (padded) column of the pair's first cell
            // all TX columns of the tile are core cells inside the window: rows that are core rows inside the window
            // take the vector store without per-cell tests
            bool const colsInside = x0 >= 1 && x0 >= W.i0 && x0 + TX <= A.nx + 1 && x0 + TX <= W.i1;
            double* out = A.dst + size_t(y0 + r0) * A.pitchElems + gi;
#pragma unroll
            for(int r = 0; r < RPT; ++r)
            {
                double const* q = p + size_t(r + 2) * BOX_X;
                double2 const nl = lds128(q), nc = lds128(q + 2), nr = lds128(q + 4);
                double const v0 = ftcs(cc.x, cl.y, cc.y, up.x, nc.x, A.k, A.rX, A.rY);
                double const v1 = ftcs(cc.y, cc.x, cr.x, up.y, nc.y, A.k, A.rX, A.rY);
                uint32_t const gj = y0 + r0 + r;
                bool const rowInside = gj >= 1 && gj <= A.ny && gj >= W.j0 && gj < W.j1;
                if(colsInside && rowInside)
                {
                    stg2<HINT>(out, v0, v1);
                    if(W.strip)
                    {
                        // fused halo exchange, row part: this row is the neighbour's ghost row
                        if(gj == 1u && A.peerDst[0] != nullptr)
                            stg2<0>(A.peerDst[0] + size_t(A.ny + 1u) * A.pitchElems + gi, v0, v1);
                        if(gj == A.ny && A.peerDst[1] != nullptr)
                            stg2<0>(A.peerDst[1] + gi, v0, v1);
                    }
                }
                else
                {
                    bool w0 = false, w1 = false;
                    bool core0 = false, core1 = false;
                    double o0 = v0, o1 = v1;
                    if(gj >= W.j0 && gj < W.j1 && gj <= A.ny + 1)
                    {
                        bool const jCore = gj >= 1 && gj <= A.ny;
                        if(gi >= W.i0 && gi < W.i1 && gi <= A.nx + 1)
                        {
                            core0 = jCore && gi >= 1 && gi <= A.nx;
                            w0 = core0 ? true : ringValue(A, gj, gi, o0);
                        }
                        if(gi + 1 >= W.i0 && gi + 1 < W.i1 && gi + 1 <= A.nx + 1)
                        {
                            core1 = jCore && gi + 1 >= 1 && gi + 1 <= A.nx;
                            w1 = core1 ? true : ringValue(A, gj, gi + 1, o1);
                        }
                    }
                    if(w0 && w1)
                        stg2<HINT>(out, o0, o1);
                    else if(w0)
                        out[0] = o0;
                    else if(w1)
                        out[1] = o1;
                    if(W.strip && (core0 || core1))
                    {
                        // fused halo exchange: the border cells of this tile are the neighbours' ghost cells. Peer stores
                        // (NVLink) straight from the registers that hold the fresh values; tiles have equal extents, so
                        // the neighbour's padded coordinates mirror ours.
                        if(gj == 1u && A.peerDst[0] != nullptr)
                        {
                            double* row = A.peerDst[0] + size_t(A.ny + 1u) * A.pitchElems + gi;
                            if(core0)
                                row[0] = v0;
                            if(core1)
                                row[1] = v1;
                        }
                        if(gj == A.ny && A.peerDst[1] != nullptr)
                        {
                            double* row = A.peerDst[1] + gi;
                            if(core0)
                                row[0] = v0;
                            if(core1)
                                row[1] = v1;
                        }
                        if(core1 && gi + 1u == 1u && A.peerDst[2] != nullptr)
                            A.peerDst[2][size_t(gj) * A.pitchElems + A.nx + 1u] = v1;
                        if(core0 && gi == 1u && A.peerDst[2] != nullptr)
                            A.peerDst[2][size_t(gj) * A.pitchElems + A.nx + 1u] = v0;
                        if(core0 && gi == A.nx && A.peerDst[3] != nullptr)
                            A.peerDst[3][size_t(gj) * A.pitchElems] = v0;
                        if(core1 && gi + 1u == A.nx && A.peerDst[3] != nullptr)
                            A.peerDst[3][size_t(gj) * A.pitchElems] = v1;
                    }
                }
                up = cc;
                cl = nl;
                cc = nc;
                cr = nr;
                out += A.pitchElems;
            }

            // this warp is done reading stage s: hand it back to the producer
            __syncwarp();
            if((tid & 31) == 0)
                mbarArrive(&empty[s]);
            if(++s == stages)
            {
                s = 0;
                parity ^= 1u;
            }

            if(W.strip && A.stripCounter != nullptr)
            {
                // all consumer warps have issued this strip tile's (peer) stores; count the tile, and let whoever
                // finishes the LAST strip tile of the launch publish the time level to the neighbours
                asm volatile("bar.sync 1, %0;" ::"n"(kConsumers) : "memory");
                if(tid == 0)
                    publishWhenLastStrip(A.stripCounter, A.stripTiles, A.peerFlag, A.step);
            }
        }
    }

    // ------------------------------------------------------------------------------------------------------------
    // TWO time levels per launch (temporal blocking). A stand-alone field is HBM-bound at one read + one write per cell
    // per step (16 B); fusing two steps reads level s and writes level s+2 only, i.e. 8 B per cell per step, below
    // what the one-step roofline allows. The arithmetic per cell and per level is the reference's, so results stay
    // bit-identical: level s+1 is never stored, it lives in registers.
    //   * tile TYT x 128 outputs; its (TYT+4) x 132 input box (halo 2) arrives by ONE TMA copy, zero-filled outside the
    //     field; one tile per CTA (enough CTAs are resident per SM to hide the copy).
    //   * a thread owns a column pair (c, c+1) and walks RPT rows. Per output row it loads the next level-s row as
    //     three 16-byte LDS (columns c-2..c+3), computes level s+1 at the FOUR columns c-1..c+2 of the row below
    //     (its two neighbours' values are recomputed instead of exchanged: no shared-memory round trip, no barrier),
    //     then level s+2 at its own pair from the three level-(s+1) rows it keeps in registers.
    //     FP64 work: (4 (RPT+1) + 2 RPT) / RPT stencils per pair-row = 6.25 at RPT 16 (a one-step kernel needs 2 per
    //     step), which the FP64 pipe of sm_100 (64 lanes/clk/SM) still hides behind HBM; shared-memory traffic per two
    //     steps equals the one-step kernel's per ONE step.
    //   * ring cells of the intermediate level take tf1 * (sx[i] + sy[j]) (BoundaryKernel of step s+1), ring cells of
    //     the output tf2 * (...); only tiles touching the field edge run the tested path (EDGE), interior tiles carry
    //     no per-cell tests at all.
    // Only for stand-alone fields (all four sides physical boundaries): a decomposed tile would need ghost cells two
    // deep.
    template<int TYT>
    struct Step2Geom
    {
        static constexpr int kBoxY = TYT + 4;
        static constexpr int kBoxBytes = BOX_X * kBoxY * 8;
    };

    // Rows of the array: [0, ny + 2 padY). padY = 1: the reference layout (ring rows 0 and ny+1). padY = 2: a row SLAB of
    // a field decomposed over several GPUs, ghost rows two deep (rows 0,1 and ny+2,ny+3) because two time levels are
    // advanced per exchange; on a physical side the inner one (1 / ny+2) is the ring and the outer one is unused.
    // Columns always carry the reference's one-cell ring (slabs keep the full width).
    struct Heat2Args
    {
        double* dst;
        size_t pitchElems;
        uint32_t ny, nx;
        uint32_t loY, hiY; // first / last core row = padY, ny + padY - 1
        uint32_t rows; // ny + 2 padY
        uint32_t tilesX;
        double k, rX, rY, tf1, tf2;
        double const* sx;
        double const* sy;
        uint32_t ghostTop, ghostBottom; // 1: that side has a neighbour (its rows next to the core are ghost cells)
        // tile order: strip tile rows first (top: tile row 0; bottom: tile rows tyBot..), then the interior
        uint32_t nTop, nBot, tyBot;
        // ---- fused halo exchange of a slab; null / 0 for a stand-alone field
        double* peerDst[2]; // [top, bottom] neighbour's destination buffer of this launch
        uint32_t* peerFlag[2];
        uint32_t* myFlags; // slots [0] top, [1] bottom, set by the neighbours
        uint32_t* stripCounter;
        uint32_t* status;
        uint64_t waitNs;
        uint32_t stripTiles;
        uint32_t step; // 1-based index of this launch
        uint32_t sendRows; // border rows stored into each neighbour = depth of the ghost rows (>= 2)
    };

    struct Row6
    {
        double2 a, b, c; // columns c-2,c-1 | c,c+1 | c+2,c+3
    };

    __device__ __forceinline__ Row6 ldsRow(double const* q)
    {
        return Row6{lds128(q), lds128(q + 2), lds128(q + 4)};
    }

    // exactSolution on the ring at a given time factor; corners, ghost cells and everything outside the field: 0
    // (never consumed)
    __device__ __forceinline__ double ringOrZero(Heat2Args const& A, uint32_t j, uint32_t i, double tf)
    {
        bool const iCore = i >= 1 && i <= A.nx;
        bool const jCore = j >= A.loY && j <= A.hiY;
        bool const rowRing = (j + 1u == A.loY && !A.ghostTop) || (j == A.hiY + 1u && !A.ghostBottom);
        bool const colRing = i == 0 || i == A.nx + 1;
        if((rowRing && iCore) || (colRing && jCore))
            return __dmul_rn(tf, __dadd_rn(__ldg(A.sx + i), __ldg(A.sy + j)));
        return 0.0;
    }

    // level s+1 at row gj, columns gi-1 .. gi+2, from the level-s rows above (up), at (cur) and below (dn) it
    template<bool EDGE>
    __device__ __forceinline__ void level1Row(
        Heat2Args const& A,
        Row6 const& up,
        Row6 const& cur,
        Row6 const& dn,
        uint32_t gj,
        uint32_t gi,
        double (&o)[4])
    {
        o[0] = ftcs(cur.a.y, cur.a.x, cur.b.x, up.a.y, dn.a.y, A.k, A.rX, A.rY);
        o[1] = ftcs(cur.b.x, cur.a.y, cur.b.y, up.b.x, dn.b.x, A.k, A.rX, A.rY);
        o[2] = ftcs(cur.b.y, cur.b.x, cur.c.x, up.b.y, dn.b.y, A.k, A.rX, A.rY);
        o[3] = ftcs(cur.c.x, cur.b.y, cur.c.y, up.c.x, dn.c.x, A.k, A.rX, A.rY);
        if constexpr(EDGE)
        {
            // the stencil applies on core rows and on the neighbours' rows next to them (ghost rows, two deep, hold
            // level s); (uint32 wrap of gj = y0 - 1 at y0 = 0 fails the range test, as it must)
            bool const jStencil = gj + A.ghostTop >= A.loY && gj <= A.hiY + A.ghostBottom;
#pragma unroll
            for(int q = 0; q < 4; ++q)
            {
                uint32_t const i = gi + uint32_t(q) - 1u;
                if(!(jStencil && i >= 1 && i <= A.nx))
                    o[q] = ringOrZero(A, gj, i, A.tf1);
            }
        }
    }

    template<int HINT, int TYT, int RPT, bool EDGE>
    __device__ __forceinline__ void step2Rows(Heat2Args const& A, double const* box, uint32_t y0, uint32_t x0, int cp, int r0)
    {
        // box(row, col): row = tile row + 2, col = tile column + 2; p points at tile row r0-2, column 2cp-2
        double const* p = box + size_t(r0) * BOX_X + 2 * cp;
        uint32_t const gi = x0 + 2u * uint32_t(cp);
        uint32_t gj = y0 + uint32_t(r0);

        Row6 prev = ldsRow(p); // level s, tile row r0-2
        Row6 cur = ldsRow(p + BOX_X); //          r0-1
        Row6 next = ldsRow(p + 2 * BOX_X); //     r0
        double l1[4];
        level1Row<EDGE>(A, prev, cur, next, gj - 1u, gi, l1); // level s+1, row r0-1 (only columns c, c+1 are used)
        double up0 = l1[1], up1 = l1[2];
        prev = cur;
        cur = next;
        next = ldsRow(p + 3 * BOX_X); // r0+1
        double c1[4];
        level1Row<EDGE>(A, prev, cur, next, gj, gi, c1); // level s+1, row r0
        prev = cur;
        cur = next;

        double* out = A.dst + size_t(gj) * A.pitchElems + gi;
#pragma unroll
        for(int r = 0; r < RPT; ++r)
        {
            next = ldsRow(p + size_t(r + 4) * BOX_X); // level s, tile row r0+r+2
            double n1[4];
            level1Row<EDGE>(A, prev, cur, next, gj + 1u, gi, n1); // level s+1, row r0+r+1
            double v0 = ftcs(c1[1], c1[0], c1[2], up0, n1[1], A.k, A.rX, A.rY);
            double v1 = ftcs(c1[2], c1[1], c1[3], up1, n1[2], A.k, A.rX, A.rY);
            if constexpr(!EDGE)
            {
                stg2<HINT>(out, v0, v1);
            }
            else
            {
                bool const jCore = gj >= A.loY && gj <= A.hiY;
                bool const jRing = (gj + 1u == A.loY && !A.ghostTop) || (gj == A.hiY + 1u && !A.ghostBottom);
                bool w0 = false, w1 = false;
                if(gi <= A.nx + 1u)
                {
                    bool const iCore = gi >= 1 && gi <= A.nx;
                    bool const core = jCore && iCore;
                    bool const ring = (jCore && !iCore) || (jRing && iCore);
                    if(ring)
                        v0 = ringOrZero(A, gj, gi, A.tf2);
                    w0 = core || ring;
                }
                if(gi + 1u <= A.nx + 1u)
                {
                    bool const iCore = gi + 1u <= A.nx; // gi + 1 >= 1 always
                    bool const core = jCore && iCore;
                    bool const ring = (jCore && !iCore) || (jRing && iCore);
                    if(ring)
                        v1 = ringOrZero(A, gj, gi + 1u, A.tf2);
                    w1 = core || ring;
                }
                if(w0 && w1)
                    stg2<HINT>(out, v0, v1);
                else if(w0)
                    out[0] = v0;
                else if(w1)
                    out[1] = v1;
                // fused halo exchange: my first / last `sendRows` core rows (ring columns included) are the neighbour's ghost
                // rows; peer stores (NVLink) from the registers holding the fresh values. Slabs have equal heights, so
                // my row gj is the upper neighbour's row gj + ny and the lower neighbour's row gj - ny.
                if(jCore && (w0 || w1))
                {
                    double* peer = nullptr;
                    if(A.peerDst[0] != nullptr && gj < A.loY + A.sendRows)
                        peer = A.peerDst[0] + size_t(gj + A.ny) * A.pitchElems + gi;
                    else if(A.peerDst[1] != nullptr && gj + A.sendRows > A.hiY)
                        peer = A.peerDst[1] + size_t(gj - A.ny) * A.pitchElems + gi;
                    if(peer != nullptr)
                    {
                        if(w0 && w1)
                            stg2<0>(peer, v0, v1);
                        else if(w0)
                            peer[0] = v0;
                        else
                            peer[1] = v1;
                    }
                }
            }
            up0 = c1[1];
            up1 = c1[2];
#pragma unroll
            for(int q = 0; q < 4; ++q)
                c1[q] = n1[q];
            prev = cur;
            cur = next;
            out += A.pitchElems;
            ++gj;
        }
    }

    template<int HINT, int TYT, int RPT>
    __global__ void __launch_bounds__((TX / 2) * (TYT / RPT)) heatStep2Kernel(const __grid_constant__ CUtensorMap mapSrc, Heat2Args const A)
    {
        extern __shared__ __align__(128) unsigned char smem[];
        __shared__ uint64_t full;
        int const tid = threadIdx.x;
        uint32_t const ord = blockIdx.x / A.tilesX; // tile row in launch order: strips first
        uint32_t const tx = blockIdx.x - ord * A.tilesX;
        bool const strip = ord < A.nTop + A.nBot;
        uint32_t const ty = ord < A.nTop ? ord : (strip ? A.tyBot + (ord - A.nTop) : A.nTop + (ord - A.nTop - A.nBot));
        uint32_t const y0 = ty * TYT, x0 = tx * TX;
        if(tid == 0)
        {
            mbarInit(&full, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            // strip tiles read ghost rows and overwrite the neighbours' ghost rows: wait for the neighbours' previous launch
            if(strip && A.myFlags != nullptr)
                waitForNeighbours(A.peerDst, A.myFlags, A.step, A.status, A.waitNs);
            mbarExpectTx(&full, Step2Geom<TYT>::kBoxBytes);
            tmaLoad2d(smem, &mapSrc, int32_t(x0) - 2, int32_t(y0) - 2, &full);
        }
        __syncthreads();
        mbarWait(&full, 0);

        int const cp = tid % (TX / 2);
        int const r0 = (tid / (TX / 2)) * RPT;
        double const* box = reinterpret_cast<double const*>(smem);
        // every level-(s+1) cell this tile computes, rows y0-1..y0+TYT and columns x0-1..x0+TX, is a core cell, and no
        // row of the tile travels to a neighbour
        bool const interior = !strip && y0 >= A.loY + 1u && y0 + TYT <= A.hiY && x0 >= 2 && x0 + TX <= A.nx;
        if(interior)
            step2Rows<HINT, TYT, RPT, false>(A, box, y0, x0, cp, r0);
        else
            step2Rows<HINT, TYT, RPT, true>(A, box, y0, x0, cp, r0);

        if(strip && A.stripCounter != nullptr)
        {
            // every thread has issued this strip tile's (peer) stores: count the tile; whoever finishes the LAST strip tile
            // of the launch publishes the launch index to the neighbours
            __syncthreads();
            if(tid == 0)
                publishWhenLastStrip(A.stripCounter, A.stripTiles, A.peerFlag, A.step);
        }
    }


    // ------------------------------------------------------------------------------------------------------------
    // S time levels per launch, S = 3 or 4 (deeper temporal blocking; 4 is the default depth). Recomputing the neighbours'
    // columns as the two-level kernel does costs (S+1) S stencils per pair-row and would make three levels FP64-bound, so
    // here every thread computes ONLY its own column pair at every level and takes what it needs from the adjacent lanes
    // by warp shuffle. A warp owns a window of 64 columns; what the edge lanes cannot obtain becomes invalid one column per
    // level, so after S levels the warp stores the inner 64 - 2M columns (M = 2 (S/2): pairs stay aligned) and neighbouring
    // warps' windows overlap by 2M columns -- 6.7 % redundant work at S = 3, 14 % at S = 4.
    //   * tile = NWY x RPT rows by NWX x (64 - 2M) columns; its (rows + 2S) x (columns + 2M + 4) input box arrives by ONE
    //     TMA copy (S = 3: 128 columns = 1 KB per row), zero-filled outside the field;
    //   * a thread walks down its rows once: per input row 3 LDS.128 (own pair + the pairs left and right of it), then one
    //     pair of stencils per level as soon as three rows of the level below exist; rows roll through registers;
    //   * every product v*rX / v*rY is computed ONCE per value and shared by the two cells / two rows whose stencils use it
    //     (RowN): 3 multiplications + 4 additions per cell and level instead of 5 + 4, the same IEEE products and the same
    //     order of additions, hence the same bits. Lanes exchange the products (two 64-bit shuffles per level and row).
    //     FP64 instructions per output pair-row at S = 4, RPT 16: about 62 for FOUR steps (the one-step kernel issues 18 per
    //     step); that is what made four levels per launch faster than three.
    //   * ring cells of intermediate level k take tf[k] * (sx + sy); only tiles touching the field edge test per cell.
    // Stand-alone fields (padY = 1) and row slabs whose ghost rows are at least S deep (padY >= S, fused halo exchange).
    constexpr int kMaxLevels = 8; // the tile kernel below: 3 or 4; the walker further down: 4, 6 or 8

    template<int S>
    struct StepNGeom
    {
        static constexpr int M = 2 * (S / 2); // columns a warp window loses on each side
        static constexpr int NWX = 2; // warps side by side
        static constexpr int WOUT_WARP = 64 - 2 * M;
        static constexpr int WOUT = NWX * WOUT_WARP; // output columns per tile
        static constexpr int BOXX = WOUT + 2 * M + 4; // + M each side + 2 each side so that pairs stay 16-byte aligned
    };

    // Rows of the array: [0, ny + 2 padY). padY = 1: the reference layout. padY = S: a row slab whose ghost rows are S deep
    // (see Heat2Args); columns always carry the one-cell ring.
    struct HeatNArgs
    {
        double* dst;
        size_t pitchElems;
        uint32_t ny, nx;
        int32_t loY, hiY; // first / last core row = padY, ny + padY - 1
        uint32_t tilesX;
        double k, rX, rY;
        double tf[kMaxLevels]; // tf[l-1]: time factor of the l-th level of this launch
        double const* sx;
        double const* sy;
        int32_t ghostTop, ghostBottom; // 1: that side has a neighbour
        uint32_t nTop, nBot, tyBot; // tile order: strip tile rows first, as in Heat2Args
        // ---- fused halo exchange of a slab; null / 0 for a stand-alone field
        double* peerDst[2];
        uint32_t* peerFlag[2];
        uint32_t* myFlags;
        uint32_t* stripCounter;
        uint32_t* status;
        uint64_t waitNs;
        uint32_t stripTiles;
        uint32_t step; // 1-based index of this launch
        int32_t sendRows; // border rows stored into each neighbour = depth of the ghost rows (>= S)
    };

    // One row of one level as a thread holds it: its own pair and the PRODUCTS the stencil needs from it. Every product
    // v*rX serves two cells (as the right term of the cell on its left and the left term of the cell on its right), every
    // v*rY two rows (down term of the row above, up term of the row below): computing each once leaves 3 multiplications
    // per cell instead of 5 -- the same IEEE products, so the sums are bit-identical -- and lanes exchange products, not
    // values.
    struct RowN
    {
        double x, y; // own pair
        double xh, yh; // x*rX, y*rX
        double xv, yv; // x*rY, y*rY
        double lh, rh; // left neighbour's y*rX, right neighbour's x*rX
    };

    // SQ: rX == rY bit for bit (square cells, dx == dy: every BASELINE.json heat configuration). Then v*rY IS v*rX -- the
    // same IEEE product -- and one multiplication serves both directions: 2 multiplications + 4 additions per cell and level.
    template<bool SQ>
    __device__ __forceinline__ RowN makeRowN(double x, double y, double rX, double rY)
    {
        RowN r;
        r.x = x;
        r.y = y;
        r.xh = __dmul_rn(x, rX);
        r.yh = __dmul_rn(y, rX);
        r.xv = SQ ? r.xh : __dmul_rn(x, rY);
        r.yv = SQ ? r.yh : __dmul_rn(y, rY);
        r.lh = r.rh = 0.0;
        return r;
    }

    // StencilKernel.hpp:84-86 on precomputed products: ((((c*k + l*rX) + r*rX) + u*rY) + d*rY)
    __device__ __forceinline__ double ftcsP(double c, double k, double lh, double rh, double uv, double dv)
    {
        double t = __dmul_rn(c, k);
        t = __dadd_rn(t, lh);
        t = __dadd_rn(t, rh);
        t = __dadd_rn(t, uv);
        t = __dadd_rn(t, dv);
        return t;
    }

    __device__ __forceinline__ double shflUp1(double v)
    {
        return __shfl_up_sync(0xffffffffu, v, 1);
    }

    __device__ __forceinline__ double shflDown1(double v)
    {
        return __shfl_down_sync(0xffffffffu, v, 1);
    }

    // A cell of a level that the stencil does not produce: exactSolution where it is a ring cell -- the physical ring
    // rows over core columns, the ring columns over the rows [jLo, jHi] on which this level is defined (core rows, and for a
    // slab the neighbour's rows it still needs) --, 0 anywhere else (never consumed). j, i may be out of range.
    __device__ __forceinline__ double ringOrZeroN(HeatNArgs const& A, int32_t j, int32_t i, int32_t jLo, int32_t jHi, double tf)
    {
        bool const iCore = i >= 1 && i <= int32_t(A.nx);
        bool const jDefined = j >= jLo && j <= jHi;
        bool const rowRing = (j == A.loY - 1 && !A.ghostTop) || (j == A.hiY + 1 && !A.ghostBottom);
        bool const colRing = i == 0 || i == int32_t(A.nx) + 1;
        if((rowRing && iCore) || (colRing && jDefined))
            return __dmul_rn(tf, __dadd_rn(__ldg(A.sx + i), __ldg(A.sy + j)));
        return 0.0;
    }

    template<int S, int RPT, int NWY, bool EDGE, bool SQ>
    __device__ __forceinline__ void stepNRows(HeatNArgs const& A, double const* box, int32_t y0, int32_t x0, int wx, int wy, int lane)
    {
        using G = StepNGeom<S>;
        int const r0 = wy * RPT;
        // box(row, col): row = tile row + S, col = global column - (x0 - M - 2)
        int const bc = 2 + wx * G::WOUT_WARP + 2 * lane; // box column of the own pair
        int32_t const gi = x0 - G::M + wx * G::WOUT_WARP + 2 * lane; // its global column
        double const* p = box + size_t(r0) * G::BOXX + (bc - 2);
        bool const storeLane = lane >= G::M / 2 && lane < 32 - G::M / 2;

        double U[S][2]; // level l: the row two above the newest one -- only its x*rY, y*rY are still needed
        RowN C[S]; //              the row one above the newest one
#pragma unroll
        for(int i = 0; i < RPT + 2 * S; ++i)
        {
            // level 0, tile row r0 - S + i
            double2 const b = lds128(p + size_t(i) * G::BOXX + 2);
            RowN N = makeRowN<SQ>(b.x, b.y, A.rX, A.rY);
            if constexpr(S % 2 == 0)
            {
                // The neighbours' products come from the adjacent lanes, as at every other level: one LDS.128 per row and no
                // extra multiplication. The window then loses one column per side at level 1 already, S columns in all --
                // exactly the M = S columns an even S gives up anyway (lanes M/2 .. 31 - M/2 store).
                N.lh = shflUp1(N.yh);
                N.rh = shflDown1(N.xh);
            }
            else
            {
                // odd S: M = S - 1, the window cannot afford to lose a column at level 1; the cells left and right of the pair
                // are read from the box (two 8-byte loads at a 16-byte lane stride: 2-way bank conflicts, ncu r02)
                double2 const a = lds128(p + size_t(i) * G::BOXX), c = lds128(p + size_t(i) * G::BOXX + 4);
                N.lh = __dmul_rn(a.y, A.rX);
                N.rh = __dmul_rn(c.x, A.rX);
            }
#pragma unroll
            for(int l = 0; l < S; ++l)
            {
                // N is row (r0 - S + i - l) of level l. With three rows of level l, its middle row yields level l+1.
                if(i < 2 * (l + 1))
                {
                    // not enough rows yet: just roll
                    U[l][0] = C[l].xv;
                    U[l][1] = C[l].yv;
                    C[l] = N;
                    break;
                }
                double vx = ftcsP(C[l].x, A.k, C[l].lh, C[l].yh, U[l][0], N.xv);
                double vy = ftcsP(C[l].y, A.k, C[l].xh, C[l].rh, U[l][1], N.yv);
                U[l][0] = C[l].xv;
                U[l][1] = C[l].yv;
                C[l] = N;
                int32_t const gj = y0 + r0 - S + i - (l + 1); // row of the new level-(l+1) values
                // rows on which level l+1 is defined: the core rows, and on a side with a neighbour the S-(l+1) rows
                // beyond them that deeper levels still need
                int32_t const jLo = A.loY - (S - (l + 1)) * A.ghostTop, jHi = A.hiY + (S - (l + 1)) * A.ghostBottom;
                if(l + 1 < S)
                {
                    if constexpr(EDGE)
                    {
                        bool const jDefined = gj >= jLo && gj <= jHi;
                        if(!(jDefined && gi >= 1 && gi <= int32_t(A.nx)))
                            vx = ringOrZeroN(A, gj, gi, jLo, jHi, A.tf[l]);
                        if(!(jDefined && gi + 1 >= 1 && gi + 1 <= int32_t(A.nx)))
                            vy = ringOrZeroN(A, gj, gi + 1, jLo, jHi, A.tf[l]);
                    }
                    N = makeRowN<SQ>(vx, vy, A.rX, A.rY);
                    N.lh = shflUp1(N.yh);
                    N.rh = shflDown1(N.xh);
                }
                else
                {
                    // level S: the output row
                    double* out = A.dst + int64_t(gj) * int64_t(A.pitchElems) + gi;
                    if constexpr(!EDGE)
                    {
                        if(storeLane)
                            stg2<1>(out, vx, vy);
                    }
                    else if(storeLane)
                    {
                        bool const jCore = gj >= A.loY && gj <= A.hiY;
                        bool const jRing = (gj == A.loY - 1 && !A.ghostTop) || (gj == A.hiY + 1 && !A.ghostBottom);
                        bool w0 = false, w1 = false;
                        if((jCore || jRing) && gi >= 0 && gi <= int32_t(A.nx) + 1)
                        {
                            bool const iCore = gi >= 1 && gi <= int32_t(A.nx);
                            if(!(jCore && iCore))
                                vx = ringOrZeroN(A, gj, gi, jLo, jHi, A.tf[S - 1]);
                            w0 = jCore || iCore; // core or ring; corners have neither
                        }
                        if((jCore || jRing) && gi + 1 >= 0 && gi + 1 <= int32_t(A.nx) + 1)
                        {
                            bool const iCore = gi + 1 >= 1 && gi + 1 <= int32_t(A.nx);
                            if(!(jCore && iCore))
                                vy = ringOrZeroN(A, gj, gi + 1, jLo, jHi, A.tf[S - 1]);
                            w1 = jCore || iCore;
                        }
                        if(w0 && w1)
                            stg2<1>(out, vx, vy);
                        else if(w0)
                            out[0] = vx;
                        else if(w1)
                            out[1] = vy;
                        // fused halo exchange: my first / last `sendRows` core rows (ring columns included) are the neighbour's
                        // ghost rows; slabs have equal heights, so my row gj is the upper neighbour's row gj + ny and the lower
                        // neighbour's row gj - ny
                        if(jCore && (w0 || w1))
                        {
                            double* peer = nullptr;
                            if(A.peerDst[0] != nullptr && gj < A.loY + A.sendRows)
                                peer = A.peerDst[0] + int64_t(gj + int32_t(A.ny)) * int64_t(A.pitchElems) + gi;
                            else if(A.peerDst[1] != nullptr && gj > A.hiY - A.sendRows)
                                peer = A.peerDst[1] + int64_t(gj - int32_t(A.ny)) * int64_t(A.pitchElems) + gi;
                            if(peer != nullptr)
                            {
                                if(w0 && w1)
                                    stg2<0>(peer, vx, vy);
                                else if(w0)
                                    peer[0] = vx;
                                else
                                    peer[1] = vy;
                            }
                        }
                    }
                }
            }
        }
    }

    template<int S, int RPT, int NWY, bool SQ, int MINB = 8 / NWY>
    // minimum CTAs per SM pinned so that no variant exceeds 128 registers (4 CTAs of 128 threads / 2 of 256: what the
    // 40 KB / 72 KB boxes admit anyway); MINB = 5 (heat.stepn_ctas = 5, square cells, default shape) trades 96 registers + a few spills for a fifth resident CTA
    __global__ void __launch_bounds__(32 * StepNGeom<S>::NWX * NWY, MINB) heatStepNKernel(const __grid_constant__ CUtensorMap mapSrc, HeatNArgs const A)
    {
        using G = StepNGeom<S>;
        constexpr int TYT = NWY * RPT;
        constexpr uint32_t kBoxBytes = uint32_t(G::BOXX) * (TYT + 2 * S) * 8;
        extern __shared__ __align__(128) unsigned char smem[];
        __shared__ uint64_t full;
        int const tid = threadIdx.x;
        uint32_t const ord = blockIdx.x / A.tilesX; // tile row in launch order: strips first
        uint32_t const tx = blockIdx.x - ord * A.tilesX;
        bool const strip = ord < A.nTop + A.nBot;
        uint32_t const ty = ord < A.nTop ? ord : (strip ? A.tyBot + (ord - A.nTop) : A.nTop + (ord - A.nTop - A.nBot));
        int32_t const y0 = int32_t(ty) * TYT, x0 = int32_t(tx) * G::WOUT;
        if(tid == 0)
        {
            mbarInit(&full, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            // strip tiles read ghost rows and overwrite the neighbours' ghost rows: wait for the neighbours' previous launch
            if(strip && A.myFlags != nullptr)
                waitForNeighbours(A.peerDst, A.myFlags, A.step, A.status, A.waitNs);
            mbarExpectTx(&full, kBoxBytes);
            tmaLoad2d(smem, &mapSrc, x0 - G::M - 2, y0 - S, &full);
        }
        __syncthreads();
        mbarWait(&full, 0);

        int const warp = tid / 32, lane = tid % 32;
        int const wx = warp % G::NWX, wy = warp / G::NWX;
        double const* box = reinterpret_cast<double const*>(smem);
        // every intermediate cell this tile computes -- rows y0-(S-1) .. y0+TYT+S-2, columns x0-M .. x0+WOUT+M-1 -- is a
        // core cell, and no row of the tile travels to a neighbour
        bool const interior = !strip && y0 - (S - 1) >= A.loY && y0 + TYT + S - 2 <= A.hiY && x0 - G::M >= 1
                              && x0 + G::WOUT + G::M - 1 <= int32_t(A.nx);
        if(interior)
            stepNRows<S, RPT, NWY, false, SQ>(A, box, y0, x0, wx, wy, lane);
        else
            stepNRows<S, RPT, NWY, true, SQ>(A, box, y0, x0, wx, wy, lane);

        if(strip && A.stripCounter != nullptr)
        {
            __syncthreads();
            if(tid == 0)
                publishWhenLastStrip(A.stripCounter, A.stripTiles, A.peerFlag, A.step);
        }
    }

    // ------------------------------------------------------------------------------------------------------------
    // S time levels per launch by WALKERS, S even (4, 6, 8): the round-2 replacement of the tile kernel above for even S.
    // ncu on heatStepNKernel<4,16,2> (profiles/r02): FP64 pipe 70 %, DRAM 66 %, 1.45 x the minimal DP instructions -- a
    // 16-row tile recomputes (16 + 2S)/16 of the rows of the lower levels and reloads as many input rows; and with two
    // columns per lane the four SHFL.32 per level, the register moves around them and the addressing take as many issue slots
    // as the DP instructions themselves (an SMSP issues one instruction per cycle, the FP64 pipe takes one per two).
    // Here ONE WARP walks down a 128-column window (FOUR columns per lane) over a tall row segment and never looks at a row
    // twice: the row redundancy is 2S per SEGMENT, the column redundancy 128/(128 - 8 ceil(S/4)), and the shuffles are
    // shared by twice as many cells.
    //   * no CTA-wide synchronisation at all: every warp owns a ring of ST shared-memory stages of R rows x 128 columns
    //     (1 KB per row), filled by its own TMA copies (one elected lane, one mbarrier per stage, refilled as soon as
    //     the warp has consumed the stage); a CTA is just four independent walkers side by side;
    //   * per input row two LDS.128 per lane, then one "arrival" per level: the state a level keeps between rows is 2 doubles
    //     per cell -- the partial sum q = ((c*k + l*rX) + r*rX) + u*rY of the row waiting for the row below it, and that
    //     row's product with rY for its successor -- instead of the 4 of the tile kernel (RowN + U), because everything a
    //     stencil takes from its own row and the row above is folded in when the row is CREATED; the row below then
    //     completes it with one addition. Same IEEE products, same order of additions (StencilKernel.hpp:84-86 left to
    //     right), same bits. The dependent chain from level to level is add -> mul (was mul -> 4 add);
    //   * 6 DP instructions per cell and level on square cells (8 mul + 16 add per lane quad), two 64-bit shuffles per
    //     level and lane (the outer products of the quad);
    //   * rows of a level the walk has not produced yet (the first 2S input rows of a segment) flow through as finite
    //     garbage that no stored cell depends on: no prologue code, the segment's row range guards the stores. Chunks whose
    //     rows are all core rows run bare (interior windows) or with the ring columns patched per lane (edge windows; the
    //     row factors sy[j] roll through registers one row ahead, so no load sits in the dependent chain); rows next to the
    //     ring, ghost zones and slab strips take the CAREFUL path (ring values, store guards, peer stores);
    //   * slabs: the strip rows (my first / last G core rows = the neighbours' ghost rows) are short segments of their own at
    //     the front of the grid, so the flags go out while the interior walkers are still under way.
    struct HeatWArgs : HeatNArgs
    {
        uint32_t nWin; // column windows
        uint32_t nEdgeRight; // windows at the right end that hold ring columns or columns beyond the field (window 0 is the left one)
        uint32_t nSegAll; // segments per window, strips included
        uint32_t nFront; // segments at the front of the segment order: slab strips (rows sent to a neighbour) and, when the
                         // interior runs in the bare-only kernel, the S-row bands next to a physical top / bottom ring
        uint32_t frontIsStrip; // bit k: front segment k is a strip (flag wait, peer stores, counted for the flag)
        int32_t frontY0[2], frontY1[2]; // their output rows [y0, y1)
        int32_t intY0, intY1, segRows; // interior output rows [intY0, intY1) in segments of segRows
        uint32_t nWalkers;
        int32_t rows; // rows of the array (sy has that many entries)
        // columns of the array: [0, nx + 2 padX). padX = 1: the reference layout / row slabs (ring columns 0 and nx+1).
        // padX = G: a 2-D tile whose ghost columns are G deep (b200_heat2d_tile_plan_create), as the rows of a slab
        int32_t loX, hiX; // first / last core column = padX, nx + padX - 1
        int32_t ghostLeft, ghostRight; // 1: that side has a neighbour
        uint32_t const* colFlags[2]; // my flag words set by the left / right neighbour's column exchange, or null
        int32_t align32; // 1: rows of dst (and of the peers' arrays) are 32-byte aligned -> 256-bit stores
        double sxLeft, sxRight; // sx[0], sx[nx+1]: the ring columns' factors
    };

    // ringOrZeroN with the columns generalised: a cell of a level the stencil does not produce -- exactSolution where it is a
    // boundary cell (physical ring rows over the columns [iLo, iHi] on which the level is defined, physical ring columns
    // over the rows [jLo, jHi]), 0 anywhere else (never consumed).
    __device__ __forceinline__ double ringOrZeroW(HeatWArgs const& A, int32_t j, int32_t i, int32_t jLo, int32_t jHi, int32_t iLo, int32_t iHi, double tf)
    {
        bool const iDefined = i >= iLo && i <= iHi;
        bool const jDefined = j >= jLo && j <= jHi;
        bool const rowRing = (j == A.loY - 1 && !A.ghostTop) || (j == A.hiY + 1 && !A.ghostBottom);
        bool const colRing = (i == A.loX - 1 && !A.ghostLeft) || (i == A.hiX + 1 && !A.ghostRight);
        if((rowRing && iDefined) || (colRing && jDefined))
            return __dmul_rn(tf, __dadd_rn(__ldg(A.sx + i), __ldg(A.sy + j)));
        return 0.0;
    }

    constexpr int kWalkCols = 4; // columns per lane

    template<int S>
    struct WalkGeom
    {
        static constexpr int LOST = (S + 3) / 4; // lanes per side whose columns the levels invalidate
        static constexpr int WW = 32 * kWalkCols - 2 * kWalkCols * LOST; // output columns per window
    };

    template<int S>
    struct WalkState
    {
        double q[S][kWalkCols]; // the waiting row of level l: ((c*k + left*rX) + right*rX) + up*rY, lacking only down*rY
        double u[S][kWalkCols]; // its products with rY: the up terms of the row that arrives next
    };

    // A new row v[] of some level arrives: it completes the waiting row (-> one row of the next level, returned in v) and
    // becomes the waiting row itself.
    template<bool SQ>
    __device__ __forceinline__ void walkArrive(double (&v)[kWalkCols], double k, double rX, double rY, double (&q)[kWalkCols], double (&u)[kWalkCols])
    {
        double h[kWalkCols], w[kWalkCols], o[kWalkCols];
#pragma unroll
        for(int i = 0; i < kWalkCols; ++i)
        {
            h[i] = __dmul_rn(v[i], rX);
            w[i] = SQ ? h[i] : __dmul_rn(v[i], rY);
        }
        double const lh = shflUp1(h[kWalkCols - 1]), rh = shflDown1(h[0]);
#pragma unroll
        for(int i = 0; i < kWalkCols; ++i)
        {
            o[i] = __dadd_rn(q[i], w[i]);
            double const left = i == 0 ? lh : h[i - 1], right = i == kWalkCols - 1 ? rh : h[i + 1];
            double const p = __dadd_rn(__dadd_rn(__dmul_rn(v[i], k), left), right);
            q[i] = __dadd_rn(p, u[i]);
            u[i] = w[i];
        }
#pragma unroll
        for(int i = 0; i < kWalkCols; ++i)
            v[i] = o[i];
    }

    enum WalkMode
    {
        kWalkBare, // interior window, all rows core rows: no tests but the segment's row range
        kWalkEdgeCols, // edge window, all rows core rows: ring columns per lane
        kWalkCareful // anything: ring rows, ghost zones, slab strips with peer stores
    };

    template<int HINT>
    __device__ __forceinline__ void storeQuad(double* p, double const (&v)[kWalkCols], bool align32)
    {
        if(align32)
        {
            if constexpr(HINT)
                asm volatile("st.global.cs.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]) : "memory");
            else
                asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]) : "memory");
        }
        else
        {
            stg2<HINT>(p, v[0], v[1]);
            stg2<HINT>(p + 2, v[2], v[3]);
        }
    }

    // One input row j0 (level 0) through all S levels; stores row j0 - S of level S into `out` (= its address for this
    // lane's quad). [ya, yb) are the output rows of the walker's segment. ringMask / validMask: bit i = column gi + i is a
    // ring column / lies in [0, nx+1] (edge windows). syw[l] = sy[j0 - 1 - l] (edge windows, all rows core).
    template<int S, bool SQ, WalkMode MODE>
    __device__ __forceinline__ void walkRow(HeatWArgs const& A, WalkState<S>& st, double (&v)[kWalkCols], int32_t j0, int32_t gi, bool storeLane, int32_t ya, int32_t yb, double* out, uint32_t ringMask, uint32_t validMask, double const (&syw)[S])
    {
#pragma unroll
        for(int l = 0; l < S; ++l)
        {
            walkArrive<SQ>(v, A.k, A.rX, A.rY, st.q[l], st.u[l]);
            int32_t const gj = j0 - (l + 1); // row of the new level-(l+1) values, now in v
            if constexpr(MODE == kWalkEdgeCols)
            {
                // rows are core rows: the only cells the stencil does not produce are the ring columns
                if(ringMask != 0u)
                {
#pragma unroll
                    for(int i = 0; i < kWalkCols; ++i)
                        if(ringMask & (1u << i))
                            v[i] = __dmul_rn(A.tf[l], __dadd_rn(gi + i < A.loX ? A.sxLeft : A.sxRight, syw[l]));
                }
            }
            if(l + 1 < S)
            {
                if constexpr(MODE == kWalkCareful)
                {
                    // rows on which level l+1 is defined: the core rows, and on a side with a neighbour the S-(l+1) rows
                    // beyond them that deeper levels still need
                    int32_t const jLo = A.loY - (S - (l + 1)) * A.ghostTop, jHi = A.hiY + (S - (l + 1)) * A.ghostBottom;
                    int32_t const iLo = A.loX - (S - (l + 1)) * A.ghostLeft, iHi = A.hiX + (S - (l + 1)) * A.ghostRight;
                    bool const jDefined = gj >= jLo && gj <= jHi;
#pragma unroll
                    for(int i = 0; i < kWalkCols; ++i)
                        if(!(jDefined && gi + i >= iLo && gi + i <= iHi))
                            v[i] = ringOrZeroW(A, gj, gi + i, jLo, jHi, iLo, iHi, A.tf[l]);
                }
            }
            else if(storeLane && gj >= ya && gj < yb)
            {
                if constexpr(MODE == kWalkBare)
                    storeQuad<1>(out, v, A.align32 != 0);
                else if constexpr(MODE == kWalkEdgeCols)
                {
                    if(validMask == 0xfu)
                        storeQuad<1>(out, v, A.align32 != 0);
                    else
                    {
#pragma unroll
                        for(int i = 0; i < kWalkCols; ++i)
                            if(validMask & (1u << i))
                                out[i] = v[i];
                    }
                }
                else
                {
                    bool const jCore = gj >= A.loY && gj <= A.hiY;
                    bool const jRing = (gj == A.loY - 1 && !A.ghostTop) || (gj == A.hiY + 1 && !A.ghostBottom);
                    uint32_t wr = 0u; // cells to write: core or ring; corners and cells beyond the field have neither
#pragma unroll
                    for(int i = 0; i < kWalkCols; ++i)
                    {
                        int32_t const c = gi + i;
                        bool const iCore = c >= A.loX && c <= A.hiX;
                        bool const iRing = (c == A.loX - 1 && !A.ghostLeft) || (c == A.hiX + 1 && !A.ghostRight);
                        if((jCore || jRing) && (iCore || iRing))
                        {
                            if(!(jCore && iCore))
                                v[i] = ringOrZeroW(A, gj, c, A.loY, A.hiY, A.loX, A.hiX, A.tf[S - 1]);
                            if(jCore || iCore)
                                wr |= 1u << i;
                        }
                    }
                    // fused halo exchange, as in the tile kernel: my first / last `sendRows` core rows (ring columns
                    // included) are the neighbours' ghost rows
                    double* peer = nullptr;
                    if(jCore && wr != 0u)
                    {
                        if(A.peerDst[0] != nullptr && gj < A.loY + A.sendRows)
                            peer = A.peerDst[0] + int64_t(gj + int32_t(A.ny)) * int64_t(A.pitchElems) + gi;
                        else if(A.peerDst[1] != nullptr && gj > A.hiY - A.sendRows)
                            peer = A.peerDst[1] + int64_t(gj - int32_t(A.ny)) * int64_t(A.pitchElems) + gi;
                    }
                    if(wr == 0xfu)
                    {
                        storeQuad<1>(out, v, A.align32 != 0);
                        if(peer != nullptr)
                            storeQuad<0>(peer, v, A.align32 != 0);
                    }
                    else
                    {
#pragma unroll
                        for(int i = 0; i < kWalkCols; ++i)
                            if(wr & (1u << i))
                            {
                                out[i] = v[i];
                                if(peer != nullptr)
                                    peer[i] = v[i];
                            }
                    }
                }
            }
        }
    }

    constexpr int kWalkWarps = 4; // walkers per CTA

    // MINB: CTAs per SM the register allocation is held to (128 threads: 4 -> 128 registers, 3 -> 168, 2 -> 255)
    // SWAP: lanes 4-7, 12-15, ... load the two halves of their quad in the opposite order, which makes both LDS.128 of a row
    // conflict-free (a 32-byte lane stride is 2-way conflicted otherwise) at the price of four 64-bit selects per row
    // BAREONLY: the interior windows over rows whose every level is a core row, nothing else -- no ring, edge or strip code
    // in the register allocation (128 registers, a fourth CTA per SM). The full kernel then keeps the edge windows, the
    // S-row bands next to a physical ring and the strips, and lets this one start next to it (programmatic dependent launch).
    template<int S, int R, int ST, bool SQ, int MINB, bool SWAP = false, bool BAREONLY = false>
    __global__ void __launch_bounds__(32 * kWalkWarps, MINB) heatWalkKernel(const __grid_constant__ CUtensorMap mapSrc, HeatWArgs const A)
    {
        static_assert(S % 2 == 0 && S >= 2 && S <= kMaxLevels, "even S");
        using G = WalkGeom<S>;
        constexpr int BOXX = 32 * kWalkCols;
        constexpr uint32_t kStageBytes = uint32_t(BOXX) * R * 8u;
        extern __shared__ __align__(128) unsigned char smem[];
        __shared__ uint64_t full[kWalkWarps][ST];
        int const warp = threadIdx.x / 32, lane = threadIdx.x % 32;
        uint32_t const walker = blockIdx.x * kWalkWarps + warp;
        if(walker >= A.nWalkers)
            return; // (no CTA-wide barrier anywhere in this kernel)
        uint32_t const nEdge = 1u + A.nEdgeRight;
        uint32_t seg, w;
        bool strip = false;
        int32_t ya, yb;
        if constexpr(BAREONLY)
        {
            uint32_t const nInner = A.nWin - nEdge;
            seg = walker / nInner;
            w = 1u + (walker - seg * nInner);
            ya = A.intY0 + int32_t(seg) * A.segRows;
            yb = min(ya + A.segRows, A.intY1);
        }
        else
        {
            // the dependent bare-only launch (if any) may start as soon as every CTA of this grid is under way
            asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
            // walker order: the edge windows of every segment first (they run the slower ring-column path; dispatched last
            // they would be the tail of the launch), then the interior windows segment by segment (front segments first)
            uint32_t const nEdgeWalkers = nEdge * A.nSegAll;
            if(walker < nEdgeWalkers)
            {
                seg = walker / nEdge;
                uint32_t const e = walker - seg * nEdge;
                w = e == 0u ? 0u : A.nWin - e;
            }
            else
            {
                uint32_t const idx = walker - nEdgeWalkers, nInner = A.nWin - nEdge;
                seg = idx / nInner;
                w = 1u + (idx - seg * nInner);
            }
            if(seg < A.nFront)
            {
                strip = (A.frontIsStrip >> seg) & 1u;
                ya = A.frontY0[seg];
                yb = A.frontY1[seg];
            }
            else
            {
                ya = A.intY0 + int32_t(seg - A.nFront) * A.segRows;
                yb = min(ya + A.segRows, A.intY1);
            }
        }
        unsigned char* const stages = smem + size_t(warp) * ST * kStageBytes;
        uint64_t* const bar = full[warp];
        int32_t const cx = int32_t(w) * G::WW - kWalkCols * G::LOST; // global column of lane 0's quad
        int32_t const gi = cx + kWalkCols * lane;
        int32_t const rowStart = ya - S; // first input row
        int32_t const nChunks = (yb - ya + 2 * S + R - 1) / R;
        if(lane == 0)
        {
#pragma unroll
            for(int s = 0; s < ST; ++s)
                mbarInit(&bar[s], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            if constexpr(!BAREONLY)
            {
                // strip walkers read ghost rows and overwrite the neighbours' ghost rows: wait for the neighbours' previous launch
                if(strip && A.myFlags != nullptr)
                    waitForNeighbours(A.peerDst, A.myFlags, A.step, A.status, A.waitNs);
                // 2-D tiles: walkers whose window or rows touch ghost COLUMNS (every window that is not bare, and the strips,
                // whose corners come with them) wait for the column exchange that followed the neighbours' previous launch
                bool const bareWin = cx >= A.loX - (A.ghostLeft ? A.loX : 0) && cx + BOXX - 1 <= A.hiX + (A.ghostRight ? A.loX : 0)
                                     && cx >= A.loX && cx + BOXX - 1 <= A.hiX;
                if(strip || !bareWin)
                {
                    bool waited = false;
                    for(int side = 0; side < 2; ++side)
                        if(A.colFlags[side] != nullptr)
                        {
                            if(!b200::waitFlagAtLeast(A.colFlags[side], A.step, 1u, A.waitNs))
                                atomicExch(A.status, 3u + uint32_t(side));
                            waited = true;
                        }
                    if(waited)
                        asm volatile("fence.proxy.async;" ::: "memory");
                }
            }
#pragma unroll
            for(int s = 0; s < ST; ++s)
                if(s < nChunks)
                {
                    mbarExpectTx(&bar[s], kStageBytes);
                    tmaLoad2d(stages + s * kStageBytes, &mapSrc, cx, rowStart + s * R, &bar[s]);
                }
        }
        __syncwarp();

        bool const storeLane = lane >= G::LOST && lane < 32 - G::LOST;
        // bare window: every column holds field data with no ring column among them (core columns, or a neighbour's ghost
        // columns), and every STORED column is a core column
        bool const colInterior = cx >= A.loX - (A.ghostLeft ? A.loX : 0) && cx + BOXX - 1 <= A.hiX + (A.ghostRight ? A.loX : 0)
                                 && cx + kWalkCols * G::LOST >= A.loX && cx + BOXX - 1 - kWalkCols * G::LOST <= A.hiX;
        uint32_t ringMask = 0u, validMask = 0u; // bit i: column gi + i is a physical ring column / a column this rank owns
#pragma unroll
        for(int i = 0; i < kWalkCols; ++i)
        {
            int32_t const c = gi + i;
            bool const ring = (c == A.loX - 1 && !A.ghostLeft) || (c == A.hiX + 1 && !A.ghostRight);
            if(ring)
                ringMask |= 1u << i;
            if(ring || (c >= A.loX && c <= A.hiX))
                validMask |= 1u << i;
        }
        WalkState<S> st;
#pragma unroll
        for(int l = 0; l < S; ++l)
#pragma unroll
            for(int i = 0; i < kWalkCols; ++i)
                st.q[l][i] = st.u[l][i] = 0.0;
        // sy[j] one row ahead of its use (edge windows): syw[l] = sy[j0 - 1 - l] at input row j0
        auto const syAt = [&](int32_t j) { return __ldg(A.sy + min(max(j, 0), A.rows - 1)); };
        double syw[S];
        if(!colInterior)
        {
#pragma unroll
            for(int l = 0; l < S; ++l)
                syw[l] = syAt(rowStart - 1 - l);
        }
        else
        {
#pragma unroll
            for(int l = 0; l < S; ++l)
                syw[l] = 0.0;
        }

        int stage = 0;
        uint32_t parity = 0;
        int64_t const pitch = int64_t(A.pitchElems);
        double* outRow = A.dst + int64_t(rowStart - S) * pitch + gi; // output address of the row that input row rowStart completes
        bool const swapHalves = SWAP && ((lane >> 2) & 1);
        int const firstHalf = swapHalves ? 2 : 0;
        auto const loadQuad = [&](double const* q, double (&v)[kWalkCols])
        {
            double2 const a = lds128(q + firstHalf), b = lds128(q + (2 - firstHalf));
            if constexpr(SWAP)
            {
                v[0] = swapHalves ? b.x : a.x;
                v[1] = swapHalves ? b.y : a.y;
                v[2] = swapHalves ? a.x : b.x;
                v[3] = swapHalves ? a.y : b.y;
            }
            else
            {
                v[0] = a.x;
                v[1] = a.y;
                v[2] = b.x;
                v[3] = b.y;
            }
        };
        for(int32_t c = 0; c < nChunks; ++c)
        {
            int32_t const r0 = rowStart + c * R; // first input row of the chunk
            mbarWait(&bar[stage], parity);
            double const* rowp = reinterpret_cast<double const*>(stages + stage * kStageBytes) + kWalkCols * lane;
            // every row this chunk produces at any level (r0 - S .. r0 + R - 2) is a core row: nothing but the stencil there,
            // except in the ring columns of an edge window. (Rows of the segment's prologue hold finite garbage that no
            // stored cell depends on; the row range [ya, yb) guards the stores.)
            bool const rowsCore = BAREONLY || (!strip && r0 - S >= A.loY && r0 + R - 2 <= A.hiY);
            if(BAREONLY || (rowsCore && colInterior))
            {
#pragma unroll
                for(int r = 0; r < R; ++r)
                {
                    double v[kWalkCols];
                    loadQuad(rowp + r * BOXX, v);
                    walkRow<S, SQ, kWalkBare>(A, st, v, r0 + r, gi, storeLane, ya, yb, outRow + r * pitch, ringMask, validMask, syw);
                }
            }
            else if(rowsCore)
            {
#pragma unroll
                for(int r = 0; r < R; ++r)
                {
                    double const syNext = syAt(r0 + r); // sy[j0]: the newest entry of the NEXT row's window
                    double v[kWalkCols];
                    loadQuad(rowp + r * BOXX, v);
                    walkRow<S, SQ, kWalkEdgeCols>(A, st, v, r0 + r, gi, storeLane, ya, yb, outRow + r * pitch, ringMask, validMask, syw);
#pragma unroll
                    for(int l = S - 1; l > 0; --l)
                        syw[l] = syw[l - 1];
                    syw[0] = syNext;
                }
            }
            else
            {
#pragma unroll 1
                for(int r = 0; r < R; ++r)
                {
                    double const syNext = syAt(r0 + r);
                    double v[kWalkCols];
                    loadQuad(rowp + r * BOXX, v);
                    walkRow<S, SQ, kWalkCareful>(A, st, v, r0 + r, gi, storeLane, ya, yb, outRow + r * pitch, ringMask, validMask, syw);
#pragma unroll
                    for(int l = S - 1; l > 0; --l)
                        syw[l] = syw[l - 1];
                    syw[0] = syNext;
                }
            }
            outRow += R * pitch;
            __syncwarp();
            if(lane == 0 && c + ST < nChunks)
            {
                mbarExpectTx(&bar[stage], kStageBytes);
                tmaLoad2d(stages + stage * kStageBytes, &mapSrc, cx, r0 + ST * R, &bar[stage]);
            }
            if(++stage == ST)
            {
                stage = 0;
                parity ^= 1u;
            }
        }

        if constexpr(!BAREONLY)
        {
            if(strip && A.stripCounter != nullptr)
            {
                __syncwarp();
                if(lane == 0)
                    publishWhenLastStrip(A.stripCounter, A.stripTiles, A.peerFlag, A.step);
            }
        }
    }

    // ---- the COLUMN exchange of a 2-D tile with ghost cells G deep (second phase of the halo exchange; the rows travel inside
    // the walker launch). After launch L every rank stores its first / last G core columns -- over ALL rows that hold level
    // data, the ghost rows just received from the vertical neighbours included, which is how the corner blocks reach the
    // diagonal neighbours without a third partner -- into the left / right neighbour's ghost columns and publishes L in their
    // column flag words. Block 0..n-1 share the rows; the first thread of every block waits for the vertical neighbours'
    // row flags of THIS launch before anything is read.
    struct HaloColsArgs
    {
        double const* src; // the buffer launch L wrote
        size_t pitchElems;
        int32_t rowLo, rowHi; // inclusive row range that holds level data
        int32_t loX, hiX, G;
        double* peer[2]; // [left, right] neighbour's same buffer, or null
        uint32_t* peerFlag[2]; // their flag word for my side
        uint32_t const* rowFlags[2]; // my flag words set by the top / bottom neighbour's launch, or null
        uint32_t const* colFlags[2]; // my flag words set by the left / right neighbour's column kernel; waited for at the END
        uint32_t* counter; // blocks finished (reset by the last one)
        uint32_t* status;
        uint64_t waitNs;
        uint32_t step;
    };

    __global__ void __launch_bounds__(256) haloColsKernel(HaloColsArgs const A)
    {
        if(threadIdx.x == 0)
        {
            for(int side = 0; side < 2; ++side)
                if(A.rowFlags[side] != nullptr && !b200::waitFlagAtLeast(A.rowFlags[side], A.step, 0u, A.waitNs))
                    atomicExch(A.status, 1u + uint32_t(side));
        }
        __syncthreads();
        int64_t const nRows = int64_t(A.rowHi) - A.rowLo + 1;
        int64_t const perSide = nRows * A.G;
        for(int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; t < 2 * perSide; t += int64_t(gridDim.x) * blockDim.x)
        {
            int const side = t >= perSide ? 1 : 0;
            if(A.peer[side] == nullptr)
                continue;
            int64_t const u = t - side * perSide;
            int64_t const row = A.rowLo + u / A.G;
            int32_t const c = int32_t(u % A.G);
            // my first G core columns are the left neighbour's right ghost columns, my last G its left ones
            int32_t const from = side == 0 ? A.loX + c : A.hiX - A.G + 1 + c;
            int32_t const to = side == 0 ? A.hiX + 1 + c : A.loX - A.G + c;
            // (ld.volatile: the ghost rows were written by peers during this launch; the flag wait above ordered them)
            double const v = *reinterpret_cast<double const volatile*>(A.src + row * int64_t(A.pitchElems) + from);
            A.peer[side][row * int64_t(A.pitchElems) + to] = v;
        }
        __syncthreads();
        if(threadIdx.x == 0)
        {
            __threadfence_system();
            uint32_t const done = atomicAdd(A.counter, 1u);
            if(done == gridDim.x - 1u)
            {
                __threadfence_system();
                *A.counter = 0u;
                for(int side = 0; side < 2; ++side)
                    if(A.peerFlag[side] != nullptr)
                        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(A.peerFlag[side]), "r"(A.step) : "memory");
                // ... and stay until the horizontal neighbours' columns of this launch have arrived here: the kernel is the
                // pairwise barrier of the column exchange, so the next walker launch starts with its ghost columns complete
                // and none of its walkers has to wait for them (waiting inside the walker cost 150 us per launch on two GPUs,
                // profiles/r02/heat_deep_probe_n2.log)
                for(int side = 0; side < 2; ++side)
                    if(A.colFlags[side] != nullptr && !b200::waitFlagAtLeast(A.colFlags[side], A.step, 0u, A.waitNs))
                        atomicExch(A.status, 3u + uint32_t(side));
            }
        }
    }

    // Ring only (BoundaryKernel.hpp:63-84): top/bottom rows i = 1..nx, left/right columns j = 1..ny, corners untouched.
    // One thread per ring cell: [0,nx) top, [nx,2nx) bottom, [2nx,2nx+ny) left, [2nx+ny, 2nx+2ny) right.
    __global__ void __launch_bounds__(256) heatBoundaryKernel(HeatArgs const A)
    {
        uint64_t const t = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
        uint64_t const nx = A.nx, ny = A.ny;
        if(t >= 2 * nx + 2 * ny)
            return;
        uint32_t j, i;
        int side;
        if(t < nx)
        {
            j = 0;
            i = uint32_t(t) + 1;
            side = B200_EDGE_TOP;
        }
        else if(t < 2 * nx)
        {
            j = A.ny + 1;
            i = uint32_t(t - nx) + 1;
            side = B200_EDGE_BOTTOM;
        }
        else if(t < 2 * nx + ny)
        {
            j = uint32_t(t - 2 * nx) + 1;
            i = 0;
            side = B200_EDGE_LEFT;
        }
        else
        {
            j = uint32_t(t - 2 * nx - ny) + 1;
            i = A.nx + 1;
            side = B200_EDGE_RIGHT;
        }
        if(A.edges & side)
            A.dst[size_t(j) * A.pitchElems + i] = __dmul_rn(A.tf, __dadd_rn(__ldg(A.sx + i), __ldg(A.sy + j)));
    }

    // ---- host side
    using EncodeTiledFn = CUresult (*)(
        CUtensorMap*,
        CUtensorMapDataType,
        cuuint32_t,
        void*,
        cuuint64_t const*,
        cuuint64_t const*,
        cuuint32_t const*,
        cuuint32_t const*,
        CUtensorMapInterleave,
        CUtensorMapSwizzle,
        CUtensorMapL2promotion,
        CUtensorMapFloatOOBfill);

    EncodeTiledFn encoder()
    {
        static EncodeTiledFn fn = nullptr;
        static std::once_flag once;
        std::call_once(
            once,
            []
            {
                void* p = nullptr;
                cudaDriverEntryPointQueryResult q;
                if(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess
                   && q == cudaDriverEntryPointSuccess)
                    fn = reinterpret_cast<EncodeTiledFn>(p);
                else
                    (void) cudaGetLastError();
            });
        return fn;
    }

    // TMA descriptor of one padded field: rows x (nx+2) doubles at `pitchBytes`, box BOX_X x boxY
    bool encodeFieldMap(EncodeTiledFn enc, CUtensorMap* map, double* base, size_t pitchBytes, uint64_t rows, uint32_t nx, int boxY, int boxX = BOX_X)
    {
        int64_t const promoSel = b200::tune("heat.l2promo", 256);
        CUtensorMapL2promotion const promo = promoSel == 0     ? CU_TENSOR_MAP_L2_PROMOTION_NONE
                                             : promoSel == 64  ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                             : promoSel == 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
                                                               : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
        cuuint64_t const dims[2] = {cuuint64_t(nx) + 2, cuuint64_t(rows)};
        cuuint64_t const strides[1] = {cuuint64_t(pitchBytes)};
        cuuint32_t const box[2] = {cuuint32_t(boxX), cuuint32_t(boxY)};
        cuuint32_t const estr[2] = {1, 1};
        return enc(
                   map,
                   CU_TENSOR_MAP_DATA_TYPE_FLOAT64,
                   2,
                   base,
                   dims,
                   strides,
                   box,
                   estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE,
                   promo,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)
               == CUDA_SUCCESS;
    }
} // namespace

struct b200_heat2d_plan_st
{
    int dev;
    double* u[2];
    size_t pitchBytes;
    uint32_t ny, nx;
    int edges;
    double* sx;
    double* sy;
    CUtensorMap map[2];
    CUtensorMap map2[2]; // two-level kernel: box (TYT+4) x 132 at tile height map2Tyt (0 = not built yet)
    int map2Tyt = 0;
    CUtensorMap mapN[2]; // N-level kernel: box keyed by mapNKey = levels * 1000 + tile rows (0 = not built yet)
    int mapNKey = 0;
    CUtensorMap mapW[2]; // walker kernel: box 128 columns x R rows, keyed by mapWKey = R (0 = not built yet)
    int mapWKey = 0;
    double sxLeft = 0.0, sxRight = 0.0; // sx[0], sx[nx+1] (host copies for the walker's edge windows)
    uint32_t padY = 1; // 1: reference layout (ny+2 rows); G: row slab / tile with ghost rows G deep (ny+2G rows)
    uint32_t padX = 1; // 1: the reference's one-cell ring in the columns; G: 2-D tile with ghost columns G deep (nx+2G columns)
    // fused halo exchange (b200_heat2d_plan_set_halo)
    bool hasHalo = false;
    b200_heat2d_halo halo{};
    uint32_t* haloScratch = nullptr; // [0] strip-tile counter, [1] status
};

namespace
{
    // ---- the walker form of an S-level launch (heatWalkKernel), S = 4, 6, 8
    // ---- host-side planning of a walker launch: windows, edge windows, front segments (slab strips / physical bands), interior
    // segments. Pure arithmetic on the geometry already in A (loY, hiY, loX, hiX, ghost*, sendRows) -- no device needed, which is
    // what b200_heat2d_walk_plan_query exposes to the CPU property tests. Fills the decomposition fields of A; `rowsInt`,
    // `nInner` and `split` go back to the launcher. false: too many walkers.
    struct WalkPlanExtra
    {
        int64_t rowsInt;
        uint32_t nInner, nStrip;
        bool split;
    };

    // Segments per window for `rowsInt` rows on `nslots` resident walkers: maximise (share of the slots used over all waves) x
    // (share of a walker's rows that are not its 2S-row prologue) x waves / (waves + tail) -- the last factor models the tail a
    // slower walker of the last wave leaves; it favours a few waves over exactly one.
    int64_t pickWalkSegRows(int64_t rowsInt, int64_t windows, int64_t nslots, double tail, int S)
    {
        int64_t best = rowsInt;
        double bestScore = -1.0;
        int64_t const maxSeg = (rowsInt + 15) / 16;
        for(int64_t k = 1; k <= maxSeg && k <= 8192; ++k)
        {
            int64_t const segRows = (rowsInt + k - 1) / k;
            int64_t const segs = (rowsInt + segRows - 1) / segRows;
            double const walkers = double(segs) * double(windows);
            double const waves = double((int64_t(walkers) + nslots - 1) / nslots);
            double const score = walkers / (waves * double(nslots)) * double(segRows) / double(segRows + 2 * S) * waves / (waves + tail);
            if(score > bestScore + 1e-9)
            {
                bestScore = score;
                best = segRows;
            }
        }
        return best;
    }

    bool planWalk(HeatWArgs& A, int S, int64_t slots, bool splitOk, WalkPlanExtra& X)
    {
        int const LOST = (S + 3) / 4;
        int const BOXX = 32 * kWalkCols;
        int const WW = BOXX - 2 * kWalkCols * LOST;
        // windows cover the columns 0 .. hiX + 1 (window w stores columns [w WW, (w+1) WW))
        A.nWin = (uint32_t(A.hiX) + 2 + uint32_t(WW) - 1) / uint32_t(WW);
        // bare windows (the kernel's colInterior): field data in all 128 columns, no ring column among them, core columns
        // stored. The others -- window 0 and the last few -- are the edge windows, dispatched first.
        auto const isBare = [&](uint32_t w)
        {
            int64_t const cx = int64_t(w) * WW - kWalkCols * LOST;
            return cx >= A.loX - (A.ghostLeft ? A.loX : 0) && cx + BOXX - 1 <= A.hiX + (A.ghostRight ? A.loX : 0)
                   && cx + kWalkCols * LOST >= A.loX && cx + BOXX - 1 - kWalkCols * LOST <= A.hiX;
        };
        A.nEdgeRight = 0;
        while(A.nEdgeRight + 1 < A.nWin && !isBare(A.nWin - 1 - A.nEdgeRight))
            ++A.nEdgeRight;
        bool middleAllBare = true; // (window 0 is never bare: its first lanes lie left of column 0)
        for(uint32_t w = 1; w + A.nEdgeRight < A.nWin; ++w)
            middleAllBare = middleAllBare && isBare(w);
        uint32_t const nEdge = 1u + A.nEdgeRight, nInner = A.nWin - nEdge;
        // Split: the interior windows' interior rows go to the bare-only kernel (128 registers, 4 CTAs per SM), this kernel
        // keeps the edge windows, the strips and the S-row bands next to a physical ring. Needs rows for both bands.
        // Measured (profiles/r02/heat_walk_probe_split.log, 16384^2, 960 steps): 4 levels 215 us per step split (225 without the
        // dependent-launch overlap) against 196 in one kernel, 6 levels 185 against 183 -- a fourth CTA per SM does not pay
        // for the second launch and the shallower stage ring, so the split stays OFF by default.
        bool const split = splitOk && nInner > 0 && middleAllBare && b200::tune("heat.walk_split", 0) != 0
                           && int64_t(A.hiY) - A.loY + 1 >= int64_t(4 * S) + 2 * A.sendRows;
        // output rows: the core rows, plus the ring row on a physical side
        int32_t outLo = A.loY - (A.ghostTop ? 0 : 1), outHi = A.hiY + (A.ghostBottom ? 0 : 1); // inclusive
        A.nFront = 0;
        A.frontIsStrip = 0;
        uint32_t nStrip = 0;
        if(A.ghostTop)
        {
            // strip: the rows sent to the upper neighbour
            A.frontY0[A.nFront] = A.loY;
            A.frontY1[A.nFront] = A.loY + A.sendRows;
            A.frontIsStrip |= 1u << A.nFront;
            outLo = A.loY + A.sendRows;
            ++A.nFront;
            ++nStrip;
        }
        else if(split)
        {
            // band: the ring row and the S - 1 core rows whose lower levels touch it
            A.frontY0[A.nFront] = outLo;
            A.frontY1[A.nFront] = A.loY + S - 1;
            outLo = A.loY + S - 1;
            ++A.nFront;
        }
        if(A.ghostBottom)
        {
            A.frontY0[A.nFront] = A.hiY + 1 - A.sendRows;
            A.frontY1[A.nFront] = A.hiY + 1;
            A.frontIsStrip |= 1u << A.nFront;
            outHi = A.hiY - A.sendRows;
            ++A.nFront;
            ++nStrip;
        }
        else if(split)
        {
            A.frontY0[A.nFront] = A.hiY - S + 2;
            A.frontY1[A.nFront] = outHi + 1;
            outHi = A.hiY - S + 1;
            ++A.nFront;
        }
        A.intY0 = outLo;
        A.intY1 = outHi + 1;
        int64_t const rowsInt = int64_t(A.intY1) - A.intY0;
        int64_t const forced = b200::tune("heat.walk_seg_rows", 0);
        uint32_t nSeg = 0;
        A.segRows = 1;
        if(rowsInt > 0)
        {
            // in split mode this kernel walks the interior rows with the edge windows only: short segments, so that it is over
            // long before the bare-only kernel that runs next to it
            A.segRows = int32_t(forced > 0 ? forced : (split ? std::min<int64_t>(rowsInt, 128) : pickWalkSegRows(rowsInt, A.nWin, slots, 0.3, S)));
            nSeg = uint32_t((rowsInt + A.segRows - 1) / A.segRows);
        }
        A.nSegAll = A.nFront + nSeg;
        uint64_t const walkers = split ? uint64_t(nEdge) * A.nSegAll + uint64_t(nInner) * A.nFront : uint64_t(A.nSegAll) * A.nWin;
        if(walkers > 0x7fffffffull)
            return false;
        A.nWalkers = uint32_t(walkers);
        A.stripTiles = nStrip * A.nWin; // strip WALKERS: each counts itself once
        X.rowsInt = rowsInt;
        X.nInner = nInner;
        X.nStrip = nStrip;
        X.split = split;
        (void) walkers;
        return true;
    }

    // SPLIT_OK: this shape also exists as a bare-only kernel for the interior (instantiated for the default shapes only)
    template<int S, int R, int ST, int MINB, bool SWAP = false, bool SPLIT_OK = false>
    int launchWalkShape(b200_heat2d_plan_t plan, cudaStream_t s, int src_index, HeatWArgs& A, bool sq)
    {
        constexpr int WW = WalkGeom<S>::WW;
        constexpr int BOXX = 32 * kWalkCols;
        constexpr size_t smemBytes = size_t(kWalkWarps) * ST * R * BOXX * 8;
        auto* const kSq = heatWalkKernel<S, R, ST, true, MINB, SWAP>;
        auto* const kGen = heatWalkKernel<S, R, ST, false, MINB, SWAP>;
        auto* const kernel = sq ? kSq : kGen;
        static std::mutex mtx;
        static int slotsPerSm[2][64] = {}; // [sq][device]: resident walkers per SM (0 = not asked yet)
        int slots;
        {
            std::lock_guard<std::mutex> lock(mtx);
            int& cached = slotsPerSm[sq ? 1 : 0][plan->dev & 63];
            if(cached == 0)
            {
                B200_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smemBytes)));
                int ctas = 0;
                B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, kernel, 32 * kWalkWarps, smemBytes));
                cached = (ctas > 0 ? ctas : 1) * kWalkWarps;
            }
            slots = cached * b200::smCount(plan->dev);
        }
        uint32_t const rows = plan->ny + 2 * plan->padY;
        if(plan->mapWKey != R)
        {
            EncodeTiledFn const enc = encoder();
            if(!enc)
                return b200::fail(B200_ENODEV, "cuTensorMapEncodeTiled entry point", __FILE__, __LINE__);
            for(int b = 0; b < 2; ++b)
                if(!encodeFieldMap(enc, &plan->mapW[b], plan->u[b], plan->pitchBytes, rows, plan->nx + 2 * plan->padX - 2, R, BOXX))
                    return b200::fail(B200_EINVAL, "cuTensorMapEncodeTiled (walker box)", __FILE__, __LINE__);
            plan->mapWKey = R;
        }
        A.rows = int32_t(rows);
        A.sxLeft = plan->sxLeft;
        A.sxRight = plan->sxRight;
        A.loX = int32_t(plan->padX);
        A.hiX = int32_t(plan->nx + plan->padX - 1);
        A.ghostLeft = (plan->edges & B200_EDGE_LEFT) ? 0 : 1;
        A.ghostRight = (plan->edges & B200_EDGE_RIGHT) ? 0 : 1;
        A.colFlags[0] = A.colFlags[1] = nullptr;
        if(A.myFlags != nullptr && plan->padX > 1 && b200::tune("heat.tile_wait_in_walker", 0) != 0)
        {
            // 2-D tile: the ghost columns arrive by the column exchange that follows every launch (haloColsKernel)
            if(A.ghostLeft)
                A.colFlags[0] = plan->halo.my_flags + 2;
            if(A.ghostRight)
                A.colFlags[1] = plan->halo.my_flags + 3;
        }
        auto aligned32 = [](void const* p) { return p == nullptr || reinterpret_cast<uintptr_t>(p) % 32 == 0; };
        A.align32 = plan->pitchBytes % 32 == 0 && aligned32(A.dst) && aligned32(A.peerDst[0]) && aligned32(A.peerDst[1]) ? 1 : 0;
        WalkPlanExtra X{};
        if(!planWalk(A, S, slots, SPLIT_OK, X))
            return b200::fail(B200_ERANGE, "heat walker: too many walkers", __FILE__, __LINE__);
        int64_t const rowsInt = X.rowsInt;
        uint32_t const nInner = X.nInner;
        bool const split = X.split;
        uint64_t const walkers = A.nWalkers;
        int64_t const forced = b200::tune("heat.walk_seg_rows", 0);
        auto const pickSegRows = [&](int64_t rows_, int64_t windows, int64_t nslots, double tail) { return pickWalkSegRows(rows_, windows, nslots, tail, S); };
        if(walkers > 0)
        {
            unsigned const grid = unsigned((walkers + kWalkWarps - 1) / kWalkWarps);
            kernel<<<grid, 32 * kWalkWarps, smemBytes, s>>>(plan->mapW[src_index], A);
            B200_LAUNCH_CHECK();
        }
        if constexpr(SPLIT_OK)
        {
            if(split && rowsInt > 0)
            {
                // the bare-only kernel: 3 stages of R rows (48 KB per CTA at R = 4: four CTAs per SM), its own segmentation
                constexpr int STB = 3;
                constexpr size_t smemBare = size_t(kWalkWarps) * STB * R * BOXX * 8;
                constexpr int MINBB = S == 4 ? 4 : 3; // 128 / 168 registers
                auto* const bSq = heatWalkKernel<S, R, STB, true, MINBB, SWAP, true>;
                auto* const bGen = heatWalkKernel<S, R, STB, false, MINBB, SWAP, true>;
                auto* const bare = sq ? bSq : bGen;
                static int bareSlotsPerSm[2][64] = {};
                int bareSlots;
                {
                    std::lock_guard<std::mutex> lock(mtx);
                    int& cached = bareSlotsPerSm[sq ? 1 : 0][plan->dev & 63];
                    if(cached == 0)
                    {
                        B200_CUDA(cudaFuncSetAttribute(bare, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smemBare)));
                        int ctas = 0;
                        B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, bare, 32 * kWalkWarps, smemBare));
                        cached = (ctas > 0 ? ctas : 1) * kWalkWarps;
                    }
                    bareSlots = cached * b200::smCount(plan->dev);
                }
                HeatWArgs B = A;
                B.segRows = int32_t(forced > 0 ? forced : pickSegRows(rowsInt, nInner, bareSlots, 0.1));
                uint64_t const bareWalkers = uint64_t((rowsInt + B.segRows - 1) / B.segRows) * nInner;
                B200_REQUIRE(bareWalkers <= 0x7fffffffull, B200_ERANGE);
                B.nWalkers = uint32_t(bareWalkers);
                cudaLaunchConfig_t cfg{};
                cfg.gridDim = dim3(unsigned((bareWalkers + kWalkWarps - 1) / kWalkWarps));
                cfg.blockDim = dim3(32 * kWalkWarps);
                cfg.dynamicSmemBytes = smemBare;
                cfg.stream = s;
                cudaLaunchAttribute attr{};
                attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
                attr.val.programmaticStreamSerializationAllowed = 1;
                cfg.attrs = &attr;
                cfg.numAttrs = (walkers > 0 && b200::tune("heat.walk_pdl", 1) != 0) ? 1 : 0;
                B200_CUDA(cudaLaunchKernelEx(&cfg, bare, plan->mapW[src_index], B));
                b200::countLaunch();
            }
        }
        return 0;
    }

    int launchWalk(b200_heat2d_plan_t plan, cudaStream_t s, int src_index, HeatNArgs const& base, int levels, bool sq)
    {
        HeatWArgs A{};
        static_cast<HeatNArgs&>(A) = base;
        // R * 10 + ST; 0 = the default: 4 rows x 4 stages = 64 KB per CTA (measured best at every depth, 16384^2 over 960 steps:
        // 8 levels 174.5 / 175.1 / 182.5 us per step with 3 / 4 / 6 stages; profiles/r02/heat_walk_probe.log)
        int shape = int(b200::tune("heat.walk_shape", 0));
        if(shape == 0)
            shape = 44;
        // CTAs per SM the registers are budgeted for: 0 = the default of the depth (4 levels: 3, 6: 2, 8: 2)
        int minb = int(b200::tune("heat.walk_minb", 0));
        if(minb == 0)
            minb = levels == 4 ? 3 : 2;
        auto go = [&]<int S_, int R_, int ST_>() -> int
        {
            // The default shape (R4 x ST4) exists at both register budgets, with conflict-free loads (heat.walk_lds_swap) and
            // with the interior split off (SPLIT_OK; 4 and 6 levels); the three-stage shape at the default budget only.
            // (Six-stage rings and two-row stages were measured slower and are no longer instantiated:
            // profiles/r02/heat_walk_probe.log.)
            constexpr bool kDefaultShape = R_ == 4 && ST_ == 4;
            bool const swap = kDefaultShape && b200::tune("heat.walk_lds_swap", 0) != 0;
            if constexpr(S_ == 4)
            {
                if constexpr(kDefaultShape)
                {
                    if(minb >= 4)
                        return launchWalkShape<S_, R_, ST_, 4>(plan, s, src_index, A, sq);
                    if(swap)
                        return launchWalkShape<S_, R_, ST_, 3, true>(plan, s, src_index, A, sq);
                    return launchWalkShape<S_, R_, ST_, 3, false, true>(plan, s, src_index, A, sq);
                }
                else
                    return launchWalkShape<S_, R_, ST_, 3>(plan, s, src_index, A, sq);
            }
            else
            {
                if constexpr(kDefaultShape)
                {
                    if(minb >= 3)
                        return launchWalkShape<S_, R_, ST_, 3>(plan, s, src_index, A, sq);
                    if(swap)
                        return launchWalkShape<S_, R_, ST_, 2, true>(plan, s, src_index, A, sq);
                    return launchWalkShape<S_, R_, ST_, 2, false, S_ == 6>(plan, s, src_index, A, sq);
                }
                else
                    return launchWalkShape<S_, R_, ST_, 2>(plan, s, src_index, A, sq);
            }
        };
        switch(levels * 100 + shape)
        {
        case 443:
            return go.template operator()<4, 4, 3>();
        case 444:
            return go.template operator()<4, 4, 4>();
        case 643:
            return go.template operator()<6, 4, 3>();
        case 644:
            return go.template operator()<6, 4, 4>();
        case 843:
            return go.template operator()<8, 4, 3>();
        case 844:
            return go.template operator()<8, 4, 4>();
        default:
            return b200::fail(B200_EINVAL, "heat.walk_shape: supported 43, 44 (rows per stage x 10 + stages)", __FILE__, __LINE__);
        }
    }
} // namespace

extern "C"
{
    namespace
    {
        int createPlan(
            int dev,
            double* u0,
            double* u1,
            size_t pitch_bytes,
            uint32_t ny,
            uint32_t nx,
            double const* sx_host,
            double const* sy_host,
            int edges,
            uint32_t padY,
            b200_heat2d_plan_t* out,
            uint32_t padX = 1)
        {
            B200_REQUIRE(out && u0 && u1 && sx_host && sy_host, B200_EINVAL);
            B200_REQUIRE(ny >= 1 && nx >= 1 && (edges & ~B200_EDGE_ALL) == 0, B200_EINVAL);
            B200_REQUIRE(pitch_bytes >= (size_t(nx) + 2 * padX) * 8, B200_EINVAL);
            B200_REQUIRE(pitch_bytes % 16 == 0, B200_EALIGN);
            B200_REQUIRE(reinterpret_cast<uintptr_t>(u0) % 16 == 0 && reinterpret_cast<uintptr_t>(u1) % 16 == 0, B200_EALIGN);
            B200_REQUIRE(uint64_t(ny) + 4 + 64 < 0x7fffffffull && uint64_t(nx) + 2 + TX < 0x7fffffffull, B200_ERANGE);
            B200_CUDA(cudaSetDevice(dev));
            EncodeTiledFn const enc = encoder();
            if(!enc)
                return b200::fail(B200_ENODEV, "cuTensorMapEncodeTiled entry point", __FILE__, __LINE__);

            auto* plan = new b200_heat2d_plan_st{};
            plan->dev = dev;
            plan->u[0] = u0;
            plan->u[1] = u1;
            plan->pitchBytes = pitch_bytes;
            plan->ny = ny;
            plan->nx = nx;
            plan->edges = edges;
            plan->padY = padY;
            plan->padX = padX;
            plan->sxLeft = sx_host[padX - 1];
            plan->sxRight = sx_host[size_t(nx) + padX];
            uint64_t const rows = uint64_t(ny) + 2 * padY;
            for(int b = 0; b < 2; ++b)
            {
                if(!encodeFieldMap(enc, &plan->map[b], plan->u[b], pitch_bytes, rows, nx + 2 * padX - 2, BOX_Y))
                {
                    delete plan;
                    return b200::fail(B200_EINVAL, "cuTensorMapEncodeTiled", __FILE__, __LINE__);
                }
            }
            size_t const bx = (size_t(nx) + 2 * padX) * 8, by = size_t(rows) * 8;
            cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&plan->sx), bx);
            if(e == cudaSuccess)
                e = cudaMalloc(reinterpret_cast<void**>(&plan->sy), by);
            if(e == cudaSuccess)
                e = cudaMemcpy(plan->sx, sx_host, bx, cudaMemcpyHostToDevice);
            if(e == cudaSuccess)
                e = cudaMemcpy(plan->sy, sy_host, by, cudaMemcpyHostToDevice);
            auto optIn = [&](auto* kernel, int bytes = kMaxStages * STAGE_BYTES)
            {
                if(e == cudaSuccess)
                    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
                // (measured: forcing three 72 KB tiles per SM -- 80 registers + the maximum shared-memory carve-out -- is
                // 2 % slower than the two the driver's default carve-out admits: profiles/r01/heat_step2_probe*.log)
            };
            optIn(heatStepKernel<0, 8>);
            optIn(heatStepKernel<1, 8>);
            optIn(heatStepKernel<0, 4>);
            optIn(heatStepKernel<1, 4>);
            optIn(heatStep2Kernel<1, 32, 8>, Step2Geom<32>::kBoxBytes);
            optIn(heatStep2Kernel<1, 32, 16>, Step2Geom<32>::kBoxBytes);
            optIn(heatStep2Kernel<1, 32, 32>, Step2Geom<32>::kBoxBytes);
            optIn(heatStep2Kernel<1, 64, 16>, Step2Geom<64>::kBoxBytes);
            optIn(heatStep2Kernel<1, 64, 32>, Step2Geom<64>::kBoxBytes);
            optIn(heatStepNKernel<3, 16, 4, false>, StepNGeom<3>::BOXX * (64 + 6) * 8);
            optIn(heatStepNKernel<3, 16, 4, true>, StepNGeom<3>::BOXX * (64 + 6) * 8);
            optIn(heatStepNKernel<3, 16, 2, false>, StepNGeom<3>::BOXX * (32 + 6) * 8);
            optIn(heatStepNKernel<3, 16, 2, true>, StepNGeom<3>::BOXX * (32 + 6) * 8);
            optIn(heatStepNKernel<3, 32, 2, false>, StepNGeom<3>::BOXX * (64 + 6) * 8);
            optIn(heatStepNKernel<3, 32, 2, true>, StepNGeom<3>::BOXX * (64 + 6) * 8);
            optIn(heatStepNKernel<4, 16, 4, false>, StepNGeom<4>::BOXX * (64 + 8) * 8);
            optIn(heatStepNKernel<4, 16, 4, true>, StepNGeom<4>::BOXX * (64 + 8) * 8);
            optIn(heatStepNKernel<4, 16, 2, false>, StepNGeom<4>::BOXX * (32 + 8) * 8);
            optIn(heatStepNKernel<4, 16, 2, true>, StepNGeom<4>::BOXX * (32 + 8) * 8);
            optIn(heatStepNKernel<4, 16, 2, true, 5>, StepNGeom<4>::BOXX * (32 + 8) * 8);
            optIn(heatStepNKernel<4, 32, 2, false>, StepNGeom<4>::BOXX * (64 + 8) * 8);
            optIn(heatStepNKernel<4, 32, 2, true>, StepNGeom<4>::BOXX * (64 + 8) * 8);
            if(e != cudaSuccess)
            {
                cudaFree(plan->sx);
                cudaFree(plan->sy);
                delete plan;
                return b200::cudaFail(e, "heat2d plan setup", __FILE__, __LINE__);
            }
            *out = plan;
            return 0;
        }
    } // namespace

    int b200_heat2d_plan_create(
        int dev,
        double* u0,
        double* u1,
        size_t pitch_bytes,
        uint32_t ny,
        uint32_t nx,
        double const* sx_host,
        double const* sy_host,
        int edges,
        b200_heat2d_plan_t* out)
    {
        return createPlan(dev, u0, u1, pitch_bytes, ny, nx, sx_host, sy_host, edges, 1, out);
    }

    int b200_heat2d_slab_plan_create(
        int dev,
        double* u0,
        double* u1,
        size_t pitch_bytes,
        uint32_t ny,
        uint32_t nx,
        double const* sx_host,
        double const* sy_host,
        int edges,
        uint32_t ghost_rows,
        b200_heat2d_plan_t* out)
    {
        // a slab keeps the full width: left and right are physical boundaries; the border rows sent up and down must be
        // distinct rows
        B200_REQUIRE((edges & B200_EDGE_LEFT) && (edges & B200_EDGE_RIGHT), B200_EINVAL);
        B200_REQUIRE(ghost_rows >= 2 && ghost_rows <= uint32_t(kMaxLevels) && ny >= 2 * ghost_rows, B200_EINVAL);
        return createPlan(dev, u0, u1, pitch_bytes, ny, nx, sx_host, sy_host, edges, ghost_rows, out);
    }

    int b200_heat2d_plan_destroy(b200_heat2d_plan_t plan)
    {
        if(!plan)
            return 0;
        B200_CUDA(cudaSetDevice(plan->dev));
        B200_CUDA(cudaFree(plan->sx));
        B200_CUDA(cudaFree(plan->sy));
        if(plan->haloScratch != nullptr)
            B200_CUDA(cudaFree(plan->haloScratch));
        delete plan;
        return 0;
    }

    namespace
    {
        // appends window [j0,j1) x [i0,i1) to the launch description (skipped when empty)
        void addWindow(HeatArgs& A, uint32_t j0, uint32_t j1, uint32_t i0, uint32_t i1, uint32_t strip)
        {
            if(j0 >= j1 || i0 >= i1)
                return;
            HeatWindow& W = A.win[A.nWin++];
            W.j0 = j0;
            W.j1 = j1;
            W.i0 = i0;
            W.i1 = i1;
            W.iw0 = i0 & ~1u;
            W.tilesX = (i1 - W.iw0 + TX - 1) / TX;
            W.tileBegin = A.totalTiles;
            W.strip = strip;
            uint32_t const tilesY = (j1 - j0 + TY - 1) / TY;
            A.totalTiles += W.tilesX * tilesY;
            if(strip)
                A.stripTiles += W.tilesX * tilesY;
        }

        HeatArgs baseArgs(b200_heat2d_plan_t plan, int src_index, double rx, double ry, double time_factor)
        {
            HeatArgs A{};
            A.dst = plan->u[1 - src_index];
            A.pitchElems = plan->pitchBytes / 8;
            A.ny = plan->ny;
            A.nx = plan->nx;
            // StencilKernel.hpp:84: (1.0 - 2.0 * rX - 2.0 * rY), evaluated left to right in IEEE double on the host
            // (this TU is built with -ffp-contract=off for host code)
            A.rX = rx;
            A.rY = ry;
            A.k = 1.0 - 2.0 * rx - 2.0 * ry;
            A.tf = time_factor;
            A.sx = plan->sx;
            A.sy = plan->sy;
            A.edges = plan->edges;
            // heat.ctas_per_sm = 0 (default): ONE tile per CTA, a single stage, residency (6 CTAs of 288 threads per SM)
            // hides the latency. > 0: persistent CTAs striding over the tiles with a `heat.stages`-deep TMA ring.
            bool const persistent = b200::tune("heat.ctas_per_sm", 0) > 0;
            int stages = int(b200::tune("heat.stages", persistent ? 2 : 1));
            A.stages = stages < 1 ? 1 : (stages > kMaxStages ? kMaxStages : stages);
            return A;
        }

        int launchHeat(b200_heat2d_plan_t plan, b200_stream_t stream, int src_index, HeatArgs const& A)
        {
            if(A.totalTiles == 0)
                return 0;
            int const ctasPerSm = int(b200::tune("heat.ctas_per_sm", 0));
            int const hint = int(b200::tune("heat.hint", 1));
            uint64_t grid = ctasPerSm > 0 ? uint64_t(b200::smCount(plan->dev)) * ctasPerSm : uint64_t(A.totalTiles);
            // heat.grid_cap: upper bound on CTAs per launch. Needed when several decomposed tiles share ONE device
            // (tests): their kernels wait for each other's flags, so all of them must be resident at the same time.
            int64_t const cap = b200::tune("heat.grid_cap", 0);
            if(cap > 0 && grid > uint64_t(cap))
                grid = uint64_t(cap);
            if(grid > A.totalTiles)
                grid = A.totalTiles;
            auto const s = reinterpret_cast<cudaStream_t>(stream);
            size_t const smemBytes = size_t(A.stages) * STAGE_BYTES;
            int const rpt = int(b200::tune("heat.rpt", 8));
            auto launch = [&](auto* kernel, int threads) { kernel<<<unsigned(grid), threads, smemBytes, s>>>(plan->map[src_index], A); };
            switch(rpt * 2 + (hint ? 1 : 0))
            {
            case 17:
                launch(heatStepKernel<1, 8>, consumerThreads(8) + 32);
                break;
            case 16:
                launch(heatStepKernel<0, 8>, consumerThreads(8) + 32);
                break;
            case 9:
                launch(heatStepKernel<1, 4>, consumerThreads(4) + 32);
                break;
            case 8:
                launch(heatStepKernel<0, 4>, consumerThreads(4) + 32);
                break;
            default:
                return b200::fail(B200_EINVAL, "heat.rpt must be 4 or 8", __FILE__, __LINE__);
            }
            B200_LAUNCH_CHECK();
            return 0;
        }
    } // namespace

    int b200_heat2d_step_window_f64(
        b200_heat2d_plan_t plan,
        b200_stream_t stream,
        int src_index,
        double rx,
        double ry,
        double time_factor,
        uint32_t j0,
        uint32_t j1,
        uint32_t i0,
        uint32_t i1)
    {
        B200_REQUIRE(plan && (src_index == 0 || src_index == 1) && plan->padY == 1 && plan->padX == 1, B200_EINVAL);
        B200_REQUIRE(j1 <= plan->ny + 2 && i1 <= plan->nx + 2, B200_EINVAL);
        B200_CUDA(cudaSetDevice(plan->dev));
        HeatArgs A = baseArgs(plan, src_index, rx, ry, time_factor);
        addWindow(A, j0, j1, i0, i1, 0);
        return launchHeat(plan, stream, src_index, A);
    }

    int b200_heat2d_boundary_f64(b200_heat2d_plan_t plan, b200_stream_t stream, int dst_index, double time_factor)
    {
        B200_REQUIRE(plan && (dst_index == 0 || dst_index == 1) && plan->padY == 1 && plan->padX == 1, B200_EINVAL);
        B200_CUDA(cudaSetDevice(plan->dev));
        HeatArgs A{};
        A.dst = plan->u[dst_index];
        A.pitchElems = plan->pitchBytes / 8;
        A.ny = plan->ny;
        A.nx = plan->nx;
        A.tf = time_factor;
        A.sx = plan->sx;
        A.sy = plan->sy;
        A.edges = plan->edges;
        uint64_t const cells = 2 * (uint64_t(plan->nx) + plan->ny);
        heatBoundaryKernel<<<unsigned((cells + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(A);
        B200_LAUNCH_CHECK();
        return 0;
    }

    int b200_heat2d_step_f64(b200_heat2d_plan_t plan, b200_stream_t s, int src_index, double rx, double ry, double time_factor)
    {
        B200_REQUIRE(plan, B200_EINVAL);
        return b200_heat2d_step_window_f64(plan, s, src_index, rx, ry, time_factor, 0, plan->ny + 2, 0, plan->nx + 2);
    }

    namespace
    {
        // `haloStep` = 0: no exchange (stand-alone field); >= 1: the 1-based launch index of a connected slab
        int launchStep2(b200_heat2d_plan_t plan, b200_stream_t stream, int src_index, double rx, double ry, double tf1, double tf2, uint32_t haloStep)
        {
            B200_CUDA(cudaSetDevice(plan->dev));
            int const tyt = int(b200::tune("heat.step2_ty", 64));
            int const rpt = int(b200::tune("heat.step2_rpt", 16));
            B200_REQUIRE(tyt == 32 || tyt == 64, B200_EINVAL);
            uint32_t const rows = plan->ny + 2 * plan->padY;
            if(plan->map2Tyt != tyt)
            {
                EncodeTiledFn const enc = encoder();
                if(!enc)
                    return b200::fail(B200_ENODEV, "cuTensorMapEncodeTiled entry point", __FILE__, __LINE__);
                for(int b = 0; b < 2; ++b)
                    if(!encodeFieldMap(enc, &plan->map2[b], plan->u[b], plan->pitchBytes, rows, plan->nx, tyt + 4))
                        return b200::fail(B200_EINVAL, "cuTensorMapEncodeTiled (two-level box)", __FILE__, __LINE__);
                plan->map2Tyt = tyt;
            }
            Heat2Args A{};
            A.dst = plan->u[1 - src_index];
            A.pitchElems = plan->pitchBytes / 8;
            A.ny = plan->ny;
            A.nx = plan->nx;
            A.loY = plan->padY;
            A.hiY = plan->ny + plan->padY - 1;
            A.rows = rows;
            A.tilesX = (plan->nx + 2 + TX - 1) / TX;
            A.rX = rx;
            A.rY = ry;
            A.k = 1.0 - 2.0 * rx - 2.0 * ry; // StencilKernel.hpp:84, as in baseArgs
            A.tf1 = tf1;
            A.tf2 = tf2;
            A.sx = plan->sx;
            A.sy = plan->sy;
            A.ghostTop = (plan->edges & B200_EDGE_TOP) ? 0u : 1u;
            A.ghostBottom = (plan->edges & B200_EDGE_BOTTOM) ? 0u : 1u;
            uint32_t const tilesY = (rows + uint32_t(tyt) - 1) / uint32_t(tyt);
            // strip tile rows: the one holding ghost rows 0,1 and border rows 2,3; the ones holding rows hiY-1.. (border and
            // ghost rows at the bottom). They come first in the launch so their rows travel while the interior is computed.
            A.sendRows = plan->padY;
            A.nTop = A.ghostTop ? 1u : 0u;
            A.tyBot = A.ghostBottom ? (A.hiY + 1u - plan->padY) / uint32_t(tyt) : tilesY;
            if(A.tyBot < A.nTop)
                A.tyBot = A.nTop;
            A.nBot = tilesY - A.tyBot;
            if(haloStep != 0)
            {
                int const dstIndex = 1 - src_index;
                for(int side = 0; side < 2; ++side)
                {
                    A.peerDst[side] = plan->halo.peer_u[side][dstIndex];
                    A.peerFlag[side] = plan->halo.peer_flag[side];
                }
                A.myFlags = plan->halo.my_flags;
                A.stripCounter = plan->haloScratch;
                A.status = plan->haloScratch + 1;
                A.waitNs = b200::waitLimitNs();
                A.stripTiles = (A.nTop + A.nBot) * A.tilesX;
                A.step = haloStep;
                // heat.halo_debug (measurement only, results become wrong): 1 = no peer stores, 2 = no flag wait
                int64_t const dbg = b200::tune("heat.halo_debug", 0);
                if(dbg & 1)
                    A.peerDst[0] = A.peerDst[1] = nullptr;
                if(dbg & 2)
                    A.myFlags = nullptr;
            }
            uint64_t const grid = uint64_t(tilesY) * A.tilesX;
            B200_REQUIRE(grid <= 0x7fffffffull, B200_ERANGE);
            auto const s = reinterpret_cast<cudaStream_t>(stream);
            // (the opt-in for more than 48 KB of dynamic shared memory was made per device when the plan was created)
            auto launch = [&](auto* kernel, int threads, size_t smemBytes) { kernel<<<unsigned(grid), threads, smemBytes, s>>>(plan->map2[src_index], A); };
            switch(tyt * 100 + rpt)
            {
            case 3208:
                launch(heatStep2Kernel<1, 32, 8>, 64 * 4, Step2Geom<32>::kBoxBytes);
                break;
            case 3216:
                launch(heatStep2Kernel<1, 32, 16>, 64 * 2, Step2Geom<32>::kBoxBytes);
                break;
            case 3232:
                launch(heatStep2Kernel<1, 32, 32>, 64, Step2Geom<32>::kBoxBytes);
                break;
            case 6416:
                launch(heatStep2Kernel<1, 64, 16>, 64 * 4, Step2Geom<64>::kBoxBytes);
                break;
            case 6432:
                launch(heatStep2Kernel<1, 64, 32>, 64 * 2, Step2Geom<64>::kBoxBytes);
                break;
            default:
                return b200::fail(B200_EINVAL, "heat.step2_ty/heat.step2_rpt: supported 32/8, 32/16, 32/32, 64/16, 64/32", __FILE__, __LINE__);
            }
            B200_LAUNCH_CHECK();
            return 0;
        }
    } // namespace

    int b200_heat2d_step2_f64(
        b200_heat2d_plan_t plan,
        b200_stream_t stream,
        int src_index,
        double rx,
        double ry,
        double time_factor_1,
        double time_factor_2)
    {
        B200_REQUIRE(plan && (src_index == 0 || src_index == 1), B200_EINVAL);
        // ghost sides need the neighbour's intermediate level: stand-alone fields only (slabs: b200_heat2d_step2_halo_f64)
        B200_REQUIRE(plan->edges == B200_EDGE_ALL && !plan->hasHalo && plan->padY == 1, B200_EINVAL);
        return launchStep2(plan, stream, src_index, rx, ry, time_factor_1, time_factor_2, 0);
    }

    namespace
    {
        int launchStepN(b200_heat2d_plan_t plan, b200_stream_t stream, int src_index, double rx, double ry, int levels, double const* tfs, uint32_t haloStep)
        {
            B200_CUDA(cudaSetDevice(plan->dev));
            // even depths from 4 on: the walker kernel (heat.walk = 0 keeps the round-1 tile kernel for 4 levels)
            bool const walk = levels >= 4 && levels % 2 == 0 && (levels > 4 || plan->padX > 1 || b200::tune("heat.walk", 1) != 0);
            B200_REQUIRE(walk || plan->padX == 1, B200_EINVAL); // ghost columns: the walker kernel only
            int const rpt = int(b200::tune("heat.stepn_rpt", 16));
            int const nwy = int(b200::tune("heat.stepn_nwy", 2));
            int const tyt = rpt * nwy;
            int const boxX = levels == 3 ? StepNGeom<3>::BOXX : StepNGeom<4>::BOXX;
            int const wout = levels == 3 ? StepNGeom<3>::WOUT : StepNGeom<4>::WOUT;
            uint32_t const rows = plan->ny + 2 * plan->padY;
            int const key = levels * 1000 + tyt;
            if(!walk && plan->mapNKey != key)
            {
                EncodeTiledFn const enc = encoder();
                if(!enc)
                    return b200::fail(B200_ENODEV, "cuTensorMapEncodeTiled entry point", __FILE__, __LINE__);
                for(int b = 0; b < 2; ++b)
                    if(!encodeFieldMap(enc, &plan->mapN[b], plan->u[b], plan->pitchBytes, rows, plan->nx, tyt + 2 * levels, boxX))
                        return b200::fail(B200_EINVAL, "cuTensorMapEncodeTiled (N-level box)", __FILE__, __LINE__);
                plan->mapNKey = key;
            }
            HeatNArgs A{};
            A.dst = plan->u[1 - src_index];
            A.pitchElems = plan->pitchBytes / 8;
            A.ny = plan->ny;
            A.nx = plan->nx;
            A.loY = int32_t(plan->padY);
            A.hiY = int32_t(plan->ny + plan->padY - 1);
            A.tilesX = (plan->nx + 2 + uint32_t(wout) - 1) / uint32_t(wout);
            A.rX = rx;
            A.rY = ry;
            A.k = 1.0 - 2.0 * rx - 2.0 * ry; // StencilKernel.hpp:84, as in baseArgs
            for(int l = 0; l < levels; ++l)
                A.tf[l] = tfs[l];
            A.sx = plan->sx;
            A.sy = plan->sy;
            A.ghostTop = (plan->edges & B200_EDGE_TOP) ? 0 : 1;
            A.ghostBottom = (plan->edges & B200_EDGE_BOTTOM) ? 0 : 1;
            uint32_t const tilesY = (rows + uint32_t(tyt) - 1) / uint32_t(tyt);
            // strip tile rows: tile row 0 (ghost rows 0..S-1, border rows S..2S-1) and the tile rows from the one holding the
            // first of the last S core rows on (tile rows are at least 2S rows tall)
            A.sendRows = int32_t(plan->padY);
            A.nTop = A.ghostTop ? 1u : 0u;
            A.tyBot = A.ghostBottom ? uint32_t(A.hiY + 1 - int32_t(plan->padY)) / uint32_t(tyt) : tilesY;
            if(A.tyBot < A.nTop)
                A.tyBot = A.nTop;
            A.nBot = tilesY - A.tyBot;
            if(haloStep != 0)
            {
                int const dstIndex = 1 - src_index;
                for(int side = 0; side < 2; ++side)
                {
                    A.peerDst[side] = plan->halo.peer_u[side][dstIndex];
                    A.peerFlag[side] = plan->halo.peer_flag[side];
                }
                A.myFlags = plan->halo.my_flags;
                A.stripCounter = plan->haloScratch;
                A.status = plan->haloScratch + 1;
                A.waitNs = b200::waitLimitNs();
                A.stripTiles = (A.nTop + A.nBot) * A.tilesX;
                A.step = haloStep;
                int64_t const dbg = b200::tune("heat.halo_debug", 0);
                if(dbg & 1)
                    A.peerDst[0] = A.peerDst[1] = nullptr;
                if(dbg & 2)
                    A.myFlags = nullptr;
            }
            uint64_t const grid = uint64_t(tilesY) * A.tilesX;
            B200_REQUIRE(grid <= 0x7fffffffull, B200_ERANGE);
            auto const s = reinterpret_cast<cudaStream_t>(stream);
            if(walk)
                return launchWalk(plan, s, src_index, A, levels, rx == ry && b200::tune("heat.stepn_sq", 1) != 0);
            size_t const smemBytes = size_t(boxX) * size_t(tyt + 2 * levels) * 8;
            auto launch = [&](auto* kernel, int threads) { kernel<<<unsigned(grid), threads, smemBytes, s>>>(plan->mapN[src_index], A); };
            // square cells (dx == dy): rX == rY bit for bit, one product serves both directions (makeRowN<SQ>)
            bool const sq = rx == ry && b200::tune("heat.stepn_sq", 1) != 0;
            auto go = [&]<int S_, int RPT_, int NWY_>(int threads)
            {
                if(sq)
                    launch(heatStepNKernel<S_, RPT_, NWY_, true>, threads);
                else
                    launch(heatStepNKernel<S_, RPT_, NWY_, false>, threads);
            };
            switch(levels * 10000 + rpt * 100 + nwy)
            {
            case 31604:
                go.template operator()<3, 16, 4>(256);
                break;
            case 31602:
                go.template operator()<3, 16, 2>(128);
                break;
            case 33202:
                go.template operator()<3, 32, 2>(128);
                break;
            case 41604:
                go.template operator()<4, 16, 4>(256);
                break;
            case 41602:
                if(sq && b200::tune("heat.stepn_ctas", 4) == 5)
                    launch(heatStepNKernel<4, 16, 2, true, 5>, 128);
                else
                    go.template operator()<4, 16, 2>(128);
                break;
            case 43202:
                go.template operator()<4, 32, 2>(128);
                break;
            default:
                return b200::fail(B200_EINVAL, "heat.stepn_rpt/heat.stepn_nwy: supported 16/2, 16/4, 32/2", __FILE__, __LINE__);
            }
            B200_LAUNCH_CHECK();
            return 0;
        }
    } // namespace

    int b200_heat2d_stepn_f64(
        b200_heat2d_plan_t plan,
        b200_stream_t stream,
        int src_index,
        double rx,
        double ry,
        int levels,
        double const* time_factors)
    {
        B200_REQUIRE(plan && (src_index == 0 || src_index == 1) && time_factors, B200_EINVAL);
        B200_REQUIRE(levels == 3 || levels == 4 || levels == 6 || levels == 8, B200_EINVAL);
        B200_REQUIRE(plan->edges == B200_EDGE_ALL && !plan->hasHalo && plan->padY == 1, B200_EINVAL);
        return launchStepN(plan, stream, src_index, rx, ry, levels, time_factors, 0);
    }

    int b200_heat2d_stepn_halo_f64(
        b200_heat2d_plan_t plan,
        b200_stream_t stream,
        int src_index,
        double rx,
        double ry,
        int levels,
        double const* time_factors,
        uint32_t step)
    {
        B200_REQUIRE(plan && (src_index == 0 || src_index == 1) && time_factors && step >= 1, B200_EINVAL);
        B200_REQUIRE(levels == 3 || levels == 4 || levels == 6 || levels == 8, B200_EINVAL);
        // the ghost rows must be at least as deep as the launch advances (all of them are refreshed by every launch)
        B200_REQUIRE(plan->hasHalo && plan->padY >= uint32_t(levels) && plan->padX == 1, B200_EINVAL);
        return launchStepN(plan, stream, src_index, rx, ry, levels, time_factors, step);
    }

    int b200_heat2d_step2_halo_f64(
        b200_heat2d_plan_t plan,
        b200_stream_t stream,
        int src_index,
        double rx,
        double ry,
        double time_factor_1,
        double time_factor_2,
        uint32_t step)
    {
        B200_REQUIRE(plan && (src_index == 0 || src_index == 1) && step >= 1, B200_EINVAL);
        // the ghost rows must be at least as deep as the launch advances (all of them are refreshed by every launch)
        B200_REQUIRE(plan->hasHalo && plan->padY >= 2 && plan->padX == 1, B200_EINVAL);
        return launchStep2(plan, stream, src_index, rx, ry, time_factor_1, time_factor_2, step);
    }

    int b200_heat2d_tile_plan_create(
        int dev,
        double* u0,
        double* u1,
        size_t pitch_bytes,
        uint32_t ny,
        uint32_t nx,
        double const* sx_host,
        double const* sy_host,
        int edges,
        uint32_t ghost,
        b200_heat2d_plan_t* out)
    {
        // ghost cells `ghost` deep on every side; the border rows / columns sent to the neighbours must be distinct
        B200_REQUIRE(ghost >= 4 && ghost <= uint32_t(kMaxLevels) && ny >= 2 * ghost && nx >= 2 * ghost, B200_EINVAL);
        return createPlan(dev, u0, u1, pitch_bytes, ny, nx, sx_host, sy_host, edges, ghost, out, ghost);
    }

    int b200_heat2d_stepn_tile_f64(
        b200_heat2d_plan_t plan,
        b200_stream_t stream,
        int src_index,
        double rx,
        double ry,
        int levels,
        double const* time_factors,
        uint32_t step)
    {
        B200_REQUIRE(plan && (src_index == 0 || src_index == 1) && time_factors && step >= 1, B200_EINVAL);
        B200_REQUIRE(levels == 4 || levels == 6 || levels == 8, B200_EINVAL);
        B200_REQUIRE(plan->hasHalo && plan->padX > 1 && plan->padX == plan->padY && plan->padY >= uint32_t(levels), B200_EINVAL);
        // phase 1: the walker launch (rows travel to the vertical neighbours from its strips)
        if(int const rc = launchStepN(plan, stream, src_index, rx, ry, levels, time_factors, step))
            return rc;
        // phase 2: the columns, ghost rows included
        bool const left = !(plan->edges & B200_EDGE_LEFT), right = !(plan->edges & B200_EDGE_RIGHT);
        if(!left && !right)
            return 0;
        int const dstIndex = 1 - src_index;
        bool const top = !(plan->edges & B200_EDGE_TOP), bottom = !(plan->edges & B200_EDGE_BOTTOM);
        HaloColsArgs H{};
        H.src = plan->u[dstIndex];
        H.pitchElems = plan->pitchBytes / 8;
        int32_t const G = int32_t(plan->padY);
        H.G = G;
        H.loX = G;
        H.hiX = int32_t(plan->nx) + G - 1;
        H.rowLo = top ? 0 : G - 1; // ghost rows above, or the ring row
        H.rowHi = int32_t(plan->ny) + G + (bottom ? G - 1 : 0);
        H.peer[0] = left ? plan->halo.peer_u[2][dstIndex] : nullptr;
        H.peer[1] = right ? plan->halo.peer_u[3][dstIndex] : nullptr;
        H.peerFlag[0] = left ? plan->halo.peer_flag[2] : nullptr;
        H.peerFlag[1] = right ? plan->halo.peer_flag[3] : nullptr;
        H.rowFlags[0] = top ? plan->halo.my_flags + 0 : nullptr;
        H.rowFlags[1] = bottom ? plan->halo.my_flags + 1 : nullptr;
        bool const waitHere = b200::tune("heat.tile_wait_in_walker", 0) == 0;
        H.colFlags[0] = left && waitHere ? plan->halo.my_flags + 2 : nullptr;
        H.colFlags[1] = right && waitHere ? plan->halo.my_flags + 3 : nullptr;
        H.counter = plan->haloScratch + 2;
        H.status = plan->haloScratch + 1;
        H.waitNs = b200::waitLimitNs();
        H.step = step;
        int64_t const dbg = b200::tune("heat.halo_debug", 0);
        if(dbg & 2)
            H.rowFlags[0] = H.rowFlags[1] = H.colFlags[0] = H.colFlags[1] = nullptr;
        if(dbg & 1)
            H.peer[0] = H.peer[1] = nullptr; // (the flags are still published)
        if(dbg & 4)
            return 0; // measurement only: no column kernel at all (the next launch must run with bit 2 set)
        int64_t const cells = 2 * (int64_t(H.rowHi) - H.rowLo + 1) * G;
        unsigned const grid = unsigned(std::min<int64_t>(64, (cells + 255) / 256));
        haloColsKernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(H);
        B200_LAUNCH_CHECK();
        return 0;
    }

    int b200_heat2d_walk_plan_query(
        uint32_t ny,
        uint32_t nx,
        uint32_t pad_y,
        uint32_t pad_x,
        int edges,
        int levels,
        int resident_walkers,
        b200_heat2d_walk_plan* out)
    {
        // the host-side decomposition of a walker launch, without a device (CPU property tests; tools)
        B200_REQUIRE(out && ny >= 1 && nx >= 1 && pad_y >= 1 && pad_x >= 1 && resident_walkers >= 1, B200_EINVAL);
        B200_REQUIRE((levels == 4 || levels == 6 || levels == 8) && (edges & ~B200_EDGE_ALL) == 0, B200_EINVAL);
        HeatWArgs A{};
        A.ny = ny;
        A.nx = nx;
        A.loY = int32_t(pad_y);
        A.hiY = int32_t(ny + pad_y - 1);
        A.loX = int32_t(pad_x);
        A.hiX = int32_t(nx + pad_x - 1);
        A.ghostTop = (edges & B200_EDGE_TOP) ? 0 : 1;
        A.ghostBottom = (edges & B200_EDGE_BOTTOM) ? 0 : 1;
        A.ghostLeft = (edges & B200_EDGE_LEFT) ? 0 : 1;
        A.ghostRight = (edges & B200_EDGE_RIGHT) ? 0 : 1;
        A.sendRows = int32_t(pad_y);
        WalkPlanExtra X{};
        if(!planWalk(A, levels, resident_walkers, levels != 8, X))
            return b200::fail(B200_ERANGE, "heat walker: too many walkers", __FILE__, __LINE__);
        int const lost = (levels + 3) / 4;
        out->window_columns = uint32_t(32 * kWalkCols - 2 * kWalkCols * lost);
        out->n_windows = A.nWin;
        out->n_edge_right = A.nEdgeRight;
        out->n_front = A.nFront;
        out->front_is_strip = A.frontIsStrip;
        for(int k = 0; k < 2; ++k)
        {
            out->front_y0[k] = A.frontY0[k];
            out->front_y1[k] = A.frontY1[k];
        }
        out->interior_y0 = A.intY0;
        out->interior_y1 = A.intY1;
        out->segment_rows = A.segRows;
        out->n_segments = A.nSegAll;
        out->n_walkers = A.nWalkers;
        out->split = X.split ? 1 : 0;
        return 0;
    }

    int b200_heat2d_plan_set_halo(b200_heat2d_plan_t plan, b200_heat2d_halo const* halo)
    {
        B200_REQUIRE(plan && halo && halo->my_flags, B200_EINVAL);
        B200_CUDA(cudaSetDevice(plan->dev));
        for(int side = 0; side < 4; ++side)
        {
            bool const hasNeighbour = halo->peer_u[side][0] != nullptr;
            B200_REQUIRE(hasNeighbour == (halo->peer_u[side][1] != nullptr), B200_EINVAL);
            B200_REQUIRE(hasNeighbour == (halo->peer_flag[side] != nullptr), B200_EINVAL);
            // a side is either a physical boundary (plan `edges`) or has a neighbour, never both
            int const bit = side == 0 ? B200_EDGE_TOP : side == 1 ? B200_EDGE_BOTTOM : side == 2 ? B200_EDGE_LEFT : B200_EDGE_RIGHT;
            B200_REQUIRE(hasNeighbour != ((plan->edges & bit) != 0), B200_EINVAL);
        }
        if(plan->haloScratch == nullptr)
        {
            B200_CUDA(cudaMalloc(reinterpret_cast<void**>(&plan->haloScratch), 64));
            B200_CUDA(cudaMemset(plan->haloScratch, 0, 64));
        }
        plan->halo = *halo;
        plan->hasHalo = true;
        return 0;
    }

    int b200_heat2d_step_halo_f64(b200_heat2d_plan_t plan, b200_stream_t stream, int src_index, double rx, double ry, double time_factor, uint32_t step)
    {
        B200_REQUIRE(plan && plan->hasHalo && (src_index == 0 || src_index == 1) && step >= 1 && plan->padY == 1, B200_EINVAL);
        B200_CUDA(cudaSetDevice(plan->dev));
        HeatArgs A = baseArgs(plan, src_index, rx, ry, time_factor);
        uint32_t const H = plan->ny + 2, Wd = plan->nx + 2;
        // Edge strips first (their border cells travel to the neighbours while the interior is computed), each one
        // tile thick so that no tile is fetched for a sliver; then the interior. Small tiles: everything is "strip".
        if(H > 2 * TY && Wd > 2 * TX)
        {
            addWindow(A, 0, TY, 0, Wd, 1); // top
            addWindow(A, H - TY, H, 0, Wd, 1); // bottom
            addWindow(A, TY, H - TY, 0, TX, 1); // left
            addWindow(A, TY, H - TY, Wd - TX, Wd, 1); // right
            addWindow(A, TY, H - TY, TX, Wd - TX, 0); // interior
        }
        else
        {
            addWindow(A, 0, H, 0, Wd, 1);
        }
        int const dstIndex = 1 - src_index;
        for(int side = 0; side < 4; ++side)
        {
            A.peerDst[side] = plan->halo.peer_u[side][dstIndex];
            A.peerFlag[side] = plan->halo.peer_flag[side];
        }
        A.myFlags = plan->halo.my_flags;
        // heat.halo_debug (measurement only, results become wrong): 1 = no peer stores, 2 = no flag wait
        int64_t const dbg = b200::tune("heat.halo_debug", 0);
        if(dbg & 1)
            for(int side = 0; side < 4; ++side)
                A.peerDst[side] = nullptr;
        if(dbg & 2)
            A.myFlags = nullptr;
        A.stripCounter = plan->haloScratch;
        A.status = plan->haloScratch + 1;
        A.waitNs = b200::waitLimitNs();
        A.step = step;
        return launchHeat(plan, stream, src_index, A);
    }

    int b200_heat2d_halo_status(b200_heat2d_plan_t plan, uint32_t* status)
    {
        B200_REQUIRE(plan && status, B200_EINVAL);
        *status = 0;
        if(plan->haloScratch == nullptr)
            return 0;
        B200_CUDA(cudaSetDevice(plan->dev));
        B200_CUDA(cudaMemcpy(status, plan->haloScratch + 1, sizeof(uint32_t), cudaMemcpyDeviceToHost));
        return 0;
    }
}
