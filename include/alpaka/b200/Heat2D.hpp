// include/alpaka/b200/Heat2D.hpp -- first-class fused FTCS step for heatEquation2D on alpaka buffers and queues.
//
// The reference expresses one time step as two kernel launches, StencilKernel then BoundaryKernel
// (example/heatEquation2D/src/heatEquation2D.cpp:141-168). Those two functors are recognised individually by
// alpaka/b200/Native.hpp; this class is the ONE-launch form (b200_heat2d_step_f64: TMA-pipelined stencil with the
// boundary ring fused in), for drivers that can call it directly. Same field layout as the reference driver:
// (ny+2) x (nx+2) doubles with a ring of boundary cells, row pitch from getPitchesInBytes. Optional `edges` /
// index offsets describe a sub-domain of a 2-D decomposition (ghost sides are left to the halo exchange).
#pragma once

#include "Kernel.hpp"

#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

namespace alpaka::b200
{
    //! Time levels of the next launch when `n` steps are left and a launch may advance up to `depth`: the deepest depth
    //! the kernels support (1, 2, 3 tile kernels; 4, 6, 8 walker kernel) that does not leave a single step behind
    //! (4 = 2 + 2 rather than 3 + 1); `minDepth` = 2 for slabs. 0 if nothing fits.
    [[nodiscard]] inline auto heatNextDepth(std::uint32_t n, int depth, int minDepth = 1) -> std::uint32_t
    {
        for(int k : {8, 6, 4, 3, 2, 1})
        {
            if(k > depth || k < minDepth || static_cast<std::uint32_t>(k) > n)
                continue;
            std::uint32_t const left = n - static_cast<std::uint32_t>(k);
            if(left == 1 && (k > 2 || minDepth > 1))
                continue;
            return static_cast<std::uint32_t>(k);
        }
        return 0;
    }

    class Heat2DStepper
    {
    public:
        //! \param bufA,bufB the ping-pong pair (identical extents and pitches), bufA holds the current field
        //! \param dx,dy,dt grid spacing and time step of the GLOBAL problem
        //! \param edges which sides of this field are physical boundaries (B200_EDGE_*)
        //! \param jOffset,iOffset global index of this field's [0][0] cell (0 for an undecomposed field)
        template<typename TIdx>
        Heat2DStepper(
            BufB200<double, DimInt<2u>, TIdx>& bufA,
            BufB200<double, DimInt<2u>, TIdx>& bufB,
            double dx,
            double dy,
            double dt,
            int edges = B200_EDGE_ALL,
            std::uint64_t jOffset = 0,
            std::uint64_t iOffset = 0)
            : m_dt(dt)
            , m_rX(dt / (dx * dx)) // StencilKernel.hpp:70-71
            , m_rY(dt / (dy * dy))
        {
            auto const ext = getExtents(bufA);
            if(ext != getExtents(bufB) || getPitchesInBytes(bufA) != getPitchesInBytes(bufB) || ext[0] < 3 || ext[1] < 3)
                throw std::runtime_error("Heat2DStepper: the two fields must have identical extents (>= 3x3) and pitches");
            auto const ny = static_cast<std::uint32_t>(ext[0] - 2);
            auto const nx = static_cast<std::uint32_t>(ext[1] - 2);
            // four levels per launch by default, eight on fields tall enough for long walks (as alpaka_b200.heat2d.Heat2D)
            m_defaultDepth = ny >= 12288u && nx >= 4096u ? 8 : 4;
            constexpr double pi = math::constants::pi;
            // boundary factors on the HOST with the C library (the reference CPU back-end's values):
            // exactSolution(x, y, t) = exp(-pi*pi*t) * (sin(pi*x) + sin(pi*y)), analyticalSolution.hpp:17-21
            std::vector<double> sx(nx + 2u), sy(ny + 2u);
            for(std::uint32_t i = 0; i < nx + 2u; ++i)
                sx[i] = std::sin(pi * (static_cast<double>(i + iOffset) * dx));
            for(std::uint32_t j = 0; j < ny + 2u; ++j)
                sy[j] = std::sin(pi * (static_cast<double>(j + jOffset) * dy));
            check(b200_heat2d_plan_create(
                getDev(bufA).getNativeHandle(),
                std::data(bufA),
                std::data(bufB),
                static_cast<std::size_t>(getPitchesInBytes(bufA)[0]),
                ny,
                nx,
                sx.data(),
                sy.data(),
                edges,
                &m_plan));
        }
        Heat2DStepper(Heat2DStepper const&) = delete;
        auto operator=(Heat2DStepper const&) -> Heat2DStepper& = delete;
        ~Heat2DStepper()
        {
            checkNoexcept(b200_heat2d_plan_destroy(m_plan));
        }

        //! one fused step (stencil + boundary for time level `stepsDone()+1`) in queue order; swaps the roles of the buffers
        template<typename TQueue>
        void step(TQueue& queue)
        {
            constexpr double pi = math::constants::pi;
            ++m_step;
            double const tf = std::exp(-pi * pi * (m_step * m_dt));
            check(b200_heat2d_step_f64(m_plan, queue.getNativeHandle(), m_cur, m_rX, m_rY, tf));
            m_cur ^= 1;
            queue.afterEnqueue();
        }

        //! two fused steps in ONE launch (b200_heat2d_step2_f64): the intermediate time level stays in registers, HBM
        //! traffic is one read + one write per cell for both steps, the field is bit-identical to two step() calls.
        //! The roles of the buffers swap ONCE. Stand-alone fields only (all four sides physical boundaries).
        template<typename TQueue>
        void step2(TQueue& queue)
        {
            constexpr double pi = math::constants::pi;
            double const tf1 = std::exp(-pi * pi * ((m_step + 1) * m_dt));
            double const tf2 = std::exp(-pi * pi * ((m_step + 2) * m_dt));
            check(b200_heat2d_step2_f64(m_plan, queue.getNativeHandle(), m_cur, m_rX, m_rY, tf1, tf2));
            m_step += 2;
            m_cur ^= 1;
            queue.afterEnqueue();
        }

        //! `levels` (3, 4, 6 or 8) fused steps in ONE launch (b200_heat2d_stepn_f64: own column pair per thread and level, the
        //! horizontal neighbours by warp shuffle; 4, 6, 8: the walker kernel); same bits, the roles of the buffers swap ONCE.
        template<typename TQueue>
        void stepN(TQueue& queue, int levels)
        {
            constexpr double pi = math::constants::pi;
            double tf[8] = {};
            for(int l = 0; l < levels && l < 8; ++l)
                tf[l] = std::exp(-pi * pi * ((m_step + 1u + static_cast<std::uint32_t>(l)) * m_dt));
            check(b200_heat2d_stepn_f64(m_plan, queue.getNativeHandle(), m_cur, m_rX, m_rY, levels, tf));
            m_step += static_cast<std::uint32_t>(levels);
            m_cur ^= 1;
            queue.afterEnqueue();
        }

        //! `n` steps with up to `depth` (1..8; 0 = defaultDepth()) time levels per launch; a remainder runs in shallower
        //! launches (4 = 2 + 2 rather than 3 + 1)
        template<typename TQueue>
        void steps(TQueue& queue, std::uint32_t n, int depth = 0)
        {
            if(depth == 0)
                depth = m_defaultDepth;
            if(depth < 1 || depth > 8)
                throw std::runtime_error("Heat2DStepper::steps: between 1 and 8 time levels per launch");
            while(n > 0)
            {
                auto const k = heatNextDepth(n, depth);
                if(k == 1)
                    step(queue);
                else if(k == 2)
                    step2(queue);
                else
                    stepN(queue, static_cast<int>(k));
                n -= k;
            }
        }

        //! 0 if the current field is bufA, 1 if it is bufB
        [[nodiscard]] auto currentIndex() const -> int
        {
            return m_cur;
        }
        [[nodiscard]] auto stepsDone() const -> std::uint32_t
        {
            return m_step;
        }
        //! time levels per launch that steps() aims for when it is not told
        [[nodiscard]] auto defaultDepth() const -> int
        {
            return m_defaultDepth;
        }

    private:
        b200_heat2d_plan_t m_plan = nullptr;
        double m_dt, m_rX, m_rY;
        int m_defaultDepth = 4;
        int m_cur = 0;
        std::uint32_t m_step = 0;
    };

    //! K row slabs of ONE heat field driven from one host thread: slab k lives on devs[k] (devices may repeat -- several
    //! slabs on one device, which is how the parity tests run on a single GPU). Every launch advances up to `levels`
    //! (2..4) time levels and exchanges `levels` ghost rows per side straight from the kernel into the neighbour's
    //! array (peer stores + flag words: b200_heat2d_slab_plan_create / b200_heat2d_step2_halo_f64 /
    //! b200_heat2d_stepn_halo_f64). The multi-process form of the same thing is alpaka_b200.multi.HeatSlab.
    //! Several slabs on ONE device are meant for tests on small fields: the launches of different slabs wait for each
    //! other's flag words and therefore have to be co-resident, which holds as long as the strip tiles of one launch
    //! (two tile rows) do not fill the device.
    //! Host fields are (NY+2) x (NX+2) doubles, rows unpadded, as the reference driver's host buffer.
    class Heat2DSlabs
    {
    public:
        using Idx = std::uint32_t;
        using Buf2 = BufB200<double, DimInt<2u>, Idx>;
        using BufFlags = BufB200<std::uint32_t, DimInt<1u>, Idx>;
        using Queue = QueueB200<NonBlocking>;

        Heat2DSlabs(std::vector<DevB200> const& devs, Idx NY, Idx NX, double dx, double dy, double dt, int levels = 4)
            : m_NY(NY)
            , m_NX(NX)
            , m_G(static_cast<Idx>(levels))
            , m_dt(dt)
            , m_rX(dt / (dx * dx))
            , m_rY(dt / (dy * dy))
        {
            auto const K = static_cast<Idx>(devs.size());
            if(K == 0 || !(levels == 2 || levels == 3 || levels == 4 || levels == 6 || levels == 8) || NY % K != 0 || NY / K < 2 * m_G)
                throw std::runtime_error("Heat2DSlabs: NY must divide into slabs of at least 2*levels rows, levels 2, 3, 4, 6 or 8");
            m_ny = NY / K;
            bool distinct = false;
            for(auto const& d : devs)
                distinct = distinct || d.getNativeHandle() != devs[0].getNativeHandle();
            if(distinct)
            {
                int pairs = 0;
                check(b200_enable_peer_all(&pairs));
            }
            constexpr double pi = math::constants::pi;
            std::vector<double> sx(NX + 2u);
            for(Idx i = 0; i < NX + 2u; ++i)
                sx[i] = std::sin(pi * (static_cast<double>(i) * dx));
            m_slabs.reserve(K);
            for(Idx k = 0; k < K; ++k)
            {
                Vec<DimInt<2u>, Idx> const ext{m_ny + 2u * m_G, NX + 2u};
                Slab sl{devs[k], Queue{devs[k]}, allocBuf<double, Idx>(devs[k], ext), allocBuf<double, Idx>(devs[k], ext),
                        allocBuf<std::uint32_t, Idx>(devs[k], Vec<DimInt<1u>, Idx>{16u}), nullptr,
                        static_cast<std::int64_t>(k) * m_ny - (static_cast<std::int64_t>(m_G) - 1)};
                alpaka::memset(sl.queue, sl.flags, 0);
                alpaka::memset(sl.queue, sl.u[0], 0);
                alpaka::memset(sl.queue, sl.u[1], 0);
                std::vector<double> sy(ext[0]);
                for(Idx j = 0; j < ext[0]; ++j)
                    sy[j] = std::sin(pi * (static_cast<double>(sl.g0 + static_cast<std::int64_t>(j)) * dy));
                int const edges = B200_EDGE_LEFT | B200_EDGE_RIGHT | (k == 0 ? B200_EDGE_TOP : 0) | (k == K - 1 ? B200_EDGE_BOTTOM : 0);
                check(b200_heat2d_slab_plan_create(
                    devs[k].getNativeHandle(),
                    std::data(sl.u[0]),
                    std::data(sl.u[1]),
                    static_cast<std::size_t>(getPitchesInBytes(sl.u[0])[0]),
                    m_ny,
                    NX,
                    sx.data(),
                    sy.data(),
                    edges,
                    m_G,
                    &sl.plan));
                m_slabs.push_back(std::move(sl));
            }
            for(Idx k = 0; k < K; ++k)
            {
                b200_heat2d_halo halo{};
                if(k > 0)
                {
                    halo.peer_u[0][0] = std::data(m_slabs[k - 1].u[0]);
                    halo.peer_u[0][1] = std::data(m_slabs[k - 1].u[1]);
                    halo.peer_flag[0] = std::data(m_slabs[k - 1].flags) + 1; // its slot for ITS bottom side
                }
                if(k + 1 < K)
                {
                    halo.peer_u[1][0] = std::data(m_slabs[k + 1].u[0]);
                    halo.peer_u[1][1] = std::data(m_slabs[k + 1].u[1]);
                    halo.peer_flag[1] = std::data(m_slabs[k + 1].flags) + 0; // its slot for ITS top side
                }
                halo.my_flags = std::data(m_slabs[k].flags);
                wait(m_slabs[k].queue);
                check(b200_heat2d_plan_set_halo(m_slabs[k].plan, &halo));
            }
        }
        Heat2DSlabs(Heat2DSlabs const&) = delete;
        auto operator=(Heat2DSlabs const&) -> Heat2DSlabs& = delete;
        ~Heat2DSlabs()
        {
            for(auto& sl : m_slabs)
                checkNoexcept(b200_heat2d_plan_destroy(sl.plan));
        }

        //! both ping-pong arrays of every slab start from its window of the field (ghost rows included)
        void upload(double const* hostField)
        {
            for(auto& sl : m_slabs)
                for(int b = 0; b < 2; ++b)
                    copyRows(sl, b, 0, m_ny + 2u * m_G, const_cast<double*>(hostField), B200_COPY_H2D);
            waitAll();
            m_cur = 0;
        }

        //! `n` steps: launches of `levels` time levels, a remainder in shallower ones (never a single level: n != 1)
        void steps(std::uint32_t n)
        {
            if(n == 1 || (m_G == 2 && n % 2 != 0))
                throw std::runtime_error("Heat2DSlabs::steps: a slab cannot advance a single time level");
            constexpr double pi = math::constants::pi;
            while(n > 0)
            {
                auto const k = heatNextDepth(n, static_cast<int>(m_G), 2);
                if(k == 0)
                    throw std::runtime_error("Heat2DSlabs::steps: the remaining steps cannot be covered");
                double tf[8] = {};
                for(Idx l = 0; l < k; ++l)
                    tf[l] = std::exp(-pi * pi * ((m_step + 1u + l) * m_dt));
                ++m_launch;
                for(auto& sl : m_slabs)
                {
                    if(k == 2)
                        check(b200_heat2d_step2_halo_f64(sl.plan, sl.queue.getNativeHandle(), m_cur, m_rX, m_rY, tf[0], tf[1], m_launch));
                    else
                        check(b200_heat2d_stepn_halo_f64(sl.plan, sl.queue.getNativeHandle(), m_cur, m_rX, m_rY, static_cast<int>(k), tf, m_launch));
                    sl.queue.afterEnqueue();
                }
                m_step += k;
                m_cur ^= 1;
                n -= k;
            }
        }

        //! the rows every slab OWNS (core rows, plus the physical ring row on the first / last slab) into the host field
        void download(double* hostField)
        {
            waitAll();
            for(std::size_t k = 0; k < m_slabs.size(); ++k)
            {
                Idx const j0 = k == 0 ? m_G - 1u : m_G;
                Idx const j1 = k + 1 == m_slabs.size() ? m_ny + m_G + 1u : m_ny + m_G;
                copyRows(m_slabs[k], m_cur, j0, j1, hostField, B200_COPY_D2H);
            }
            waitAll();
            for(auto& sl : m_slabs)
            {
                std::uint32_t status = 0;
                check(b200_heat2d_halo_status(sl.plan, &status));
                if(status != 0)
                    throw std::runtime_error("Heat2DSlabs: a neighbour's flag never arrived (side " + std::to_string(status - 1) + ")");
            }
        }

        void waitAll()
        {
            for(auto& sl : m_slabs)
                wait(sl.queue);
        }

        [[nodiscard]] auto stepsDone() const -> std::uint32_t
        {
            return m_step;
        }
        [[nodiscard]] auto launches() const -> std::uint32_t
        {
            return m_launch;
        }

    private:
        struct Slab
        {
            DevB200 dev;
            Queue queue;
            Buf2 u[2];
            BufFlags flags;
            b200_heat2d_plan_t plan;
            std::int64_t g0; // global padded row of local row 0
        };

        //! local rows [j0, j1) of array `b` <-> the same global rows of the host field, clipped to the field
        void copyRows(Slab& sl, int b, Idx j0, Idx j1, double* hostField, int kind)
        {
            std::int64_t lo = sl.g0 + j0, hi = sl.g0 + j1;
            lo = lo < 0 ? 0 : lo;
            hi = hi > static_cast<std::int64_t>(m_NY) + 2 ? static_cast<std::int64_t>(m_NY) + 2 : hi;
            if(lo >= hi)
                return;
            auto const pitch = static_cast<std::size_t>(getPitchesInBytes(sl.u[b])[0]);
            auto* const devPtr = reinterpret_cast<char*>(std::data(sl.u[b])) + static_cast<std::size_t>(lo - sl.g0) * pitch;
            double* const hostPtr = hostField + static_cast<std::size_t>(lo) * (m_NX + 2u);
            std::size_t const rowBytes = (static_cast<std::size_t>(m_NX) + 2u) * sizeof(double);
            void* const dst = kind == B200_COPY_H2D ? static_cast<void*>(devPtr) : static_cast<void*>(hostPtr);
            void const* const src = kind == B200_COPY_H2D ? static_cast<void const*>(hostPtr) : static_cast<void const*>(devPtr);
            check(b200_memcpy2d_async(
                sl.dev.getNativeHandle(),
                dst,
                kind == B200_COPY_H2D ? pitch : rowBytes,
                src,
                kind == B200_COPY_H2D ? rowBytes : pitch,
                rowBytes,
                static_cast<std::size_t>(hi - lo),
                kind,
                sl.queue.getNativeHandle()));
        }

        Idx m_NY, m_NX, m_G, m_ny = 0;
        double m_dt, m_rX, m_rY;
        std::vector<Slab> m_slabs;
        int m_cur = 0;
        std::uint32_t m_step = 0, m_launch = 0;
    };
    //! Py x Px TILES of ONE heat field driven from one host thread, advanced `levels` (4, 6 or 8) time levels per launch: the
    //! C++ form of alpaka_b200.multi.HeatTileDeep. Tile r = cy * Px + cx lives on devs[r] (devices may repeat). Ghost cells
    //! `levels` deep on all four sides; per launch the rows travel inside the walker kernel and the columns (with the
    //! corners) in the column kernel that follows it (b200_heat2d_tile_plan_create / b200_heat2d_stepn_tile_f64).
    //! Host fields are (NY+2) x (NX+2) doubles, rows unpadded, as the reference driver's host buffer.
    class Heat2DTiles
    {
    public:
        using Idx = std::uint32_t;
        using Buf2 = BufB200<double, DimInt<2u>, Idx>;
        using BufFlags = BufB200<std::uint32_t, DimInt<1u>, Idx>;
        using Queue = QueueB200<NonBlocking>;

        Heat2DTiles(std::vector<DevB200> const& devs, Idx Py, Idx Px, Idx NY, Idx NX, double dx, double dy, double dt, int levels = 4)
            : m_NY(NY)
            , m_NX(NX)
            , m_Py(Py)
            , m_Px(Px)
            , m_G(static_cast<Idx>(levels))
            , m_dt(dt)
            , m_rX(dt / (dx * dx))
            , m_rY(dt / (dy * dy))
        {
            if(Py == 0 || Px == 0 || devs.size() != static_cast<std::size_t>(Py) * Px || !(levels == 4 || levels == 6 || levels == 8)
               || NY % Py != 0 || NX % Px != 0 || NY / Py < 2 * m_G || NX / Px < 2 * m_G)
                throw std::runtime_error(
                    "Heat2DTiles: one device per tile, NY x NX must divide into tiles of at least 2*levels cells per side, levels 4, 6 or 8");
            m_ny = NY / Py;
            m_nx = NX / Px;
            bool distinct = false;
            for(auto const& d : devs)
                distinct = distinct || d.getNativeHandle() != devs[0].getNativeHandle();
            if(distinct)
            {
                int pairs = 0;
                check(b200_enable_peer_all(&pairs));
            }
            constexpr double pi = math::constants::pi;
            auto const K = Py * Px;
            m_tiles.reserve(K);
            for(Idx r = 0; r < K; ++r)
            {
                Idx const cy = r / Px, cx = r % Px;
                Vec<DimInt<2u>, Idx> const ext{m_ny + 2u * m_G, m_nx + 2u * m_G};
                Tile t{devs[r], Queue{devs[r]}, allocBuf<double, Idx>(devs[r], ext), allocBuf<double, Idx>(devs[r], ext),
                       allocBuf<std::uint32_t, Idx>(devs[r], Vec<DimInt<1u>, Idx>{16u}), nullptr,
                       static_cast<std::int64_t>(cy) * m_ny - (static_cast<std::int64_t>(m_G) - 1),
                       static_cast<std::int64_t>(cx) * m_nx - (static_cast<std::int64_t>(m_G) - 1),
                       (cy == 0 ? B200_EDGE_TOP : 0) | (cy == Py - 1 ? B200_EDGE_BOTTOM : 0) | (cx == 0 ? B200_EDGE_LEFT : 0)
                           | (cx == Px - 1 ? B200_EDGE_RIGHT : 0)};
                alpaka::memset(t.queue, t.flags, 0);
                alpaka::memset(t.queue, t.u[0], 0);
                alpaka::memset(t.queue, t.u[1], 0);
                std::vector<double> sx(ext[1]), sy(ext[0]);
                for(Idx i = 0; i < ext[1]; ++i)
                    sx[i] = std::sin(pi * (static_cast<double>(t.gi0 + static_cast<std::int64_t>(i)) * dx));
                for(Idx j = 0; j < ext[0]; ++j)
                    sy[j] = std::sin(pi * (static_cast<double>(t.gj0 + static_cast<std::int64_t>(j)) * dy));
                check(b200_heat2d_tile_plan_create(
                    devs[r].getNativeHandle(),
                    std::data(t.u[0]),
                    std::data(t.u[1]),
                    static_cast<std::size_t>(getPitchesInBytes(t.u[0])[0]),
                    m_ny,
                    m_nx,
                    sx.data(),
                    sy.data(),
                    t.edges,
                    m_G,
                    &t.plan));
                m_tiles.push_back(std::move(t));
            }
            for(Idx r = 0; r < K; ++r)
            {
                Idx const cy = r / Px, cx = r % Px;
                b200_heat2d_halo halo{};
                // side 0 top, 1 bottom, 2 left, 3 right; the neighbour's flag slot is the one for the OPPOSITE side
                auto const wire = [&](int side, Idx nb, int opposite)
                {
                    halo.peer_u[side][0] = std::data(m_tiles[nb].u[0]);
                    halo.peer_u[side][1] = std::data(m_tiles[nb].u[1]);
                    halo.peer_flag[side] = std::data(m_tiles[nb].flags) + opposite;
                };
                if(cy > 0)
                    wire(0, r - Px, 1);
                if(cy + 1 < Py)
                    wire(1, r + Px, 0);
                if(cx > 0)
                    wire(2, r - 1, 3);
                if(cx + 1 < Px)
                    wire(3, r + 1, 2);
                halo.my_flags = std::data(m_tiles[r].flags);
                wait(m_tiles[r].queue);
                check(b200_heat2d_plan_set_halo(m_tiles[r].plan, &halo));
            }
        }
        Heat2DTiles(Heat2DTiles const&) = delete;
        auto operator=(Heat2DTiles const&) -> Heat2DTiles& = delete;
        ~Heat2DTiles()
        {
            for(auto& t : m_tiles)
            {
                try
                {
                    wait(t.queue);
                }
                catch(...)
                {
                }
                checkNoexcept(b200_heat2d_plan_destroy(t.plan));
            }
        }

        //! every tile's window of the host field (ghost cells included; cells outside the field stay 0) into both buffers
        void upload(double const* hostField)
        {
            for(auto& t : m_tiles)
                for(int b = 0; b < 2; ++b)
                    copyBlock(t, b, 0, m_ny + 2u * m_G, 0, m_nx + 2u * m_G, const_cast<double*>(hostField), B200_COPY_H2D);
            waitAll();
            m_cur = 0;
        }

        //! `n` steps in launches of 8, 6 or 4 levels, none deeper than the ghost cells; n must be coverable (e.g. a multiple of 4)
        void steps(std::uint32_t n)
        {
            constexpr double pi = math::constants::pi;
            while(n > 0)
            {
                std::uint32_t k = 0;
                for(std::uint32_t d : {8u, 6u, 4u})
                    if(d <= m_G && d <= n && (n - d == 0 || n - d >= 4) && (n - d) % 2 == 0)
                    {
                        k = d;
                        break;
                    }
                if(k == 0)
                    throw std::runtime_error("Heat2DTiles::steps: the remaining steps cannot be covered by launches of 4, 6 or 8 levels");
                double tf[8] = {};
                for(Idx l = 0; l < k; ++l)
                    tf[l] = std::exp(-pi * pi * ((m_step + 1u + l) * m_dt));
                ++m_launch;
                for(auto& t : m_tiles)
                {
                    check(b200_heat2d_stepn_tile_f64(t.plan, t.queue.getNativeHandle(), m_cur, m_rX, m_rY, static_cast<int>(k), tf, m_launch));
                    t.queue.afterEnqueue();
                }
                m_step += k;
                m_cur ^= 1;
                n -= k;
            }
        }

        //! the cells every tile OWNS (core cells, plus the physical ring on boundary sides) into the host field
        void download(double* hostField)
        {
            waitAll();
            for(auto& t : m_tiles)
            {
                Idx const j0 = (t.edges & B200_EDGE_TOP) ? m_G - 1u : m_G, j1 = (t.edges & B200_EDGE_BOTTOM) ? m_ny + m_G + 1u : m_ny + m_G;
                Idx const i0 = (t.edges & B200_EDGE_LEFT) ? m_G - 1u : m_G, i1 = (t.edges & B200_EDGE_RIGHT) ? m_nx + m_G + 1u : m_nx + m_G;
                copyBlock(t, m_cur, j0, j1, i0, i1, hostField, B200_COPY_D2H);
            }
            waitAll();
            for(auto& t : m_tiles)
            {
                std::uint32_t status = 0;
                check(b200_heat2d_halo_status(t.plan, &status));
                if(status != 0)
                    throw std::runtime_error("Heat2DTiles: a neighbour's flag never arrived (side " + std::to_string(status - 1) + ")");
            }
        }

        void waitAll()
        {
            for(auto& t : m_tiles)
                wait(t.queue);
        }

        [[nodiscard]] auto stepsDone() const -> std::uint32_t
        {
            return m_step;
        }
        [[nodiscard]] auto launches() const -> std::uint32_t
        {
            return m_launch;
        }

    private:
        struct Tile
        {
            DevB200 dev;
            Queue queue;
            Buf2 u[2];
            BufFlags flags;
            b200_heat2d_plan_t plan;
            std::int64_t gj0, gi0; // global padded row / column of local cell (0, 0)
            int edges;
        };

        //! local cells [j0, j1) x [i0, i1) of array `b` <-> the same global cells of the host field, clipped to the field
        void copyBlock(Tile& t, int b, Idx j0, Idx j1, Idx i0, Idx i1, double* hostField, int kind)
        {
            auto const clip = [](std::int64_t v, std::int64_t hi) { return v < 0 ? std::int64_t{0} : (v > hi ? hi : v); };
            std::int64_t const jl = clip(t.gj0 + j0, static_cast<std::int64_t>(m_NY) + 2), jh = clip(t.gj0 + j1, static_cast<std::int64_t>(m_NY) + 2);
            std::int64_t const il = clip(t.gi0 + i0, static_cast<std::int64_t>(m_NX) + 2), ih = clip(t.gi0 + i1, static_cast<std::int64_t>(m_NX) + 2);
            if(jl >= jh || il >= ih)
                return;
            auto const pitch = static_cast<std::size_t>(getPitchesInBytes(t.u[b])[0]);
            auto* const devPtr = reinterpret_cast<char*>(std::data(t.u[b])) + static_cast<std::size_t>(jl - t.gj0) * pitch
                                 + static_cast<std::size_t>(il - t.gi0) * sizeof(double);
            double* const hostPtr = hostField + static_cast<std::size_t>(jl) * (m_NX + 2u) + static_cast<std::size_t>(il);
            std::size_t const hostPitch = (static_cast<std::size_t>(m_NX) + 2u) * sizeof(double);
            std::size_t const rowBytes = static_cast<std::size_t>(ih - il) * sizeof(double);
            void* const dst = kind == B200_COPY_H2D ? static_cast<void*>(devPtr) : static_cast<void*>(hostPtr);
            void const* const src = kind == B200_COPY_H2D ? static_cast<void const*>(hostPtr) : static_cast<void const*>(devPtr);
            check(b200_memcpy2d_async(
                t.dev.getNativeHandle(),
                dst,
                kind == B200_COPY_H2D ? pitch : hostPitch,
                src,
                kind == B200_COPY_H2D ? hostPitch : pitch,
                rowBytes,
                static_cast<std::size_t>(jh - jl),
                kind,
                t.queue.getNativeHandle()));
        }

        Idx m_NY, m_NX, m_Py, m_Px, m_G, m_ny = 0, m_nx = 0;
        double m_dt, m_rX, m_rY;
        std::vector<Tile> m_tiles;
        int m_cur = 0;
        std::uint32_t m_step = 0, m_launch = 0;
    };
} // namespace alpaka::b200
