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
#include <vector>

namespace alpaka::b200
{
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

        //! `levels` (3 or 4) fused steps in ONE launch (b200_heat2d_stepn_f64: own column pair per thread and level, the
        //! horizontal neighbours by warp shuffle); same bits, the roles of the buffers swap ONCE.
        template<typename TQueue>
        void stepN(TQueue& queue, int levels)
        {
            constexpr double pi = math::constants::pi;
            double tf[4] = {};
            for(int l = 0; l < levels && l < 4; ++l)
                tf[l] = std::exp(-pi * pi * ((m_step + 1u + static_cast<std::uint32_t>(l)) * m_dt));
            check(b200_heat2d_stepn_f64(m_plan, queue.getNativeHandle(), m_cur, m_rX, m_rY, levels, tf));
            m_step += static_cast<std::uint32_t>(levels);
            m_cur ^= 1;
            queue.afterEnqueue();
        }

        //! `n` steps with up to `depth` (1..4; measured best: 3) time levels per launch; a remainder runs in shallower
        //! launches (4 = 2 + 2 rather than 3 + 1)
        template<typename TQueue>
        void steps(TQueue& queue, std::uint32_t n, int depth = 3)
        {
            if(depth < 1 || depth > 4)
                throw std::runtime_error("Heat2DStepper::steps: between 1 and 4 time levels per launch");
            while(n > 0)
            {
                auto k = static_cast<std::uint32_t>(depth) < n ? static_cast<std::uint32_t>(depth) : n;
                if(k > 2 && n - k == 1)
                    --k;
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

    private:
        b200_heat2d_plan_t m_plan = nullptr;
        double m_dt, m_rX, m_rY;
        int m_cur = 0;
        std::uint32_t m_step = 0;
    };
} // namespace alpaka::b200
