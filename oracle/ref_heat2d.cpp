/* oracle/_ref harness, part 3/3: example/heatEquation2D on the reference's own CPU back-ends.
 *
 * TEST INFRASTRUCTURE ONLY (see ref_babelstream.cpp). StencilKernel, BoundaryKernel, exactSolution,
 * validateSolution, initalizeBuffer and getElementPtr are the reference's files included verbatim from
 * -I/root/reference/example/heatEquation2D/src. The shipped driver hard-codes a 64x64 grid, 4000 steps
 * and tMax = 0.1 (heatEquation2D.cpp:54-59); the time loop of heatEquation2D.cpp:101-182 is re-issued
 * here with extent / steps / dx / dy / dt as arguments. Chunk size (16x16), shared tile ((16+2)^2), work
 * division and launch order (Stencil, Boundary, swap) are the driver's.
 *
 * Note on corners: BoundaryKernel never writes the four corner cells (BoundaryKernel.hpp:63-84) and the
 * stencil never reads them, while the driver's uNext buffer starts uninitialised. To make whole-buffer
 * comparisons well defined BOTH buffers start as copies of the caller's field here (the corners keep their
 * initial value for ever); every other cell is identical to what the shipped driver computes.
 */
#include "BoundaryKernel.hpp" // reference file
#include "StencilKernel.hpp" // reference file
#include "analyticalSolution.hpp" // reference file

#include <alpaka/alpaka.hpp>

#include <chrono>
#include <cstdint>
#include <cstring>
#include <utility>

namespace
{
    using HDim = alpaka::DimInt<2u>;
    using HIdx = std::uint32_t;
    using Vec2 = alpaka::Vec<HDim, HIdx>;

    template<typename TAcc>
    int runOn(
        double* u,
        HIdx ny,
        HIdx nx,
        std::uint32_t stepFirst,
        std::uint32_t numSteps,
        double dx,
        double dy,
        double dt,
        double* seconds)
    {
        Vec2 const numNodes{ny, nx};
        Vec2 const extent = numNodes + Vec2{2, 2};

        auto const devHost = alpaka::getDevByIdx(alpaka::PlatformCpu{}, 0);
        auto const devAcc = alpaka::getDevByIdx(alpaka::Platform<TAcc>{}, 0);

        auto uCurrBufAcc = alpaka::allocBuf<double, HIdx>(devAcc, extent);
        auto uNextBufAcc = alpaka::allocBuf<double, HIdx>(devAcc, extent);
        auto const pitchCurrAcc{alpaka::getPitchesInBytes(uCurrBufAcc)};
        auto const pitchNextAcc{alpaka::getPitchesInBytes(uNextBufAcc)};

        std::size_t const bytes = static_cast<std::size_t>(extent[0]) * extent[1] * sizeof(double);
        std::memcpy(uCurrBufAcc.data(), u, bytes);
        std::memcpy(uNextBufAcc.data(), u, bytes);

        using QueueAcc = alpaka::Queue<TAcc, alpaka::NonBlocking>;
        QueueAcc computeQueue{devAcc};

        // heatEquation2D.cpp:101-138
        constexpr Vec2 elemPerThread{1, 1};
        constexpr HIdx xSize = 16u;
        constexpr HIdx ySize = 16u;
        constexpr HIdx halo = 2u;
        constexpr Vec2 chunkSize{ySize, xSize};
        constexpr auto sharedMemSize = (ySize + halo) * (xSize + halo);
        if(numNodes[0] % chunkSize[0] != 0 || numNodes[1] % chunkSize[1] != 0)
            return -1; // "Domain must be divisible by chunk size" (heatEquation2D.cpp:114-116)
        Vec2 const numChunks{
            alpaka::core::divCeil(numNodes[0], chunkSize[0]),
            alpaka::core::divCeil(numNodes[1], chunkSize[1]),
        };

        StencilKernel<sharedMemSize> stencilKernel;
        BoundaryKernel boundaryKernel;

        auto const kernelFunctionAttributes = alpaka::getFunctionAttributes<TAcc>(
            devAcc,
            stencilKernel,
            uCurrBufAcc.data(),
            uNextBufAcc.data(),
            chunkSize,
            pitchCurrAcc,
            pitchNextAcc,
            dx,
            dy,
            dt);
        auto const maxThreadsPerBlock = static_cast<HIdx>(kernelFunctionAttributes.maxThreadsPerBlock);
        auto const threadsPerBlock = maxThreadsPerBlock < chunkSize.prod() ? Vec2{maxThreadsPerBlock, 1} : chunkSize;
        alpaka::WorkDivMembers<HDim, HIdx> workDiv_manual{numChunks, threadsPerBlock, elemPerThread};

        auto const t0 = std::chrono::high_resolution_clock::now();
        for(std::uint32_t step = stepFirst; step < stepFirst + numSteps; ++step)
        {
            alpaka::exec<TAcc>(
                computeQueue,
                workDiv_manual,
                stencilKernel,
                uCurrBufAcc.data(),
                uNextBufAcc.data(),
                chunkSize,
                pitchCurrAcc,
                pitchNextAcc,
                dx,
                dy,
                dt);
            alpaka::exec<TAcc>(
                computeQueue,
                workDiv_manual,
                boundaryKernel,
                uNextBufAcc.data(),
                chunkSize,
                pitchNextAcc,
                step,
                dx,
                dy,
                dt);
            std::swap(uNextBufAcc, uCurrBufAcc);
        }
        alpaka::wait(computeQueue);
        auto const t1 = std::chrono::high_resolution_clock::now();
        if(seconds)
            *seconds = std::chrono::duration<double>(t1 - t0).count();

        std::memcpy(u, uCurrBufAcc.data(), bytes);
        (void) devHost;
        return 0;
    }

    //! Host view over caller memory with the two members the reference helpers use.
    struct HostField
    {
        double* p;
        Vec2 ext;

        [[nodiscard]] auto data() const -> double*
        {
            return p;
        }
    };
} // namespace

namespace alpaka::trait
{
    template<>
    struct GetExtents<HostField>
    {
        auto operator()(HostField const& f) const -> Vec2
        {
            return f.ext;
        }
    };

    template<>
    struct DimType<HostField>
    {
        using type = HDim;
    };

    template<>
    struct IdxType<HostField>
    {
        using type = HIdx;
    };
} // namespace alpaka::trait

extern "C"
{
    //! Advance `u` ((ny+2) x (nx+2) doubles, row-major, unpadded) by numSteps FTCS steps numbered
    //! stepFirst .. stepFirst+numSteps-1 (the driver starts at 1). acc: 0 AccCpuSerial, 1 AccCpuOmp2Blocks.
    int ref_heat2d_run(
        int acc,
        double* u,
        std::uint32_t ny,
        std::uint32_t nx,
        std::uint32_t stepFirst,
        std::uint32_t numSteps,
        double dx,
        double dy,
        double dt,
        double* seconds)
    {
        try
        {
            if(acc == 1)
                return runOn<alpaka::AccCpuOmp2Blocks<HDim, HIdx>>(u, ny, nx, stepFirst, numSteps, dx, dy, dt, seconds);
            return runOn<alpaka::AccCpuSerial<HDim, HIdx>>(u, ny, nx, stepFirst, numSteps, dx, dy, dt, seconds);
        }
        catch(...)
        {
            return -2;
        }
    }

    //! initalizeBuffer (analyticalSolution.hpp:58-71) over caller memory.
    void ref_heat2d_init(double* u, std::uint32_t ny, std::uint32_t nx, double dx, double dy)
    {
        HostField f{u, Vec2{ny + 2, nx + 2}};
        initalizeBuffer(f, dx, dy);
    }

    //! validateSolution (analyticalSolution.hpp:31-51): returns the max-abs error against the analytic field.
    double ref_heat2d_validate(double const* u, std::uint32_t ny, std::uint32_t nx, double dx, double dy, double tMax)
    {
        HostField f{const_cast<double*>(u), Vec2{ny + 2, nx + 2}};
        return validateSolution(f, f.ext, dx, dy, tMax).second;
    }

    //! exactSolution (analyticalSolution.hpp:17-21).
    double ref_heat2d_exact(double x, double y, double t)
    {
        return exactSolution(x, y, t);
    }
}
