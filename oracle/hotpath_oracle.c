/* oracle/hotpath_oracle.c -- CPU restatement of the reference's hot path. TEST INFRASTRUCTURE ONLY.
 * See hotpath_oracle.h for scope, parity status and who may call this. Paths cited are relative to
 * /root/reference. Build: oracle/Makefile (gcc -std=c11 -O2 -fopenmp -ffp-contract=off). */
#include "hotpath_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#    define M_PI 3.14159265358979323846
#endif

/* ------------------------------------------------------------------------------------------------
 * BabelStream element-wise kernels. One reference "thread" per element (getIdx<Grid,Threads>), so the
 * result of element i depends on i only; the loops below are order-free.
 * ---------------------------------------------------------------------------------------------- */
#define ORC_STREAMS(T, SFX)                                                                                           \
    /* InitKernel: babelStreamMainTest.cpp:61-68 */                                                                   \
    void orc_init_##SFX(T* a, T* b, T* c, T initA, uint64_t n)                                                        \
    {                                                                                                                 \
        _Pragma("omp parallel for") for(uint64_t i = 0; i < n; ++i)                                                   \
        {                                                                                                             \
            a[i] = initA;                                                                                             \
            b[i] = (T) 0.0;                                                                                           \
            c[i] = (T) 0.0;                                                                                           \
        }                                                                                                             \
    }                                                                                                                 \
    /* CopyKernel: :80-85 */                                                                                          \
    void orc_copy_##SFX(T const* a, T* b, uint64_t n)                                                                 \
    {                                                                                                                 \
        _Pragma("omp parallel for") for(uint64_t i = 0; i < n; ++i) b[i] = a[i];                                      \
    }                                                                                                                 \
    /* MultKernel: :97-103 (reference scalar is fixed to scalarVal = 2.0f, babelStreamCommon.hpp:31) */               \
    void orc_mul_##SFX(T const* a, T* b, T scalar, uint64_t n)                                                        \
    {                                                                                                                 \
        _Pragma("omp parallel for") for(uint64_t i = 0; i < n; ++i) b[i] = scalar * a[i];                             \
    }                                                                                                                 \
    /* AddKernel: :116-121 */                                                                                         \
    void orc_add_##SFX(T const* a, T const* b, T* c, uint64_t n)                                                      \
    {                                                                                                                 \
        _Pragma("omp parallel for") for(uint64_t i = 0; i < n; ++i) c[i] = a[i] + b[i];                               \
    }                                                                                                                 \
    /* TriadKernel: :134-140; product rounded, then sum rounded (no FMA: -ffp-contract=off) */                        \
    void orc_triad_##SFX(T const* a, T const* b, T* c, T scalar, uint64_t n)                                          \
    {                                                                                                                 \
        _Pragma("omp parallel for") for(uint64_t i = 0; i < n; ++i) c[i] = a[i] + scalar * b[i];                      \
    }                                                                                                                 \
    /* Nstream: not in the reference; upstream BabelStream a[i] += b[i] + scalar*c[i], i.e.                           \
     * a = a + (b + (scalar*c)) as C evaluates `a += expr`. */                                                         \
    void orc_nstream_##SFX(T* a, T const* b, T const* c, T scalar, uint64_t n)                                        \
    {                                                                                                                 \
        _Pragma("omp parallel for") for(uint64_t i = 0; i < n; ++i) a[i] += b[i] + scalar * c[i];                     \
    }

ORC_STREAMS(double, f64)
ORC_STREAMS(float, f32)

/* ------------------------------------------------------------------------------------------------
 * Dot. DotKernel (babelStreamMainTest.cpp:154-180): thread (g,t) accumulates a[i]*b[i] for
 * i = g*B+t, +G*B, ... (:164-167), stores into tbSum[t] (:168), then a halving tree over the block with
 * a barrier per level (:170-176), thread 0 writes sum[g] (:178-179). Host: std::reduce(sum, sum+G, T{0})
 * (:402-403), which libstdc++ evaluates 4 elements at a time as init + ((p0+p1)+(p2+p3)).
 * ---------------------------------------------------------------------------------------------- */
#define ORC_DOT(T, SFX)                                                                                               \
    T orc_dot_##SFX(T const* a, T const* b, uint64_t n, uint32_t G, uint32_t B, T* partials)                          \
    {                                                                                                                 \
        T* p = partials ? partials : (T*) malloc((size_t) G * sizeof(T));                                             \
        uint64_t const total = (uint64_t) G * B;                                                                      \
        _Pragma("omp parallel for") for(uint32_t g = 0; g < G; ++g)                                                   \
        {                                                                                                             \
            T tb[1024];                                                                                               \
            for(uint32_t t = 0; t < B; ++t)                                                                           \
            {                                                                                                         \
                T threadSum = 0;                                                                                      \
                for(uint64_t i = (uint64_t) g * B + t; i < n; i += total)                                             \
                    threadSum += a[i] * b[i];                                                                         \
                tb[t] = threadSum;                                                                                    \
            }                                                                                                         \
            for(uint32_t offset = B / 2; offset > 0; offset /= 2)                                                     \
                for(uint32_t t = 0; t < offset; ++t)                                                                  \
                    tb[t] += tb[t + offset];                                                                          \
            p[g] = tb[0];                                                                                             \
        }                                                                                                             \
        T init = 0;                                                                                                   \
        uint32_t k = 0;                                                                                               \
        for(; G - k >= 4; k += 4)                                                                                     \
        {                                                                                                             \
            T v1 = p[k] + p[k + 1];                                                                                   \
            T v2 = p[k + 2] + p[k + 3];                                                                               \
            T v3 = v1 + v2;                                                                                           \
            init = init + v3;                                                                                         \
        }                                                                                                             \
        for(; k < G; ++k)                                                                                             \
            init = init + p[k];                                                                                       \
        if(!partials)                                                                                                 \
            free(p);                                                                                                  \
        return init;                                                                                                  \
    }

ORC_DOT(double, f64)
ORC_DOT(float, f32)

/* ------------------------------------------------------------------------------------------------
 * example/reduce. reduce.cpp:55-63 picks the launch shape; ReduceKernel (kernel.hpp:58-131) runs once
 * over the input with blockCount blocks and once with one block over the blockCount partials, in place
 * (reduce.cpp:75-98).
 * ---------------------------------------------------------------------------------------------- */
uint32_t orc_reduce_block_count(uint64_t n, uint32_t multiProcessorCount, uint32_t blockSize)
{
    uint32_t blockCount = multiProcessorCount * 8u; /* reduce.cpp:59 */
    uint32_t maxBlockCount = (uint32_t) ((((n + 1) / 2) - 1) / blockSize + 1); /* reduce.cpp:60 */
    if(blockCount > maxBlockCount)
        blockCount = maxBlockCount;
    return blockCount;
}

#define ORC_REDUCE(T, SFX)                                                                                            \
    /* one launch of ReduceKernel<blockSize,T,add> with `blocks` blocks; returns -1 on an access the                  \
     * reference would make out of bounds */                                                                         \
    static int orc_reduce_launch_##SFX(                                                                               \
        T const* source,                                                                                              \
        T* destination,                                                                                               \
        uint64_t n,                                                                                                   \
        uint32_t blocks,                                                                                              \
        uint32_t blockSize,                                                                                           \
        int iterator)                                                                                                 \
    {                                                                                                                 \
        int bad = 0;                                                                                                  \
        uint32_t const gridSize = blocks * blockSize; /* kernel.hpp:74: gridDimension * TBlockSize */                 \
        T* results = (T*) malloc((size_t) blocks * sizeof(T));                                                        \
        _Pragma("omp parallel for reduction(| : bad)") for(uint32_t blockIndex = 0; blockIndex < blocks; ++blockIndex)\
        {                                                                                                             \
            T* sdata = (T*) calloc(blockSize, sizeof(T));                                                             \
            for(uint32_t threadIndex = 0; threadIndex < blockSize; ++threadIndex)                                     \
            {                                                                                                         \
                uint32_t const lin = blockIndex * blockSize + threadIndex; /* kernel.hpp:72 */                        \
                uint64_t idx, end, stride;                                                                            \
                if(iterator == 0)                                                                                     \
                { /* IteratorCpu ctor, iterator.hpp:126-138 (uint32 casts kept) */                                    \
                    uint64_t const m = (uint64_t) gridSize < n ? (uint64_t) gridSize : n;                             \
                    idx = (uint32_t) ((n * lin) / m);                                                                 \
                    end = (uint32_t) ((n * ((uint64_t) lin + 1)) / m);                                                \
                    stride = 1;                                                                                       \
                }                                                                                                     \
                else                                                                                                  \
                { /* IteratorGpu, iterator.hpp:247-257: begin = lin, ++ adds gridSize, end = n */                     \
                    idx = lin;                                                                                        \
                    end = n;                                                                                          \
                    stride = gridSize;                                                                                \
                }                                                                                                     \
                T result = 0;                                                                                         \
                if(threadIndex < n)                                                                                   \
                { /* kernel.hpp:80-82: result = *(it++) */                                                            \
                    if(idx >= n)                                                                                      \
                    {                                                                                                 \
                        bad = 1;                                                                                      \
                        continue;                                                                                     \
                    }                                                                                                 \
                    result = source[idx];                                                                             \
                    idx += stride;                                                                                    \
                }                                                                                                     \
                /* kernel.hpp:90-94: while(it + 3 < it.end()) 4x unrolled */                                          \
                while(idx + 3 * stride < end)                                                                         \
                {                                                                                                     \
                    T const x0 = source[idx], x1 = source[idx + stride];                                              \
                    T const x2 = source[idx + 2 * stride], x3 = source[idx + 3 * stride];                             \
                    result = (T) ((T) ((T) (result + (T) (x0 + x1)) + x2) + x3);                                      \
                    idx += 4 * stride;                                                                                \
                }                                                                                                     \
                /* kernel.hpp:97-98 */                                                                                \
                while(idx < end)                                                                                      \
                {                                                                                                     \
                    result = (T) (result + source[idx]);                                                              \
                    idx += stride;                                                                                    \
                }                                                                                                     \
                if(threadIndex < n)                                                                                   \
                    sdata[threadIndex] = result; /* kernel.hpp:100-101 */                                             \
            }                                                                                                         \
            /* kernel.hpp:109-126: halving tree, barrier per level */                                                 \
            for(uint32_t cbs = blockSize, up = (blockSize + 1) / 2; cbs > 1; cbs = cbs / 2, up = (cbs + 1) / 2)       \
            {                                                                                                         \
                for(uint32_t threadIndex = 0; threadIndex < blockSize; ++threadIndex)                                 \
                {                                                                                                     \
                    int const cond = threadIndex < up && (threadIndex + up) < blockSize                               \
                                     && ((uint64_t) (uint32_t) (blockIndex * blockSize + threadIndex + up)) < n       \
                                     && threadIndex < n;                                                              \
                    if(cond)                                                                                          \
                        sdata[threadIndex] = (T) (sdata[threadIndex] + sdata[threadIndex + up]);                      \
                }                                                                                                     \
            }                                                                                                         \
            results[blockIndex] = sdata[0];                                                                           \
            free(sdata);                                                                                              \
        }                                                                                                             \
        /* kernel.hpp:129-130 (threadIndex == 0 < n always holds for n >= 1); written after all blocks have          \
         * read their input because the second launch runs in place on `destination`. */                             \
        for(uint32_t blockIndex = 0; blockIndex < blocks; ++blockIndex)                                               \
            destination[blockIndex] = results[blockIndex];                                                            \
        free(results);                                                                                                \
        return bad ? -1 : 0;                                                                                          \
    }                                                                                                                 \
    int orc_reduce_##SFX(T const* src, uint64_t n, uint32_t blockCount, uint32_t blockSize, int iterator, T* out)     \
    {                                                                                                                 \
        if(n == 0 || blockCount == 0 || blockSize == 0)                                                               \
            return -1;                                                                                                \
        if(iterator == 0 && n > 0xffffffffull)                                                                        \
            return -1; /* IteratorCpu's uint32 casts wrap */                                        \
        T* dest = (T*) calloc(blockCount, sizeof(T));                                                                 \
        int rc = orc_reduce_launch_##SFX(src, dest, n, blockCount, blockSize, iterator);                              \
        if(rc == 0)                                                                                                   \
            rc = orc_reduce_launch_##SFX(dest, dest, (uint64_t) blockCount, 1u, blockSize, iterator);                 \
        *out = dest[0];                                                                                               \
        free(dest);                                                                                                   \
        return rc;                                                                                                    \
    }

ORC_REDUCE(uint32_t, u32)
ORC_REDUCE(int32_t, i32)
ORC_REDUCE(uint64_t, u64)
ORC_REDUCE(float, f32)
ORC_REDUCE(double, f64)

/* ------------------------------------------------------------------------------------------------
 * example/heatEquation2D.
 * ---------------------------------------------------------------------------------------------- */
/* analyticalSolution.hpp:17-21: exp(-pi*pi*t) * (sin(pi*x) + sin(pi*y)); unary minus binds first. */
double orc_heat2d_exact(double x, double y, double t)
{
    double const pi = M_PI;
    return exp(-pi * pi * t) * (sin(pi * x) + sin(pi * y));
}

/* analyticalSolution.hpp:58-71 */
void orc_heat2d_init(double* u, uint32_t ny, uint32_t nx, size_t pitch, double dx, double dy)
{
    for(uint32_t j = 0; j < ny + 2; ++j)
        for(uint32_t i = 0; i < nx + 2; ++i)
            u[(size_t) j * pitch + i] = orc_heat2d_exact(i * dx, j * dy, 0.0);
}

/* analyticalSolution.hpp:31-51 (core cells only) */
double orc_heat2d_validate(double const* u, uint32_t ny, uint32_t nx, size_t pitch, double dx, double dy, double tMax)
{
    double maxError = 0.0;
    for(uint32_t j = 1; j < ny + 1; ++j)
        for(uint32_t i = 1; i < nx + 1; ++i)
        {
            double const error = fabs(u[(size_t) j * pitch + i] - orc_heat2d_exact(i * dx, j * dy, tMax));
            maxError = error > maxError ? error : maxError;
        }
    return maxError;
}

void orc_heat2d_boundary_tables(double* sx, double* sy, uint32_t ny, uint32_t nx, double dx, double dy)
{
    double const pi = M_PI;
    for(uint32_t i = 0; i < nx + 2; ++i)
        sx[i] = sin(pi * (i * dx));
    for(uint32_t j = 0; j < ny + 2; ++j)
        sy[j] = sin(pi * (j * dy));
}

double orc_heat2d_time_factor(uint32_t step, double dt)
{
    double const pi = M_PI;
    return exp(-pi * pi * (step * dt));
}

void orc_heat2d_step(
    double const* uCurr,
    double* uNext,
    uint32_t ny,
    uint32_t nx,
    size_t pitch,
    uint32_t step,
    double dx,
    double dy,
    double dt)
{
    /* StencilKernel.hpp:70-71 */
    double const rX = dt / (dx * dx);
    double const rY = dt / (dy * dy);
    /* StencilKernel.hpp:84-86. sdata[localIdx1D -/+ 1] are the x neighbours, -/+ (chunkSize[1]+halo[1]) the
     * y neighbours; C evaluates the sum left to right. */
#pragma omp parallel for
    for(uint32_t j = 1; j <= ny; ++j)
    {
        double const* c = uCurr + (size_t) j * pitch;
        double const* up = c - pitch;
        double const* dn = c + pitch;
        double* o = uNext + (size_t) j * pitch;
        for(uint32_t i = 1; i <= nx; ++i)
            o[i] = c[i] * (1.0 - 2.0 * rX - 2.0 * rY) + c[i - 1] * rX + c[i + 1] * rX + up[i] * rY + dn[i] * rY;
    }
    /* BoundaryKernel.hpp:52-84: top/bottom rows for x index 1..nx, left/right columns for y index 1..ny,
     * value exactSolution(idx2D[1]*dx, idx2D[0]*dy, step*dt) with uint32 indices (:58). */
    for(uint32_t i = 1; i <= nx; ++i)
    {
        uNext[i] = orc_heat2d_exact(i * dx, 0u * dy, step * dt);
        uNext[(size_t) (ny + 1) * pitch + i] = orc_heat2d_exact(i * dx, (ny + 1) * dy, step * dt);
    }
    for(uint32_t j = 1; j <= ny; ++j)
    {
        uNext[(size_t) j * pitch] = orc_heat2d_exact(0u * dx, j * dy, step * dt);
        uNext[(size_t) j * pitch + nx + 1] = orc_heat2d_exact((nx + 1) * dx, j * dy, step * dt);
    }
}

int orc_heat2d_run(
    double* u,
    uint32_t ny,
    uint32_t nx,
    uint32_t stepFirst,
    uint32_t numSteps,
    double dx,
    double dy,
    double dt)
{
    size_t const pitch = (size_t) nx + 2;
    size_t const bytes = pitch * ((size_t) ny + 2) * sizeof(double);
    double* curr = (double*) malloc(bytes);
    double* next = (double*) malloc(bytes);
    if(!curr || !next)
    {
        free(curr);
        free(next);
        return -1;
    }
    memcpy(curr, u, bytes);
    memcpy(next, u, bytes);
    for(uint32_t step = stepFirst; step < stepFirst + numSteps; ++step)
    {
        orc_heat2d_step(curr, next, ny, nx, pitch, step, dx, dy, dt);
        double* t = curr; /* std::swap(uNextBufAcc, uCurrBufAcc), heatEquation2D.cpp:181 */
        curr = next;
        next = t;
    }
    memcpy(u, curr, bytes);
    free(curr);
    free(next);
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * Seeded synthetic inputs: splitmix64 of (seed, global index) -> any shard is reproducible on its own.
 * ---------------------------------------------------------------------------------------------- */
static inline uint64_t orc_mix(uint64_t seed, uint64_t i)
{
    uint64_t z = seed + (i + 1) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

void orc_fill_uniform_f64(double* x, uint64_t first, uint64_t count, uint64_t seed)
{
#pragma omp parallel for
    for(uint64_t k = 0; k < count; ++k)
        x[k] = (double) (orc_mix(seed, first + k) >> 11) * (1.0 / 4503599627370496.0) - 1.0; /* 2^-52 * 53 bits */
}

void orc_fill_uniform_f32(float* x, uint64_t first, uint64_t count, uint64_t seed)
{
#pragma omp parallel for
    for(uint64_t k = 0; k < count; ++k)
        x[k] = (float) (orc_mix(seed, first + k) >> 40) * (1.0f / 8388608.0f) - 1.0f; /* 24 bits * 2^-23 */
}

void orc_fill_hash_u32(uint32_t* x, uint64_t first, uint64_t count, uint64_t seed)
{
#pragma omp parallel for
    for(uint64_t k = 0; k < count; ++k)
        x[k] = (uint32_t) (orc_mix(seed, first + k) >> 32);
}

void orc_fill_bernoulli_f32(float* x, uint64_t first, uint64_t count, uint64_t seed)
{
#pragma omp parallel for
    for(uint64_t k = 0; k < count; ++k)
        x[k] = (float) (orc_mix(seed, first + k) >> 63);
}
