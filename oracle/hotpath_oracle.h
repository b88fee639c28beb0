/* oracle/hotpath_oracle.h -- CPU restatement of the reference's hot path. TEST INFRASTRUCTURE ONLY.
 *
 * Plain C11 restatement of the algorithms on the path named by BASELINE.json `north_star`
 * (alpaka: benchmarks/babelstream, example/reduce, example/heatEquation2D). Every function cites the
 * reference file:line it follows (paths relative to /root/reference). It is the CHECKER for the CUDA
 * product path: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it; nothing under alpaka_b200/ or include/ may call it.
 *
 * Parity status: PINNED. tests/test_oracle_vs_ref.py compares every function here bit-for-bit against
 * the unmodified reference compiled into oracle/_ref/libalpaka_ref.so (oracle/Makefile), and
 * tests/golden/ holds vectors generated from that build (tests/golden/make_golden.py), plus the
 * reference's own known answers (A=1,B=2,C=5, Dot=2N: babelStreamMainTest.cpp:353-355,405; reduce closed
 * form: reduce.cpp:148; heat analytic max-abs < 1e-4 at 64x64/4000 steps: analyticalSolution.hpp:49).
 * Exception: Nstream has no reference implementation at this commit ("parity unpinned" for that one
 * kernel); it is pinned only against the same functor run through the reference
 * CPU back-end in oracle/ref_babelstream.cpp.
 *
 * Floating point: built with -ffp-contract=off; all expressions keep the reference's operation order.
 */
#ifndef B200_HOTPATH_ORACLE_H
#define B200_HOTPATH_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C"
{
#endif

    /* ---- BabelStream element-wise kernels (babelStreamMainTest.cpp:53-141) ---- */
    void orc_init_f64(double* a, double* b, double* c, double initA, uint64_t n);
    void orc_copy_f64(double const* a, double* b, uint64_t n);
    void orc_mul_f64(double const* a, double* b, double scalar, uint64_t n);
    void orc_add_f64(double const* a, double const* b, double* c, uint64_t n);
    void orc_triad_f64(double const* a, double const* b, double* c, double scalar, uint64_t n);
    void orc_nstream_f64(double* a, double const* b, double const* c, double scalar, uint64_t n);
    void orc_init_f32(float* a, float* b, float* c, float initA, uint64_t n);
    void orc_copy_f32(float const* a, float* b, uint64_t n);
    void orc_mul_f32(float const* a, float* b, float scalar, uint64_t n);
    void orc_add_f32(float const* a, float const* b, float* c, uint64_t n);
    void orc_triad_f32(float const* a, float const* b, float* c, float scalar, uint64_t n);
    void orc_nstream_f32(float* a, float const* b, float const* c, float scalar, uint64_t n);

    /* ---- Dot (babelStreamMainTest.cpp:145-181 device part, :402-403 host finish) ----
     * Work division {gridBlocks, blockThreads, 1}; blockThreads must be a power of two <= 1024.
     * partials (gridBlocks entries) may be NULL. Returns the host-side std::reduce of the partials in
     * libstdc++ 13 order (4-way grouped, bits/stl_numeric / <numeric> reduce for random-access iterators). */
    double orc_dot_f64(double const* a, double const* b, uint64_t n, uint32_t gridBlocks, uint32_t blockThreads, double* partials);
    float orc_dot_f32(float const* a, float const* b, uint64_t n, uint32_t gridBlocks, uint32_t blockThreads, float* partials);

    /* ---- example/reduce (kernel.hpp:58-131 run twice as reduce.cpp:75-98) ----
     * iterator: 0 = IteratorCpu (contiguous chunk per thread, iterator.hpp:115-138),
     *           1 = IteratorGpu (grid-strided, iterator.hpp:239-257).
     * blockCount/blockSize are the launch shape of the main kernel (reduce.cpp:55-63,75); use
     * orc_reduce_block_count() for the reference's choice. Returns 0, or -1 if the shape would make the
     * reference read out of bounds. */
    uint32_t orc_reduce_block_count(uint64_t n, uint32_t multiProcessorCount, uint32_t blockSize);
    int orc_reduce_u32(uint32_t const* src, uint64_t n, uint32_t blockCount, uint32_t blockSize, int iterator, uint32_t* out);
    int orc_reduce_i32(int32_t const* src, uint64_t n, uint32_t blockCount, uint32_t blockSize, int iterator, int32_t* out);
    int orc_reduce_u64(uint64_t const* src, uint64_t n, uint32_t blockCount, uint32_t blockSize, int iterator, uint64_t* out);
    int orc_reduce_f32(float const* src, uint64_t n, uint32_t blockCount, uint32_t blockSize, int iterator, float* out);
    int orc_reduce_f64(double const* src, uint64_t n, uint32_t blockCount, uint32_t blockSize, int iterator, double* out);

    /* ---- example/heatEquation2D ----
     * Field layout: (ny+2) x (nx+2) doubles, row-major, row pitch `pitchElems` doubles (>= nx+2);
     * index [j][i], j = y (slow), i = x (fast); ring of boundary cells around ny x nx core cells
     * (heatEquation2D.cpp:54-56). */
    double orc_heat2d_exact(double x, double y, double t); /* analyticalSolution.hpp:17-21 */
    void orc_heat2d_init(double* u, uint32_t ny, uint32_t nx, size_t pitchElems, double dx, double dy); /* :58-71 */
    double orc_heat2d_validate(double const* u, uint32_t ny, uint32_t nx, size_t pitchElems, double dx, double dy, double tMax); /* :31-51 */
    /* One FTCS step: StencilKernel.hpp:69-87 into uNext core cells, then BoundaryKernel.hpp:52-84 onto the
     * ring of uNext (corners untouched) for time level `step`. */
    void orc_heat2d_step(double const* uCurr, double* uNext, uint32_t ny, uint32_t nx, size_t pitchElems, uint32_t step, double dx, double dy, double dt);
    /* numSteps steps numbered stepFirst.. (driver loop heatEquation2D.cpp:141-182); u is in/out, both
     * internal buffers start as copies of u (see oracle/ref_heat2d.cpp note on corners). */
    int orc_heat2d_run(double* u, uint32_t ny, uint32_t nx, uint32_t stepFirst, uint32_t numSteps, double dx, double dy, double dt);
    /* The separable boundary factors the CUDA path is fed with (host, glibc): sx[i] = sin(pi*(i*dx)),
     * sy[j] = sin(pi*(j*dy)), e(step) = exp(-pi*pi*(step*dt)); exact == e * (sx[i] + sy[j]) bit-for-bit. */
    void orc_heat2d_boundary_tables(double* sx, double* sy, uint32_t ny, uint32_t nx, double dx, double dy);
    double orc_heat2d_time_factor(uint32_t step, double dt);

    /* ---- seeded synthetic inputs: counter-based hash so shards generate independently */
    void orc_fill_uniform_f64(double* x, uint64_t first, uint64_t count, uint64_t seed); /* U[-1,1) */
    void orc_fill_uniform_f32(float* x, uint64_t first, uint64_t count, uint64_t seed); /* U[-1,1) */
    void orc_fill_hash_u32(uint32_t* x, uint64_t first, uint64_t count, uint64_t seed);
    void orc_fill_bernoulli_f32(float* x, uint64_t first, uint64_t count, uint64_t seed); /* {0,1} */

#ifdef __cplusplus
}
#endif
#endif
