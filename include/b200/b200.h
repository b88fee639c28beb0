/* include/b200/b200.h -- C ABI of libalpaka_b200.so: the B200-native runtime + hand-written sm_100a kernels
 * underneath alpaka's accelerator/trait API.
 *
 * This is the drop-in boundary. The reference (alpaka) reaches CUDA through a struct of static inline
 * wrappers, `alpaka::ApiCudaRt` (include/alpaka/core/ApiCudaRt.hpp:107-396), called from its trait
 * specialisations; its only device entry point is the generic trampoline
 * `alpaka::detail::gpuKernel` (include/alpaka/kernel/TaskKernelGpuUniformCudaHipRt.hpp:61-76).
 * Here the same trait specialisations (include/alpaka/, C++20) call the functions below instead: plain
 * pointers, sizes and opaque handles, no C++ or torch types. Each entry cites the reference interface it
 * replaces. Paths are relative to the reference root.
 *
 * Conventions
 *  - every function returns 0 on success, a positive cudaError_t value for CUDA failures, or a negative
 *    B200_E* code; b200_last_error_string() describes the last failure on the calling thread (the
 *    reference throws std::runtime_error with the same text from ALPAKA_UNIFORM_CUDA_HIP_RT_CHECK,
 *    core/UniformCudaHip.hpp:23-112; the C++ layer above re-throws).
 *  - like the reference, every call selects its device first (cudaSetDevice; e.g.
 *    kernel/TaskKernelGpuUniformCudaHipRt.hpp:265); kernel entries take the stream and run on the device
 *    the stream belongs to.
 *  - kernel entries are asynchronous: they enqueue on `stream` and return.
 *  - there is NO CPU fallback anywhere: without a CUDA device every entry that needs one fails.
 */
#ifndef B200_B200_H
#define B200_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C"
{
#endif

#define B200_ABI_VERSION 1

    /* negative (non-CUDA) error codes */
#define B200_EINVAL (-1) /* bad argument (null pointer, zero size where not allowed, bad enum) */
#define B200_EALIGN (-2) /* pointer / pitch alignment the kernel requires is not met */
#define B200_ENODEV (-3) /* no CUDA device / driver entry point unavailable */
#define B200_ERANGE (-4) /* extent does not fit the kernel's index range */

    typedef struct b200_stream_st* b200_stream_t; /* == cudaStream_t */
    typedef struct b200_event_st* b200_event_t; /* == cudaEvent_t */

    int b200_abi_version(void);
    char const* b200_last_error_string(void);
    /* name of a positive (CUDA) or negative (B200_E*) code; never NULL */
    char const* b200_error_name(int code);

    /* ---------------------------------------------------------------------------------------------
     * Platform / device  (replaces platform/PlatformUniformCudaHipRt.hpp:41-106,
     * dev/DevUniformCudaHipRt.hpp:109-266, acc/AccGpuUniformCudaHipRt.hpp:113-187)
     * ------------------------------------------------------------------------------------------- */
    typedef struct b200_device_props
    {
        char name[256];
        int32_t cc_major, cc_minor;
        int32_t multi_processor_count; /* AccDevProps::m_multiProcessorCount */
        int32_t max_grid_dim[3]; /* x,y,z -> m_gridBlockExtentMax (reversed into alpaka's slow..fast order above) */
        int32_t max_block_dim[3]; /* x,y,z -> m_blockThreadExtentMax */
        int32_t max_threads_per_block; /* m_blockThreadCountMax */
        int32_t warp_size;
        uint64_t shared_mem_per_block; /* default limit (48 KiB), m_sharedMemSizeBytes */
        uint64_t shared_mem_per_block_optin; /* opt-in limit (227 KiB) used internally by the native kernels */
        uint64_t total_global_mem; /* m_globalMemSizeBytes */
        uint64_t free_global_mem;
        int32_t l2_cache_bytes;
        int32_t memory_pools_supported;
    } b200_device_props;

    int b200_device_count(int* count);
    int b200_device_props_get(int dev, b200_device_props* out);
    int b200_device_mem_info(int dev, uint64_t* free_bytes, uint64_t* total_bytes);
    int b200_device_sync(int dev); /* trait::CurrentThreadWaitFor<Dev> */
    int b200_device_reset(int dev); /* trait::Reset */
    /* cudaDeviceEnablePeerAccess all-to-all, and read/write access of every device to every other device's
     * stream-ordered pool (cudaMemPoolSetAccess: peer access proper covers cudaMalloc memory only), so that kernels may
     * load and store through pointers of buffers on other devices. The reference never enables peer access (SURVEY.md
     * section 2.1) and relies on UVA staging. n_pairs_enabled may be NULL. */
    int b200_enable_peer_all(int* n_pairs_enabled);

    /* ---------------------------------------------------------------------------------------------
     * Queue = stream  (replaces queue/cuda_hip/QueueUniformCudaHipRt.hpp:40-242)
     * ------------------------------------------------------------------------------------------- */
    int b200_stream_create(int dev, b200_stream_t* out); /* cudaStreamNonBlocking, as :60-61 */
    int b200_stream_destroy(int dev, b200_stream_t s); /* syncs first, as the queue dtor :67-76 */
    int b200_stream_sync(b200_stream_t s); /* wait(queue) :166-177 */
    int b200_stream_query(b200_stream_t s, int* is_empty); /* empty(queue) :145-163 */
    typedef void (*b200_host_fn)(void* user);
    int b200_launch_host_func(b200_stream_t s, b200_host_fn fn, void* user); /* :194-230 */

    /* ---------------------------------------------------------------------------------------------
     * Event  (replaces event/EventUniformCudaHipRt.hpp:26-260). timing=0 matches the reference
     * (cudaEventDisableTiming, :50-52); timing=1 is an extension used by the benchmark harness.
     * ------------------------------------------------------------------------------------------- */
    int b200_event_create(int dev, int timing, b200_event_t* out);
    int b200_event_destroy(b200_event_t e);
    int b200_event_record(b200_event_t e, b200_stream_t s); /* enqueue(queue, event) */
    int b200_event_query(b200_event_t e, int* is_complete); /* isComplete(event) */
    int b200_event_sync(b200_event_t e); /* wait(event) */
    int b200_stream_wait_event(b200_stream_t s, b200_event_t e); /* wait(queue, event) */
    int b200_device_wait_event(int dev, b200_event_t e); /* wait(dev, event): all current streams of dev */
    int b200_event_elapsed_ms(b200_event_t start, b200_event_t stop, float* ms);

    /* ---------------------------------------------------------------------------------------------
     * Buffers  (replaces mem/buf/BufUniformCudaHipRt.hpp:205-360). allocBuf and allocAsyncBuf are both
     * served from a per-device stream-ordered pool (cudaMallocAsync, release threshold = keep
     * everything) instead of cudaMalloc/cudaMallocPitch; 2-D rows are padded to B200_ROW_ALIGN bytes
     * because pools have no pitched API and the TMA stencil needs 16-byte-multiple row strides.
     * ------------------------------------------------------------------------------------------- */
#define B200_ROW_ALIGN 128u
    int b200_malloc_async(int dev, b200_stream_t s, size_t bytes, void** out);
    int b200_free_async(int dev, b200_stream_t s, void* ptr);
    int b200_malloc_pitched_async(int dev, b200_stream_t s, size_t width_bytes, size_t height, void** out, size_t* pitch_bytes);
    size_t b200_pitch_for_width(size_t width_bytes);
    /* plain cudaMalloc / cudaFree: memory that must be exportable to other processes (CUDA IPC halos) */
    int b200_malloc_device(int dev, size_t bytes, void** out);
    int b200_free_device(int dev, void* ptr);
    int b200_host_alloc_pinned(size_t bytes, void** out); /* allocMappedBuf :338-360 */
    int b200_host_free_pinned(void* ptr);
    int b200_host_register(void* ptr, size_t bytes);
    int b200_host_unregister(void* ptr);
    int b200_pool_stats(int dev, uint64_t* reserved_bytes, uint64_t* used_bytes);
    int b200_pool_trim(int dev, size_t keep_bytes);

    /* copies / sets  (replaces mem/buf/uniformCudaHip/Copy.hpp:32-495, Set.hpp). `dev` = device whose
     * context issues the copy (the destination device in the reference, Copy.hpp:143). */
    enum
    {
        B200_COPY_H2H = 0,
        B200_COPY_H2D = 1,
        B200_COPY_D2H = 2,
        B200_COPY_D2D = 3,
        B200_COPY_DEFAULT = 4
    };
    int b200_memcpy_async(int dev, void* dst, void const* src, size_t bytes, int kind, b200_stream_t s);
    int b200_memcpy2d_async(int dev, void* dst, size_t dpitch, void const* src, size_t spitch, size_t width_bytes, size_t height, int kind, b200_stream_t s);
    int b200_memcpy_peer_async(void* dst, int dst_dev, void const* src, int src_dev, size_t bytes, b200_stream_t s);
    int b200_memset_async(int dev, void* dst, int byte_value, size_t bytes, b200_stream_t s);
    int b200_memset2d_async(int dev, void* dst, size_t pitch, int byte_value, size_t width_bytes, size_t height, b200_stream_t s);

    /* CUDA IPC (new; the reference is single-process). Handles are 64 opaque bytes. */
    int b200_ipc_get_mem_handle(int dev, void* dev_ptr, unsigned char handle_out[64]);
    int b200_ipc_open_mem_handle(int dev, unsigned char const handle[64], void** out);
    int b200_ipc_close_mem_handle(int dev, void* ptr);
    int b200_ipc_event_create(int dev, b200_event_t* out, unsigned char handle_out[64]);
    int b200_ipc_event_open(int dev, unsigned char const handle[64], b200_event_t* out);

    /* ---------------------------------------------------------------------------------------------
     * Generic kernel launch  (replaces kernel/TaskKernelGpuUniformCudaHipRt.hpp:188-366). `func` is the
     * host address of a __global__ function registered with the SAME (shared) cudart instance.
     * grid/block are CUDA order x,y,z.
     * ------------------------------------------------------------------------------------------- */
    typedef struct b200_func_attributes
    {
        int32_t max_threads_per_block;
        int32_t num_regs;
        uint64_t shared_size_bytes, const_size_bytes, local_size_bytes;
        int32_t max_dynamic_shared_size_bytes;
        int32_t ptx_version, binary_version;
    } b200_func_attributes;
    int b200_func_attributes_get(int dev, void const* func, b200_func_attributes* out);
    /* The device allocation of THIS library (b200_malloc_async / b200_malloc_pitched_async / b200_malloc_device) that
     * contains `ptr`: *base / *bytes, or *base = NULL when the pointer is foreign. New: the generic launch path uses it to
     * prove that a kernel's pointer arguments cannot alias (pairwise distinct allocations) before it picks the
     * restrict-qualified, block-coarsened trampoline (include/alpaka/b200/Kernel.hpp); the reference has no counterpart
     * (kernel/TaskKernelGpuUniformCudaHipRt.hpp:61-76 always runs one element per thread). */
    int b200_mem_range(void const* ptr, void** base, size_t* bytes);
    int b200_launch(int dev, void const* func, uint32_t const grid[3], uint32_t const block[3], size_t dyn_smem_bytes, b200_stream_t s, void** args);
    /* Device address of a __device__ / __constant__ variable registered with the shared cudart instance
     * (replaces ApiCudaRt::getSymbolAddress, core/ApiCudaRt.hpp, as used by the device-global copies in
     * mem/global/DeviceGlobalUniformCudaHipBuiltIn.hpp:43-200). */
    int b200_symbol_address(int dev, void const* symbol, void** out);

    /* Stream memory operations: `s` blocks until the 32-bit word at `addr` (device memory or pinned-mapped host
     * memory, 4-byte aligned) is >= value / writes `value` to it in stream order. Replaces the driver-API
     * cuStreamWaitValue32 call the reference's test helper makes directly
     * (include/alpaka/test/event/EventHostManualTrigger.hpp:407-417, 440-468); the driver entry points are resolved at
     * run time through cudaGetDriverEntryPoint, so libcuda is not a link dependency. B200_ENODEV if unavailable. */
    int b200_stream_wait_value32(int dev, b200_stream_t s, void* addr, uint32_t value);
    int b200_stream_write_value32(int dev, b200_stream_t s, void* addr, uint32_t value);

    /* ---------------------------------------------------------------------------------------------
     * Work-division selection (host logic, no device needed)
     * (replaces workdiv/WorkDivHelpers.hpp:133-309 subDivideGridElems and :406-549 isValidWorkDiv).
     * Vectors are in alpaka order (index 0 = slowest dimension), `dim` entries, 1 <= dim <= 4.
     * ------------------------------------------------------------------------------------------- */
    typedef struct b200_acc_dev_props
    {
        uint64_t multi_processor_count;
        uint64_t grid_block_extent_max[4];
        uint64_t grid_block_count_max;
        uint64_t block_thread_extent_max[4];
        uint64_t block_thread_count_max;
        uint64_t thread_elem_extent_max[4];
        uint64_t thread_elem_count_max;
        uint64_t shared_mem_size_bytes;
        uint64_t global_mem_size_bytes;
    } b200_acc_dev_props;
    enum
    {
        B200_SUBDIV_EQUAL_EXTENT = 0,
        B200_SUBDIV_CLOSE_TO_EQUAL_EXTENT = 1,
        B200_SUBDIV_UNRESTRICTED = 2
    };
    int b200_acc_dev_props_get(int dev, int dim, b200_acc_dev_props* out); /* getAccDevProps<Acc>(dev) */
    int b200_subdivide_grid_elems(
        int dim,
        uint64_t const* grid_elem_extent,
        uint64_t const* thread_elem_extent,
        b200_acc_dev_props const* props,
        uint64_t kernel_block_thread_count_max,
        int block_thread_must_divide_grid_thread_extent,
        int restriction,
        uint64_t* grid_block_extent_out,
        uint64_t* block_thread_extent_out,
        uint64_t* thread_elem_extent_out);
    int b200_is_valid_work_div(
        int dim,
        uint64_t const* grid_block_extent,
        uint64_t const* block_thread_extent,
        uint64_t const* thread_elem_extent,
        b200_acc_dev_props const* props,
        uint64_t kernel_block_thread_count_max, /* 0 = ignore */
        int* is_valid);

    /* ---------------------------------------------------------------------------------------------
     * BabelStream kernels (replace the functors of benchmarks/babelstream/src/babelStreamMainTest.cpp:
     * Init :53-69, Copy :72-86, Mult :89-104, Add :107-122, Triad :125-141; Nstream is new, upstream
     * BabelStream semantics a[i] += b[i] + scalar*c[i]). 128/256-bit vectorised grid-stride streams;
     * FMA contraction is OFF (explicit mul then add) so results are bit-identical to the reference CPU
     * back-end built with -ffp-contract=off. Any n >= 0 and any element-aligned pointers are accepted;
     * the vector path needs 32-byte aligned pointers (every b200_malloc* result is).
     * ------------------------------------------------------------------------------------------- */
    int b200_stream_init_f64(b200_stream_t s, double* a, double* b, double* c, double init_a, uint64_t n);
    int b200_stream_copy_f64(b200_stream_t s, double const* a, double* b, uint64_t n);
    int b200_stream_mul_f64(b200_stream_t s, double const* a, double* b, double scalar, uint64_t n);
    int b200_stream_add_f64(b200_stream_t s, double const* a, double const* b, double* c, uint64_t n);
    int b200_stream_triad_f64(b200_stream_t s, double const* a, double const* b, double* c, double scalar, uint64_t n);
    int b200_stream_nstream_f64(b200_stream_t s, double* a, double const* b, double const* c, double scalar, uint64_t n);
    int b200_stream_init_f32(b200_stream_t s, float* a, float* b, float* c, float init_a, uint64_t n);
    int b200_stream_copy_f32(b200_stream_t s, float const* a, float* b, uint64_t n);
    int b200_stream_mul_f32(b200_stream_t s, float const* a, float* b, float scalar, uint64_t n);
    int b200_stream_add_f32(b200_stream_t s, float const* a, float const* b, float* c, uint64_t n);
    int b200_stream_triad_f32(b200_stream_t s, float const* a, float const* b, float* c, float scalar, uint64_t n);
    int b200_stream_nstream_f32(b200_stream_t s, float* a, float const* b, float const* c, float scalar, uint64_t n);

    /* ---------------------------------------------------------------------------------------------
     * Reductions: Dot (babelStreamMainTest.cpp:145-181 + host finish :399-405) and example/reduce
     * (kernel.hpp:42-132 launched twice, reduce.cpp:79-98), as ONE single-pass kernel each:
     * per-thread vectorised accumulation -> warp shuffle -> block shared memory -> per-block partial in
     * `scratch` -> last block (atomic ticket + __threadfence) folds the partials in a fixed order and
     * writes the scalar to out_dev[0]. Deterministic for a given device (fixed grid, fixed order).
     * `scratch` = B200_REDUCE_SCRATCH_BYTES device bytes, zeroed ONCE by the caller (b200_memset_async);
     * the kernel leaves it ready for the next call. One scratch per concurrently running reduction.
     * ------------------------------------------------------------------------------------------- */
#define B200_REDUCE_SCRATCH_BYTES 65536u
    int b200_dot_f64(b200_stream_t s, double const* a, double const* b, uint64_t n, double* out_dev, void* scratch);
    int b200_dot_f32(b200_stream_t s, float const* a, float const* b, uint64_t n, float* out_dev, void* scratch);
    /* Reference-shaped Dot output: n_partials per-"block" sums whose std::reduce is the dot product, for
     * the unmodified driver (sum[256], babelStreamMainTest.cpp:378-405). */
    int b200_dot_partials_f64(b200_stream_t s, double const* a, double const* b, uint64_t n, double* partials_dev, uint32_t n_partials, void* scratch);
    int b200_dot_partials_f32(b200_stream_t s, float const* a, float const* b, uint64_t n, float* partials_dev, uint32_t n_partials, void* scratch);
    int b200_reduce_sum_u32(b200_stream_t s, uint32_t const* in, uint64_t n, uint32_t* out_dev, void* scratch);
    int b200_reduce_sum_i32(b200_stream_t s, int32_t const* in, uint64_t n, int32_t* out_dev, void* scratch);
    int b200_reduce_sum_u64(b200_stream_t s, uint64_t const* in, uint64_t n, uint64_t* out_dev, void* scratch);
    int b200_reduce_sum_f32(b200_stream_t s, float const* in, uint64_t n, float* out_dev, void* scratch);
    int b200_reduce_sum_f64(b200_stream_t s, double const* in, uint64_t n, double* out_dev, void* scratch);

    /* ---- Dot / reduce over SEVERAL GPUs with the exchange step fused into the launch (new; SURVEY.md section 8e: the
     * slab-sharded Dot and reduce exchange ONE scalar per rank). The last block of the single-pass reduction stores the
     * device's scalar straight into every rank's slot array through peer pointers (CUDA-IPC mappings, one process per
     * GPU, or plain pointers after b200_enable_peer_all inside one process), publishes the call number `step` in their
     * flag words, waits (bounded) for the other ranks' flags and folds the slots left to right in RANK ORDER -- the
     * combination order does not depend on a collective library, so the result is bit-identical on every rank
     * (SURVEY.md section 7.3-10). out_dev[0] holds the all-ranks value when the launch completes. No NCCL, no host step.
     *   base[r]  rank r's exchange buffer (B200_EXCHANGE_BYTES of plain device memory from b200_malloc_device, zeroed
     *            once) as seen from this device; base[rank] is the own one
     *   step     1, 2, 3, ... : the same number on every rank for the same collective call
     * Slots are double-buffered by the parity of `step`, which suffices because a rank cannot finish call k+1 before
     * every rank has entered it, i.e. has finished reading call k's slots. */
#define B200_EXCHANGE_MAX_RANKS 16
#define B200_EXCHANGE_BYTES 512u
    typedef struct b200_exchange
    {
        void* base[B200_EXCHANGE_MAX_RANKS];
        uint32_t world;
        uint32_t rank;
    } b200_exchange;
    int b200_dot_allranks_f64(b200_stream_t s, double const* a, double const* b, uint64_t n, double* out_dev, void* scratch, b200_exchange const* ex, uint32_t step);
    int b200_dot_allranks_f32(b200_stream_t s, float const* a, float const* b, uint64_t n, float* out_dev, void* scratch, b200_exchange const* ex, uint32_t step);
    int b200_reduce_sum_allranks_u32(b200_stream_t s, uint32_t const* in, uint64_t n, uint32_t* out_dev, void* scratch, b200_exchange const* ex, uint32_t step);
    int b200_reduce_sum_allranks_f32(b200_stream_t s, float const* in, uint64_t n, float* out_dev, void* scratch, b200_exchange const* ex, uint32_t step);
    int b200_reduce_sum_allranks_f64(b200_stream_t s, double const* in, uint64_t n, double* out_dev, void* scratch, b200_exchange const* ex, uint32_t step);
    /* 0, or 1 + rank whose flag never arrived in some call. Synchronous. */
    int b200_exchange_status(int dev, void const* own_exchange_buffer, uint32_t* status);

    /* ---------------------------------------------------------------------------------------------
     * heatEquation2D: one fused FTCS step = StencilKernel (StencilKernel.hpp:31-89) + BoundaryKernel
     * (BoundaryKernel.hpp:24-86) in a single persistent TMA-pipelined kernel.
     *
     * Field: (ny+2) x (nx+2) doubles, row-major, row pitch `pitch_bytes` (multiple of 16), base 16-byte
     * aligned; [j][i] with j = y. Core cells 1..ny x 1..nx get
     *   c*(1-2rX-2rY) + l*rX + r*rX + u*rY + d*rY   (left-to-right, no FMA; rX = dt/dx^2, rY = dt/dy^2),
     * ring cells (except the four corners, which are never written, as in the reference) get
     *   time_factor * (sx[i] + sy[j])  ==  exactSolution(i*dx, j*dy, step*dt)  bit-for-bit when the caller
     * fills sx[i] = sin(pi*(i*dx)), sy[j] = sin(pi*(j*dy)), time_factor = exp(-pi*pi*(step*dt)) on the
     * host with the same libm the reference CPU back-end uses.
     *
     * Sub-domain form for the 2-D decomposition: `edges` selects which sides of this tile are physical
     * boundaries (get the analytic value); the other sides are ghost cells owned by a neighbour and are
     * left untouched (filled by the halo exchange). sx/sy are indexed by LOCAL i/j.
     * ------------------------------------------------------------------------------------------- */
    enum
    {
        B200_EDGE_TOP = 1, /* j = 0 row is a physical boundary */
        B200_EDGE_BOTTOM = 2, /* j = ny+1 */
        B200_EDGE_LEFT = 4, /* i = 0 */
        B200_EDGE_RIGHT = 8, /* i = nx+1 */
        B200_EDGE_ALL = 15
    };
    typedef struct b200_heat2d_plan_st* b200_heat2d_plan_t;
    /* A plan owns the TMA descriptors for a pair of ping-pong buffers and the device copies of sx/sy. */
    int b200_heat2d_plan_create(
        int dev,
        double* u0,
        double* u1,
        size_t pitch_bytes,
        uint32_t ny,
        uint32_t nx,
        double const* sx_host, /* nx+2 */
        double const* sy_host, /* ny+2 */
        int edges,
        b200_heat2d_plan_t* out);
    int b200_heat2d_plan_destroy(b200_heat2d_plan_t plan);
    /* One step reading buffer `src_index` (0 = u0, 1 = u1) and writing the other one. */
    int b200_heat2d_step_f64(b200_heat2d_plan_t plan, b200_stream_t s, int src_index, double rx, double ry, double time_factor);
    /* TWO steps in one launch (temporal blocking; new -- the reference launches Stencil + Boundary per step,
     * heatEquation2D.cpp:141-168): reads level s from buffer `src_index`, writes level s+2 into the OTHER buffer; level
     * s+1 exists only in registers, its ring cells take time_factor_1 * (sx + sy), the ring of the result
     * time_factor_2 * (sx + sy). HBM traffic is one read + one write per cell per TWO steps; the arithmetic per cell and
     * level is unchanged, so the field is bit-identical to two b200_heat2d_step_f64 calls. Stand-alone fields only
     * (plan `edges` == B200_EDGE_ALL, no halo): a decomposed tile would need ghost cells two deep. */
    int b200_heat2d_step2_f64(b200_heat2d_plan_t plan, b200_stream_t s, int src_index, double rx, double ry, double time_factor_1, double time_factor_2);
    /* `levels` (3, 4, 6 or 8) steps in one launch: deeper temporal blocking. Threads compute only their own column pair per level
     * and take the horizontal neighbours from adjacent lanes by warp shuffle (no recomputation, no shared-memory round trip);
     * time_factors[l] is the boundary factor of the l-th level of the launch (time_factors[levels-1] the result's).
     * levels = 3: tiles of rows; levels = 4, 6, 8: one warp WALKS down a 64-column window over a tall row segment and keeps
     * every level as partial sums in registers, so no row is loaded or recomputed twice (heatWalkKernel; tunables heat.walk,
     * heat.walk_shape, heat.walk_seg_rows). Same conditions and the same bit-identical result as b200_heat2d_step2_f64. */
    int b200_heat2d_stepn_f64(b200_heat2d_plan_t plan, b200_stream_t s, int src_index, double rx, double ry, int levels, double const* time_factors);
    /* Restrict a step to a row/column window of OUTPUT cells [j0,j1) x [i0,i1) in padded coordinates
     * (used to split interior / edge strips for halo overlap). */
    int b200_heat2d_step_window_f64(b200_heat2d_plan_t plan, b200_stream_t s, int src_index, double rx, double ry, double time_factor, uint32_t j0, uint32_t j1, uint32_t i0, uint32_t i1);

    /* ---- 2-D decomposition with the halo exchange FUSED into the step kernel (new; the reference is single-device,
     * SURVEY.md section 8e). Every rank owns a tile of identical extents (ny+2) x (nx+2) and pitch; a side is either a
     * physical boundary (plan `edges`) or has a neighbour. One launch per step: the edge strips are computed first and
     * their border cells are ALSO stored straight into the neighbour's ghost cells through peer pointers (NVLink P2P
     * stores from the registers holding the fresh values), then the interior; the CTA that finishes the last strip
     * tile publishes the time level in the neighbours' flag words (st.release.sys), and a launch starts by waiting
     * (ld.acquire.sys, bounded) until its own flag words say the ghosts of the previous time level have arrived. No
     * host synchronisation, no staging copies, no NCCL. Pointers may be CUDA-IPC mappings (one process per GPU) or plain
     * peer-enabled device pointers (one process, several GPUs).
     *   peer_u[side][b]  neighbour's ping-pong buffer b (same order as this plan's u0/u1), side = 0 top (j=0), 1 bottom,
     *                    2 left (i=0), 3 right; NULL on a physical boundary
     *   peer_flag[side]  the word in the NEIGHBOUR's flag array that this rank sets (its slot for the opposite side)
     *   my_flags         this rank's 4 words, zero-initialised, written by the neighbours (slot order as `side`)
     * `step` is the 1-based time level the launch produces; ghosts of the initial field count as level 0. */
    typedef struct b200_heat2d_halo
    {
        double* peer_u[4][2];
        uint32_t* peer_flag[4];
        uint32_t* my_flags;
    } b200_heat2d_halo;
    int b200_heat2d_plan_set_halo(b200_heat2d_plan_t plan, b200_heat2d_halo const* halo);
    int b200_heat2d_step_halo_f64(b200_heat2d_plan_t plan, b200_stream_t s, int src_index, double rx, double ry, double time_factor, uint32_t step);
    /* ---- Row-SLAB decomposition with G = 2 .. 8 time levels per launch and per exchange (new). Temporal blocking
     * needs ghost cells G deep; with row slabs (every rank keeps the full width) only rows are exchanged, they are
     * contiguous, and no diagonal neighbour exists. Array layout of a slab: (ny+2G) x (nx+2) doubles -- rows 0..G-1 and
     * ny+G..ny+2G-1 are ghost rows on a side with a neighbour; on a physical side row G-1 / ny+G is the ring and the rows
     * beyond it are unused; core rows are G..ny+G-1; columns carry the reference's one-cell ring. sy_host has ny+2G
     * entries (sy[j] for local row j, ghost rows included: their ring-column cells are boundary cells of the
     * neighbour's rows). `edges` must contain LEFT and RIGHT; ny >= 2G; all slabs of a field have the same ny, nx, pitch.
     * b200_heat2d_plan_set_halo wires the neighbours (sides 0 = top, 1 = bottom; left/right stay NULL). One launch of
     * b200_heat2d_step2_halo_f64 (2 levels) / b200_heat2d_stepn_halo_f64 (levels = 3, 4, 6 or 8), levels <= G, reads level s
     * (ghost rows included) from buffer `src_index`, writes level s+levels of the core rows and physical ring into the
     * other buffer, stores its first / last G core rows straight into the neighbours' ghost rows (peer stores from
     * registers; always all G, so launches of different depths can follow each other, e.g. 1000 = 332 x 3 + 2 x 2) and
     * publishes `step` (1-based launch index) in their flag words; strip tiles come first and wait for the
     * neighbours' flags of launch step-1, as in the one-level form. */
    int b200_heat2d_slab_plan_create(
        int dev,
        double* u0,
        double* u1,
        size_t pitch_bytes,
        uint32_t ny,
        uint32_t nx,
        double const* sx_host, /* nx+2 */
        double const* sy_host, /* ny+2*ghost_rows */
        int edges,
        uint32_t ghost_rows, /* G */
        b200_heat2d_plan_t* out);
    int b200_heat2d_step2_halo_f64(b200_heat2d_plan_t plan, b200_stream_t s, int src_index, double rx, double ry, double time_factor_1, double time_factor_2, uint32_t step);
    int b200_heat2d_stepn_halo_f64(b200_heat2d_plan_t plan, b200_stream_t s, int src_index, double rx, double ry, int levels, double const* time_factors, uint32_t step);
    /* ---- 2-D TILES with G = 4 .. 8 time levels per launch (new; closes the gap between the 2-D decomposition of
     * b200_heat2d_step_halo_f64, one level per launch, and the row slabs above). Ghost cells G deep on all four sides:
     * array (ny+2G) x (nx+2G) doubles, core cells [G, ny+G) x [G, nx+G), on a physical side the row / column next to them is
     * the ring and the cells beyond are unused; sx_host has nx+2G entries, sy_host ny+2G (local indices, ghosts included);
     * ny, nx >= 2G; all tiles of a field have the same extents and pitch. b200_heat2d_plan_set_halo wires up to four
     * neighbours. One b200_heat2d_stepn_tile_f64 (levels = 4, 6 or 8 <= G; the walker kernel) is TWO launches: the step
     * itself, whose strips store the first / last G core rows into the vertical neighbours' ghost rows and publish `step` in
     * their row flags exactly as the slab form does; then a small column kernel that waits for the vertical neighbours' row
     * flags of this step and stores the first / last G core columns -- over every row holding level data, the ghost rows just
     * received included, so the corner blocks reach the diagonal neighbours through the vertical ones -- into the horizontal
     * neighbours' ghost columns and publishes `step` in their column flags. The next step's walkers that touch ghost columns
     * wait for those. `step` is the 1-based launch index. */
    int b200_heat2d_tile_plan_create(
        int dev,
        double* u0,
        double* u1,
        size_t pitch_bytes,
        uint32_t ny,
        uint32_t nx,
        double const* sx_host, /* nx+2*ghost */
        double const* sy_host, /* ny+2*ghost */
        int edges,
        uint32_t ghost, /* G */
        b200_heat2d_plan_t* out);
    int b200_heat2d_stepn_tile_f64(b200_heat2d_plan_t plan, b200_stream_t s, int src_index, double rx, double ry, int levels, double const* time_factors, uint32_t step);
    /* The host-side decomposition of a walker launch (b200_heat2d_stepn_f64 / _halo_ / _tile_ at 4, 6, 8 levels) for a field of
     * ny x nx core cells with ghost / ring depths pad_y, pad_x and `resident_walkers` walker slots on the device: column
     * windows of `window_columns` stored columns (window w stores array columns [w * window_columns, ...)), of which window 0
     * and the last n_edge_right are edge windows; per window n_front front segments (slab strips, bit k of front_is_strip;
     * physical bands when the interior is split off) with output rows [front_y0[k], front_y1[k]) and the interior rows
     * [interior_y0, interior_y1) cut into segments of segment_rows. No device needed: the CPU tests check that these segments
     * cover every output row exactly once for any geometry. */
    typedef struct b200_heat2d_walk_plan
    {
        uint32_t window_columns, n_windows, n_edge_right;
        uint32_t n_front, front_is_strip;
        int32_t front_y0[2], front_y1[2];
        int32_t interior_y0, interior_y1, segment_rows;
        uint32_t n_segments, n_walkers;
        int32_t split;
    } b200_heat2d_walk_plan;
    int b200_heat2d_walk_plan_query(uint32_t ny, uint32_t nx, uint32_t pad_y, uint32_t pad_x, int edges, int levels, int resident_walkers, b200_heat2d_walk_plan* out);
    /* 0, or 1 + side of the first flag wait that timed out (a neighbour stopped making progress). Synchronous. */
    int b200_heat2d_halo_status(b200_heat2d_plan_t plan, uint32_t* status);

    /* The ring alone, as its own small launch: BoundaryKernel (BoundaryKernel.hpp:24-86) for callers that keep the
     * reference's two-launch step (the alpaka API layer running the unmodified driver). Writes the analytic value to
     * the ring of buffer `dst_index` on the sides flagged in the plan's `edges`. */
    int b200_heat2d_boundary_f64(b200_heat2d_plan_t plan, b200_stream_t s, int dst_index, double time_factor);

    /* ---------------------------------------------------------------------------------------------
     * Tuning / introspection (bench + profiling only)
     * ------------------------------------------------------------------------------------------- */
    int b200_tune_set(char const* key, int64_t value);
    int b200_tune_get(char const* key, int64_t* value);
    /* number of kernel launches issued by this library since load (bench.py's gpu_launches) */
    uint64_t b200_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif
