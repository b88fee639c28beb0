// alpaka_b200/csrc/b200_rt.cu -- runtime half of the C ABI (include/b200/b200.h): devices, streams, events,
// stream-ordered pool allocation, copies, IPC, generic launch, tuning registry, error policy.
//
// Replaces the reference's vendor shim `alpaka::ApiCudaRt` (include/alpaka/core/ApiCudaRt.hpp:107-396) and the
// cudart call sequences inside its trait specialisations (cited per function in b200.h).
#include "b200_common.cuh"

#include <cstdlib>
#include <map>
#include <mutex>
#include <vector>

namespace b200
{
    std::string& lastError()
    {
        thread_local std::string s;
        return s;
    }

    int fail(int code, char const* what, char const* file, int line)
    {
        lastError() = std::string(file) + "(" + std::to_string(line) + ") '" + what + "' failed : '"
                      + b200_error_name(code) + "'!";
        return code;
    }

    int cudaFail(cudaError_t e, char const* cmd, char const* file, int line)
    {
        // same shape as the reference's message (core/UniformCudaHip.hpp:62-82)
        lastError() = std::string(file) + "(" + std::to_string(line) + ") '" + cmd + "' returned error : '"
                      + cudaGetErrorName(e) + "': '" + cudaGetErrorString(e) + "'!";
        (void) cudaGetLastError(); // clear the non-sticky error like the reference does
        return static_cast<int>(e);
    }

    std::atomic<uint64_t> g_launchCount{0};

    namespace
    {
        std::mutex g_tuneMutex;
        // Tunables: defaults live at the call sites; overrides come from b200_tune_set() or, once at load, from the
        // environment: B200_TUNE="heat.stages=2,stream.block=256".
        std::map<std::string, int64_t>& tuneMap()
        {
            static std::map<std::string, int64_t> m = []
            {
                std::map<std::string, int64_t> init;
                if(char const* env = std::getenv("B200_TUNE"))
                {
                    std::string s(env);
                    size_t pos = 0;
                    while(pos < s.size())
                    {
                        size_t const end = s.find(',', pos);
                        std::string const item = s.substr(pos, end == std::string::npos ? std::string::npos : end - pos);
                        size_t const eq = item.find('=');
                        if(eq != std::string::npos)
                            init[item.substr(0, eq)] = std::strtoll(item.c_str() + eq + 1, nullptr, 10);
                        if(end == std::string::npos)
                            break;
                        pos = end + 1;
                    }
                }
                return init;
            }();
            return m;
        }

        std::mutex g_devMutex;
        std::vector<int> g_smCount;
        std::vector<char> g_poolReady;
    } // namespace

    int64_t tune(char const* key, int64_t dflt)
    {
        std::lock_guard<std::mutex> l(g_tuneMutex);
        auto const it = tuneMap().find(key);
        return it == tuneMap().end() ? dflt : it->second;
    }

    // The kernel entries of the ABI take a stream, not a device (b200_stream_*, b200_dot_*, b200_reduce_*): make the
    // stream's device current before the launch, so that one process can drive several devices (the reference sets the
    // device in front of every call, e.g. kernel/TaskKernelGpuUniformCudaHipRt.hpp:265). The NULL stream keeps the
    // current device.
    int useDeviceOf(cudaStream_t s)
    {
        if(s == nullptr)
            return 0;
        int dev = -1;
        cudaError_t e = cudaStreamGetDevice(s, &dev);
        if(e == cudaSuccess && dev != currentDevice())
            e = cudaSetDevice(dev);
        if(e != cudaSuccess)
            return cudaFail(e, "cudaStreamGetDevice / cudaSetDevice", __FILE__, __LINE__);
        return 0;
    }

    int currentDevice()
    {
        int d = 0;
        if(cudaGetDevice(&d) != cudaSuccess)
        {
            (void) cudaGetLastError();
            return 0;
        }
        return d;
    }

    int smCount(int dev)
    {
        std::lock_guard<std::mutex> l(g_devMutex);
        if(dev >= int(g_smCount.size()))
            g_smCount.resize(size_t(dev) + 1, 0);
        if(g_smCount[size_t(dev)] == 0)
        {
            int v = 0;
            if(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0)
            {
                (void) cudaGetLastError();
                v = 148; // B200
            }
            g_smCount[size_t(dev)] = v;
        }
        return g_smCount[size_t(dev)];
    }

    namespace
    {
        // One-time pool configuration per device: keep freed memory in the pool (release threshold = max) so
        // steady-state allocBuf/free never reaches the OS allocator.
        int ensurePool(int dev)
        {
            std::lock_guard<std::mutex> l(g_devMutex);
            if(dev >= int(g_poolReady.size()))
                g_poolReady.resize(size_t(dev) + 1, 0);
            if(!g_poolReady[size_t(dev)])
            {
                cudaMemPool_t pool;
                B200_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
                uint64_t threshold = UINT64_MAX;
                B200_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold));
                g_poolReady[size_t(dev)] = 1;
            }
            return 0;
        }

        cudaStream_t cs(b200_stream_t s)
        {
            return reinterpret_cast<cudaStream_t>(s);
        }

        cudaEvent_t ce(b200_event_t e)
        {
            return reinterpret_cast<cudaEvent_t>(e);
        }

        cudaMemcpyKind kindOf(int k)
        {
            switch(k)
            {
            case B200_COPY_H2H:
                return cudaMemcpyHostToHost;
            case B200_COPY_H2D:
                return cudaMemcpyHostToDevice;
            case B200_COPY_D2H:
                return cudaMemcpyDeviceToHost;
            case B200_COPY_D2D:
                return cudaMemcpyDeviceToDevice;
            default:
                return cudaMemcpyDefault;
            }
        }
    } // namespace
} // namespace b200

using namespace b200;

namespace
{
    // ---- registry of the device allocations this library handed out: base -> size. The generic launch path asks it whether
    // the pointer arguments of a kernel lie in pairwise distinct allocations (then no in-bounds access through one can alias
    // an access through another, and the kernel may be compiled with restrict-qualified parameters; Kernel.hpp).
    std::mutex g_allocMutex;
    std::map<uintptr_t, size_t>& allocMap()
    {
        static std::map<uintptr_t, size_t> m;
        return m;
    }

    void registerAlloc(void* p, size_t bytes)
    {
        if(p == nullptr)
            return;
        std::lock_guard<std::mutex> l(g_allocMutex);
        allocMap()[reinterpret_cast<uintptr_t>(p)] = bytes;
    }

    void unregisterAlloc(void* p)
    {
        std::lock_guard<std::mutex> l(g_allocMutex);
        allocMap().erase(reinterpret_cast<uintptr_t>(p));
    }
} // namespace

extern "C"
{
    int b200_abi_version(void)
    {
        return B200_ABI_VERSION;
    }

    char const* b200_last_error_string(void)
    {
        return lastError().c_str();
    }

    char const* b200_error_name(int code)
    {
        switch(code)
        {
        case 0:
            return "success";
        case B200_EINVAL:
            return "B200_EINVAL";
        case B200_EALIGN:
            return "B200_EALIGN";
        case B200_ENODEV:
            return "B200_ENODEV";
        case B200_ERANGE:
            return "B200_ERANGE";
        default:
            return code > 0 ? cudaGetErrorName(static_cast<cudaError_t>(code)) : "B200_EUNKNOWN";
        }
    }

    // ------------------------------------------------------------------ platform / device
    int b200_device_count(int* count)
    {
        B200_REQUIRE(count, B200_EINVAL);
        *count = 0;
        cudaError_t const e = cudaGetDeviceCount(count);
        if(e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver)
        {
            // platform/PlatformUniformCudaHipRt.hpp:56-66: "no device" is a count of zero, not an error
            (void) cudaGetLastError();
            *count = 0;
            return 0;
        }
        B200_CUDA(e);
        return 0;
    }

    int b200_device_props_get(int dev, b200_device_props* out)
    {
        B200_REQUIRE(out, B200_EINVAL);
        std::memset(out, 0, sizeof(*out));
        cudaDeviceProp p;
        B200_CUDA(cudaGetDeviceProperties(&p, dev));
        std::strncpy(out->name, p.name, sizeof(out->name) - 1);
        out->cc_major = p.major;
        out->cc_minor = p.minor;
        out->multi_processor_count = p.multiProcessorCount;
        for(int i = 0; i < 3; ++i)
        {
            out->max_grid_dim[i] = p.maxGridSize[i];
            out->max_block_dim[i] = p.maxThreadsDim[i];
        }
        out->max_threads_per_block = p.maxThreadsPerBlock;
        out->warp_size = p.warpSize;
        out->shared_mem_per_block = p.sharedMemPerBlock;
        out->shared_mem_per_block_optin = p.sharedMemPerBlockOptin;
        out->total_global_mem = p.totalGlobalMem;
        out->l2_cache_bytes = p.l2CacheSize;
        out->memory_pools_supported = p.memoryPoolsSupported;
        uint64_t freeB = 0, totalB = 0;
        int const rc = b200_device_mem_info(dev, &freeB, &totalB);
        if(rc != 0)
            return rc;
        out->free_global_mem = freeB;
        return 0;
    }

    int b200_device_mem_info(int dev, uint64_t* free_bytes, uint64_t* total_bytes)
    {
        B200_CUDA(cudaSetDevice(dev));
        size_t f = 0, t = 0;
        B200_CUDA(cudaMemGetInfo(&f, &t));
        if(free_bytes)
            *free_bytes = f;
        if(total_bytes)
            *total_bytes = t;
        return 0;
    }

    int b200_device_sync(int dev)
    {
        B200_CUDA(cudaSetDevice(dev));
        B200_CUDA(cudaDeviceSynchronize());
        return 0;
    }

    int b200_device_reset(int dev)
    {
        B200_CUDA(cudaSetDevice(dev));
        B200_CUDA(cudaDeviceReset());
        std::lock_guard<std::mutex> l(g_devMutex);
        if(dev < int(g_poolReady.size()))
            g_poolReady[size_t(dev)] = 0;
        return 0;
    }

    int b200_enable_peer_all(int* n_pairs_enabled)
    {
        int n = 0;
        int rc = b200_device_count(&n);
        if(rc != 0)
            return rc;
        int enabled = 0;
        for(int i = 0; i < n; ++i)
        {
            B200_CUDA(cudaSetDevice(i));
            for(int j = 0; j < n; ++j)
            {
                if(i == j)
                    continue;
                int can = 0;
                B200_CUDA(cudaDeviceCanAccessPeer(&can, i, j));
                if(!can)
                    continue;
                cudaError_t const e = cudaDeviceEnablePeerAccess(j, 0);
                if(e == cudaErrorPeerAccessAlreadyEnabled)
                    (void) cudaGetLastError();
                else
                    B200_CUDA(e);
                // cudaDeviceEnablePeerAccess covers cudaMalloc memory only: buffers come from the stream-ordered pool
                // (b200_malloc_async), whose access is granted per pool. Device i may now read and write device j's pool.
                cudaMemPool_t pool;
                B200_CUDA(cudaDeviceGetDefaultMemPool(&pool, j));
                cudaMemAccessDesc desc{};
                desc.location.type = cudaMemLocationTypeDevice;
                desc.location.id = i;
                desc.flags = cudaMemAccessFlagsProtReadWrite;
                B200_CUDA(cudaMemPoolSetAccess(pool, &desc, 1));
                ++enabled;
            }
        }
        if(n_pairs_enabled)
            *n_pairs_enabled = enabled;
        return 0;
    }

    int b200_acc_dev_props_get(int dev, int dim, b200_acc_dev_props* out)
    {
        // acc/AccGpuUniformCudaHipRt.hpp:113-187: nine cudaDeviceGetAttribute reads (faster than the full
        // cudaGetDeviceProperties) + cudaMemGetInfo; x is the LAST component in alpaka's vector order.
        B200_REQUIRE(out && dim >= 1 && dim <= 4, B200_EINVAL);
        std::memset(out, 0, sizeof(*out));
        int mp = 0, grid[3] = {}, block[3] = {}, threads = 0, smem = 0;
        B200_CUDA(cudaDeviceGetAttribute(&mp, cudaDevAttrMultiProcessorCount, dev));
        B200_CUDA(cudaDeviceGetAttribute(&grid[0], cudaDevAttrMaxGridDimX, dev));
        B200_CUDA(cudaDeviceGetAttribute(&grid[1], cudaDevAttrMaxGridDimY, dev));
        B200_CUDA(cudaDeviceGetAttribute(&grid[2], cudaDevAttrMaxGridDimZ, dev));
        B200_CUDA(cudaDeviceGetAttribute(&block[0], cudaDevAttrMaxBlockDimX, dev));
        B200_CUDA(cudaDeviceGetAttribute(&block[1], cudaDevAttrMaxBlockDimY, dev));
        B200_CUDA(cudaDeviceGetAttribute(&block[2], cudaDevAttrMaxBlockDimZ, dev));
        B200_CUDA(cudaDeviceGetAttribute(&threads, cudaDevAttrMaxThreadsPerBlock, dev));
        B200_CUDA(cudaDeviceGetAttribute(&smem, cudaDevAttrMaxSharedMemoryPerBlock, dev));
        uint64_t freeB = 0, totalB = 0;
        int const rc = b200_device_mem_info(dev, &freeB, &totalB);
        if(rc != 0)
            return rc;
        out->multi_processor_count = uint64_t(mp);
        for(int i = 0; i < dim; ++i)
        {
            int const cudaAxis = dim - 1 - i; // component i (slow..fast) <- CUDA axis (x = fastest)
            out->grid_block_extent_max[i] = cudaAxis < 3 ? uint64_t(grid[cudaAxis]) : 1u;
            out->block_thread_extent_max[i] = cudaAxis < 3 ? uint64_t(block[cudaAxis]) : 1u;
            out->thread_elem_extent_max[i] = UINT64_MAX;
        }
        out->grid_block_count_max = UINT64_MAX;
        out->block_thread_count_max = uint64_t(threads);
        out->thread_elem_count_max = UINT64_MAX;
        out->shared_mem_size_bytes = uint64_t(smem);
        out->global_mem_size_bytes = totalB;
        return 0;
    }

    // ------------------------------------------------------------------ streams
    int b200_stream_create(int dev, b200_stream_t* out)
    {
        B200_REQUIRE(out, B200_EINVAL);
        B200_CUDA(cudaSetDevice(dev));
        cudaStream_t s;
        B200_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
        *out = reinterpret_cast<b200_stream_t>(s);
        return 0;
    }

    int b200_stream_destroy(int dev, b200_stream_t s)
    {
        B200_CUDA(cudaSetDevice(dev));
        B200_CUDA(cudaStreamSynchronize(cs(s)));
        B200_CUDA(cudaStreamDestroy(cs(s)));
        return 0;
    }

    int b200_stream_sync(b200_stream_t s)
    {
        B200_CUDA(cudaStreamSynchronize(cs(s)));
        return 0;
    }

    int b200_stream_query(b200_stream_t s, int* is_empty)
    {
        B200_REQUIRE(is_empty, B200_EINVAL);
        cudaError_t const e = cudaStreamQuery(cs(s));
        if(e == cudaErrorNotReady)
        {
            (void) cudaGetLastError();
            *is_empty = 0;
            return 0;
        }
        B200_CUDA(e);
        *is_empty = 1;
        return 0;
    }

    int b200_launch_host_func(b200_stream_t s, b200_host_fn fn, void* user)
    {
        B200_REQUIRE(fn, B200_EINVAL);
        B200_CUDA(cudaLaunchHostFunc(cs(s), reinterpret_cast<cudaHostFn_t>(fn), user));
        return 0;
    }

    // ------------------------------------------------------------------ events
    int b200_event_create(int dev, int timing, b200_event_t* out)
    {
        B200_REQUIRE(out, B200_EINVAL);
        B200_CUDA(cudaSetDevice(dev));
        cudaEvent_t e;
        B200_CUDA(cudaEventCreateWithFlags(&e, timing ? cudaEventDefault : cudaEventDisableTiming));
        *out = reinterpret_cast<b200_event_t>(e);
        return 0;
    }

    int b200_event_destroy(b200_event_t e)
    {
        B200_CUDA(cudaEventDestroy(ce(e)));
        return 0;
    }

    int b200_event_record(b200_event_t e, b200_stream_t s)
    {
        B200_CUDA(cudaEventRecord(ce(e), cs(s)));
        return 0;
    }

    int b200_event_query(b200_event_t e, int* is_complete)
    {
        B200_REQUIRE(is_complete, B200_EINVAL);
        cudaError_t const r = cudaEventQuery(ce(e));
        if(r == cudaErrorNotReady)
        {
            (void) cudaGetLastError();
            *is_complete = 0;
            return 0;
        }
        B200_CUDA(r);
        *is_complete = 1;
        return 0;
    }

    int b200_event_sync(b200_event_t e)
    {
        B200_CUDA(cudaEventSynchronize(ce(e)));
        return 0;
    }

    int b200_stream_wait_event(b200_stream_t s, b200_event_t e)
    {
        B200_CUDA(cudaStreamWaitEvent(cs(s), ce(e), 0));
        return 0;
    }

    int b200_device_wait_event(int dev, b200_event_t e)
    {
        // event/EventUniformCudaHipRt.hpp:228-249: wait(dev, event) makes the legacy default stream wait,
        // which orders all blocking streams of the device behind the event.
        B200_CUDA(cudaSetDevice(dev));
        B200_CUDA(cudaStreamWaitEvent(nullptr, ce(e), 0));
        return 0;
    }

    int b200_event_elapsed_ms(b200_event_t start, b200_event_t stop, float* ms)
    {
        B200_REQUIRE(ms, B200_EINVAL);
        B200_CUDA(cudaEventElapsedTime(ms, ce(start), ce(stop)));
        return 0;
    }

    // ------------------------------------------------------------------ memory
    size_t b200_pitch_for_width(size_t width_bytes)
    {
        return (width_bytes + (B200_ROW_ALIGN - 1)) / B200_ROW_ALIGN * B200_ROW_ALIGN;
    }

    int b200_malloc_async(int dev, b200_stream_t s, size_t bytes, void** out)
    {
        B200_REQUIRE(out, B200_EINVAL);
        *out = nullptr;
        B200_CUDA(cudaSetDevice(dev));
        int const rc = ensurePool(dev);
        if(rc != 0)
            return rc;
        if(bytes == 0)
            return 0; // zero-sized buffers are legal (test/unit/mem/buf BufTest "zero-size")
        B200_CUDA(cudaMallocAsync(out, bytes, cs(s)));
        registerAlloc(*out, bytes);
        return 0;
    }

    int b200_free_async(int dev, b200_stream_t s, void* ptr)
    {
        if(!ptr)
            return 0;
        B200_CUDA(cudaSetDevice(dev));
        unregisterAlloc(ptr);
        B200_CUDA(cudaFreeAsync(ptr, cs(s)));
        return 0;
    }

    int b200_malloc_pitched_async(int dev, b200_stream_t s, size_t width_bytes, size_t height, void** out, size_t* pitch_bytes)
    {
        B200_REQUIRE(out && pitch_bytes, B200_EINVAL);
        size_t const pitch = b200_pitch_for_width(width_bytes);
        *pitch_bytes = pitch;
        return b200_malloc_async(dev, s, pitch * height, out);
    }

    int b200_malloc_device(int dev, size_t bytes, void** out)
    {
        B200_REQUIRE(out, B200_EINVAL);
        *out = nullptr;
        B200_CUDA(cudaSetDevice(dev));
        if(bytes == 0)
            return 0;
        B200_CUDA(cudaMalloc(out, bytes));
        registerAlloc(*out, bytes);
        return 0;
    }

    int b200_free_device(int dev, void* ptr)
    {
        if(!ptr)
            return 0;
        B200_CUDA(cudaSetDevice(dev));
        unregisterAlloc(ptr);
        B200_CUDA(cudaFree(ptr));
        return 0;
    }

    int b200_mem_range(void const* ptr, void** base, size_t* bytes)
    {
        B200_REQUIRE(base && bytes, B200_EINVAL);
        *base = nullptr;
        *bytes = 0;
        if(ptr == nullptr)
            return 0;
        auto const a = reinterpret_cast<uintptr_t>(ptr);
        std::lock_guard<std::mutex> l(g_allocMutex);
        auto& m = allocMap();
        auto it = m.upper_bound(a);
        if(it == m.begin())
            return 0;
        --it;
        if(a < it->first + it->second)
        {
            *base = reinterpret_cast<void*>(it->first);
            *bytes = it->second;
        }
        return 0;
    }

    int b200_host_alloc_pinned(size_t bytes, void** out)
    {
        B200_REQUIRE(out, B200_EINVAL);
        *out = nullptr;
        if(bytes == 0)
            return 0;
        B200_CUDA(cudaHostAlloc(out, bytes, cudaHostAllocPortable | cudaHostAllocMapped));
        return 0;
    }

    int b200_host_free_pinned(void* ptr)
    {
        if(!ptr)
            return 0;
        B200_CUDA(cudaFreeHost(ptr));
        return 0;
    }

    int b200_host_register(void* ptr, size_t bytes)
    {
        B200_REQUIRE(ptr && bytes, B200_EINVAL);
        B200_CUDA(cudaHostRegister(ptr, bytes, cudaHostRegisterPortable));
        return 0;
    }

    int b200_host_unregister(void* ptr)
    {
        B200_REQUIRE(ptr, B200_EINVAL);
        B200_CUDA(cudaHostUnregister(ptr));
        return 0;
    }

    int b200_pool_stats(int dev, uint64_t* reserved_bytes, uint64_t* used_bytes)
    {
        cudaMemPool_t pool;
        B200_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
        uint64_t r = 0, u = 0;
        B200_CUDA(cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &r));
        B200_CUDA(cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &u));
        if(reserved_bytes)
            *reserved_bytes = r;
        if(used_bytes)
            *used_bytes = u;
        return 0;
    }

    int b200_pool_trim(int dev, size_t keep_bytes)
    {
        cudaMemPool_t pool;
        B200_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
        B200_CUDA(cudaMemPoolTrimTo(pool, keep_bytes));
        return 0;
    }

    int b200_memcpy_async(int dev, void* dst, void const* src, size_t bytes, int kind, b200_stream_t s)
    {
        if(bytes == 0)
            return 0;
        B200_REQUIRE(dst && src, B200_EINVAL);
        B200_CUDA(cudaSetDevice(dev));
        B200_CUDA(cudaMemcpyAsync(dst, src, bytes, kindOf(kind), cs(s)));
        return 0;
    }

    int b200_memcpy2d_async(int dev, void* dst, size_t dpitch, void const* src, size_t spitch, size_t width_bytes, size_t height, int kind, b200_stream_t s)
    {
        if(width_bytes == 0 || height == 0)
            return 0;
        B200_REQUIRE(dst && src, B200_EINVAL);
        B200_CUDA(cudaSetDevice(dev));
        B200_CUDA(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width_bytes, height, kindOf(kind), cs(s)));
        return 0;
    }

    int b200_memcpy_peer_async(void* dst, int dst_dev, void const* src, int src_dev, size_t bytes, b200_stream_t s)
    {
        if(bytes == 0)
            return 0;
        B200_REQUIRE(dst && src, B200_EINVAL);
        B200_CUDA(cudaMemcpyPeerAsync(dst, dst_dev, src, src_dev, bytes, cs(s)));
        return 0;
    }

    int b200_memset_async(int dev, void* dst, int byte_value, size_t bytes, b200_stream_t s)
    {
        if(bytes == 0)
            return 0;
        B200_REQUIRE(dst, B200_EINVAL);
        B200_CUDA(cudaSetDevice(dev));
        B200_CUDA(cudaMemsetAsync(dst, byte_value, bytes, cs(s)));
        return 0;
    }

    int b200_memset2d_async(int dev, void* dst, size_t pitch, int byte_value, size_t width_bytes, size_t height, b200_stream_t s)
    {
        if(width_bytes == 0 || height == 0)
            return 0;
        B200_REQUIRE(dst, B200_EINVAL);
        B200_CUDA(cudaSetDevice(dev));
        B200_CUDA(cudaMemset2DAsync(dst, pitch, byte_value, width_bytes, height, cs(s)));
        return 0;
    }

    // ------------------------------------------------------------------ IPC
    static_assert(sizeof(cudaIpcMemHandle_t) == 64 && sizeof(cudaIpcEventHandle_t) == 64, "handle size");

    int b200_ipc_get_mem_handle(int dev, void* dev_ptr, unsigned char handle_out[64])
    {
        B200_REQUIRE(dev_ptr && handle_out, B200_EINVAL);
        B200_CUDA(cudaSetDevice(dev));
        cudaIpcMemHandle_t h;
        B200_CUDA(cudaIpcGetMemHandle(&h, dev_ptr));
        std::memcpy(handle_out, &h, 64);
        return 0;
    }

    int b200_ipc_open_mem_handle(int dev, unsigned char const handle[64], void** out)
    {
        B200_REQUIRE(handle && out, B200_EINVAL);
        B200_CUDA(cudaSetDevice(dev));
        cudaIpcMemHandle_t h;
        std::memcpy(&h, handle, 64);
        B200_CUDA(cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess));
        return 0;
    }

    int b200_ipc_close_mem_handle(int dev, void* ptr)
    {
        B200_REQUIRE(ptr, B200_EINVAL);
        B200_CUDA(cudaSetDevice(dev));
        B200_CUDA(cudaIpcCloseMemHandle(ptr));
        return 0;
    }

    int b200_ipc_event_create(int dev, b200_event_t* out, unsigned char handle_out[64])
    {
        B200_REQUIRE(out && handle_out, B200_EINVAL);
        B200_CUDA(cudaSetDevice(dev));
        cudaEvent_t e;
        B200_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming | cudaEventInterprocess));
        cudaIpcEventHandle_t h;
        B200_CUDA(cudaIpcGetEventHandle(&h, e));
        std::memcpy(handle_out, &h, 64);
        *out = reinterpret_cast<b200_event_t>(e);
        return 0;
    }

    int b200_ipc_event_open(int dev, unsigned char const handle[64], b200_event_t* out)
    {
        B200_REQUIRE(handle && out, B200_EINVAL);
        B200_CUDA(cudaSetDevice(dev));
        cudaIpcEventHandle_t h;
        std::memcpy(&h, handle, 64);
        cudaEvent_t e;
        B200_CUDA(cudaIpcOpenEventHandle(&e, h));
        *out = reinterpret_cast<b200_event_t>(e);
        return 0;
    }

    // ------------------------------------------------------------------ generic launch
    int b200_func_attributes_get(int dev, void const* func, b200_func_attributes* out)
    {
        B200_REQUIRE(func && out, B200_EINVAL);
        B200_CUDA(cudaSetDevice(dev));
        cudaFuncAttributes a;
        B200_CUDA(cudaFuncGetAttributes(&a, func));
        out->max_threads_per_block = a.maxThreadsPerBlock;
        out->num_regs = a.numRegs;
        out->shared_size_bytes = a.sharedSizeBytes;
        out->const_size_bytes = a.constSizeBytes;
        out->local_size_bytes = a.localSizeBytes;
        out->max_dynamic_shared_size_bytes = a.maxDynamicSharedSizeBytes;
        out->ptx_version = a.ptxVersion;
        out->binary_version = a.binaryVersion;
        return 0;
    }

    int b200_launch(int dev, void const* func, uint32_t const grid[3], uint32_t const block[3], size_t dyn_smem_bytes, b200_stream_t s, void** args)
    {
        B200_REQUIRE(func && grid && block, B200_EINVAL);
        B200_CUDA(cudaSetDevice(dev));
        if(dyn_smem_bytes > 48u * 1024u)
        {
            // opt in to the large carve-out (the reference API cannot express this)
            B200_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, int(dyn_smem_bytes)));
        }
        B200_CUDA(cudaLaunchKernel(func, dim3(grid[0], grid[1], grid[2]), dim3(block[0], block[1], block[2]), args, dyn_smem_bytes, cs(s)));
        countLaunch();
        return 0;
    }

    int b200_symbol_address(int dev, void const* symbol, void** out)
    {
        B200_REQUIRE(symbol && out, B200_EINVAL);
        B200_CUDA(cudaSetDevice(dev));
        B200_CUDA(cudaGetSymbolAddress(out, symbol));
        return 0;
    }

    // ------------------------------------------------------------------ stream memory operations
    namespace
    {
        // CUresult (*)(CUstream, CUdeviceptr, cuuint32_t, unsigned flags); 0 == CUDA_SUCCESS
        using StreamValue32Fn = int (*)(cudaStream_t, unsigned long long, uint32_t, unsigned);

        StreamValue32Fn driverEntry(char const* name)
        {
            void* fn = nullptr;
            cudaDriverEntryPointQueryResult status = cudaDriverEntryPointSymbolNotFound;
            if(cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &status) != cudaSuccess
               || status != cudaDriverEntryPointSuccess)
            {
                (void) cudaGetLastError();
                return nullptr;
            }
            return reinterpret_cast<StreamValue32Fn>(fn);
        }
    } // namespace

    int b200_stream_wait_value32(int dev, b200_stream_t s, void* addr, uint32_t value)
    {
        B200_REQUIRE(addr, B200_EINVAL);
        B200_REQUIRE((reinterpret_cast<uintptr_t>(addr) & 3u) == 0, B200_EALIGN);
        B200_CUDA(cudaSetDevice(dev));
        static StreamValue32Fn const fn = driverEntry("cuStreamWaitValue32");
        B200_REQUIRE(fn, B200_ENODEV);
        constexpr unsigned waitGeq = 0x0; // CU_STREAM_WAIT_VALUE_GEQ
        int const rc = fn(cs(s), static_cast<unsigned long long>(reinterpret_cast<uintptr_t>(addr)), value, waitGeq);
        B200_REQUIRE(rc == 0, B200_ENODEV);
        return 0;
    }

    int b200_stream_write_value32(int dev, b200_stream_t s, void* addr, uint32_t value)
    {
        B200_REQUIRE(addr, B200_EINVAL);
        B200_REQUIRE((reinterpret_cast<uintptr_t>(addr) & 3u) == 0, B200_EALIGN);
        B200_CUDA(cudaSetDevice(dev));
        static StreamValue32Fn const fn = driverEntry("cuStreamWriteValue32");
        B200_REQUIRE(fn, B200_ENODEV);
        int const rc = fn(cs(s), static_cast<unsigned long long>(reinterpret_cast<uintptr_t>(addr)), value, 0u);
        B200_REQUIRE(rc == 0, B200_ENODEV);
        return 0;
    }

    // ------------------------------------------------------------------ tuning / introspection
    int b200_tune_set(char const* key, int64_t value)
    {
        B200_REQUIRE(key, B200_EINVAL);
        std::lock_guard<std::mutex> l(g_tuneMutex);
        tuneMap()[key] = value;
        return 0;
    }

    int b200_tune_get(char const* key, int64_t* value)
    {
        B200_REQUIRE(key && value, B200_EINVAL);
        std::lock_guard<std::mutex> l(g_tuneMutex);
        auto const it = tuneMap().find(key);
        if(it == tuneMap().end())
            return B200_EINVAL;
        *value = it->second;
        return 0;
    }

    uint64_t b200_launch_count(void)
    {
        return g_launchCount.load();
    }
}
