// mm_core.cu -- error reporting, device queries, host-side GLL tables.
#include <cstdarg>
#include <cstdio>
#include <mutex>

#include "mm_common.cuh"

static thread_local char g_err[512] = "";

void mm_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

int mm_cuda_fail(cudaError_t e, const char *what, const char *file, int line)
{
    mm_set_error("CUDA error %d (%s) in `%s` at %s:%d", (int)e, cudaGetErrorString(e), what, file,
                 line);
    return MM_ERR_CUDA;
}

// ---- library-owned stream-ordered memory pool, one per device ------------------------------------
// Index builds allocate and free a few buffers per call; a pool that keeps freed blocks avoids paying
// cudaMalloc/cudaFree every time.  The pool is PRIVATE to this library: the device's default pool --
// shared with the host process (e.g. torch) -- is never reconfigured.  mm_pool_trim() hands the cached
// blocks back to the driver.
namespace {
constexpr int MM_MAX_DEVICES = 64;
std::mutex g_pool_mutex;
cudaMemPool_t g_pools[MM_MAX_DEVICES] = {};
}  // namespace

cudaError_t mm_pool_alloc(void **p, size_t bytes, cudaStream_t st)
{
    int dev = 0;
    cudaError_t rc = cudaGetDevice(&dev);
    if (rc != cudaSuccess) return rc;
    if (dev < 0 || dev >= MM_MAX_DEVICES) return cudaErrorInvalidDevice;
    cudaMemPool_t pool;
    {
        std::lock_guard<std::mutex> lock(g_pool_mutex);
        if (!g_pools[dev]) {
            cudaMemPoolProps props = {};
            props.allocType = cudaMemAllocationTypePinned;
            props.handleTypes = cudaMemHandleTypeNone;
            props.location.type = cudaMemLocationTypeDevice;
            props.location.id = dev;
            rc = cudaMemPoolCreate(&g_pools[dev], &props);
            if (rc != cudaSuccess) return rc;
            uint64_t keep = UINT64_MAX;  // freed blocks stay in OUR pool until mm_pool_trim()
            cudaMemPoolSetAttribute(g_pools[dev], cudaMemPoolAttrReleaseThreshold, &keep);
        }
        pool = g_pools[dev];
    }
    return cudaMallocFromPoolAsync(p, bytes ? bytes : 16, pool, st);
}

void mm_pool_free(void *p, cudaStream_t st)
{
    if (p) cudaFreeAsync(p, st);
}

extern "C" int mm_pool_trim(void)
{
    int dev = 0;
    MM_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_pool_mutex);
    if (dev >= 0 && dev < MM_MAX_DEVICES && g_pools[dev]) {
        MM_CUDA(cudaDeviceSynchronize());
        MM_CUDA(cudaMemPoolTrimTo(g_pools[dev], 0));
    }
    return MM_OK;
}

int mm_num_sms()
{
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    return n;
}

// Host-side table construction.  The host compiler is told -ffp-contract=off, so every operation
// here is one IEEE rounding: c_i = 1 / (((z_i - z_a) * (z_i - z_b)) * ...), j ascending.
int mm_make_table(int order, mm_gll_table *t)
{
    int m = 0;
    for (int i = 0; i < MM_MAXM; ++i) t->z[i] = t->c[i] = 0.0;
    switch (order) {
    case 1: t->z[0] = -1.0; t->z[1] = 1.0; m = 2; break;
    case 2: t->z[0] = -1.0; t->z[1] = 0.0; t->z[2] = 1.0; m = 3; break;
    case 4:
        t->z[0] = -1.0; t->z[1] = -0x1.4f2ec413cb52ap-1; t->z[2] = 0.0;
        t->z[3] = 0x1.4f2ec413cb52ap-1; t->z[4] = 1.0;  // sqrt(3/7)
        m = 5;
        break;
    default: return 0;
    }
    for (int i = 0; i < m; ++i) {
        volatile double prod = 1.0;
        for (int j = 0; j < m; ++j)
            if (j != i) prod = prod * (t->z[i] - t->z[j]);
        t->c[i] = 1.0 / prod;
    }
    return m;
}

extern "C" int mm_version(void) { return MM_VERSION; }
extern "C" const char *mm_last_error(void) { return g_err; }
