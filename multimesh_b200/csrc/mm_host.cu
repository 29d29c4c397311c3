// mm_host.cu -- end-to-end entry point over HOST buffers (what the e2e benchmark times and what a
// ctypes/cffi caller without torch would use): H2D of the source mesh and the target points,
// K0 geometry, K1 index build, the fused K1->K2->K3 pipeline, D2H of the results.
//
// Device buffers and two streams live in a per-thread pool that is re-used between calls
// (mm_host_release frees it).  The copy stream moves the target points and the field blocks
// while the main stream computes the geometry and builds the index from the nodes.
#include "mm_common.cuh"

namespace {

struct buf_t {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        cudaError_t rc = cudaMalloc(&p, bytes);
        if (rc == cudaSuccess) cap = bytes;
        return rc;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T> T *as() { return static_cast<T *>(p); }
};

struct host_pool {
    int device = -1;
    cudaStream_t main = nullptr, copy = nullptr;
    cudaEvent_t copied = nullptr, fields_copied = nullptr;
    buf_t nodes, fields, pts, cent, aabb, pre, elem, xi, out, nf, ws;
    void release()
    {
        for (buf_t *b : {&nodes, &fields, &pts, &cent, &aabb, &pre, &elem, &xi, &out, &nf, &ws}) b->release();
        if (copied) cudaEventDestroy(copied);
        if (fields_copied) cudaEventDestroy(fields_copied);
        if (main) cudaStreamDestroy(main);
        if (copy) cudaStreamDestroy(copy);
        copied = fields_copied = nullptr;
        main = copy = nullptr;
        device = -1;
    }
};

thread_local host_pool g_pool;

struct index_holder {
    mm_index_t *ix = nullptr;
    ~index_holder() { if (ix) mm_index_destroy(ix); }
};

}  // namespace

#define MM_TRY(call)                  \
    do {                              \
        int _rc = (call);             \
        if (_rc != MM_OK) return _rc; \
    } while (0)

extern "C" int mm_host_release(void)
{
    g_pool.release();
    return MM_OK;
}

extern "C" int mm_interpolate_host(int order, int dim, int64_t E, const double *nodes, int F,
                                   const double *fields, int64_t N, const double *pts, int k,
                                   int gll_points_form, const mm_locate_params *params,
                                   double *values, int32_t *elem, double *xi, int64_t *num_failed)
{
    MM_REQUIRE(mm_valid_order(order), MM_ERR_INVALID, "mm_interpolate_host: order %d", order);
    MM_REQUIRE(dim == 2 || dim == 3, MM_ERR_INVALID, "mm_interpolate_host: dim %d", dim);
    MM_REQUIRE(E > 0 && F >= 1 && N >= 0 && k >= 1 && k <= 64, MM_ERR_INVALID,
               "mm_interpolate_host: sizes");
    MM_REQUIRE(nodes && fields && params && values && (N == 0 || pts), MM_ERR_INVALID,
               "mm_interpolate_host: null buffer");
    if (num_failed) *num_failed = 0;
    if (N == 0) return MM_OK;
    const int P = mm_pow(order + 1, dim);
    host_pool &pl = g_pool;
    int dev = 0;
    MM_CUDA(cudaGetDevice(&dev));
    if (pl.device != dev) {
        pl.release();
        MM_CUDA(cudaStreamCreateWithFlags(&pl.main, cudaStreamNonBlocking));
        MM_CUDA(cudaStreamCreateWithFlags(&pl.copy, cudaStreamNonBlocking));
        MM_CUDA(cudaEventCreateWithFlags(&pl.copied, cudaEventDisableTiming));
        MM_CUDA(cudaEventCreateWithFlags(&pl.fields_copied, cudaEventDisableTiming));
        pl.device = dev;
    }
    MM_CUDA(pl.nodes.ensure(sizeof(double) * E * P * dim));
    MM_CUDA(pl.fields.ensure(sizeof(double) * E * F * P));
    MM_CUDA(pl.pts.ensure(sizeof(double) * N * dim));
    MM_CUDA(pl.cent.ensure(sizeof(double) * E * dim));
    MM_CUDA(pl.aabb.ensure(sizeof(double) * E * 2 * dim));
    MM_CUDA(pl.pre.ensure(sizeof(double) * E * (2 * dim + dim * dim)));
    MM_CUDA(pl.out.ensure(sizeof(double) * N * F));
    MM_CUDA(pl.nf.ensure(sizeof(int64_t)));
    if (elem) MM_CUDA(pl.elem.ensure(sizeof(int32_t) * N));
    if (xi) MM_CUDA(pl.xi.ensure(sizeof(double) * N * dim));

    // main stream: nodes -> geometry -> index;   copy stream: target points, field blocks
    MM_CUDA(cudaMemcpyAsync(pl.nodes.p, nodes, sizeof(double) * E * P * dim, cudaMemcpyHostToDevice, pl.main));
    MM_CUDA(cudaMemcpyAsync(pl.pts.p, pts, sizeof(double) * N * dim, cudaMemcpyHostToDevice, pl.copy));
    MM_CUDA(cudaEventRecord(pl.copied, pl.copy));  // target points on the device
    MM_CUDA(cudaMemcpyAsync(pl.fields.p, fields, sizeof(double) * E * F * P, cudaMemcpyHostToDevice, pl.copy));
    MM_CUDA(cudaEventRecord(pl.fields_copied, pl.copy));  // only K3 needs the fields
    MM_TRY(mm_element_geometry(order, dim, E, pl.nodes.as<double>(), pl.cent.as<double>(),
                               pl.aabb.as<double>(), pl.main));
    MM_TRY(mm_element_presolve(order, dim, E, pl.nodes.as<double>(), pl.pre.as<double>(), pl.main));
    index_holder ih;
    if (gll_points_form)
        MM_TRY(mm_index_create(&ih.ix, dim, E * P, pl.nodes.as<double>(), pl.main));
    else
        MM_TRY(mm_index_create(&ih.ix, dim, E, pl.cent.as<double>(), pl.main));
    const size_t ws_bytes = mm_interpolate_workspace_bytes(ih.ix, dim, N, k);
    MM_CUDA(pl.ws.ensure(ws_bytes));
    MM_CUDA(cudaStreamWaitEvent(pl.main, pl.copied, 0));
    MM_TRY(mm_interpolate_impl(ih.ix, gll_points_form ? P : 1, order, dim, E, pl.nodes.as<double>(),
                          pl.cent.as<double>(), pl.aabb.as<double>(), pl.pre.as<double>(), F,
                          pl.fields.as<double>(), N,
                          pl.pts.as<double>(), k, params, pl.out.as<double>(),
                          elem ? pl.elem.as<int32_t>() : nullptr, xi ? pl.xi.as<double>() : nullptr,
                          nullptr, pl.nf.as<int64_t>(), pl.ws.p, pl.ws.cap, pl.main, pl.fields_copied));
    MM_CUDA(cudaMemcpyAsync(values, pl.out.p, sizeof(double) * N * F, cudaMemcpyDeviceToHost, pl.main));
    if (elem) MM_CUDA(cudaMemcpyAsync(elem, pl.elem.p, sizeof(int32_t) * N, cudaMemcpyDeviceToHost, pl.main));
    if (xi) MM_CUDA(cudaMemcpyAsync(xi, pl.xi.p, sizeof(double) * N * dim, cudaMemcpyDeviceToHost, pl.main));
    int64_t nf = 0;
    MM_CUDA(cudaMemcpyAsync(&nf, pl.nf.p, sizeof(int64_t), cudaMemcpyDeviceToHost, pl.main));
    MM_CUDA(cudaStreamSynchronize(pl.main));
    if (num_failed) *num_failed = nf;
    return MM_OK;
}
