// mm_host.cu -- end-to-end entry point over HOST buffers (what the e2e benchmark times and what a
// ctypes/cffi caller without torch would use): H2D of the source mesh and the target points,
// K0 geometry, K1 index build + k-NN, K2 locate, K3 gather, D2H of the results.
#include "mm_common.cuh"

namespace {
struct dev_buf {
    void *p = nullptr;
    ~dev_buf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 16); }
    template <typename T> T *as() { return static_cast<T *>(p); }
};
struct index_holder {
    mm_index_t *ix = nullptr;
    ~index_holder() { if (ix) mm_index_destroy(ix); }
};
}  // namespace

#define MM_TRY(call)            \
    do {                        \
        int _rc = (call);       \
        if (_rc != MM_OK) return _rc; \
    } while (0)

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
    cudaStream_t stream = nullptr;
    dev_buf d_nodes, d_fields, d_pts, d_cent, d_aabb, d_elem, d_xi, d_out, d_nf, d_ws;
    MM_CUDA(d_nodes.alloc(sizeof(double) * E * P * dim));
    MM_CUDA(d_fields.alloc(sizeof(double) * E * F * P));
    MM_CUDA(d_pts.alloc(sizeof(double) * N * dim));
    MM_CUDA(d_cent.alloc(sizeof(double) * E * dim));
    MM_CUDA(d_aabb.alloc(sizeof(double) * E * 2 * dim));
    MM_CUDA(d_out.alloc(sizeof(double) * N * F));
    MM_CUDA(d_nf.alloc(sizeof(int64_t)));
    if (elem) MM_CUDA(d_elem.alloc(sizeof(int32_t) * N));
    if (xi) MM_CUDA(d_xi.alloc(sizeof(double) * N * dim));
    MM_CUDA(cudaMemcpyAsync(d_nodes.p, nodes, sizeof(double) * E * P * dim, cudaMemcpyHostToDevice, stream));
    MM_CUDA(cudaMemcpyAsync(d_pts.p, pts, sizeof(double) * N * dim, cudaMemcpyHostToDevice, stream));
    MM_TRY(mm_element_geometry(order, dim, E, d_nodes.as<double>(), d_cent.as<double>(),
                               d_aabb.as<double>(), stream));
    index_holder ih;
    if (gll_points_form)
        MM_TRY(mm_index_create(&ih.ix, dim, E * P, d_nodes.as<double>(), stream));
    else
        MM_TRY(mm_index_create(&ih.ix, dim, E, d_cent.as<double>(), stream));
    MM_CUDA(cudaMemcpyAsync(d_fields.p, fields, sizeof(double) * E * F * P, cudaMemcpyHostToDevice, stream));
    const size_t ws_bytes = mm_interpolate_workspace_bytes(ih.ix, dim, N, k);
    MM_CUDA(d_ws.alloc(ws_bytes));
    MM_TRY(mm_interpolate(ih.ix, gll_points_form ? P : 1, order, dim, E, d_nodes.as<double>(),
                          d_cent.as<double>(), d_aabb.as<double>(), F, d_fields.as<double>(), N,
                          d_pts.as<double>(), k, params, d_out.as<double>(),
                          elem ? d_elem.as<int32_t>() : nullptr, xi ? d_xi.as<double>() : nullptr,
                          nullptr, d_nf.as<int64_t>(), d_ws.p, ws_bytes, stream));
    MM_CUDA(cudaMemcpyAsync(values, d_out.p, sizeof(double) * N * F, cudaMemcpyDeviceToHost, stream));
    if (elem) MM_CUDA(cudaMemcpyAsync(elem, d_elem.p, sizeof(int32_t) * N, cudaMemcpyDeviceToHost, stream));
    if (xi) MM_CUDA(cudaMemcpyAsync(xi, d_xi.p, sizeof(double) * N * dim, cudaMemcpyDeviceToHost, stream));
    int64_t nf = 0;
    MM_CUDA(cudaMemcpyAsync(&nf, d_nf.p, sizeof(int64_t), cudaMemcpyDeviceToHost, stream));
    MM_CUDA(cudaStreamSynchronize(stream));
    if (num_failed) *num_failed = nf;
    return MM_OK;
}
