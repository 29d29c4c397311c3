// mm_source.cu -- a source mesh RESIDENT on the device, and the chunked host-buffer pipeline.
//
// The reference reloads the source model, rebuilds its KD-tree and walks every point again for every
// call (components/interpolator.py:660-760).  In the workflows it is used in (sum of gradients, model
// updates on a fixed mesh pair) the source mesh is the same for many calls, so the handle below keeps
//   nodes [E][P][d], fields [E][F][P], centroid, AABB, affine pre-solve, the spatial index (+ site table)
// in HBM; a call then moves only the target points in and the values out.
//
// mm_source_interpolate_host cuts the target points into chunks and runs a three-stream pipeline
//   h2d stream : points of chunk i+1           (PCIe, host -> device)
//   main stream: mm_interpolate on chunk i     (K1 -> K2 -> K3, no host synchronisation inside)
//   d2h stream : values of chunk i-1           (PCIe, device -> host; full duplex with the h2d stream)
// with two device buffers per direction, so both copy engines and the SMs are busy at the same time.
// Results do not depend on the chunking: every point's result is a pure function of (point, source).
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "mm_common.cuh"

struct mm_source {
    int device = -1;
    int order = 0, dim = 0, P = 0, F = 0;
    int64_t E = 0;
    int gll_form = 0;
    bool owns_mesh = false;           // nodes / fields were uploaded by us (host constructor)
    double *nodes = nullptr, *fields = nullptr;
    double *cent = nullptr, *aabb = nullptr, *pre = nullptr;
    mm_index_t *index = nullptr;
    cudaStream_t main = nullptr, h2d = nullptr, d2h = nullptr;
    cudaEvent_t fields_ready = nullptr, ev_alloc = nullptr;
    bool fields_pending = false;      // an asynchronous field upload has been enqueued on h2d
    // per-call staging, grown on demand and re-used
    struct buf_t {
        void *p = nullptr;
        size_t cap = 0;
    };
    buf_t pts[2], out[2], elem[2], xi[2], ws, nf;
    cudaEvent_t ev_in[2] = {}, ev_comp[2] = {}, ev_out[2] = {};
    size_t resident_bytes = 0;
};

namespace {

#define MM_TRY(call)                  \
    do {                              \
        int _rc = (call);             \
        if (_rc != MM_OK) return _rc; \
    } while (0)

// All device memory of a source comes from the library's stream-ordered pool (mm_pool_alloc), allocated and
// freed on the handle's main stream: a one-shot call (create -> interpolate -> destroy) then re-uses the blocks
// of the previous call instead of paying cudaMalloc / cudaFree for gigabytes every time.  The other two streams
// are ordered behind the allocations with an event (sync_allocations).
cudaError_t ensure(mm_source *s, mm_source::buf_t &b, size_t bytes)
{
    if (bytes <= b.cap) return cudaSuccess;
    if (b.p) mm_pool_free(b.p, s->main);
    b.p = nullptr;
    b.cap = 0;
    cudaError_t rc = mm_pool_alloc(&b.p, bytes, s->main);
    if (rc == cudaSuccess) b.cap = bytes;
    return rc;
}

void release(mm_source *s, mm_source::buf_t &b)
{
    if (b.p) mm_pool_free(b.p, s->main);
    b.p = nullptr;
    b.cap = 0;
}

cudaError_t sync_allocations(mm_source *s)
{
    cudaError_t rc = cudaEventRecord(s->ev_alloc, s->main);
    if (rc == cudaSuccess) rc = cudaStreamWaitEvent(s->h2d, s->ev_alloc, 0);
    if (rc == cudaSuccess) rc = cudaStreamWaitEvent(s->d2h, s->ev_alloc, 0);
    return rc;
}

int64_t host_chunk_points(int64_t N)
{
    int64_t c = (int64_t)1 << 20;  // 1 M points: 24 MB in, 40 MB out (F = 5) per chunk (S2: 20.0 ms; 2 M: 20.4, 4 M: 20.9)
    if (const char *e = getenv("MM_HOST_CHUNK")) {
        long long v = atoll(e);
        if (v > 0) c = v;
    }
    return std::max<int64_t>(1, std::min<int64_t>(c, N));
}

int init_common(mm_source *s, int order, int dim, int64_t E, int F, int gll_form)
{
    MM_REQUIRE(mm_valid_order(order), MM_ERR_INVALID, "mm_source: order %d (supported 1, 2, 4)", order);
    MM_REQUIRE(dim == 2 || dim == 3, MM_ERR_INVALID, "mm_source: dim %d", dim);
    MM_REQUIRE(E > 0 && E <= (int64_t)INT32_MAX, MM_ERR_INVALID, "mm_source: E=%lld", (long long)E);
    MM_REQUIRE(F >= 0, MM_ERR_INVALID, "mm_source: F=%d", F);
    s->order = order;
    s->dim = dim;
    s->E = E;
    s->F = F;
    s->P = mm_pow(order + 1, dim);
    s->gll_form = gll_form ? 1 : 0;
    MM_REQUIRE(!gll_form || E * (int64_t)s->P <= (int64_t)INT32_MAX, MM_ERR_INVALID,
               "mm_source: E*P=%lld GLL points do not fit the int32 ids of the index", (long long)(E * s->P));
    MM_CUDA(cudaGetDevice(&s->device));
    MM_CUDA(cudaStreamCreateWithFlags(&s->main, cudaStreamNonBlocking));
    MM_CUDA(cudaStreamCreateWithFlags(&s->h2d, cudaStreamNonBlocking));
    MM_CUDA(cudaStreamCreateWithFlags(&s->d2h, cudaStreamNonBlocking));
    MM_CUDA(cudaEventCreateWithFlags(&s->fields_ready, cudaEventDisableTiming));
    MM_CUDA(cudaEventCreateWithFlags(&s->ev_alloc, cudaEventDisableTiming));
    for (int i = 0; i < 2; ++i) {
        MM_CUDA(cudaEventCreateWithFlags(&s->ev_in[i], cudaEventDisableTiming));
        MM_CUDA(cudaEventCreateWithFlags(&s->ev_comp[i], cudaEventDisableTiming));
        MM_CUDA(cudaEventCreateWithFlags(&s->ev_out[i], cudaEventDisableTiming));
    }
    return MM_OK;
}

// geometry + pre-solve + index (+ site table) from nodes that are (or will be, in stream order) on the device
int build_derived(mm_source *s, cudaStream_t st)
{
    const int d = s->dim;
    MM_CUDA(mm_pool_alloc((void **)&s->cent, sizeof(double) * s->E * d, st));
    MM_CUDA(mm_pool_alloc((void **)&s->aabb, sizeof(double) * s->E * 2 * d, st));
    MM_CUDA(mm_pool_alloc((void **)&s->pre, sizeof(double) * s->E * (2 * d + d * d), st));
    s->resident_bytes += sizeof(double) * s->E * (size_t)(d + 2 * d + 2 * d + d * d);
    MM_TRY(mm_element_geometry(s->order, d, s->E, s->nodes, s->cent, s->aabb, st));
    MM_TRY(mm_element_presolve(s->order, d, s->E, s->nodes, s->pre, st));
    if (s->gll_form) {
        MM_TRY(mm_index_create(&s->index, d, s->E * s->P, s->nodes, st));
        MM_TRY(mm_index_prepare_sites(s->index, st));
    } else {
        MM_TRY(mm_index_create(&s->index, d, s->E, s->cent, st));
    }
    int64_t info[8];
    MM_TRY(mm_index_info(s->index, info, nullptr));
    s->resident_bytes += (size_t)info[6];
    return MM_OK;
}

}  // namespace

extern "C" int mm_source_destroy(mm_source_t *s)
{
    if (!s) return MM_OK;
    int prev = -1;
    cudaGetDevice(&prev);
    if (s->device >= 0) cudaSetDevice(s->device);
    cudaDeviceSynchronize();  // callers of mm_source_interpolate may still have work in flight on their own streams
    if (s->index) mm_index_destroy(s->index);
    if (s->owns_mesh) {
        mm_pool_free(s->nodes, s->main);
        mm_pool_free(s->fields, s->main);
    }
    for (double *p : {s->cent, s->aabb, s->pre}) mm_pool_free(p, s->main);
    for (int i = 0; i < 2; ++i) {
        release(s, s->pts[i]);
        release(s, s->out[i]);
        release(s, s->elem[i]);
        release(s, s->xi[i]);
        if (s->ev_in[i]) cudaEventDestroy(s->ev_in[i]);
        if (s->ev_comp[i]) cudaEventDestroy(s->ev_comp[i]);
        if (s->ev_out[i]) cudaEventDestroy(s->ev_out[i]);
    }
    release(s, s->ws);
    release(s, s->nf);
    if (s->main) cudaStreamSynchronize(s->main);  // the frees above are stream-ordered
    if (s->fields_ready) cudaEventDestroy(s->fields_ready);
    if (s->ev_alloc) cudaEventDestroy(s->ev_alloc);
    if (s->main) cudaStreamDestroy(s->main);
    if (s->h2d) cudaStreamDestroy(s->h2d);
    if (s->d2h) cudaStreamDestroy(s->d2h);
    delete s;
    if (prev >= 0) cudaSetDevice(prev);
    return MM_OK;
}

// internal: host constructor; with defer_fields the field upload is left in flight on the h2d stream
// (the one-shot entry point overlaps it with the index build and the first chunks' K1/K2)
int mm_source_create_host_impl(mm_source_t **out, int order, int dim, int64_t E, const double *nodes, int F,
                               const double *fields, int gll_points_form, bool defer_fields)
{
    MM_REQUIRE(out, MM_ERR_INVALID, "mm_source_create_host: null out");
    *out = nullptr;
    MM_REQUIRE(nodes, MM_ERR_INVALID, "mm_source_create_host: null nodes");
    MM_REQUIRE(F == 0 || fields, MM_ERR_INVALID, "mm_source_create_host: F=%d but fields is null", F);
    mm_source *s = new mm_source();
    struct guard_t {
        mm_source *p;
        ~guard_t() { if (p) mm_source_destroy(p); }
    } guard{s};
    MM_TRY(init_common(s, order, dim, E, F, gll_points_form));
    s->owns_mesh = true;
    const size_t nb = sizeof(double) * E * s->P * dim, fb = sizeof(double) * E * (size_t)F * s->P;
    MM_CUDA(mm_pool_alloc((void **)&s->nodes, nb, s->main));
    if (F) MM_CUDA(mm_pool_alloc((void **)&s->fields, fb, s->main));
    MM_CUDA(sync_allocations(s));
    s->resident_bytes = nb + fb;
    // nodes first (the index build needs them), then the fields behind them on the same copy stream
    MM_CUDA(cudaMemcpyAsync(s->nodes, nodes, nb, cudaMemcpyHostToDevice, s->h2d));
    MM_CUDA(cudaEventRecord(s->ev_in[0], s->h2d));
    if (F) {
        MM_CUDA(cudaMemcpyAsync(s->fields, fields, fb, cudaMemcpyHostToDevice, s->h2d));
        MM_CUDA(cudaEventRecord(s->fields_ready, s->h2d));
        s->fields_pending = true;
    }
    MM_CUDA(cudaStreamWaitEvent(s->main, s->ev_in[0], 0));
    MM_TRY(build_derived(s, s->main));  // synchronises s->main
    if (!defer_fields) {
        MM_CUDA(cudaStreamSynchronize(s->h2d));  // the caller may free its host buffers now
        s->fields_pending = false;
    }
    guard.p = nullptr;
    *out = s;
    return MM_OK;
}

extern "C" int mm_source_create_host(mm_source_t **out, int order, int dim, int64_t E, const double *nodes,
                                     int F, const double *fields, int gll_points_form)
{
    return mm_source_create_host_impl(out, order, dim, E, nodes, F, fields, gll_points_form, false);
}

extern "C" int mm_source_create_device(mm_source_t **out, int order, int dim, int64_t E, const double *nodes,
                                       int F, const double *fields, int gll_points_form, void *stream_)
{
    MM_REQUIRE(out, MM_ERR_INVALID, "mm_source_create_device: null out");
    *out = nullptr;
    MM_REQUIRE(nodes, MM_ERR_INVALID, "mm_source_create_device: null nodes");
    MM_REQUIRE(((uintptr_t)nodes & 15) == 0 && ((uintptr_t)fields & 15) == 0, MM_ERR_INVALID,
               "mm_source_create_device: nodes / fields must be 16-byte aligned");
    mm_source *s = new mm_source();
    struct guard_t {
        mm_source *p;
        ~guard_t() { if (p) mm_source_destroy(p); }
    } guard{s};
    MM_TRY(init_common(s, order, dim, E, fields ? F : 0, gll_points_form));
    s->owns_mesh = false;  // borrowed: the caller keeps nodes / fields alive
    s->nodes = const_cast<double *>(nodes);
    s->fields = const_cast<double *>(fields);
    // what the caller enqueued on `stream` (e.g. an NCCL broadcast of the mesh) must be visible to our streams
    cudaStream_t st = (cudaStream_t)stream_;
    MM_CUDA(cudaEventRecord(s->ev_in[0], st));
    MM_CUDA(cudaStreamWaitEvent(s->main, s->ev_in[0], 0));
    MM_TRY(build_derived(s, s->main));
    guard.p = nullptr;
    *out = s;
    return MM_OK;
}

extern "C" int mm_source_set_fields_host(mm_source_t *s, int F, const double *fields)
{
    MM_REQUIRE(s && fields && F >= 1, MM_ERR_INVALID, "mm_source_set_fields_host: arguments");
    MM_REQUIRE(s->owns_mesh, MM_ERR_INVALID,
               "mm_source_set_fields_host: the source borrows device arrays; update them directly");
    MM_CUDA(cudaSetDevice(s->device));
    const size_t fb = sizeof(double) * s->E * (size_t)F * s->P;
    MM_CUDA(cudaStreamSynchronize(s->main));  // no kernel may still be reading the old fields
    if (F != s->F) {
        mm_pool_free(s->fields, s->main);
        s->fields = nullptr;
        s->resident_bytes -= sizeof(double) * s->E * (size_t)s->F * s->P;
        s->F = 0;
        MM_CUDA(mm_pool_alloc((void **)&s->fields, fb, s->main));
        MM_CUDA(sync_allocations(s));
        s->F = F;
        s->resident_bytes += fb;
    }
    MM_CUDA(cudaMemcpyAsync(s->fields, fields, fb, cudaMemcpyHostToDevice, s->h2d));
    MM_CUDA(cudaStreamSynchronize(s->h2d));
    s->fields_pending = false;
    return MM_OK;
}

extern "C" int mm_source_info(const mm_source_t *s, int64_t info[8])
{
    MM_REQUIRE(s && info, MM_ERR_INVALID, "mm_source_info: null");
    info[0] = s->E;
    info[1] = s->P;
    info[2] = s->dim;
    info[3] = s->F;
    info[4] = s->order;
    info[5] = s->gll_form;
    info[6] = (int64_t)s->resident_bytes;
    info[7] = s->device;
    return MM_OK;
}

extern "C" const mm_index_t *mm_source_index(const mm_source_t *s) { return s ? s->index : nullptr; }

// device-pointer form: one mm_interpolate on the caller's stream, workspace owned by the handle
extern "C" int mm_source_interpolate(mm_source_t *s, int64_t N, const double *pts, int k,
                                     const mm_locate_params *params, double *out, int32_t *elem, double *xi,
                                     uint8_t *status, int64_t *num_failed, void *stream)
{
    MM_REQUIRE(s && params, MM_ERR_INVALID, "mm_source_interpolate: null source/params");
    MM_REQUIRE(N >= 0 && N <= (int64_t)INT32_MAX, MM_ERR_INVALID, "mm_source_interpolate: N=%lld outside [0, 2^31)",
               (long long)N);
    if (N == 0) {
        if (num_failed) MM_CUDA(cudaMemsetAsync(num_failed, 0, sizeof(int64_t), (cudaStream_t)stream));
        return MM_OK;
    }
    const size_t need = mm_interpolate_workspace_bytes(s->index, s->dim, N, k);
    if (need > s->ws.cap) {
        // the previous workspace may still be in use by work enqueued earlier on any stream
        MM_CUDA(cudaDeviceSynchronize());
        MM_CUDA(ensure(s, s->ws, need));
        MM_CUDA(cudaStreamSynchronize(s->main));  // allocated on main, used on the caller's stream
    }
    void *fr = nullptr;
    if (s->fields_pending) fr = s->fields_ready;
    return mm_interpolate_impl(s->index, s->gll_form ? s->P : 1, s->order, s->dim, s->E, s->nodes, s->cent, s->aabb,
                               s->pre, s->F, out ? s->fields : nullptr, N, pts, k, params, out, elem, xi, status,
                               num_failed, s->ws.p, s->ws.cap, stream, fr);
}

// host-pointer form: chunked three-stream pipeline (see the file header)
extern "C" int mm_source_interpolate_host(mm_source_t *s, int64_t N, const double *pts, int k,
                                          const mm_locate_params *params, double *values, int32_t *elem,
                                          double *xi, int64_t *num_failed)
{
    MM_REQUIRE(s && params, MM_ERR_INVALID, "mm_source_interpolate_host: null source/params");
    MM_REQUIRE(N >= 0, MM_ERR_INVALID, "mm_source_interpolate_host: N");
    MM_REQUIRE(k >= 1 && k <= 64, MM_ERR_INVALID, "mm_source_interpolate_host: k=%d outside [1, 64]", k);
    MM_REQUIRE(!values || s->F >= 1, MM_ERR_INVALID, "mm_source_interpolate_host: the source has no fields");
    if (num_failed) *num_failed = 0;
    if (N == 0) return MM_OK;
    MM_REQUIRE(pts && (values || elem || xi), MM_ERR_INVALID, "mm_source_interpolate_host: null buffer");
    int prev = -1;
    MM_CUDA(cudaGetDevice(&prev));
    MM_CUDA(cudaSetDevice(s->device));
    struct restore_t {
        int dev;
        ~restore_t() { cudaSetDevice(dev); }
    } restore{prev};

    const int d = s->dim, F = s->F;
    const int64_t C = host_chunk_points(N);  // any N: each chunk is one mm_interpolate of <= C < 2^31 points
    // chunk schedule: the d2h stream is the bottleneck of the pipeline (F values out per point against d coordinates
    // in) and idles until the first chunk is through h2d + K1-K3, so the first chunks are small (C/8, C/4, C/2) and
    // the copy back starts after ~0.4 ms instead of ~1.7 ms
    std::vector<int64_t> chunk_at;
    {
        const char *e = getenv("MM_HOST_RAMP");
        const bool ramp = !(e && atoi(e) == 0);
        int64_t n = ramp ? std::min<int64_t>(C, std::max<int64_t>(C / 8, 65536)) : C;
        for (int64_t at = 0; at < N; at += n, n = std::min<int64_t>(C, n * 2)) {
            chunk_at.push_back(at);
            if (at + n >= N) break;
        }
        chunk_at.push_back(N);
    }
    const int64_t nchunks = (int64_t)chunk_at.size() - 1;
    const size_t ws_bytes = mm_interpolate_workspace_bytes(s->index, d, C, k);
    MM_CUDA(ensure(s, s->ws, ws_bytes));
    MM_CUDA(ensure(s, s->nf, sizeof(int64_t) * (size_t)nchunks));
    for (int i = 0; i < 2; ++i) {
        MM_CUDA(ensure(s, s->pts[i], sizeof(double) * C * d));
        if (values) MM_CUDA(ensure(s, s->out[i], sizeof(double) * C * F));
        if (elem) MM_CUDA(ensure(s, s->elem[i], sizeof(int32_t) * C));
        if (xi) MM_CUDA(ensure(s, s->xi[i], sizeof(double) * C * d));
    }
    MM_CUDA(sync_allocations(s));
    int64_t *nf_d = static_cast<int64_t *>(s->nf.p);
    void *fr = s->fields_pending ? (void *)s->fields_ready : nullptr;

    for (int64_t c = 0; c < nchunks; ++c) {
        const int b = (int)(c & 1);
        const int64_t at = chunk_at[c], n = chunk_at[c + 1] - at;
        // h2d: the points buffer is free once the kernels of chunk c-2 are done
        if (c >= 2) MM_CUDA(cudaStreamWaitEvent(s->h2d, s->ev_comp[b], 0));
        MM_CUDA(cudaMemcpyAsync(s->pts[b].p, pts + at * d, sizeof(double) * n * d, cudaMemcpyHostToDevice, s->h2d));
        MM_CUDA(cudaEventRecord(s->ev_in[b], s->h2d));
        // main: needs the points of chunk c and the output buffers of chunk c-2 drained
        MM_CUDA(cudaStreamWaitEvent(s->main, s->ev_in[b], 0));
        if (c >= 2) MM_CUDA(cudaStreamWaitEvent(s->main, s->ev_out[b], 0));
        MM_TRY(mm_interpolate_impl(s->index, s->gll_form ? s->P : 1, s->order, d, s->E, s->nodes, s->cent, s->aabb,
                                   s->pre, F, values ? s->fields : nullptr, n, static_cast<double *>(s->pts[b].p), k,
                                   params, values ? static_cast<double *>(s->out[b].p) : nullptr,
                                   elem ? static_cast<int32_t *>(s->elem[b].p) : nullptr,
                                   xi ? static_cast<double *>(s->xi[b].p) : nullptr, nullptr, nf_d + c, s->ws.p,
                                   s->ws.cap, s->main, fr));
        MM_CUDA(cudaEventRecord(s->ev_comp[b], s->main));
        // d2h
        MM_CUDA(cudaStreamWaitEvent(s->d2h, s->ev_comp[b], 0));
        if (values)
            MM_CUDA(cudaMemcpyAsync(values + at * F, s->out[b].p, sizeof(double) * n * F, cudaMemcpyDeviceToHost,
                                    s->d2h));
        if (elem)
            MM_CUDA(cudaMemcpyAsync(elem + at, s->elem[b].p, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, s->d2h));
        if (xi)
            MM_CUDA(cudaMemcpyAsync(xi + at * d, s->xi[b].p, sizeof(double) * n * d, cudaMemcpyDeviceToHost, s->d2h));
        MM_CUDA(cudaEventRecord(s->ev_out[b], s->d2h));
    }
    std::vector<int64_t> nf((size_t)nchunks, 0);
    MM_CUDA(cudaStreamWaitEvent(s->d2h, s->ev_comp[(nchunks - 1) & 1], 0));
    MM_CUDA(cudaMemcpyAsync(nf.data(), nf_d, sizeof(int64_t) * (size_t)nchunks, cudaMemcpyDeviceToHost, s->d2h));
    MM_CUDA(cudaStreamSynchronize(s->d2h));
    MM_CUDA(cudaStreamSynchronize(s->main));
    if (s->fields_pending) {
        MM_CUDA(cudaStreamSynchronize(s->h2d));
        s->fields_pending = false;
    }
    if (num_failed)
        for (int64_t v : nf) *num_failed += v;
    return MM_OK;
}
