// mm_pipeline.cu -- fused device pipeline K1 -> K2 -> K3 over one batch of target points.
//
//   1. counting-sort the points by index cell (coherent warps, L2 locality); K1 and K2 run in sorted order and
//      results are written back through the permutation;
//   2. progressive search: K1 writes a certified PREFIX of the canonical k-NN list (up to k1 = 8 / 4 entries; CTA-tile
//      kernel over the cells, mm_index.cu), K2 runs in "prefix" mode -- a point is final if one of the prefix's
//      candidates accepts it, which is exactly what the full list would have decided, because any prefix of the
//      canonical k-NN list IS the list of the nearest; points that exhaust the prefix are appended to a work list;
//   3. the work list (typically a few per cent) is re-run in chunks with all k candidates and
//      the variant's complete fallback logic, and scattered back; its length stays on the device
//      (the re-run kernels read it there), so the whole call is stream-ordered: no host
//      synchronisation, capturable in a CUDA graph;
//   4. K3 groups the located points by source element and gathers element by element (mm_interp_elem.cu),
//      writing out[perm[n]] and the un-permuted location outputs.
// Results are identical to mm_knn -> mm_locate -> mm_interp (tests/test_gpu_parity.py).
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "mm_common.cuh"

struct mm_profile {
    int max_calls = 0;
    int calls = 0;
    std::vector<cudaEvent_t> ev;  // [max_calls][MM_N_STAGES + 1]
};

static thread_local mm_profile *g_profile = nullptr;

namespace {

// points per re-run chunk: the number of unresolved points is only known on the device, so the host
// enqueues ceil(N / chunk) <= 4 rounds and the kernels of a round that has nothing to do exit at once
inline int64_t rerun_chunk(int64_t N)
{
    if (const char *e = getenv("MM_RERUN_CHUNK")) {  // tests: force several rounds on small inputs
        const long long v = atoll(e);
        if (v > 0) return v;
    }
    return std::max<int64_t>((int64_t)1 << 20, (N + 3) / 4);
}

inline size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

struct ws_layout {
    size_t sorted, perm, perm_c, cands1, elem, xi, status, list, counters, sort_scratch, k3_scratch;
    size_t b_pts, b_cands, b_elem, b_xi, b_status, total;
};

// first-pass list length: 8 for the GLL-point form (a shared node appears up to 8 times, and its
// copies are consecutive in the k-NN order), 4 for the centroid form
inline int first_pass_k(int k, int32_t divisor) { return std::min(k, divisor > 1 ? 8 : 4); }

ws_layout make_layout(const mm_index_t *ix, int dim, int64_t N, int k)
{
    ws_layout L{};
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t at = o; o += align_up(bytes); return at; };
    const int k1 = std::min(k, 8);  // sized for the larger first pass
    L.sorted = take(sizeof(double) * N * MM_QREC);  // 32-byte records {x, y, z, original index}
    L.perm = 0;
    L.perm_c = take(sizeof(int32_t) * N);  // the permutation once more, compact (written by K1's tile kernel)
    L.cands1 = take(sizeof(int32_t) * N * k1);
    L.elem = take(sizeof(int32_t) * N);
    L.xi = take(sizeof(double) * N * dim);
    L.status = take(N);
    L.list = take(sizeof(int32_t) * N);
    L.counters = take(64);
    L.sort_scratch = take(mm_index_sort_scratch_bytes(ix));
    // element tables of K3: the number of source elements is not an argument of the workspace query; the number of
    // indexed points (>= elements in both k-NN forms) bounds it
    L.k3_scratch = take(mm_interp_elem_scratch_bytes(mm_index_size(ix), N));
    const int64_t cb = std::min<int64_t>(N, rerun_chunk(N));
    L.b_pts = take(sizeof(double) * cb * dim);
    L.b_cands = take(sizeof(int32_t) * cb * k);
    L.b_elem = take(sizeof(int32_t) * cb);
    L.b_xi = take(sizeof(double) * cb * dim);
    L.b_status = take(cb);
    L.total = o;
    return L;
}

__global__ void __launch_bounds__(256)
gather_points_kernel(int dim, int64_t n, const long long *__restrict__ n_dev, int64_t n_off,
                     const int32_t *__restrict__ list, const double *__restrict__ pts, double *__restrict__ out)
{
    {
        const long long have = *n_dev - n_off;
        n = have < 0 ? 0 : (have < n ? have : n);
    }
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x)
        for (int c = 0; c < dim; ++c) out[i * dim + c] = pts[(int64_t)list[i] * MM_QREC + c];  // from the records
}

__global__ void __launch_bounds__(256)
scatter_results_kernel(int dim, int64_t n, const long long *__restrict__ n_dev, int64_t n_off,
                       const int32_t *__restrict__ list,
                       const int32_t *__restrict__ elem_b, const double *__restrict__ xi_b,
                       const uint8_t *__restrict__ status_b, int32_t *__restrict__ elem,
                       double *__restrict__ xi, uint8_t *__restrict__ status)
{
    {
        const long long have = *n_dev - n_off;
        n = have < 0 ? 0 : (have < n ? have : n);
    }
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t t = list[i];
        elem[t] = elem_b[i];
        status[t] = status_b[i];
        for (int c = 0; c < dim; ++c) xi[t * dim + c] = xi_b[i * dim + c];
    }
}

__global__ void __launch_bounds__(256)
unpermute_kernel(int dim, int64_t n, const int32_t *__restrict__ perm, int perm_stride,
                 const int32_t *__restrict__ elem_s, const double *__restrict__ xi_s,
                 const uint8_t *__restrict__ status_s, int32_t *__restrict__ elem,
                 double *__restrict__ xi, uint8_t *__restrict__ status)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t t = perm[i * perm_stride];
        if (elem) elem[t] = elem_s[i];
        if (status) status[t] = status_s[i];
        if (xi)
            for (int c = 0; c < dim; ++c) xi[t * dim + c] = xi_s[i * dim + c];
    }
}

int blocks_for(int64_t work)
{
    int sms = mm_num_sms() > 0 ? mm_num_sms() : 148;
    int64_t need = (work + 255) / 256;
    int64_t cap = (int64_t)sms * 16;
    return (int)(need < 1 ? 1 : (need > cap ? cap : need));
}

}  // namespace

#define MM_TRY(call)                  \
    do {                              \
        int _rc = (call);             \
        if (_rc != MM_OK) return _rc; \
    } while (0)

extern "C" size_t mm_interpolate_workspace_bytes(const mm_index_t *index, int dim, int64_t N, int k)
{
    if (!index || N < 0 || k < 1) return 0;
    return make_layout(index, dim, N, k).total;
}

// fields_ready: optional event the stream waits on right before K3 (the host entry point copies
// the field blocks on a second stream while K1/K2 run)
int mm_interpolate_impl(const mm_index_t *index, int32_t divisor, int order, int dim, int64_t E,
                        const double *nodes, const double *centroid, const double *aabb,
                        const double *presolve, int F, const double *fields, int64_t N,
                        const double *pts, int k, const mm_locate_params *params, double *out,
                        int32_t *elem, double *xi, uint8_t *status, int64_t *num_failed,
                        void *workspace, size_t workspace_bytes, void *stream_, void *fields_ready)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    MM_REQUIRE(index && params, MM_ERR_INVALID, "mm_interpolate: null index/params");
    MM_REQUIRE(k >= 1 && k <= 64, MM_ERR_INVALID, "mm_interpolate: k=%d outside [1, 64]", k);
    MM_REQUIRE(N >= 0 && N <= (int64_t)INT32_MAX, MM_ERR_INVALID,
               "mm_interpolate: N=%lld outside [0, 2^31): the permutation and the work lists are int32 -- split "
               "the target points into several calls", (long long)N);
    MM_REQUIRE(E >= 0 && E <= (int64_t)INT32_MAX, MM_ERR_INVALID, "mm_interpolate: E=%lld", (long long)E);
    if (num_failed) MM_CUDA(cudaMemsetAsync(num_failed, 0, sizeof(int64_t), stream));
    if (N == 0) return MM_OK;
    const ws_layout L = make_layout(index, dim, N, k);
    MM_REQUIRE(workspace && workspace_bytes >= L.total, MM_ERR_INVALID,
               "mm_interpolate: workspace of %zu bytes needed, %zu given", L.total, workspace_bytes);
    MM_REQUIRE(((uintptr_t)workspace & 255) == 0, MM_ERR_INVALID, "mm_interpolate: workspace must be 256-byte aligned");
    unsigned char *ws = static_cast<unsigned char *>(workspace);
    double *sorted = reinterpret_cast<double *>(ws + L.sorted);
    // the permutation is the low word of the fourth lane of every record
    const int32_t *perm = reinterpret_cast<const int32_t *>(sorted + 3);
    constexpr int PERM_STRIDE = MM_QREC * 2;
    int32_t *cands1 = reinterpret_cast<int32_t *>(ws + L.cands1);
    int32_t *elem_s = reinterpret_cast<int32_t *>(ws + L.elem);
    double *xi_s = reinterpret_cast<double *>(ws + L.xi);
    uint8_t *status_s = ws + L.status;
    int32_t *list = reinterpret_cast<int32_t *>(ws + L.list);
    int64_t *counters = reinterpret_cast<int64_t *>(ws + L.counters);  // [0] unresolved, [1] failed
    const int k1 = first_pass_k(k, divisor);

    cudaEvent_t *pev = nullptr;  // stage-boundary events of this call, if profiling is on
    if (g_profile && g_profile->calls < g_profile->max_calls)
        pev = g_profile->ev.data() + (size_t)(g_profile->calls++) * (MM_N_STAGES + 1);
    auto mark = [&](int i) {
        if (pev) cudaEventRecord(pev[i], stream);
    };
    mark(0);

    // 1. spatial sort of the target points
    // (the element array is not written before K2: it holds the per-cell ranks of the sort meanwhile)
    MM_TRY(mm_index_sort_queries(index, N, pts, sorted, reinterpret_cast<int32_t *>(ws + L.elem),
                                 ws + L.sort_scratch, stream));
    MM_CUDA(cudaMemsetAsync(counters, 0, 64, stream));

    mark(1);
    // 2. first pass: k1 nearest candidates, prefix mode (unless k1 == k: complete semantics)
    // GLL-point form: shared nodes are stored up to 8 times; search over distinct coordinates when the
    // index has a site table (mm_index_prepare_sites)
    const bool site_pass = divisor > 1 && k1 < k && mm_index_has_sites(index) && !getenv("MM_NO_SITES");
    bool tiled = false;
    if (k1 < k)  // prefix mode: the CTA-tile kernel may write shorter prefixes, the re-run below completes them
        MM_TRY(mm_knn_tile_first_pass(index, N, sorted, ws + L.sort_scratch, k1, divisor, site_pass, cands1,
                                      reinterpret_cast<int32_t *>(ws + L.perm_c), stream, &tiled));
    if (tiled) {
    } else if (site_pass) {
        MM_TRY(mm_knn_sites(index, N, sorted, MM_QREC, k1, divisor, cands1, stream));
    } else {
        MM_TRY(mm_knn_first_pass(index, N, sorted, MM_QREC, k1, divisor, cands1, stream));
    }
    mark(2);
    mm_locate_params p1 = *params;
    p1.reserved = (k1 < k) ? 1 : 0;
    // Centroid form: the first candidate is (almost always) the owner, and the points of a cell belong to up to 8
    // elements -- a warp of 32 cell-sorted points stages ~10 element blocks for ~3 lanes each (S5: 101 GB of DRAM reads
    // for a 30 GB node array, 15 of 32 lanes active).  Grouped by first candidate, a warp shares 3-4 blocks.
    const int32_t *k2_order = nullptr;
    if (k1 < k && divisor == 1 && E > 0 && E <= mm_index_size(index) && !getenv("MM_K2_NO_GROUPING")) {
        int32_t *ord = nullptr;
        MM_TRY(mm_group_by_key(E, N, cands1, k1, ws + L.k3_scratch, &ord, stream));
        k2_order = ord;
    }
    MM_TRY(mm_locate_impl(order, dim, E, nodes, centroid, aabb, presolve, N, sorted, MM_QREC, k1, cands1, &p1, elem_s,
                          xi_s, status_s, counters + 1, false, (k1 < k) ? list : nullptr,
                          (k1 < k) ? counters : nullptr, stream, nullptr, 0, k2_order));

    mark(3);
    // 3. re-run the unresolved points with the full candidate list.  Their number lives in counters[0] on the
    //    device; every kernel of a round reads it there and processes list[at, min(at + cb, count))
    if (k1 < k) {
        mm_locate_params p2 = *params;
        p2.reserved = 0;
        double *b_pts = reinterpret_cast<double *>(ws + L.b_pts);
        int32_t *b_cands = reinterpret_cast<int32_t *>(ws + L.b_cands);
        int32_t *b_elem = reinterpret_cast<int32_t *>(ws + L.b_elem);
        double *b_xi = reinterpret_cast<double *>(ws + L.b_xi);
        uint8_t *b_status = ws + L.b_status;
        const long long *n_un = reinterpret_cast<const long long *>(counters);
        const int64_t cb = std::min<int64_t>(N, rerun_chunk(N));
        for (int64_t at = 0; at < N; at += cb) {
            const int64_t nb = std::min<int64_t>(cb, N - at);
            gather_points_kernel<<<blocks_for(nb), 256, 0, stream>>>(dim, nb, n_un, at, list + at, sorted, b_pts);
            MM_TRY(mm_knn_strided(index, nb, b_pts, dim, k, divisor, b_cands, nullptr, stream, counters, at));
            MM_TRY(mm_locate_impl(order, dim, E, nodes, centroid, aabb, presolve, nb, b_pts, dim, k, b_cands, &p2,
                                  b_elem, b_xi, b_status, counters + 1, false, nullptr, nullptr, stream, counters,
                                  at));
            scatter_results_kernel<<<blocks_for(nb), 256, 0, stream>>>(
                dim, nb, n_un, at, list + at, b_elem, b_xi, b_status, elem_s, xi_s, status_s);
            MM_CUDA(cudaGetLastError());
        }
    }
    if (num_failed)
        MM_CUDA(cudaMemcpyAsync(num_failed, counters + 1, sizeof(int64_t), cudaMemcpyDeviceToDevice, stream));

    mark(4);
    // 4. gather in sorted order, written back through the permutation; the kernel also writes the
    //    un-permuted elem / xi / status when they are requested (needs elem)
    bool unpermuted = false;
    if (fields) {
        MM_REQUIRE(out, MM_ERR_INVALID, "mm_interpolate: null out");
        if (fields_ready) MM_CUDA(cudaStreamWaitEvent(stream, (cudaEvent_t)fields_ready, 0));
        const char *mode = getenv("MM_INTERP_MODE");  // "t" / "w": the point-order variants (A/B comparisons)
        if (!mode && E <= mm_index_size(index)) {
            // (K1's tile kernel left a compact copy of the permutation: 4 bytes per point instead of a record)
            const int32_t *perm3 = tiled ? reinterpret_cast<const int32_t *>(ws + L.perm_c) : perm;
            MM_TRY(mm_interp_by_element(order, dim, E, F, fields, N, elem_s, xi_s, status_s, perm3,
                                        tiled ? 1 : PERM_STRIDE, out, elem, xi, status, ws + L.k3_scratch, stream));
        } else {
            MM_TRY(mm_interp_fused(order, dim, E, F, fields, N, elem_s, xi_s, status_s, perm, PERM_STRIDE, out, elem,
                                   xi, status, stream));
        }
        unpermuted = elem != nullptr;
    }
    mark(5);
    if (!unpermuted && (elem || xi || status)) {
        unpermute_kernel<<<blocks_for(N), 256, 0, stream>>>(dim, N, perm, PERM_STRIDE, elem_s, xi_s, status_s,
                                                            elem, xi, status);
        MM_CUDA(cudaGetLastError());
    }
    mark(6);
    return MM_OK;
}

extern "C" int mm_interpolate(const mm_index_t *index, int32_t divisor, int order, int dim,
                              int64_t E, const double *nodes, const double *centroid,
                              const double *aabb, const double *presolve, int F,
                              const double *fields, int64_t N, const double *pts, int k,
                              const mm_locate_params *params, double *out, int32_t *elem,
                              double *xi, uint8_t *status, int64_t *num_failed, void *workspace,
                              size_t workspace_bytes, void *stream)
{
    return mm_interpolate_impl(index, divisor, order, dim, E, nodes, centroid, aabb, presolve, F, fields,
                               N, pts, k, params, out, elem, xi, status, num_failed, workspace,
                               workspace_bytes, stream, nullptr);
}

// ------------------------------------------------------------------------------------------------
// Order-1 nodal (Exodus HEX8) path, progressive: the same three steps as above with the trilinear
// candidate loop (mm_trilinear.cu) in the place of K2 -- sort, certified 4-prefix of the k-NN list,
// search in prefix mode, re-run of the points without acceptance with all k candidates and the
// C routine's complete logic.  Results are those of mm_knn(k) + mm_trilinear, bit for bit.
// ------------------------------------------------------------------------------------------------
namespace {
struct tri_layout {
    size_t sorted, cands1, rank, list, counters, sort_scratch, b_pts, b_cands, total;
};

constexpr int TRI_K1 = 4;

tri_layout make_tri_layout(const mm_index_t *ix, int64_t N, int k)
{
    tri_layout L{};
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t at = o; o += align_up(bytes); return at; };
    L.sorted = take(sizeof(double) * N * MM_QREC);
    L.cands1 = take(sizeof(int32_t) * N * TRI_K1);
    L.rank = take(sizeof(int32_t) * N);
    L.list = take(sizeof(int32_t) * N);
    L.counters = take(64);
    L.sort_scratch = take(mm_index_sort_scratch_bytes(ix));
    const int64_t cb = std::min<int64_t>(N, rerun_chunk(N));
    L.b_pts = take(sizeof(double) * cb * 3);
    L.b_cands = take(sizeof(int32_t) * cb * k);
    L.total = o;
    return L;
}
}  // namespace

extern "C" size_t mm_trilinear_indexed_workspace_bytes(const mm_index_t *index, int64_t N, int k)
{
    if (!index || N < 0 || k < 1) return 0;
    return make_tri_layout(index, N, k).total;
}

extern "C" int mm_trilinear_indexed(const mm_index_t *index, int64_t nelem, const int64_t *connectivity,
                                    const double *nodes, int64_t N, const double *pts, int k, int64_t *enclosing,
                                    double *weights, int64_t *num_failed, void *workspace, size_t workspace_bytes,
                                    void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    MM_REQUIRE(index && num_failed, MM_ERR_INVALID, "mm_trilinear_indexed: null index/num_failed");
    MM_REQUIRE(k >= 1 && k <= 64, MM_ERR_INVALID, "mm_trilinear_indexed: k=%d outside [1, 64]", k);
    MM_REQUIRE(N >= 0 && N <= (int64_t)INT32_MAX, MM_ERR_INVALID, "mm_trilinear_indexed: N=%lld outside [0, 2^31)",
               (long long)N);
    int64_t info[8];
    MM_TRY(mm_index_info(index, info, nullptr));
    MM_REQUIRE(info[1] == 3, MM_ERR_INVALID, "mm_trilinear_indexed: the index must be 3-D (HEX8 centroids)");
    MM_REQUIRE(nelem == mm_index_size(index), MM_ERR_INVALID,
               "mm_trilinear_indexed: the index holds %lld points for %lld elements (it must be built over the "
               "element centroids)", (long long)mm_index_size(index), (long long)nelem);
    MM_CUDA(cudaMemsetAsync(num_failed, 0, sizeof(int64_t), stream));
    if (N == 0) return MM_OK;
    MM_REQUIRE(connectivity && nodes && pts && enclosing && weights, MM_ERR_INVALID, "mm_trilinear_indexed: null buffer");
    const tri_layout L = make_tri_layout(index, N, k);
    MM_REQUIRE(workspace && workspace_bytes >= L.total, MM_ERR_INVALID,
               "mm_trilinear_indexed: workspace of %zu bytes needed, %zu given", L.total, workspace_bytes);
    MM_REQUIRE(((uintptr_t)workspace & 255) == 0, MM_ERR_INVALID, "mm_trilinear_indexed: workspace must be 256-byte aligned");
    unsigned char *ws = static_cast<unsigned char *>(workspace);
    double *sorted = reinterpret_cast<double *>(ws + L.sorted);
    int32_t *cands1 = reinterpret_cast<int32_t *>(ws + L.cands1);
    int32_t *list = reinterpret_cast<int32_t *>(ws + L.list);
    int64_t *counters = reinterpret_cast<int64_t *>(ws + L.counters);  // [0] points without acceptance in the prefix
    const int k1 = std::min(k, TRI_K1);

    MM_TRY(mm_index_sort_queries(index, N, pts, sorted, reinterpret_cast<int32_t *>(ws + L.rank), ws + L.sort_scratch,
                                 stream));
    MM_CUDA(cudaMemsetAsync(counters, 0, 64, stream));
    bool tiled = false;
    if (k1 < k)
        MM_TRY(mm_knn_tile_first_pass(index, N, sorted, ws + L.sort_scratch, k1, 1, false, cands1, nullptr, stream,
                                      &tiled));
    if (!tiled) MM_TRY(mm_knn_first_pass(index, N, sorted, MM_QREC, k1, 1, cands1, stream));
    MM_TRY(mm_trilinear_records(k1 < k, k1, N, nullptr, 0, nullptr, sorted, cands1, connectivity, nodes, enclosing,
                                weights, num_failed, list, counters, stream));
    if (k1 < k) {
        double *b_pts = reinterpret_cast<double *>(ws + L.b_pts);
        int32_t *b_cands = reinterpret_cast<int32_t *>(ws + L.b_cands);
        const long long *n_un = reinterpret_cast<const long long *>(counters);
        const int64_t cb = std::min<int64_t>(N, rerun_chunk(N));
        for (int64_t at = 0; at < N; at += cb) {
            const int64_t nb = std::min<int64_t>(cb, N - at);
            gather_points_kernel<<<blocks_for(nb), 256, 0, stream>>>(3, nb, n_un, at, list + at, sorted, b_pts);
            MM_CUDA(cudaGetLastError());
            MM_TRY(mm_knn_strided(index, nb, b_pts, 3, k, 1, b_cands, nullptr, stream, counters, at));
            MM_TRY(mm_trilinear_records(false, k, nb, counters, at, list + at, sorted, b_cands, connectivity, nodes,
                                        enclosing, weights, num_failed, nullptr, nullptr, stream));
        }
    }
    return MM_OK;
}

extern "C" int mm_profile_create(mm_profile_t **out, int max_calls)
{
    MM_REQUIRE(out && max_calls > 0, MM_ERR_INVALID, "mm_profile_create: arguments");
    mm_profile *p = new mm_profile();
    p->max_calls = max_calls;
    p->ev.resize((size_t)max_calls * (MM_N_STAGES + 1));
    for (auto &e : p->ev) {
        cudaError_t rc = cudaEventCreate(&e);
        if (rc != cudaSuccess) {
            delete p;
            return mm_cuda_fail(rc, "cudaEventCreate", __FILE__, __LINE__);
        }
    }
    *out = p;
    return MM_OK;
}

extern "C" int mm_profile_destroy(mm_profile_t *p)
{
    if (!p) return MM_OK;
    if (g_profile == p) g_profile = nullptr;
    for (auto &e : p->ev) cudaEventDestroy(e);
    delete p;
    return MM_OK;
}

extern "C" int mm_profile_begin(mm_profile_t *p)
{
    MM_REQUIRE(p, MM_ERR_INVALID, "mm_profile_begin: null");
    p->calls = 0;
    g_profile = p;
    return MM_OK;
}

extern "C" int mm_profile_end(void)
{
    g_profile = nullptr;
    return MM_OK;
}

extern "C" int mm_profile_read(mm_profile_t *p, int *n_calls, float *stage_ms)
{
    MM_REQUIRE(p && n_calls && stage_ms, MM_ERR_INVALID, "mm_profile_read: null");
    *n_calls = p->calls;
    for (int c = 0; c < p->calls; ++c) {
        cudaEvent_t *e = p->ev.data() + (size_t)c * (MM_N_STAGES + 1);
        MM_CUDA(cudaEventSynchronize(e[MM_N_STAGES]));
        for (int s = 0; s < MM_N_STAGES; ++s)
            MM_CUDA(cudaEventElapsedTime(&stage_ms[c * MM_N_STAGES + s], e[s], e[s + 1]));
    }
    return MM_OK;
}
