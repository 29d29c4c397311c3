// mm_common.cuh -- shared device helpers for the sm_100a kernels.
//
// CANONICAL ARITHMETIC: every floating-point operation below is a single
// IEEE binary64 operation in the order written.  The library is compiled with -fmad=false so the
// compiler never contracts a*b+c by itself; the only fused operations are the explicit __fma_rn
// calls in the two tensor contractions (mm_newton.cuh eval_map, mm_interp.cu contract_field);
// nothing may be re-associated.  The CPU oracle (oracle/mm_oracle.c, test infrastructure)
// restates the same order independently, with C fma() at the same places.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#include "../../include/multimesh_b200.h"

#define MM_MAXM 5
#define MM_NEWTON_MAXIT 50
#define MM_NEWTON_TOL 1e-13
#define MM_NEWTON_TOL_FAST 1e-7     // accepted when the iteration is visibly quadratic (see newton_iterate)
#define MM_NEWTON_FAST_RATIO 1e-3
#define MM_NEWTON_DIVERGE 1e10

// ------------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------------
void mm_set_error(const char *fmt, ...);
int mm_cuda_fail(cudaError_t e, const char *what, const char *file, int line);

#define MM_CUDA(call)                                                          \
    do {                                                                       \
        cudaError_t _e = (call);                                               \
        if (_e != cudaSuccess) return mm_cuda_fail(_e, #call, __FILE__, __LINE__); \
    } while (0)

#define MM_REQUIRE(cond, code, ...)  \
    do {                             \
        if (!(cond)) {               \
            mm_set_error(__VA_ARGS__); \
            return (code);           \
        }                            \
    } while (0)

int mm_num_sms();                 // SM count of the current device
// stream-ordered allocation from the library's own per-device pool (mm_core.cu); never touches the
// device's default pool, which belongs to the host process
cudaError_t mm_pool_alloc(void **p, size_t bytes, cudaStream_t st);
void mm_pool_free(void *p, cudaStream_t st);

// internal (C++) entry points shared between translation units
int mm_locate_impl(int order, int dim, int64_t E, const double *nodes, const double *centroid,
                   const double *aabb, const double *presolve, int64_t N, const double *pts,
                   int pts_stride /* doubles between consecutive points */, int k,
                   const int32_t *cands,
                   const mm_locate_params *params, int32_t *elem, double *xi, uint8_t *status,
                   int64_t *num_failed, bool zero_num_failed, int32_t *unresolved_list,
                   int64_t *unresolved_count, void *stream, const int64_t *n_dev = nullptr, int64_t n_off = 0,
                   const int32_t *point_order = nullptr);  // [N] (prefix mode only): lane m works on point point_order[m]
// order[] that groups the points 0..N-1 by key[n * key_stride] (keys outside [0, E) go last), stable inside a group
// up to atomics; scratch: mm_interp_elem_scratch_bytes(E, N) bytes (K3's tables, free until K3 runs)
int mm_group_by_key(int64_t E, int64_t N, const int32_t *key, int key_stride, void *scratch, int32_t **order,
                    void *stream);
int mm_interpolate_impl(const mm_index_t *index, int32_t divisor, int order, int dim, int64_t E,
                        const double *nodes, const double *centroid, const double *aabb,
                        const double *presolve, int F, const double *fields, int64_t N,
                        const double *pts, int k, const mm_locate_params *params, double *out,
                        int32_t *elem, double *xi, uint8_t *status, int64_t *num_failed,
                        void *workspace, size_t workspace_bytes, void *stream, void *fields_ready);
int mm_interp_fused(int order, int dim, int64_t E, int F, const double *fields, int64_t N,
                    const int32_t *elem_s, const double *xi_s, const uint8_t *status_s,
                    const int32_t *perm, int perm_stride, double *out, int32_t *elem_u, double *xi_u,
                    uint8_t *status_u, void *stream);
// K3 of the fused pipeline, element-centric (mm_interp_elem.cu): points grouped by element with a counting sort,
// then one pass over the elements in memory order; scratch: mm_interp_elem_scratch_bytes(E or an upper bound, N)
size_t mm_interp_elem_scratch_bytes(int64_t E, int64_t N);
int mm_interp_by_element(int order, int dim, int64_t E, int F, const double *fields, int64_t N,
                         const int32_t *elem_s, const double *xi_s, const uint8_t *status_s, const int32_t *perm,
                         int perm_stride, double *out, int32_t *elem_u, double *xi_u, uint8_t *status_u,
                         void *scratch, void *stream);
// trilinear candidate search over sorted query records (mm_trilinear.cu), first pass (prefix) and re-run of
// mm_trilinear_indexed
int mm_trilinear_records(bool prefix, int k, int64_t npoints, const int64_t *n_dev, int64_t n_off, const int32_t *list,
                         const double *recs, const int32_t *cands, const int64_t *connectivity, const double *nodes,
                         int64_t *enclosing, double *weights, int64_t *num_failed, int32_t *unresolved_list,
                         int64_t *unresolved_count, void *stream);
int64_t mm_index_size(const mm_index_t *ix);  // number of indexed points
size_t mm_index_sort_scratch_bytes(const mm_index_t *ix);
// site table (distinct coordinates; built by the public mm_index_prepare_sites) and the site-level first
// pass of the progressive search
bool mm_index_has_sites(const mm_index_t *ix);
struct mm_index_sites_view {
    int64_t nsites, M;
    const double4 *site_recs;    // [nsites + 1] {x, y, z, first record}
    const int32_t *site_first;   // [nsites + 1]
    const int32_t *rec_id;       // [M] point id of record t (records of one site are contiguous, ids ascending)
};
bool mm_index_sites_view_get(const mm_index_t *ix, mm_index_sites_view *out);
int mm_knn_sites(const mm_index_t *ix, int64_t N, const double *pts, int pts_stride, int kout,
                 int32_t divisor, int32_t *idx, void *stream);
// first pass of the pipeline when the CTA-tile kernel does not apply (complete k' = k semantics, sparse query sets)
int mm_knn_first_pass(const mm_index_t *ix, int64_t N, const double *pts, int pts_stride, int k,
                      int32_t divisor, int32_t *idx, void *stream);
// CTA-tile first pass over SORTED queries, prefix semantics (mm_index.cu: knn_tile_kernel); *applied = false when it
// does not apply and the caller has to take one of the two above
// perm_out (optional, [N]): the original index of every sorted query, as a compact array
int mm_knn_tile_first_pass(const mm_index_t *ix, int64_t N, const double *sorted, const void *sort_scratch, int kout,
                           int32_t divisor, bool sites, int32_t *idx, int32_t *perm_out, void *stream, bool *applied);
// mm_knn with a stride (in doubles) between consecutive query points
// n_dev (optional, device): the kernel processes points [0, min(N, *n_dev - n_off)) -- for work lists whose
// length is only known on the device (no host synchronisation)
int mm_knn_strided(const mm_index_t *ix, int64_t N, const double *pts, int pts_stride, int k,
                   int32_t divisor, int32_t *idx, double *d2, void *stream, const int64_t *n_dev = nullptr,
                   int64_t n_off = 0);
// counting sort of query points by index cell into 32-byte records {x, y, z (0 in 2-D), original
// index in the low 32 bits of the fourth lane}: one aligned full-sector store per point
constexpr int MM_QREC = 4;  // doubles per sorted-query record
int mm_index_sort_queries(const mm_index_t *ix, int64_t N, const double *pts, double *sorted_rec,
                          int32_t *rank_tmp, void *scratch, void *stream);

// Opt-in to large dynamic shared memory once per (kernel, device) and cache the occupancy query: both are
// host-side driver calls that need not be repeated for every launch (and keeps launches free of anything but
// the launch itself, e.g. under CUDA-graph capture).  One instance per kernel (function-local static).
struct mm_kernel_cfg {
    size_t smem_set[16] = {};
    size_t occ_smem[16] = {};
    int occ_threads[16] = {};
    int occ[16] = {};
    template <class K>
    cudaError_t prepare(K kern, int threads, size_t smem, int *per_sm)
    {
        int dev = 0;
        cudaError_t rc = cudaGetDevice(&dev);
        if (rc != cudaSuccess) return rc;
        const int d = dev & 15;
        if (smem > smem_set[d]) {
            rc = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (rc != cudaSuccess) return rc;
            smem_set[d] = smem;
        }
        if (per_sm) {
            if (occ[d] == 0 || occ_smem[d] != smem || occ_threads[d] != threads) {
                int v = 1;
                rc = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, kern, threads, smem);
                if (rc != cudaSuccess) return rc;
                occ[d] = v < 1 ? 1 : v;
                occ_smem[d] = smem;
                occ_threads[d] = threads;
            }
            *per_sm = occ[d];
        }
        return cudaSuccess;
    }
};

static inline bool mm_valid_order(int order) { return order == 1 || order == 2 || order == 4; }
static inline int mm_pow(int m, int dim) { return dim == 2 ? m * m : m * m * m; }

// ------------------------------------------------------------------------------------------------
// GLL table of one order, built on the host in binary64 and passed to kernels BY VALUE (kernel
// parameters live in the constant bank, so reads are uniform constant loads):
//   z : nodes;   c : 1 / prod_{j != i, ascending}(z_i - z_j)
// ------------------------------------------------------------------------------------------------
struct mm_gll_table {
    double z[MM_MAXM];
    double c[MM_MAXM];
};
int mm_make_table(int order, mm_gll_table *t);  // host; returns m = order + 1 or 0

// L_i(x) = c_i * prod_{j != i, ascending}(x - z_j)
template <int ORDER>
__device__ __forceinline__ void lagrange_values(const mm_gll_table &T, double x,
                                                double (&L)[ORDER + 1])
{
    constexpr int M = ORDER + 1;
    double d[M];
#pragma unroll
    for (int j = 0; j < M; ++j) d[j] = x - T.z[j];
#pragma unroll
    for (int i = 0; i < M; ++i) {
        double prod = T.c[i];
#pragma unroll
        for (int j = 0; j < M; ++j)
            if (j != i) prod = prod * d[j];
        L[i] = prod;
    }
}

// L_i(x) and L_i'(x) = sum_{q != i, ascending} c_i * prod_{j != i,q, ascending}(x - z_j)
template <int ORDER>
__device__ __forceinline__ void lagrange_values_derivs(const mm_gll_table &T, double x,
                                                       double (&L)[ORDER + 1],
                                                       double (&dL)[ORDER + 1])
{
    constexpr int M = ORDER + 1;
    double d[M];
#pragma unroll
    for (int j = 0; j < M; ++j) d[j] = x - T.z[j];
#pragma unroll
    for (int i = 0; i < M; ++i) {
        double prod = T.c[i];
#pragma unroll
        for (int j = 0; j < M; ++j)
            if (j != i) prod = prod * d[j];
        L[i] = prod;
        double sum = 0.0;
#pragma unroll
        for (int q = 0; q < M; ++q) {
            if (q == i) continue;
            double term = T.c[i];
#pragma unroll
            for (int j = 0; j < M; ++j)
                if (j != i && j != q) term = term * d[j];
            sum = sum + term;
        }
        dL[i] = sum;
    }
}

// ------------------------------------------------------------------------------------------------
// mbarrier + bulk-async-copy (TMA 1-D, cp.async.bulk) wrappers.  SASS: UBLKCP / SYNCS.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// global -> shared bulk copy; dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_copy_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes,
                                              uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

__device__ __forceinline__ void prefetch_l2(const void *ptr)
{
    asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr));
}

// orders generic-proxy accesses to shared memory before later async-proxy (TMA) writes
__device__ __forceinline__ void fence_proxy_async_smem()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// slot stride for per-thread shared-memory staging: >= bytes, a multiple of 16 whose
// 16-byte count is odd (at most a 2-way bank conflict for 8-byte accesses by a half-warp)
__host__ __device__ constexpr int mm_slot_bytes(int bytes)
{
    int q = (bytes + 15) / 16;
    return ((q % 2) ? q : q + 1) * 16;
}
