// mm_index.cu -- K1: GPU spatial index (counting-sorted uniform grid) and exact k-NN query.
//
// Replaces pykdtree's KDTree(data).query(pts, k) (call sites listed in include/multimesh_b200.h).
// The result is the UNIQUE k-prefix of the data points under the strict total order
// (d2, index), d2 = (dx*dx + dy*dy) + dz*dz in binary64 without FMA, so it does not depend on the
// grid resolution, on the order points are visited in, or on the atomics used while building.
//
// Layout in HBM:  recs  [M]  32-byte records {x, y, z, id}  sorted by linear cell id
//                             (x fastest, so a row of cells is ONE contiguous record range)
//                 cell_start [ncells + 1] int32
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "mm_common.cuh"
#include "mm_scan.cuh"

struct mm_index {
    int dim = 0;
    int64_t M = 0;
    double origin[3] = {0, 0, 0};
    double cell = 1.0, inv_cell = 1.0;
    int n[3] = {1, 1, 1};
    int64_t ncells = 1;
    int64_t nonempty = 0;
    double distinct_est = 0.0;  // estimated number of distinct coordinates (probe grid)
    double4 *recs = nullptr;
    int32_t *cell_start = nullptr;
    size_t bytes = 0;
    // site table (built on demand, mm_index_build_sites): records of one cell sorted by
    // (x, y, z, id), so that points with identical coordinates -- GLL nodes shared by up to 8
    // elements -- are contiguous; a "site" is one distinct coordinate
    int64_t nsites = 0;
    double4 *site_recs = nullptr;        // [nsites + 1] {x, y, z, first record of the site}
    int32_t *site_cell_start = nullptr;  // [ncells + 1] first site of each cell
    // compact tables of the warp-cooperative first pass (knn_block_kernel):
    //   recf     [M] (plain form) or [nsites] (site form): float4 {x, y, z relative to the corner of the
    //            record's own cell, record / site position}
    //   rec_id   [M] int32: point id of record t (recs[t].w), 4 bytes instead of a 32-byte record per look-up
    //   site_first [nsites + 1] int32: first record of each site
    float4 *recf = nullptr;
    bool recf_sites = false;
    int32_t *rec_id = nullptr;
    int32_t *site_first = nullptr;
    cudaStream_t stream = nullptr;       // stream the buffers were allocated on (stream-ordered pool)
};

namespace {

inline cudaError_t pool_alloc(void **p, size_t bytes, cudaStream_t st) { return mm_pool_alloc(p, bytes, st); }

constexpr int64_t MAX_CELLS = (int64_t)1 << 26;
constexpr int KNN_BLOCK = 128;

// idx / divisor for 0 <= idx < 2^31 with a host-computed multiplier (Granlund-Montgomery:
// M = ceil(2^(31+l) / d), l = ceil(log2 d)); a runtime integer division costs ~20 instructions and K1
// performs one per candidate written
struct fast_div {
    unsigned long long mul = 1;
    int shift = 0;
    int32_t d = 1;
    __host__ explicit fast_div(int32_t divisor = 1) : d(divisor)
    {
        if (divisor > 1) {
            int l = 0;
            while (((int64_t)1 << l) < divisor) ++l;
            shift = 31 + l;
            mul = ((1ull << shift) + (unsigned long long)divisor - 1) / (unsigned long long)divisor;
        }
    }
    __device__ __forceinline__ int32_t operator()(int32_t n) const
    {
        return d == 1 ? n : (int32_t)(((unsigned long long)(uint32_t)n * mul) >> shift);
    }
};
#ifndef MM_KNN_MERGED
#define MM_KNN_MERGED 1
#endif

struct grid_t {
    double origin[3];
    double cell, inv_cell;
    int n[3];
    int dim;
};

__device__ __forceinline__ int cell_coord(const grid_t &g, double x, int c)
{
    double f = floor((x - g.origin[c]) * g.inv_cell);
    double hi = (double)(g.n[c] - 1);
    if (!(f >= 0.0)) f = 0.0;  // negative or NaN (the conversion of a NaN to int is not a valid cell)
    if (f > hi) f = hi;        // clamping keeps monotonicity
    return (int)f;
}

__device__ __forceinline__ int64_t cell_of(const grid_t &g, const double *p)
{
    int cx = cell_coord(g, p[0], 0), cy = cell_coord(g, p[1], 1);
    int cz = g.dim == 3 ? cell_coord(g, p[2], 2) : 0;
    return cx + (int64_t)g.n[0] * (cy + (int64_t)g.n[1] * cz);
}

// ---- bounding box: per-block partial min/max, finished on the host ------------------------------
__global__ void __launch_bounds__(256)
bbox_kernel(int dim, int64_t M, const double *__restrict__ pts, double *__restrict__ partial)
{
    double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < M;
         i += (int64_t)gridDim.x * blockDim.x)
        for (int c = 0; c < dim; ++c) {
            double v = pts[i * dim + c];
            lo[c] = fmin(lo[c], v);
            hi[c] = fmax(hi[c], v);
        }
    __shared__ double s[6][256];
    for (int c = 0; c < 3; ++c) {
        s[c][threadIdx.x] = lo[c];
        s[3 + c][threadIdx.x] = hi[c];
    }
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (threadIdx.x < w)
            for (int c = 0; c < 3; ++c) {
                s[c][threadIdx.x] = fmin(s[c][threadIdx.x], s[c][threadIdx.x + w]);
                s[3 + c][threadIdx.x] = fmax(s[3 + c][threadIdx.x], s[3 + c][threadIdx.x + w]);
            }
        __syncthreads();
    }
    if (threadIdx.x < 6) partial[blockIdx.x * 6 + threadIdx.x] = s[threadIdx.x][0];
}

__global__ void __launch_bounds__(256)
histogram_kernel(grid_t g, int64_t M, const double *__restrict__ pts, int32_t *__restrict__ counts)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < M;
         i += (int64_t)gridDim.x * blockDim.x)
        atomicAdd(&counts[cell_of(g, pts + i * g.dim)], 1);
}

__global__ void __launch_bounds__(256)
count_nonempty_kernel(int64_t ncells, const int32_t *__restrict__ counts,
                      unsigned long long *__restrict__ nonempty)
{
    unsigned long long local = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < ncells;
         i += (int64_t)gridDim.x * blockDim.x)
        local += counts[i] > 0;
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(nonempty, local);
}

__global__ void __launch_bounds__(256)
scatter_kernel(grid_t g, int64_t M, const double *__restrict__ pts,
               const int32_t *__restrict__ cell_start, int32_t *__restrict__ cursor,
               double4 *__restrict__ recs)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < M;
         i += (int64_t)gridDim.x * blockDim.x) {
        const double *p = pts + i * g.dim;
        int64_t c = cell_of(g, p);
        int32_t pos = cell_start[c] + atomicAdd(&cursor[c], 1);
        double4 r;
        r.x = p[0];
        r.y = p[1];
        r.z = g.dim == 3 ? p[2] : 0.0;
        r.w = __longlong_as_double((long long)i);
        recs[pos] = r;
    }
}

// ---- k-NN query: one thread per query point -----------------------------------------------------
// Two list policies with the same interface:
//   smem_list   sorted top-k list in shared memory (any k <= 64); slot-major layout, so the
//               accesses are conflict-free whatever slot each lane touches
//   reg_list<K> sorted top-K list in registers (K = 4 or 8), used for small k: the first pass
//               of the progressive search (mm_pipeline.cu) and plain queries with k <= 8
struct smem_list {
    double *d2;
    int32_t *id;
    int k, cnt;
    double kth_d2;
    int32_t kth_id;

    __device__ __forceinline__ void init(unsigned char *smem, int k_)
    {
        d2 = reinterpret_cast<double *>(smem);
        id = reinterpret_cast<int32_t *>(smem + (size_t)k_ * KNN_BLOCK * sizeof(double));
        k = k_;
    }
    __device__ __forceinline__ void reset()
    {
        cnt = 0;
        kth_d2 = INFINITY;
        kth_id = 0x7fffffff;
    }
    __device__ __forceinline__ double &D(int s) { return d2[s * KNN_BLOCK + threadIdx.x]; }
    __device__ __forceinline__ int32_t &I(int s) { return id[s * KNN_BLOCK + threadIdx.x]; }
    __device__ __forceinline__ bool full() const { return cnt == k; }
    __device__ __forceinline__ double worst() const { return kth_d2; }

    __device__ __forceinline__ void insert(double nd, int32_t ni)
    {
        int pos;
        if (cnt == k) {
            if (!(nd < kth_d2 || (nd == kth_d2 && ni < kth_id))) return;
            pos = k - 1;
        } else {
            pos = cnt++;
        }
        while (pos > 0) {
            double pd = D(pos - 1);
            int32_t pi = I(pos - 1);
            if (!(nd < pd || (nd == pd && ni < pi))) break;
            D(pos) = pd;
            I(pos) = pi;
            --pos;
        }
        D(pos) = nd;
        I(pos) = ni;
        if (cnt == k) {
            kth_d2 = D(k - 1);
            kth_id = I(k - 1);
        }
    }
    __device__ __forceinline__ void write(int32_t *oi, double *od, const fast_div &divisor)
    {
        for (int t = 0; t < k; ++t) {
            bool have = t < cnt;
            oi[t] = have ? divisor(I(t)) : -1;
            if (od) od[t] = have ? D(t) : INFINITY;
        }
    }
};

template <int K>
struct reg_list {
    // ascending in slots [K - k, K); slots below hold -inf sentinels that never move, so the
    // worst kept entry is always slot K-1 (a compile-time register) for any runtime k <= K
    double d2[K];
    int32_t id[K];
    int k;

    __device__ __forceinline__ void init(unsigned char *, int k_) { k = k_; }
    __device__ __forceinline__ void reset()
    {
#pragma unroll
        for (int j = 0; j < K; ++j) {
            d2[j] = j < K - k ? -INFINITY : INFINITY;
            id[j] = j < K - k ? -1 : 0x7fffffff;
        }
    }
    __device__ __forceinline__ bool full() const { return d2[K - 1] < INFINITY; }
    __device__ __forceinline__ double worst() const { return d2[K - 1]; }

    __device__ __forceinline__ void insert(double nd, int32_t ni)
    {
        if (!((int)(nd < d2[K - 1]) | ((int)(nd == d2[K - 1]) & (int)(ni < id[K - 1])))) return;
        // branch-free: pos = number of kept entries that precede the new one; slots above pos
        // take their lower neighbour, slot pos takes the new entry
        int pos = 0;
#pragma unroll
        for (int j = 0; j < K - 1; ++j)
            pos += (int)(d2[j] < nd) | ((int)(d2[j] == nd) & (int)(id[j] < ni));
#pragma unroll
        for (int j = K - 1; j > 0; --j) {
            const bool shift = j > pos, here = j == pos;
            d2[j] = shift ? d2[j - 1] : (here ? nd : d2[j]);
            id[j] = shift ? id[j - 1] : (here ? ni : id[j]);
        }
        if (pos == 0) {
            d2[0] = nd;
            id[0] = ni;
        }
    }
    __device__ __forceinline__ void write(int32_t *oi, double *od, const fast_div &divisor)
    {
#pragma unroll
        for (int j = 0; j < K; ++j) {
            int t = j - (K - k);
            if (t >= 0) {
                bool have = d2[j] < INFINITY;
                oi[t] = have ? divisor(id[j]) : -1;
                if (od) od[t] = have ? d2[j] : INFINITY;
            }
        }
    }
};

// distance from coordinate x to the grid's extent along axis c (0 inside), shrunk by the margin
__device__ __forceinline__ double box_gap(const grid_t &g, double x, int c, double margin)
{
    const double lo = g.origin[c], hi = g.origin[c] + g.n[c] * g.cell;
    return fmax(fmax(lo - x, x - hi) - margin, 0.0);
}

template <bool POS_ID, class List>
__device__ __forceinline__ void rank_record(List &L, const double2 &cxy, const double2 &czw, int32_t j,
                                            double px, double py, double pz, bool three_d)
{
    double dx = px - cxy.x, dy = py - cxy.y;
    double s = dx * dx + dy * dy;
    if (three_d) {
        double dz = pz - czw.x;
        s = s + dz * dz;
    }
    L.insert(s, POS_ID ? j : (int32_t)__double_as_longlong(czw.y));
}

// Ranks the records [lo, hi).  The next record is always in flight while the current one is ranked.
// PINGPONG: the loop is unrolled by two with two register sets, so the prefetched record does not
// have to be moved into place every iteration (8 register moves per record otherwise); used for
// the merged ring-0/1 pass, where almost all records are ranked.
template <bool POS_ID, bool PINGPONG, class List>
__device__ __forceinline__ void scan_range(List &L, const double4 *__restrict__ recs, int32_t lo,
                                           int32_t hi, double px, double py, double pz,
                                           bool three_d)
{
    if (lo >= hi) return;
    const double2 *q = reinterpret_cast<const double2 *>(&recs[lo]);  // 2 x LDG.128 per record
    if constexpr (PINGPONG) {
        double2 axy = __ldg(q), azw = __ldg(q + 1), bxy = axy, bzw = azw;
        int32_t j = lo;
        while (true) {
            if (j + 1 < hi) {
                bxy = __ldg(q + 2);
                bzw = __ldg(q + 3);
            }
            rank_record<POS_ID>(L, axy, azw, j, px, py, pz, three_d);
            if (++j >= hi) break;
            if (j + 1 < hi) {
                axy = __ldg(q + 4);
                azw = __ldg(q + 5);
            }
            q += 4;
            rank_record<POS_ID>(L, bxy, bzw, j, px, py, pz, three_d);
            if (++j >= hi) break;
        }
    } else {
        double2 xy = __ldg(q), zw = __ldg(q + 1);
        for (int32_t j = lo; j < hi; ++j) {
            const double2 cxy = xy, czw = zw;
            if (j + 1 < hi) {  // next record in flight while this one is ranked
                q += 2;
                xy = __ldg(q);
                zw = __ldg(q + 1);
            }
            rank_record<POS_ID>(L, cxy, czw, j, px, py, pz, three_d);
        }
    }
}

// SITES = false: `recs` are the point records, the result is the k nearest point ids.
// SITES = true (first pass of the progressive search, GLL-point form): `recs` / `cell_start` are
// the site table -- one record per DISTINCT coordinate (shared GLL nodes are stored up to 8
// times), so the search needs 3-4x fewer distance evaluations and insertions.  The 4 nearest
// sites are found; the k-NN list over the points starts with all copies of site 1 (ids
// ascending), then all copies of site 2, ... as long as the site distances are strictly
// increasing; that prefix is written (up to k entries, idx / divisor, -1 padded).  At the first
// exact tie between site distances the prefix stops (copies of tied sites interleave by id) and
// the caller's full search handles the point.
struct knn_query {
    double px, py, pz;
    int ci[3];
    double out2;  // squared distance from the query to the grid's bounding box (0 inside)
};

// cell coordinates + out-of-box distance of one query; false for a non-finite / overflowing query
__device__ __forceinline__ bool knn_setup_query(const grid_t &g, double px, double py, double pz, double margin,
                                                knn_query &q)
{
    const bool three_d = g.dim == 3;
    q.px = px;
    q.py = py;
    q.pz = pz;
    // a NaN / infinite query (or one so far away that its squared distances overflow) has no nearest
    // neighbours: every comparison is false; without this guard it would walk the whole grid before
    // reporting the same thing
    if (!(fabs(px) <= 1e150 && fabs(py) <= 1e150 && fabs(pz) <= 1e150)) return false;
    q.ci[0] = cell_coord(g, px, 0);
    q.ci[1] = cell_coord(g, py, 1);
    q.ci[2] = three_d ? cell_coord(g, pz, 2) : 0;
    // squared distance from the query to the grid's bounding box (0 for queries inside it): every
    // indexed point is at least that far away, which tightens the termination bound for targets
    // that lie outside the source mesh
    const double p[3] = {px, py, pz};
    double out2 = 0.0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        if (c >= g.dim) break;
        const double o = box_gap(g, p[c], c, margin);
        out2 += o * o;
    }
    q.out2 = out2;
    return true;
}

// rings 0 and 1 merged: the 3 x 3 block of cell rows around the query's cell, each row ONE
// contiguous record range [cx-1, cx+1], own row first, then face, edge neighbours; rows the
// current k-th distance cannot reach are skipped.  (With cells of about one distinct coordinate
// this is where almost every first-pass query ends.)
template <class List, bool SITES>
__device__ __forceinline__ void knn_block_pass(const grid_t &g, const knn_query &q, List &L,
                                               const double4 *__restrict__ recs,
                                               const int32_t *__restrict__ cell_start)
{
    const bool three_d = g.dim == 3;
    const double h = g.cell, margin = h * 1e-6;
    const double px = q.px, py = q.py, pz = q.pz;
    const int xa = max(q.ci[0] - 1, 0), xb = min(q.ci[0] + 1, g.n[0] - 1);
    const int nrow = three_d ? 9 : 3;
    // record range of row `it` (empty when the row lies outside the grid); the bounds of the
    // NEXT row are loaded before the current row is scanned, so their latency is hidden
    auto row_bounds = [&](int it, int32_t &lo, int32_t &hi) {
        const int dy = (int)((0x22161u >> (2 * it)) & 3u) - 1;   // 0,-1,+1, 0, 0,-1,+1,-1,+1
        const int dz = (int)((0x28215u >> (2 * it)) & 3u) - 1;   // 0, 0, 0,-1,+1,-1,-1,+1,+1
        const int yy = q.ci[1] + dy, zz = q.ci[2] + dz;
        lo = hi = 0;
        if (it < nrow && yy >= 0 && yy < g.n[1] && zz >= 0 && zz < g.n[2]) {
            const int base = g.n[0] * (yy + g.n[1] * zz);  // ncells <= MAX_CELLS = 2^26
            lo = cell_start[base + xa];
            hi = cell_start[base + xb + 1];
        }
    };
    int32_t nlo, nhi;
    row_bounds(0, nlo, nhi);
    for (int it = 0; it < nrow; ++it) {
        const int32_t lo = nlo, hi = nhi;
        row_bounds(it + 1, nlo, nhi);
        if (lo >= hi) continue;
        if (L.full()) {
            const int dy = (int)((0x22161u >> (2 * it)) & 3u) - 1;
            const int dz = (int)((0x28215u >> (2 * it)) & 3u) - 1;
            const int yy = q.ci[1] + dy, zz = q.ci[2] + dz;
            double yl = g.origin[1] + yy * h, yh = yl + h;
            double gy = fmax(fmax(yl - py, py - yh) - margin, 0.0);
            double gz = 0.0;
            if (three_d) {
                double zl = g.origin[2] + zz * h, zh = zl + h;
                gz = fmax(fmax(zl - pz, pz - zh) - margin, 0.0);
            }
            if (L.worst() - (gy * gy + gz * gz) < 0.0) continue;
        }
        scan_range<SITES, true>(L, recs, lo, hi, px, py, pz, three_d);
    }
}

// rings r = r0, r0 + 1, ... around the query's cell until the k-th distance bounds every unvisited cell
// (r0 = 1 after knn_block_pass: ring 1 itself is done, only its termination test remains)
template <class List, bool SITES>
__device__ __forceinline__ void knn_ring_search(const grid_t &g, const knn_query &q, List &L,
                                                const double4 *__restrict__ recs,
                                                const int32_t *__restrict__ cell_start, int r0)
{
    const bool three_d = g.dim == 3;
    const double h = g.cell, margin = h * 1e-6;
    const double px = q.px, py = q.py, pz = q.pz;
    const double p[3] = {px, py, pz};
    const int *ci = q.ci;
    const double out2 = q.out2;
    for (int r = r0;; ++r) {
        const int zlo = max(ci[2] - r, 0), zhi = min(ci[2] + r, g.n[2] - 1);
        const int ylo = max(ci[1] - r, 0), yhi = min(ci[1] + r, g.n[1] - 1);
        const int xlo = max(ci[0] - r, 0), xhi = min(ci[0] + r, g.n[0] - 1);
        // rows of the shell are visited nearest-first (offsets 0, -1, +1, -2, +2, ...), so the
        // list tightens early and the farther rows are pruned
        const int nz = (MM_KNN_MERGED && r == 1) ? 0 : (three_d ? 2 * r + 1 : 1);  // r = 1: done by the block pass
        for (int iz = 0; iz < nz; ++iz) {
            const int zz = ci[2] + ((iz & 1) ? -((iz + 1) >> 1) : ((iz + 1) >> 1));
            if (zz < zlo || zz > zhi) continue;
            double gz = 0.0;
            if (three_d) {
                double zl = g.origin[2] + zz * h, zh = zl + h;
                gz = fmax(fmax(zl - pz, pz - zh) - margin, 0.0);
            }
            const bool zedge = three_d && (abs(zz - ci[2]) == r);
            for (int iy = 0; iy < 2 * r + 1; ++iy) {
                const int yy = ci[1] + ((iy & 1) ? -((iy + 1) >> 1) : ((iy + 1) >> 1));
                if (yy < ylo || yy > yhi) continue;
                double yl = g.origin[1] + yy * h, yh = yl + h;
                double gy = fmax(fmax(yl - py, py - yh) - margin, 0.0);
                const double g2 = gy * gy + gz * gz;
                int xa = xlo, xb = xhi;
                if (L.full()) {
                    // rows farther than the current k-th neighbour cannot contribute, and
                    // inside a row only the cells that the k-th-neighbour ball reaches can
                    const double w2 = L.worst() - g2;
                    if (w2 < 0.0) continue;
                    if (!SITES || r > 1) {  // site rows are short: the x restriction costs more
                        const double wx = sqrt(w2) + margin;
                        xa = max(xa, cell_coord(g, px - wx, 0));
                        xb = min(xb, cell_coord(g, px + wx, 0));
                        if (xa > xb) continue;
                    }
                }
                const int base = g.n[0] * (yy + g.n[1] * zz);  // ncells <= MAX_CELLS = 2^26
                if (zedge || abs(yy - ci[1]) == r) {
                    scan_range<SITES, false>(L, recs, cell_start[base + xa], cell_start[base + xb + 1], px,
                                      py, pz, three_d);
                } else {  // interior row of the shell: only its two end cells are new
                    const int x0 = ci[0] - r, x1 = ci[0] + r;
                    if (x0 >= xa && x0 <= xb)
                        scan_range<SITES, false>(L, recs, cell_start[base + x0], cell_start[base + x0 + 1],
                                          px, py, pz, three_d);
                    if (r > 0 && x1 >= xa && x1 <= xb)
                        scan_range<SITES, false>(L, recs, cell_start[base + x1], cell_start[base + x1 + 1],
                                          px, py, pz, three_d);
                }
            }
        }
        // every point not yet visited lies outside the block of cells [ci-r, ci+r]; its
        // distance to p is at least the gap to the nearest block face that has cells beyond it
        double bound = INFINITY;
        bool remaining = false;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            if (c >= g.dim) break;
            if (ci[c] - r > 0) {
                remaining = true;
                bound = fmin(bound, p[c] - (g.origin[c] + (ci[c] - r) * h));
            }
            if (ci[c] + r < g.n[c] - 1) {
                remaining = true;
                bound = fmin(bound, (g.origin[c] + (ci[c] + r + 1) * h) - p[c]);
            }
        }
        if (!remaining) break;
        if (!L.full()) continue;
        bound = fmax(bound - margin, 0.0);
        if (L.worst() < bound * bound) break;
        if (out2 > 0.0) {
            // query outside the grid's box: a point beyond the face of axis c is also at least
            // the out-of-box distance away along the other axes
            double bound2 = INFINITY;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                if (c >= g.dim) break;
                const double o = box_gap(g, p[c], c, margin);
                const double other = out2 - o * o;
                if (ci[c] - r > 0) {
                    const double gap = fmax(p[c] - (g.origin[c] + (ci[c] - r) * h) - margin, 0.0);
                    bound2 = fmin(bound2, gap * gap + other);
                }
                if (ci[c] + r < g.n[c] - 1) {
                    const double gap = fmax((g.origin[c] + (ci[c] + r + 1) * h) - p[c] - margin, 0.0);
                    bound2 = fmin(bound2, gap * gap + other);
                }
            }
            if (L.worst() < bound2) break;
        }
    }
}

// writes the result row of one query: the site prefix expanded to point copies (SITES) or the list itself
template <class List, bool SITES>
__device__ __forceinline__ void knn_emit(List &L, int k, const fast_div &divisor,
                                         const double4 *__restrict__ recs,
                                         const double4 *__restrict__ point_recs, int32_t *__restrict__ o,
                                         double *__restrict__ od)
{
    if constexpr (SITES) {
        // expand: copies of site j continue the prefix only while d2[j] < d2[j+1] strictly
        int c = 0;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const bool have = L.d2[j] < INFINITY;
            const bool strict = L.d2[j] < L.d2[j + 1];  // slot j+1 is +inf without such a site
            if (!(have && strict)) break;
            const int32_t r0 = (int32_t)__double_as_longlong(recs[L.id[j]].w);
            const int32_t r1 = (int32_t)__double_as_longlong(recs[L.id[j] + 1].w);
            for (int32_t t = r0; t < r1 && c < k; ++t)
                o[c++] = divisor((int32_t)__double_as_longlong(point_recs[t].w));
        }
        for (; c < k; ++c) o[c] = -1;
    } else {
        L.write(o, od, divisor);
    }
}

template <class List, bool SITES>
__device__ __forceinline__ void
knn_body(const grid_t &g, int64_t N, const double *__restrict__ pts, int pstride, int k,
         const fast_div &divisor,
         const double4 *__restrict__ recs, const int32_t *__restrict__ cell_start,
         int32_t *__restrict__ out_idx, double *__restrict__ out_d2,
         const double4 *__restrict__ point_recs, const long long *__restrict__ n_dev, int64_t n_off)
{
    extern __shared__ __align__(16) unsigned char smem[];
    if (n_dev) {  // point count known only on the device (re-run of the unresolved points): [n_off, *n_dev)
        const long long have = *n_dev - n_off;
        N = have < 0 ? 0 : (have < N ? have : N);
    }
    List L;
    L.init(smem, SITES ? 4 : k);
    const bool three_d = g.dim == 3;
    const double margin = g.cell * 1e-6;

    for (int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; n < N;
         n += (int64_t)gridDim.x * blockDim.x) {
        knn_query q;
        if (!knn_setup_query(g, pts[n * pstride + 0], pts[n * pstride + 1], three_d ? pts[n * pstride + 2] : 0.0,
                             margin, q)) {
            for (int t = 0; t < k; ++t) {
                out_idx[n * k + t] = -1;
                if (out_d2) out_d2[n * k + t] = INFINITY;
            }
            continue;
        }
        L.reset();
        if (MM_KNN_MERGED) knn_block_pass<List, SITES>(g, q, L, recs, cell_start);
        knn_ring_search<List, SITES>(g, q, L, recs, cell_start, MM_KNN_MERGED ? 1 : 0);
        knn_emit<List, SITES>(L, k, divisor, recs, point_recs, out_idx + n * k, out_d2 ? out_d2 + n * k : nullptr);
    }
}

// ================================================================================================
// First pass of the pipeline, k' <= 4 (site table of the GLL-point form, or centroid records):
// WARP-COOPERATIVE block scan with an fp32 pre-filter.
//
// The queries arrive sorted by index cell (mm_index_sort_queries), so the 32 queries of a warp sit in
// a short run of cells of one cell row and their 3 x 3 x 3 neighbourhoods overlap almost completely.
// Per run ("segment" = lanes of one cell row whose cells span at most XSPAN cells):
//   1. the warp stages the union of the neighbourhoods ONCE, from the index's compact fp32 table (16-byte
//      records {x, y, z relative to the record's own cell, position}; half the bytes of the binary64 records
//      and no conversions): the 9 (3 in 2-D) cell rows x the cells [x_first - 1, x_last + 1].  The staging
//      buffer is COLUMN-major -- all records of the 9 cells that share an x index are contiguous -- so the
//      3 x 3 x 3 neighbourhood of a query is ONE contiguous range of the buffer;
//   2. every lane scans its range in fp32 and keeps the 5 smallest keys (d2 bits truncated to 24 bits |
//      8-bit slot number: one 32-bit min/max network, no fp64, no global loads, no branches, one flat loop);
//   3. if the 5th key exceeds the 4th by more than the fp32 error bound, the SET of the 4 nearest records
//      is certain; those 4 are evaluated exactly in binary64 and ordered by the canonical (d2, id) order.
//      With ties or near-ties around the 4th distance a second fp32 sweep ranks exactly, in binary64, every
//      record whose fp32 distance does not exceed the 4th's by more than the error bound (a handful).
//      Segments with more than CAP records and lone lanes take the exact per-thread block pass.  Either way the
//      list then equals the exact top-4 of the 3^3 block, and the common termination test / outer rings follow.
//      Results are bit-identical to knn_kernel.
//
// fp32 error bound (h = cell size, h32 = fl32(h)).  A staged coordinate is X = fma(c, h32, rel) with c the
// cell offset inside the segment (|c| <= XSPAN + 1 = 13 in x, <= 1 in y / z) and rel the cell-relative coordinate
// (|rel| <= h, fp32 rounding 2^-24 h): |X - exact| <= 2^-24 (14 h) + 13 * 2^-24 h + 2^-24 h = 1.7e-6 h in x and
// <= 2.4e-7 h in y, z; the query's coordinates are formed the same way.  dx (|dx| <= 3 h): error <= 3.6e-6 h,
// dy, dz (|.| <= 2 h): <= 6e-7 h.  |dx32^2 - dx^2| <= 2 * 3 h * 3.6e-6 h = 2.2e-5 h^2, dy, dz: 2.4e-6 h^2 each,
// the three roundings of the sum <= 3 * 2^-24 * 17 h^2 = 3e-6 h^2: total <= 3e-5 h^2; KNN_EPS = 1e-4 h^2 is used.
// Truncating the key to 24 bits under-states d2 by at most 2^-15 relative.  The binary64 reference distances
// carry ~1e-16 relative error, far below the margin.
// ================================================================================================
constexpr int KB_WARPS = 4;          // warps per CTA
constexpr int KB_CAP = 256;          // staged records per segment (8-bit slot number in the key)
constexpr int KB_XSPAN = 12;         // cells of one row a segment may span
constexpr int KB_ROWLEN = 16;        // staged cell_start entries per row: (XSPAN + 2) cells + 1, rounded up
constexpr int KB_MIN_MEMBERS = 3;    // smaller segments are cheaper on the per-thread path

struct kb_smem {
    float4 rec[KB_CAP];
    int32_t cs[9][KB_ROWLEN];        // cell_start of the staged cells, per row
    int32_t dst[9][KB_ROWLEN];       // slot of the first record of cell (row, column)
    int32_t colstart[KB_ROWLEN];     // first slot of each column (+ total)
};

template <bool SITES>
__global__ void __launch_bounds__(KB_WARPS * 32, 8)
knn_block_kernel(grid_t g, int64_t N, const double *__restrict__ pts, int pstride, int k,
                 const fast_div divisor, const double4 *__restrict__ recs,
                 const int32_t *__restrict__ cell_start, const float4 *__restrict__ recf,
                 const int32_t *__restrict__ rec_id, const int32_t *__restrict__ site_first,
                 int32_t *__restrict__ out_idx)
{
    __shared__ kb_smem sm_all[KB_WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    kb_smem &sm = sm_all[warp];
    const bool three_d = g.dim == 3;
    const int nrow = three_d ? 9 : 3;
    const double h = g.cell, margin = h * 1e-6;
    const float h32 = (float)h;
    const float eps = (float)(1e-4 * h * h);
    using List = reg_list<4>;
    List L;
    L.init(nullptr, 4);

    const int64_t warps_total = (int64_t)gridDim.x * KB_WARPS;
    for (int64_t batch = (int64_t)blockIdx.x * KB_WARPS + warp; batch * 32 < N; batch += warps_total) {
        const int64_t n = batch * 32 + lane;
        const bool valid = n < N;
        knn_query q;
        bool pending = false;
        if (valid) {
            pending = knn_setup_query(g, pts[n * pstride + 0], pts[n * pstride + 1],
                                      three_d ? pts[n * pstride + 2] : 0.0, margin, q);
            if (!pending)
                for (int t = 0; t < k; ++t) out_idx[n * k + t] = -1;
        }
        if (!pending) q.ci[0] = q.ci[1] = q.ci[2] = -1;
        // the fp32 pass needs small relative coordinates: queries inside the grid's box only
        bool eligible = pending && q.out2 == 0.0;
        bool fast_done = false;
        L.reset();

        while (true) {
            const unsigned cand = __ballot_sync(0xffffffffu, eligible);
            if (!cand) break;
            const int leader = __ffs(cand) - 1;
            const int lcx = __shfl_sync(0xffffffffu, q.ci[0], leader);
            const int lcy = __shfl_sync(0xffffffffu, q.ci[1], leader);
            const int lcz = __shfl_sync(0xffffffffu, q.ci[2], leader);
            // sorted queries: the leader has the smallest cell of its row among the lanes still eligible
            const bool member = eligible && q.ci[1] == lcy && q.ci[2] == lcz && q.ci[0] >= lcx &&
                                q.ci[0] < lcx + KB_XSPAN;
            const unsigned mmask = __ballot_sync(0xffffffffu, member);
            eligible = eligible && !member;  // every lane joins at most one segment
            if (__popc(mmask) < KB_MIN_MEMBERS) continue;
            const int xhi = __reduce_max_sync(0xffffffffu, member ? q.ci[0] : lcx);
            const int xa = max(lcx - 1, 0), xb = min(xhi + 1, g.n[0] - 1);
            const int ncol = xb - xa + 1;  // staged cells per row, <= XSPAN + 2
            // 1a. cell_start rows: entries [xa, xb + 1] of each row inside the grid, zeros otherwise
#pragma unroll
            for (int i = 0; i < (9 * KB_ROWLEN + 31) / 32; ++i) {
                const int t = lane + 32 * i;
                const int r = t / KB_ROWLEN, c = t % KB_ROWLEN;
                if (r < nrow) {
                    const int dy = (int)((0x22161u >> (2 * r)) & 3u) - 1;
                    const int dz = (int)((0x28215u >> (2 * r)) & 3u) - 1;
                    const int yy = lcy + dy, zz = lcz + dz;
                    int32_t v = 0;
                    if (yy >= 0 && yy < g.n[1] && zz >= 0 && zz < g.n[2])
                        v = __ldg(&cell_start[g.n[0] * (yy + g.n[1] * zz) + xa + min(c, ncol)]);
                    sm.cs[r][c] = v;
                }
            }
            __syncwarp();
            // 1b. column-major slots: lane c owns column c
            {
                int size = 0;
                if (lane < ncol)
                    for (int r = 0; r < nrow; ++r) size += sm.cs[r][lane + 1] - sm.cs[r][lane];
                int off = size;
#pragma unroll
                for (int o = 1; o < 16; o <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, off, o);
                    if (lane >= o) off += t;
                }
                if (lane < KB_ROWLEN) sm.colstart[lane] = off - size;  // exclusive; colstart[ncol] = total
                if (lane < ncol) {
                    int at = off - size;
                    for (int r = 0; r < nrow; ++r) {
                        sm.dst[r][lane] = at;
                        at += sm.cs[r][lane + 1] - sm.cs[r][lane];
                    }
                }
            }
            __syncwarp();
            const int total = sm.colstart[ncol];
            if (total > KB_CAP) continue;  // crowded cells: per-thread path
            // 1c. records: one (row, column) cell per task; coordinates move from the cell's frame to the segment's
            //     (corner of cell (xa, lcy, lcz)) with one fma each
#pragma unroll
            for (int i = 0; i < (9 * KB_ROWLEN + 31) / 32; ++i) {
                const int t = lane + 32 * i;
                const int r = t / KB_ROWLEN, c = t % KB_ROWLEN;
                if (r < nrow && c < ncol) {
                    const int32_t lo = sm.cs[r][c], cnt = sm.cs[r][c + 1] - lo, at = sm.dst[r][c];
                    const float fx = (float)c;
                    const float fy = (float)((int)((0x22161u >> (2 * r)) & 3u) - 1);
                    const float fz = (float)((int)((0x28215u >> (2 * r)) & 3u) - 1);
                    for (int j = 0; j < cnt; ++j) {
                        float4 f = __ldg(&recf[lo + j]);
                        f.x = fmaf(fx, h32, f.x);
                        f.y = fmaf(fy, h32, f.y);
                        f.z = fmaf(fz, h32, f.z);
                        sm.rec[at + j] = f;
                    }
                }
            }
            __syncwarp();
            // 2. fp32 scan of this lane's 3 columns (one contiguous range): five smallest keys
            if (member) {
                const double ox = g.origin[0] + q.ci[0] * h, oy = g.origin[1] + lcy * h, oz = g.origin[2] + lcz * h;
                const float qx = fmaf((float)(q.ci[0] - xa), h32, (float)(q.px - ox));
                const float qy = (float)(q.py - oy);
                const float qz = three_d ? (float)(q.pz - oz) : 0.0f;
                unsigned a0 = 0xffffffffu, a1 = a0, a2 = a0, a3 = a0, a4 = a0;
                const int lo = sm.colstart[max(q.ci[0] - 1, 0) - xa];
                const int hi = sm.colstart[min(q.ci[0] + 1, g.n[0] - 1) + 1 - xa];
                for (int j = lo; j < hi; ++j) {
                    const float4 s = sm.rec[j];
                    const float dx = qx - s.x, dy = qy - s.y, dz = qz - s.z;
                    const float d = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                    const unsigned key = (__float_as_uint(d) & 0xffffff00u) | (unsigned)j;
                    // sorted insertion, all five updates independent of one another
                    const unsigned n4 = max(a3, min(a4, key)), n3 = max(a2, min(a3, key));
                    const unsigned n2 = max(a1, min(a2, key)), n1 = max(a0, min(a1, key));
                    a0 = min(a0, key);
                    a1 = n1;
                    a2 = n2;
                    a3 = n3;
                    a4 = n4;
                }
                // 3. is the set of the 4 nearest certain?  (fewer than 5 records in the block: it is all of them)
                bool certain = true;
                if (a4 != 0xffffffffu) {
                    const float d4 = __uint_as_float(a3 & 0xffffff00u), d5 = __uint_as_float(a4 & 0xffffff00u);
                    certain = d5 - eps > d4 * (1.0f + 6.2e-5f) + eps;  // 2^-15 truncation + fp32 error both ways
                }
                if (certain) {
                    const unsigned keys[4] = {a0, a1, a2, a3};
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        if (keys[t] == 0xffffffffu) break;
                        const int32_t pos = __float_as_int(sm.rec[keys[t] & 0xffu].w);
                        const double2 *src = reinterpret_cast<const double2 *>(&recs[pos]);
                        rank_record<SITES>(L, __ldg(src), __ldg(src + 1), pos, q.px, q.py, q.pz, three_d);
                    }
                } else {
                    // ties / near-ties around the 4th distance (structured meshes produce exact ones): every record
                    // that can still belong to the 4 nearest has an fp32 distance <= cut; rank exactly those
                    // (typically 5-8 of ~34) in binary64 -- the canonical (d2, id) order settles the ties
                    const float cut = __uint_as_float(a3 & 0xffffff00u) * (1.0f + 6.2e-5f) + 2.0f * eps;
                    for (int j = lo; j < hi; ++j) {
                        const float4 s = sm.rec[j];
                        const float dx = qx - s.x, dy = qy - s.y, dz = qz - s.z;
                        const float d = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                        if (d <= cut) {
                            const int32_t pos = __float_as_int(s.w);
                            const double2 *src = reinterpret_cast<const double2 *>(&recs[pos]);
                            rank_record<SITES>(L, __ldg(src), __ldg(src + 1), pos, q.px, q.py, q.pz, three_d);
                        }
                    }
                }
                fast_done = true;
            }
            __syncwarp();  // the staging buffer is re-used by the next segment
        }
        if (pending) {
            if (!fast_done) knn_block_pass<List, SITES>(g, q, L, recs, cell_start);
            knn_ring_search<List, SITES>(g, q, L, recs, cell_start, 1);
            int32_t *o = out_idx + n * k;
            if constexpr (SITES) {
                // expand: copies of site j continue the prefix only while d2[j] < d2[j+1] strictly
                int c = 0;
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    const bool have = L.d2[j] < INFINITY;
                    const bool strict = L.d2[j] < L.d2[j + 1];  // slot j+1 is +inf without such a site
                    if (!(have && strict)) break;
                    const int32_t r0 = __ldg(&site_first[L.id[j]]), r1 = __ldg(&site_first[L.id[j] + 1]);
                    for (int32_t t = r0; t < r1 && c < k; ++t) o[c++] = divisor(__ldg(&rec_id[t]));
                }
                for (; c < k; ++c) o[c] = -1;
            } else {
                L.write(o, nullptr, divisor);
            }
        }
    }
}

// ---- compact tables of the block kernel ----------------------------------------------------------
// plain form: recf[t] = {cell-relative coordinates of record t, t}, rec_id[t] = point id
__global__ void __launch_bounds__(256)
recf_plain_kernel(grid_t g, int64_t M, const double4 *__restrict__ recs, float4 *__restrict__ recf,
                  int32_t *__restrict__ rec_id)
{
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < M; t += (int64_t)gridDim.x * blockDim.x) {
        const double4 r = recs[t];
        const double p[3] = {r.x, r.y, r.z};
        float4 f;
        f.x = (float)(r.x - (g.origin[0] + cell_coord(g, p[0], 0) * g.cell));
        f.y = (float)(r.y - (g.origin[1] + cell_coord(g, p[1], 1) * g.cell));
        f.z = g.dim == 3 ? (float)(r.z - (g.origin[2] + cell_coord(g, p[2], 2) * g.cell)) : 0.0f;
        f.w = __int_as_float((int32_t)t);
        recf[t] = f;
        rec_id[t] = (int32_t)__double_as_longlong(r.w);
    }
}

// site form: recf[s] = {cell-relative coordinates of site s, s}, site_first[s] = first record of site s
__global__ void __launch_bounds__(256)
recf_sites_kernel(grid_t g, int64_t nsites, const double4 *__restrict__ site_recs, float4 *__restrict__ recf,
                  int32_t *__restrict__ site_first)
{
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t <= nsites;
         t += (int64_t)gridDim.x * blockDim.x) {
        const double4 r = site_recs[t];
        site_first[t] = (int32_t)__double_as_longlong(r.w);
        if (t == nsites) break;  // sentinel: only its first-record entry is meaningful
        const double p[3] = {r.x, r.y, r.z};
        float4 f;
        f.x = (float)(r.x - (g.origin[0] + cell_coord(g, p[0], 0) * g.cell));
        f.y = (float)(r.y - (g.origin[1] + cell_coord(g, p[1], 1) * g.cell));
        f.z = g.dim == 3 ? (float)(r.z - (g.origin[2] + cell_coord(g, p[2], 2) * g.cell)) : 0.0f;
        f.w = __int_as_float((int32_t)t);
        recf[t] = f;
    }
}

__global__ void __launch_bounds__(256)
rec_id_kernel(int64_t M, const double4 *__restrict__ recs, int32_t *__restrict__ rec_id)
{
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < M; t += (int64_t)gridDim.x * blockDim.x)
        rec_id[t] = (int32_t)__double_as_longlong(recs[t].w);
}

template <class List>
__global__ void __launch_bounds__(KNN_BLOCK)
knn_kernel(grid_t g, int64_t N, const double *__restrict__ pts, int pstride, int k,
           const fast_div divisor,
           const double4 *__restrict__ recs, const int32_t *__restrict__ cell_start,
           int32_t *__restrict__ out_idx, double *__restrict__ out_d2,
           const long long *__restrict__ n_dev, int64_t n_off)
{
    knn_body<List, false>(g, N, pts, pstride, k, divisor, recs, cell_start, out_idx, out_d2, nullptr, n_dev, n_off);
}

// site pass: 64 registers so that 8 blocks of 128 threads stay resident per SM
__global__ void __launch_bounds__(KNN_BLOCK, 8)
knn_sites_kernel(grid_t g, int64_t N, const double *__restrict__ pts, int pstride, int k,
                 const fast_div divisor,
                 const double4 *__restrict__ site_recs, const int32_t *__restrict__ site_cell_start,
                 int32_t *__restrict__ out_idx, const double4 *__restrict__ point_recs)
{
    knn_body<reg_list<4>, true>(g, N, pts, pstride, k, divisor, site_recs, site_cell_start, out_idx, nullptr,
                                point_recs, nullptr, 0);
}

// ---- counting sort of QUERY points by index cell (coherent warps in K1-K3) -----------------------
// pass 1: histogram of the query cells; the value the atomic returns is the query's rank inside its
// cell, so the placement pass needs no second round of atomics
__global__ void __launch_bounds__(256)
query_rank_kernel(grid_t g, int64_t N, const double *__restrict__ pts, int32_t *__restrict__ counts,
                  int32_t *__restrict__ rank)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < N;
         i += (int64_t)gridDim.x * blockDim.x)
        rank[i] = atomicAdd(&counts[cell_of(g, pts + i * g.dim)], 1);
}

// pass 2 (after the exclusive scan): query i goes to record start[cell] + rank[i].  A record is 32
// bytes {x, y, z (0 in 2-D), i in the low 32 bits of the fourth lane} written with one aligned
// 32-byte store: a scattered 24-byte row plus a scattered 4-byte permutation entry would be two
// partial-sector writes per point
__global__ void __launch_bounds__(256)
query_place_kernel(grid_t g, int64_t N, const double *__restrict__ pts,
                   const int32_t *__restrict__ start, const int32_t *__restrict__ rank,
                   double *__restrict__ sorted_rec)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < N;
         i += (int64_t)gridDim.x * blockDim.x) {
        const double *p = pts + i * g.dim;
        const int64_t c = cell_of(g, p);
        const int64_t pos = (int64_t)start[c] + rank[i];
        const double z = g.dim == 3 ? p[2] : 0.0;
        const double w = __longlong_as_double((long long)i);
        asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(sorted_rec + pos * MM_QREC),
                     "d"(p[0]), "d"(p[1]), "d"(z), "d"(w)
                     : "memory");
    }
}

// ---- site table ---------------------------------------------------------------------------------
__device__ __forceinline__ bool rec_less(const double4 &a, const double4 &b)
{
    if (a.x != b.x) return a.x < b.x;
    if (a.y != b.y) return a.y < b.y;
    if (a.z != b.z) return a.z < b.z;
    return __double_as_longlong(a.w) < __double_as_longlong(b.w);
}

// one thread per cell: insertion sort of the cell's records by (x, y, z, id) (cells are small),
// then count the distinct coordinates
__global__ void __launch_bounds__(128)
site_sort_kernel(int64_t ncells, const int32_t *__restrict__ cell_start, double4 *__restrict__ recs,
                 int32_t *__restrict__ sites_in_cell)
{
    for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < ncells;
         c += (int64_t)gridDim.x * blockDim.x) {
        const int32_t lo = cell_start[c], hi = cell_start[c + 1];
        for (int32_t i = lo + 1; i < hi; ++i) {
            const double4 v = recs[i];
            int32_t j = i;
            while (j > lo) {
                const double4 u = recs[j - 1];
                if (!rec_less(v, u)) break;
                recs[j] = u;
                --j;
            }
            recs[j] = v;
        }
        int32_t ns = 0;
        for (int32_t i = lo; i < hi; ++i) {
            bool head = i == lo;
            if (!head) {
                const double4 a = recs[i - 1], b = recs[i];
                head = a.x != b.x || a.y != b.y || a.z != b.z;
            }
            ns += head;
        }
        sites_in_cell[c] = ns;
    }
}

__global__ void __launch_bounds__(128)
site_fill_kernel(int64_t ncells, const int32_t *__restrict__ cell_start,
                 const double4 *__restrict__ recs, const int32_t *__restrict__ site_cell_start,
                 double4 *__restrict__ site_recs, int64_t nsites, int32_t M)
{
    for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < ncells;
         c += (int64_t)gridDim.x * blockDim.x) {
        const int32_t lo = cell_start[c], hi = cell_start[c + 1];
        int32_t s = site_cell_start[c];
        for (int32_t i = lo; i < hi; ++i) {
            const double4 b = recs[i];
            bool head = i == lo;
            if (!head) {
                const double4 a = recs[i - 1];
                head = a.x != b.x || a.y != b.y || a.z != b.z;
            }
            if (head) {
                double4 r = b;
                r.w = __longlong_as_double((long long)i);
                site_recs[s++] = r;
            }
        }
        if (c == ncells - 1) {
            double4 r;
            r.x = r.y = r.z = 0.0;
            r.w = __longlong_as_double((long long)M);
            site_recs[nsites] = r;
        }
    }
}

grid_t grid_of(const mm_index *ix)
{
    grid_t g{};
    g.dim = ix->dim;
    for (int c = 0; c < 3; ++c) {
        g.origin[c] = ix->origin[c];
        g.n[c] = ix->n[c];
    }
    g.cell = ix->cell;
    g.inv_cell = ix->inv_cell;
    return g;
}

int launch_blocks(int64_t work, int block, int per_sm)
{
    int sms = mm_num_sms();
    int64_t need = (work + block - 1) / block;
    int64_t cap = (int64_t)(sms > 0 ? sms : 148) * per_sm;
    return (int)(need < 1 ? 1 : (need > cap ? cap : need));
}

void choose_dims(const double ext[3], int dim, double h, int n[3])
{
    for (int c = 0; c < 3; ++c) {
        n[c] = 1;
        if (c < dim && ext[c] > 0.0) {
            double q = std::ceil(ext[c] / h);
            if (!(q >= 1.0)) q = 1.0;
            if (q > 4096.0) q = 4096.0;
            n[c] = (int)q;
        }
    }
}

}  // namespace

extern "C" int mm_index_destroy(mm_index_t *ix)
{
    if (!ix) return MM_OK;
    // freed on the allocating stream so that the pool can hand the blocks straight back to the
    // next build on that stream (a cross-stream free makes the pool grow instead)
    if (ix->recs) cudaFreeAsync(ix->recs, ix->stream);
    if (ix->cell_start) cudaFreeAsync(ix->cell_start, ix->stream);
    if (ix->site_recs) cudaFreeAsync(ix->site_recs, ix->stream);
    if (ix->site_cell_start) cudaFreeAsync(ix->site_cell_start, ix->stream);
    if (ix->recf) cudaFreeAsync(ix->recf, ix->stream);
    if (ix->rec_id) cudaFreeAsync(ix->rec_id, ix->stream);
    if (ix->site_first) cudaFreeAsync(ix->site_first, ix->stream);
    delete ix;
    return MM_OK;
}

extern "C" int mm_index_info(const mm_index_t *ix, int64_t info[8], double *cell_size)
{
    MM_REQUIRE(ix && info, MM_ERR_INVALID, "mm_index_info: null");
    info[0] = ix->M;
    info[1] = ix->dim;
    info[2] = ix->n[0];
    info[3] = ix->n[1];
    info[4] = ix->n[2];
    info[5] = ix->nonempty;
    info[6] = (int64_t)ix->bytes;
    info[7] = 0;
    if (cell_size) *cell_size = ix->cell;
    return MM_OK;
}

extern "C" int mm_index_create(mm_index_t **out, int dim, int64_t M, const double *points,
                               void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    MM_REQUIRE(out, MM_ERR_INVALID, "mm_index_create: null out");
    *out = nullptr;
    MM_REQUIRE(dim == 2 || dim == 3, MM_ERR_INVALID, "mm_index_create: dim %d", dim);
    MM_REQUIRE(M >= 0 && M < ((int64_t)1 << 31), MM_ERR_INVALID,
               "mm_index_create: M=%lld outside [0, 2^31)", (long long)M);
    MM_REQUIRE(M == 0 || points, MM_ERR_INVALID, "mm_index_create: null points");

    mm_index *ix = new mm_index();
    ix->dim = dim;
    ix->M = M;
    ix->stream = stream;
    struct guard_t {
        mm_index *p;
        ~guard_t() { if (p) mm_index_destroy(p); }
    } guard{ix};

    grid_t g{};
    g.dim = dim;
    g.n[0] = g.n[1] = g.n[2] = 1;
    g.cell = g.inv_cell = 1.0;
    int32_t *counts = nullptr;
    int32_t *tile_sums = nullptr;
    unsigned long long *d_nonempty = nullptr;
    double *d_partial = nullptr;
    struct scratch_t {
        int32_t *&a;
        int32_t *&b;
        unsigned long long *&c;
        double *&d;
        cudaStream_t st;
        ~scratch_t()
        {
            if (a) cudaFreeAsync(a, st);
            if (b) cudaFreeAsync(b, st);
            if (c) cudaFreeAsync(c, st);
            if (d) cudaFreeAsync(d, st);
        }
    } scratch{counts, tile_sums, d_nonempty, d_partial, stream};

    if (M > 0) {
        // 1. bounding box
        int nb = launch_blocks(M, 256, 8);
        MM_CUDA(pool_alloc((void **)&d_partial, sizeof(double) * 6 * nb, stream));
        bbox_kernel<<<nb, 256, 0, stream>>>(dim, M, points, d_partial);
        MM_CUDA(cudaGetLastError());
        std::vector<double> part(6 * (size_t)nb);
        MM_CUDA(cudaMemcpyAsync(part.data(), d_partial, sizeof(double) * 6 * nb,
                                cudaMemcpyDeviceToHost, stream));
        MM_CUDA(cudaStreamSynchronize(stream));
        double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
        for (int b = 0; b < nb; ++b)
            for (int c = 0; c < 3; ++c) {
                lo[c] = std::fmin(lo[c], part[b * 6 + c]);
                hi[c] = std::fmax(hi[c], part[b * 6 + 3 + c]);
            }
        double ext[3] = {0, 0, 0};
        int nd = 0;
        double vol = 1.0;
        for (int c = 0; c < dim; ++c) {
            MM_REQUIRE(std::isfinite(lo[c]) && std::isfinite(hi[c]), MM_ERR_INVALID,
                       "mm_index_create: non-finite coordinates");
            g.origin[c] = lo[c];
            ext[c] = hi[c] - lo[c];
            if (ext[c] > 0.0) {
                ++nd;
                vol *= ext[c];
            }
        }
        // 2. cell size.  Target: ~1.25 DISTINCT coordinates per cell by volume (measured optimum of the
        //    ring search on lattice-like and random data, tools/exp_cell.py) -- the GLL-point form
        //    stores shared nodes 2-8 times and duplicates never separate, so the number of distinct
        //    coordinates is first estimated as the number of non-empty cells of a probe grid twice as
        //    fine as "2 points per cell".  Then refine while cells stay crowded and halving the cell
        //    still separates points (clustered data).
        double h = 1.0;
        double emax = std::max(ext[0], std::max(ext[1], ext[2]));
        if (nd > 0) {
            h = std::pow(vol * 2.0 / (double)M, 1.0 / nd);
            if (!(h > 0.0) || !std::isfinite(h)) h = emax;
            h = std::max(h, emax / 4096.0);
        }
        MM_CUDA(pool_alloc((void **)&d_nonempty, sizeof(unsigned long long), stream));
        auto evaluate = [&](double hh, int64_t *nonempty) -> int {
            choose_dims(ext, dim, hh, g.n);
            g.cell = hh;
            g.inv_cell = 1.0 / hh;
            int64_t ncells = (int64_t)g.n[0] * g.n[1] * g.n[2];
            if (counts) { cudaFreeAsync(counts, stream); counts = nullptr; }
            MM_CUDA(pool_alloc((void **)&counts, sizeof(int32_t) * (size_t)(ncells + 1), stream));
            MM_CUDA(cudaMemsetAsync(counts, 0, sizeof(int32_t) * (size_t)(ncells + 1), stream));
            MM_CUDA(cudaMemsetAsync(d_nonempty, 0, sizeof(unsigned long long), stream));
            histogram_kernel<<<launch_blocks(M, 256, 8), 256, 0, stream>>>(g, M, points, counts);
            count_nonempty_kernel<<<launch_blocks(ncells, 256, 8), 256, 0, stream>>>(
                ncells, counts, d_nonempty);
            MM_CUDA(cudaGetLastError());
            unsigned long long ne = 0;
            MM_CUDA(cudaMemcpyAsync(&ne, d_nonempty, sizeof ne, cudaMemcpyDeviceToHost, stream));
            MM_CUDA(cudaStreamSynchronize(stream));
            *nonempty = (int64_t)ne;
            return MM_OK;
        };
        auto cells_at = [&](double hh) {
            int nn[3];
            choose_dims(ext, dim, hh, nn);
            return (int64_t)nn[0] * nn[1] * nn[2];
        };
        while (cells_at(h) > MAX_CELLS) h *= 1.25;
        int64_t nonempty = 0;
        int rc = MM_OK;
        double distinct = (double)M;  // estimate of the number of distinct coordinates
        if (nd > 0) {
            double hp = std::max(0.5 * h, emax / 4096.0);
            while (cells_at(hp) > MAX_CELLS) hp *= 1.1;
            int64_t ne = 0;
            rc = evaluate(hp, &ne);
            if (rc != MM_OK) return rc;
            distinct = (double)std::max<int64_t>(ne, 1);
            double ht = std::pow(vol * 1.25 / distinct, 1.0 / nd);
            if (ht > 0.0 && std::isfinite(ht)) h = std::min(std::max(ht, emax / 4096.0), emax);
            while (cells_at(h) > MAX_CELLS) h *= 1.25;
        }
        rc = evaluate(h, &nonempty);
        if (rc != MM_OK) return rc;
        for (int it = 0; it < 8 && nd > 0; ++it) {
            if ((double)M / (double)std::max<int64_t>(nonempty, 1) <= 4.0) break;
            double h2 = 0.5 * h;
            if (h2 < emax / 4096.0) break;  // keep every axis below the per-axis cell cap
            if (cells_at(h2) > MAX_CELLS || cells_at(h2) == cells_at(h)) break;
            int64_t ne2 = 0;
            rc = evaluate(h2, &ne2);
            if (rc != MM_OK) return rc;
            // accept only if the finer grid finds coordinates the probe did not know about (clusters);
            // merely separating lattice neighbours that the target density keeps together is a loss
            const bool finds_more = (double)ne2 > 1.25 * distinct;
            distinct = std::max(distinct, (double)ne2);
            if (finds_more && (double)ne2 >= 1.5 * (double)nonempty) {
                h = h2;
                nonempty = ne2;
            } else {
                rc = evaluate(h, &nonempty);  // restore the coarser grid
                if (rc != MM_OK) return rc;
                break;
            }
        }
        if (const char *e = std::getenv("MM_INDEX_CELL_SCALE")) {  // tuning experiments only
            double sc = std::atof(e);
            if (sc > 0.0 && sc != 1.0 && nd > 0 && cells_at(h * sc) <= MAX_CELLS) {
                h *= sc;
                rc = evaluate(h, &nonempty);
                if (rc != MM_OK) return rc;
            }
        }
        ix->nonempty = nonempty;
        ix->distinct_est = distinct;
    }
    for (int c = 0; c < 3; ++c) {
        ix->origin[c] = g.origin[c];
        ix->n[c] = g.n[c];
    }
    ix->cell = g.cell;
    ix->inv_cell = g.inv_cell;
    ix->ncells = (int64_t)g.n[0] * g.n[1] * g.n[2];

    // 3. exclusive scan of the histogram -> cell_start
    MM_CUDA(pool_alloc((void **)&ix->cell_start, sizeof(int32_t) * (size_t)(ix->ncells + 1), stream));
    ix->bytes += sizeof(int32_t) * (size_t)(ix->ncells + 1);
    if (M == 0) {
        MM_CUDA(cudaMemsetAsync(ix->cell_start, 0, sizeof(int32_t) * (size_t)(ix->ncells + 1),
                                stream));
    } else {
        int64_t ntiles = (ix->ncells + SCAN_TILE - 1) / SCAN_TILE;
        MM_CUDA(pool_alloc((void **)&tile_sums, sizeof(int32_t) * (size_t)ntiles, stream));
        scan_tile_sums<<<(int)ntiles, SCAN_BLOCK, 0, stream>>>(ix->ncells, counts, tile_sums);
        scan_tile_offsets<<<1, 1024, 0, stream>>>(ntiles, tile_sums);
        scan_apply<<<(int)ntiles, SCAN_BLOCK, 0, stream>>>(ix->ncells, counts, tile_sums,
                                                           ix->cell_start);
        MM_CUDA(cudaGetLastError());
        // 4. scatter the points into cell order (the histogram buffer becomes the cursor)
        MM_CUDA(pool_alloc((void **)&ix->recs, sizeof(double4) * (size_t)M, stream));
        ix->bytes += sizeof(double4) * (size_t)M;
        MM_CUDA(cudaMemsetAsync(counts, 0, sizeof(int32_t) * (size_t)(ix->ncells + 1), stream));
        scatter_kernel<<<launch_blocks(M, 256, 8), 256, 0, stream>>>(g, M, points, ix->cell_start,
                                                                     counts, ix->recs);
        MM_CUDA(cudaGetLastError());
        // compact fp32 table of the warp-cooperative first pass -- unless the data is full of duplicates (the
        // GLL-point form), where the site table of mm_index_prepare_sites takes its place
        if (ix->distinct_est >= 0.7 * (double)M) {
            MM_CUDA(pool_alloc((void **)&ix->recf, sizeof(float4) * (size_t)M, stream));
            MM_CUDA(pool_alloc((void **)&ix->rec_id, sizeof(int32_t) * (size_t)M, stream));
            ix->bytes += (sizeof(float4) + sizeof(int32_t)) * (size_t)M;
            recf_plain_kernel<<<launch_blocks(M, 256, 8), 256, 0, stream>>>(g, M, ix->recs, ix->recf, ix->rec_id);
            MM_CUDA(cudaGetLastError());
        }
    }
    MM_CUDA(cudaStreamSynchronize(stream));
    guard.p = nullptr;
    *out = ix;
    return MM_OK;
}

extern "C" int mm_knn(const mm_index_t *ix, int64_t N, const double *pts, int k, int32_t divisor,
                      int32_t *idx, double *d2, void *stream)
{
    MM_REQUIRE(ix, MM_ERR_INVALID, "mm_knn: null index");
    return mm_knn_strided(ix, N, pts, ix->dim, k, divisor, idx, d2, stream, nullptr, 0);
}

int mm_knn_strided(const mm_index_t *ix, int64_t N, const double *pts, int pts_stride, int k,
                   int32_t divisor, int32_t *idx, double *d2, void *stream, const int64_t *n_dev_,
                   int64_t n_off)
{
    const long long *n_dev = reinterpret_cast<const long long *>(n_dev_);
    MM_REQUIRE(ix, MM_ERR_INVALID, "mm_knn: null index");
    MM_REQUIRE(k >= 1 && k <= 64, MM_ERR_INVALID, "mm_knn: k=%d outside [1, 64]", k);
    MM_REQUIRE(divisor >= 1, MM_ERR_INVALID, "mm_knn: divisor %d", (int)divisor);
    MM_REQUIRE(N >= 0, MM_ERR_INVALID, "mm_knn: N");
    if (N == 0) return MM_OK;
    MM_REQUIRE(pts && idx, MM_ERR_INVALID, "mm_knn: null buffer");
    grid_t g = grid_of(ix);
    cudaStream_t st = (cudaStream_t)stream;
    if (k <= 4) {
        knn_kernel<reg_list<4>><<<launch_blocks(N, KNN_BLOCK, 8), KNN_BLOCK, 0, st>>>(
            g, N, pts, pts_stride, k, fast_div(divisor), ix->recs, ix->cell_start, idx, d2, n_dev, n_off);
    } else if (k <= 8) {
        knn_kernel<reg_list<8>><<<launch_blocks(N, KNN_BLOCK, 8), KNN_BLOCK, 0, st>>>(
            g, N, pts, pts_stride, k, fast_div(divisor), ix->recs, ix->cell_start, idx, d2, n_dev, n_off);
    } else {
        size_t smem = (size_t)k * KNN_BLOCK * (sizeof(double) + sizeof(int32_t));
        static mm_kernel_cfg kcfg;
        MM_CUDA(kcfg.prepare(knn_kernel<smem_list>, KNN_BLOCK, smem, nullptr));
        int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (200 * 1024) / smem));
        knn_kernel<smem_list><<<launch_blocks(N, KNN_BLOCK, per_sm), KNN_BLOCK, smem, st>>>(
            g, N, pts, pts_stride, k, fast_div(divisor), ix->recs, ix->cell_start, idx, d2, n_dev, n_off);
    }
    MM_CUDA(cudaGetLastError());
    return MM_OK;
}

// ------------------------------------------------------------------------------------------------
// internal: counting sort of query points by the index's cell id.  The order inside a cell depends
// on atomics and is NOT deterministic; results of the pipeline do not depend on it (every point's
// result is a pure function of the point) and are written back through the index stored in the
// records.
// scratch layout: counts[ncells + 1] | starts[ncells + 1] | tile_sums[ntiles]
// ------------------------------------------------------------------------------------------------
size_t mm_index_sort_scratch_bytes(const mm_index_t *ix)
{
    int64_t ntiles = (ix->ncells + SCAN_TILE - 1) / SCAN_TILE;
    return sizeof(int32_t) * (size_t)(2 * (ix->ncells + 1) + ntiles + 4);
}

int mm_index_sort_queries(const mm_index_t *ix, int64_t N, const double *pts, double *sorted_rec,
                          int32_t *rank_tmp, void *scratch, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    MM_REQUIRE(N < ((int64_t)1 << 31), MM_ERR_INVALID, "mm_interpolate: N=%lld >= 2^31 per call",
               (long long)N);
    MM_REQUIRE(((uintptr_t)sorted_rec & 31) == 0, MM_ERR_INVALID, "mm_interpolate: workspace alignment");
    grid_t g = grid_of(ix);
    int32_t *counts = static_cast<int32_t *>(scratch);
    int32_t *starts = counts + (ix->ncells + 1);
    int32_t *tile_sums = starts + (ix->ncells + 1);
    int64_t ntiles = (ix->ncells + SCAN_TILE - 1) / SCAN_TILE;
    MM_CUDA(cudaMemsetAsync(counts, 0, sizeof(int32_t) * (size_t)(ix->ncells + 1), stream));
    query_rank_kernel<<<launch_blocks(N, 256, 8), 256, 0, stream>>>(g, N, pts, counts, rank_tmp);
    scan_tile_sums<<<(int)ntiles, SCAN_BLOCK, 0, stream>>>(ix->ncells, counts, tile_sums);
    scan_tile_offsets<<<1, 1024, 0, stream>>>(ntiles, tile_sums);
    scan_apply<<<(int)ntiles, SCAN_BLOCK, 0, stream>>>(ix->ncells, counts, tile_sums, starts);
    query_place_kernel<<<launch_blocks(N, 256, 8), 256, 0, stream>>>(g, N, pts, starts, rank_tmp,
                                                                     sorted_rec);
    MM_CUDA(cudaGetLastError());
    return MM_OK;
}

// ------------------------------------------------------------------------------------------------
// internal: site table and the site-level first pass (see knn_sites_kernel)
// ------------------------------------------------------------------------------------------------
extern "C" int mm_index_prepare_sites(mm_index_t *ix, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    MM_REQUIRE(ix, MM_ERR_INVALID, "mm_index_prepare_sites: null index");
    if (ix->site_recs || ix->M == 0) return MM_OK;
    int32_t *per_cell = nullptr, *tile_sums = nullptr;
    const int64_t ntiles = (ix->ncells + SCAN_TILE - 1) / SCAN_TILE;
    MM_CUDA(pool_alloc((void **)&per_cell, sizeof(int32_t) * (size_t)(ix->ncells + 1), stream));
    MM_CUDA(pool_alloc((void **)&tile_sums, sizeof(int32_t) * (size_t)ntiles, stream));
    MM_CUDA(pool_alloc((void **)&ix->site_cell_start, sizeof(int32_t) * (size_t)(ix->ncells + 1), stream));
    site_sort_kernel<<<launch_blocks(ix->ncells, 128, 16), 128, 0, stream>>>(ix->ncells, ix->cell_start,
                                                                             ix->recs, per_cell);
    scan_tile_sums<<<(int)ntiles, SCAN_BLOCK, 0, stream>>>(ix->ncells, per_cell, tile_sums);
    scan_tile_offsets<<<1, 1024, 0, stream>>>(ntiles, tile_sums);
    scan_apply<<<(int)ntiles, SCAN_BLOCK, 0, stream>>>(ix->ncells, per_cell, tile_sums, ix->site_cell_start);
    MM_CUDA(cudaGetLastError());
    int32_t ns = 0;
    MM_CUDA(cudaMemcpyAsync(&ns, ix->site_cell_start + ix->ncells, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
    MM_CUDA(cudaStreamSynchronize(stream));
    ix->nsites = ns;
    MM_CUDA(pool_alloc((void **)&ix->site_recs, sizeof(double4) * (size_t)(ns + 1), stream));
    site_fill_kernel<<<launch_blocks(ix->ncells, 128, 16), 128, 0, stream>>>(
        ix->ncells, ix->cell_start, ix->recs, ix->site_cell_start, ix->site_recs, ns, (int32_t)ix->M);
    MM_CUDA(cudaGetLastError());
    cudaFreeAsync(per_cell, stream);
    cudaFreeAsync(tile_sums, stream);
    ix->bytes += sizeof(int32_t) * (size_t)(ix->ncells + 1) + sizeof(double4) * (size_t)(ns + 1);
    // compact tables of the block kernel; the in-cell sort above moved the records, so a plain-form table is stale
    if (ix->recf) {
        cudaFreeAsync(ix->recf, stream);
        ix->recf = nullptr;
        ix->bytes -= sizeof(float4) * (size_t)ix->M;
    }
    if (!ix->rec_id) {
        MM_CUDA(pool_alloc((void **)&ix->rec_id, sizeof(int32_t) * (size_t)ix->M, stream));
        ix->bytes += sizeof(int32_t) * (size_t)ix->M;
    }
    MM_CUDA(pool_alloc((void **)&ix->recf, sizeof(float4) * (size_t)std::max<int64_t>(ns, 1), stream));
    MM_CUDA(pool_alloc((void **)&ix->site_first, sizeof(int32_t) * (size_t)(ns + 1), stream));
    ix->bytes += sizeof(float4) * (size_t)ns + sizeof(int32_t) * (size_t)(ns + 1);
    ix->recf_sites = true;
    rec_id_kernel<<<launch_blocks(ix->M, 256, 8), 256, 0, stream>>>(ix->M, ix->recs, ix->rec_id);
    recf_sites_kernel<<<launch_blocks(ns + 1, 256, 8), 256, 0, stream>>>(grid_of(ix), ns, ix->site_recs, ix->recf,
                                                                         ix->site_first);
    MM_CUDA(cudaGetLastError());
    MM_CUDA(cudaStreamSynchronize(stream));  // the table is complete when this returns: any stream may use it
    return MM_OK;
}

bool mm_index_has_sites(const mm_index_t *ix) { return ix && ix->site_recs != nullptr; }

bool mm_index_sites_view_get(const mm_index_t *ix, mm_index_sites_view *out)
{
    if (!ix || !ix->site_recs || !ix->site_first || !ix->rec_id) return false;
    out->nsites = ix->nsites;
    out->M = ix->M;
    out->site_recs = ix->site_recs;
    out->site_first = ix->site_first;
    out->rec_id = ix->rec_id;
    return true;
}

// first pass over SORTED queries with the warp-cooperative block kernel, when it applies
static bool block_kernel_applies(const mm_index_t *ix)
{
    if (const char *e = getenv("MM_KNN_BLOCK"))
        if (e[0] == '0') return false;
    return ix->cell > 1e-15 && ix->cell < 1e15;  // relative coordinates and the error margin must be normal floats
}

int mm_knn_first_pass(const mm_index_t *ix, int64_t N, const double *pts, int pts_stride, int k,
                      int32_t divisor, int32_t *idx, void *stream)
{
    MM_REQUIRE(ix, MM_ERR_INVALID, "mm_knn_first_pass: null index");
    if (N == 0) return MM_OK;
    if (k == 4 && ix->recf && !ix->recf_sites && block_kernel_applies(ix)) {
        knn_block_kernel<false><<<launch_blocks(N, KB_WARPS * 32, 8), KB_WARPS * 32, 0, (cudaStream_t)stream>>>(
            grid_of(ix), N, pts, pts_stride, k, fast_div(divisor), ix->recs, ix->cell_start, ix->recf, ix->rec_id,
            nullptr, idx);
        MM_CUDA(cudaGetLastError());
        return MM_OK;
    }
    return mm_knn_strided(ix, N, pts, pts_stride, k, divisor, idx, nullptr, stream);
}

int mm_knn_sites(const mm_index_t *ix, int64_t N, const double *pts, int pts_stride, int kout,
                 int32_t divisor, int32_t *idx, void *stream)
{
    MM_REQUIRE(ix && ix->site_recs, MM_ERR_INVALID, "mm_knn_sites: site table not built");
    if (N == 0) return MM_OK;
    if (ix->recf_sites && block_kernel_applies(ix)) {
        knn_block_kernel<true><<<launch_blocks(N, KB_WARPS * 32, 8), KB_WARPS * 32, 0, (cudaStream_t)stream>>>(
            grid_of(ix), N, pts, pts_stride, kout, fast_div(divisor), ix->site_recs, ix->site_cell_start, ix->recf,
            ix->rec_id, ix->site_first, idx);
        MM_CUDA(cudaGetLastError());
        return MM_OK;
    }
    knn_sites_kernel<<<launch_blocks(N, KNN_BLOCK, 8), KNN_BLOCK, 0, (cudaStream_t)stream>>>(
        grid_of(ix), N, pts, pts_stride, kout, fast_div(divisor), ix->site_recs, ix->site_cell_start, idx,
        ix->recs);
    MM_CUDA(cudaGetLastError());
    return MM_OK;
}
