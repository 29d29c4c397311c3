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
    // compact id tables (site form; K1's CTA-tile pass and K4 read 4-byte ids instead of 32-byte records):
    //   rec_id   [M] int32: point id of record t (recs[t].w)
    //   site_first [nsites + 1] int32: first record of each site
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
// First pass of the pipeline in PREFIX mode (mm_pipeline.cu), CTA-TILE form: one CTA per tile of TX x TY x TZ index
// cells.  The queries are sorted by cell (mm_index_sort_queries), so the queries of one cell row of the tile are
// one contiguous range of the sorted records (the sort's own `starts` table gives the ranges): the tile's queries
// are a handful of ranges, and all of them look at the same halo of (TX+2) x (TY+2) x (TZ+2) cells.
//   1. the CTA stages the halo ONCE: every halo row is one contiguous run of site (or point) records, copied
//      coalesced into shared memory twice -- the binary64 record {x, y, z, first copy / id} and a binary32 record
//      {x, y, z relative to the tile centre, |.|^2} -- plus a table cs[row][column] of the first slot of every cell;
//   2. thread per query: the 3 x 3 x 3 neighbourhood is 9 (3 in 2-D) short slot ranges.  Every candidate costs one
//      LDS.128, four fp32 operations (d' = |s|^2 + (|q|^2 + bias) - 2 q.s), one LOP3 (key = d' bits with the
//      slot number in the low 10 bits) and a branch-free min/max network that keeps the NK smallest keys;
//   3. the NK candidates are evaluated exactly, in binary64 with the canonical operation order, from the staged
//      records and sorted by exact distance e_0 <= ... <= e_{NK-1}.  Site j is emitted -- all its copies, ids
//      ascending -- while   e_j < e_{j+1}  (strictly: equidistant sites interleave by id)
//                           e_j < L        (L = lower bound of the exact distance of every candidate that is NOT among
//                                           the NK: the truncated fp32 key of the NK-th, less bias and error bound)
//                           e_j < B^2      (B = distance to the nearest face of the 3^3 block that has cells beyond it:
//                                           every record outside the block is at least that far away)
//      so what is written is a PREFIX of the canonical (d2, id) list (possibly shorter than the old kernels' when
//      distances nearly tie, never different), padded with -1.  Points the prefix does not resolve are re-run with
//      the complete search by the pipeline, as are queries outside the grid and tiles whose halo does not fit.
// fp32 error bound: coordinates relative to the tile centre, |x| <= 9.5 h, |y|, |z| <= 9.5 h (2-D tiles) => terms up
// to ~200 h^2, ten roundings of 2^-24 each: 1.2e-4 h^2; coordinate roundings 2 * 1.5 h * 2 * 2^-24 * 9.5 h per axis:
// 1e-5 h^2.  KT_EPS = 3e-4 h^2 is used, and the same amount is added to every d' as a bias so that keys are positive.
// ================================================================================================
constexpr int KT_THREADS = 256;
constexpr int KT_CAP = 1024;      // staged records per tile (10-bit slot number in the key)
constexpr int KT_MAXROWS = 36;    // halo rows (TY + 2) * (TZ + 2)
constexpr int KT_MAXCOLS = 20;    // halo columns + 1
constexpr int KT_MAXQROWS = 16;   // tile rows TY * TZ
constexpr double KT_EPS = 3e-4;

struct kt_dims {
    int tx, ty, tz;     // tile size in cells
    int ntx, nty, ntz;  // tiles per axis
    int cap;            // staged records a (sub-)tile may use, <= KT_CAP (smaller in tests: forces the split paths)
};

struct kt_smem {
    double4 sd[KT_CAP];
    float4 sf[KT_CAP];
    int32_t cs[KT_MAXROWS][KT_MAXCOLS];
    int32_t rowslot[KT_MAXROWS + 1];  // first slot of each halo row (+ total)
    int32_t rowg[KT_MAXROWS];         // global index of the row's first staged record
    int32_t rowbase[KT_MAXROWS];      // linear id of the row's cell 0, or -1 outside the grid
    int32_t qpre[KT_MAXQROWS + 1];    // queries before each tile row (+ total)
    int32_t qlo[KT_MAXQROWS];         // first sorted query of each tile row
};

// one (sub-)tile of sx x sy x sz cells at cell (tx0, ty0, tz0); false: its halo does not fit the staging buffers
// (nothing written).  Called by all threads of the CTA.
template <bool SITES, int NK, int OUTK>
__device__ __forceinline__ bool
kt_run(kt_smem &sm, const grid_t &g, const int tx0, const int ty0, const int tz0, const int sx, const int sy,
       const int sz, const double4 *__restrict__ qrecs, const int32_t *__restrict__ qstart,
       const double4 *__restrict__ srecs, const int32_t *__restrict__ scs, const int32_t *__restrict__ rec_id,
       const fast_div &divisor, int32_t *__restrict__ out_idx, int32_t *__restrict__ perm_out, const bool give_up,
       const int cap)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool three_d = g.dim == 3;
    const int nx = g.n[0], ny = g.n[1], nz = g.n[2];
    const int qrows = sy * sz;
    const int hy = sy + 2, hz = three_d ? sz + 2 : 1, hrows = hy * hz;
    const int xa = max(tx0 - 1, 0), xb = min(tx0 + sx, nx - 1), ncol = xb - xa + 1;
    struct tile_dims { int tx, ty, tz; } td{sx, sy, sz};
    __syncthreads();  // the previous (sub-)tile is done with the staging buffers

    // ---- 1. query ranges of the tile rows, record ranges of the halo rows -------------------------------------
    if (tid < qrows) {
        const int y = ty0 + tid % td.ty, z = tz0 + tid / td.ty;
        int lo = 0, hi = 0;
        if (y < ny && z < nz) {
            const int base = nx * (y + ny * z);
            lo = qstart[base + tx0];
            hi = qstart[base + min(tx0 + td.tx, nx)];
        }
        sm.qlo[tid] = lo;
        sm.qpre[tid] = hi - lo;
    } else if (tid >= 32 && tid < 32 + hrows) {
        const int r = tid - 32;
        const int y = ty0 - 1 + r % hy, z = three_d ? tz0 - 1 + r / hy : 0;
        int base = -1, glo = 0, len = 0;
        if (y >= 0 && y < ny && z >= 0 && z < nz) {
            base = nx * (y + ny * z);
            glo = scs[base + xa];
            len = scs[base + xb + 1] - glo + (SITES ? 1 : 0);  // + the next site: its `first` ends the last one
        }
        sm.rowbase[r] = base;
        sm.rowg[r] = glo;
        sm.rowslot[r] = len;
    }
    __syncthreads();
    if (tid == 0) {
        int run = 0;
        for (int r = 0; r < qrows; ++r) {
            const int c = sm.qpre[r];
            sm.qpre[r] = run;
            run += c;
        }
        sm.qpre[qrows] = run;
    } else if (tid == 32) {
        int run = 0;
        for (int r = 0; r < hrows; ++r) {
            const int c = sm.rowslot[r];
            sm.rowslot[r] = run;
            run += c;
        }
        sm.rowslot[hrows] = run;
    }
    __syncthreads();
    const int T = sm.qpre[qrows];
    if (T == 0) return true;
    // sorted position of the tile's i-th query
    auto query_of = [&](int i, int &row) -> int {
        int r = 0;
#pragma unroll
        for (int step = KT_MAXQROWS / 2; step > 0; step >>= 1)
            if (r + step < qrows && sm.qpre[r + step] <= i) r += step;
        row = r;
        return sm.qlo[r] + (i - sm.qpre[r]);
    };
    if (sm.rowslot[hrows] > cap) {  // crowded halo
        if (!give_up) return false;    // the caller splits the tile
        for (int i = tid; i < T; i += KT_THREADS) {  // the complete search takes these points (re-run)
            int row;
            const int64_t n = query_of(i, row);
            int32_t *o = out_idx + n * OUTK;
#pragma unroll
            for (int c = 0; c < OUTK; ++c) o[c] = -1;
            if (perm_out) perm_out[n] = (int32_t)__double_as_longlong(__ldg(reinterpret_cast<const double *>(qrecs + n) + 3));
        }
        return true;
    }
    // ---- 2. stage the halo ---------------------------------------------------------------------------------
    const double h = g.cell;
    const double cxo = g.origin[0] + (tx0 + 0.5 * td.tx) * h, cyo = g.origin[1] + (ty0 + 0.5 * td.ty) * h;
    const double czo = three_d ? g.origin[2] + (tz0 + 0.5 * td.tz) * h : 0.0;
    for (int i = tid; i < hrows * (ncol + 1); i += KT_THREADS) {
        const int r = i / (ncol + 1), c = i - r * (ncol + 1);
        int v = sm.rowslot[r];
        const int base = sm.rowbase[r];
        if (base >= 0) v += scs[base + xa + c] - sm.rowg[r];
        sm.cs[r][c] = v;
    }
    for (int r = warp; r < hrows; r += KT_THREADS / 32) {
        const int slot0 = sm.rowslot[r], len = sm.rowslot[r + 1] - slot0;
        const double4 *src = srecs + sm.rowg[r];
        for (int j = lane; j < len; j += 32) {
            const double2 *q = reinterpret_cast<const double2 *>(src + j);
            const double2 xy = __ldg(q), zw = __ldg(q + 1);
            double4 d;
            d.x = xy.x;
            d.y = xy.y;
            d.z = zw.x;
            d.w = zw.y;
            sm.sd[slot0 + j] = d;
            float4 f;
            f.x = (float)(xy.x - cxo);
            f.y = (float)(xy.y - cyo);
            f.z = three_d ? (float)(zw.x - czo) : 0.0f;
            f.w = fmaf(f.z, f.z, fmaf(f.y, f.y, f.x * f.x));
            sm.sf[slot0 + j] = f;
        }
    }
    __syncthreads();

    // ---- 3. the tile's queries -----------------------------------------------------------------------------
    const float h32 = (float)h;
    const float eps = (float)(KT_EPS * h * h);
    // low corner of the tile relative to the centre: the faces of a query's 3^3 block in the fp32 frame
    const float fx0 = (float)(-0.5 * td.tx * h), fy0 = (float)(-0.5 * td.ty * h), fz0 = (float)(-0.5 * td.tz * h);
    const float fmargin = 1e-4f * h32;
    for (int i = tid; i < T; i += KT_THREADS) {
        int row;
        const int64_t n = query_of(i, row);
        const double2 *qp = reinterpret_cast<const double2 *>(qrecs + n);
        const double2 pxy = __ldg(qp), pzw = __ldg(qp + 1);
        const double px = pxy.x, py = pxy.y, pz = three_d ? pzw.x : 0.0;
        // the original index of the query, once more as a compact array: later passes (K3's grouping) read 4 bytes
        // per point instead of a 32-byte record
        if (perm_out) perm_out[n] = (int32_t)__double_as_longlong(pzw.y);
        const int ry = row % td.ty, rz = row / td.ty;
        const int cy = ty0 + ry, cz = tz0 + rz;
        // the query's cell along x (y and z are those of the tile row: the sort used the same cell_coord)
        const int cx = cell_coord(g, px, 0);
        // relative position inside the own cell, in cells: a query that the sort CLAMPED into the grid is not
        // within its cell -- its fp32 coordinates would not obey the error bound
        const double ux = (px - g.origin[0]) * g.inv_cell - cx, uy = (py - g.origin[1]) * g.inv_cell - cy;
        const double uz = three_d ? (pz - g.origin[2]) * g.inv_cell - cz : 0.5;
        int32_t res[OUTK];
#pragma unroll
        for (int c = 0; c < OUTK; ++c) res[c] = -1;
        const bool inside = ux >= -0.01 && ux <= 1.01 && uy >= -0.01 && uy <= 1.01 && uz >= -0.01 && uz <= 1.01;
        if (inside) {
            const float qx = (float)(px - cxo), qy = (float)(py - cyo), qz = three_d ? (float)(pz - czo) : 0.0f;
            const float q2 = fmaf(qz, qz, fmaf(qy, qy, qx * qx)) + eps;
            const float mx = -2.0f * qx, my = -2.0f * qy, mz = -2.0f * qz;
            unsigned a[NK];
#pragma unroll
            for (int u = 0; u < NK; ++u) a[u] = 0xffffffffu;
            const int ca = max(cx - 1, 0) - xa, cb = min(cx + 1, nx - 1) + 1 - xa;
            const int ly = ry + 1, lz = three_d ? rz + 1 : 0;
            for (int dz = three_d ? -1 : 0; dz <= (three_d ? 1 : 0); ++dz) {
#pragma unroll
                for (int dy = -1; dy <= 1; ++dy) {
                    const int hr = (ly + dy) + hy * (lz + dz);
                    const int lo = sm.cs[hr][ca], hi = sm.cs[hr][cb];
                    for (int j = lo; j < hi; ++j) {
                        const float4 s = sm.sf[j];
                        const float d = fmaf(mx, s.x, fmaf(my, s.y, fmaf(mz, s.z, s.w + q2)));
                        const unsigned key = (__float_as_uint(d) & 0xfffffc00u) | (unsigned)j;
#pragma unroll
                        for (int u = NK - 1; u > 0; --u) a[u] = max(a[u - 1], min(a[u], key));
                        a[0] = min(a[0], key);
                    }
                }
            }
            // exact distances of the NK candidates, canonical operation order
            double e[NK + 1];
            int sl[NK];
#pragma unroll
            for (int u = 0; u < NK; ++u) {
                sl[u] = (int)(a[u] & 0x3ffu);
                e[u] = INFINITY;
                if (a[u] != 0xffffffffu) {
                    const double4 r = sm.sd[sl[u]];
                    const double dx = px - r.x, dy = py - r.y;
                    double s2 = dx * dx + dy * dy;
                    if (three_d) {
                        const double dz = pz - r.z;
                        s2 = s2 + dz * dz;
                    }
                    e[u] = s2;
                }
            }
            e[NK] = INFINITY;
            // every record outside the NK has an exact distance of at least L
            double L = INFINITY;
            if (a[NK - 1] != 0xffffffffu)
                L = (double)(__uint_as_float(a[NK - 1] & 0xfffffc00u) - 2.01f * eps);
            // distance to the faces of the 3^3 block that have cells beyond them (fp32 frame, conservative margin)
            float B = INFINITY;
            {
                const float xl = fx0 + (float)(cx - 1 - tx0) * h32, yl = fy0 + (float)(ry - 1) * h32;
                if (cx - 1 > 0) B = fminf(B, qx - xl);
                if (cx + 1 < nx - 1) B = fminf(B, xl + 3.0f * h32 - qx);
                if (cy - 1 > 0) B = fminf(B, qy - yl);
                if (cy + 1 < ny - 1) B = fminf(B, yl + 3.0f * h32 - qy);
                if (three_d) {
                    const float zl = fz0 + (float)(rz - 1) * h32;
                    if (cz - 1 > 0) B = fminf(B, qz - zl);
                    if (cz + 1 < nz - 1) B = fminf(B, zl + 3.0f * h32 - qz);
                }
                B = fmaxf(B - fmargin, 0.0f);
            }
            const double lim = fmin(L, (double)B * (double)B);
            // sort the NK by exact distance (ties: the emission stops there anyway)
#define MM_KT_CE(I, J)                                   \
    {                                                    \
        const bool sw = e[J] < e[I];                     \
        const double e_lo = sw ? e[J] : e[I];            \
        const double e_hi = sw ? e[I] : e[J];            \
        const int s_lo = sw ? sl[J] : sl[I];             \
        const int s_hi = sw ? sl[I] : sl[J];             \
        e[I] = e_lo;                                     \
        e[J] = e_hi;                                     \
        sl[I] = s_lo;                                    \
        sl[J] = s_hi;                                    \
    }
            if constexpr (NK == 4) {
                MM_KT_CE(0, 1) MM_KT_CE(2, 3) MM_KT_CE(0, 2) MM_KT_CE(1, 3) MM_KT_CE(1, 2)
            } else {
                static_assert(NK == 5, "network sizes 4 and 5");
                MM_KT_CE(0, 1) MM_KT_CE(3, 4) MM_KT_CE(2, 4) MM_KT_CE(2, 3) MM_KT_CE(1, 4)
                MM_KT_CE(0, 3) MM_KT_CE(0, 2) MM_KT_CE(1, 3) MM_KT_CE(1, 2)
            }
#undef MM_KT_CE
            bool go = true;
            if constexpr (SITES) {
                int first[NK - 1], upto[NK - 1];  // first record of site j, entries emitted up to and including site j
                int run = 0;
#pragma unroll
                for (int u = 0; u < NK - 1; ++u) {
                    go = go && e[u] < e[u + 1] && e[u] < lim;
                    first[u] = 0;
                    if (go) {
                        const int32_t *w0 = reinterpret_cast<const int32_t *>(&sm.sd[sl[u]].w);
                        const int32_t *w1 = reinterpret_cast<const int32_t *>(&sm.sd[sl[u] + 1].w);
                        first[u] = *w0;
                        run += *w1 - *w0;
                    }
                    upto[u] = run;
                }
#pragma unroll
                for (int c = 0; c < OUTK; ++c) {
                    int tpos = -1;
#pragma unroll
                    for (int u = NK - 2; u >= 0; --u)
                        if (c < upto[u]) tpos = first[u] + c - (u > 0 ? upto[u - 1] : 0);
                    // (descending u: the smallest u with c < upto[u] wins)
                    if (tpos >= 0) res[c] = divisor(__ldg(&rec_id[tpos]));
                }
            } else {
#pragma unroll
                for (int u = 0; u < NK - 1 && u < OUTK; ++u) {
                    go = go && e[u] < e[u + 1] && e[u] < lim;
                    if (go) res[u] = divisor((int32_t)__double_as_longlong(sm.sd[sl[u]].w));
                }
            }
        }
        int4 *o = reinterpret_cast<int4 *>(out_idx + n * OUTK);
#pragma unroll
        for (int c = 0; c < OUTK / 4; ++c) o[c] = make_int4(res[4 * c], res[4 * c + 1], res[4 * c + 2], res[4 * c + 3]);
    }
    return true;
}

template <bool SITES, int NK, int OUTK>
__global__ void __launch_bounds__(KT_THREADS, 4)
knn_tile_kernel(const grid_t g, const kt_dims td, const double4 *__restrict__ qrecs,
                const int32_t *__restrict__ qstart, const double4 *__restrict__ srecs,
                const int32_t *__restrict__ scs, const int32_t *__restrict__ rec_id, const fast_div divisor,
                int32_t *__restrict__ out_idx, int32_t *__restrict__ perm_out)
{
    extern __shared__ __align__(32) unsigned char kt_raw[];
    kt_smem &sm = *reinterpret_cast<kt_smem *>(kt_raw);
    int t = blockIdx.x;
    const int tix = t % td.ntx;
    t /= td.ntx;
    const int tiy = t % td.nty, tiz = t / td.nty;
    const int tx0 = tix * td.tx, ty0 = tiy * td.ty, tz0 = tiz * td.tz;
    // pass 0: the whole tile.  A crowded tile (locally dense source) is split into eight (four in 2-D) sub-tiles,
    // each with its own, smaller halo (passes 1 .. nsub); one call site, so the body is instantiated once
    const int sx = max(td.tx / 2, 1), sy = max(td.ty / 2, 1), sz = max(td.tz / 2, 1);
    const int px = (td.tx + sx - 1) / sx, py = (td.ty + sy - 1) / sy, pz = (td.tz + sz - 1) / sz;
    for (int pass = 0; pass <= px * py * pz; ++pass) {
        int x0 = tx0, y0 = ty0, z0 = tz0, ex = td.tx, ey = td.ty, ez = td.tz;
        if (pass > 0) {
            const int q = pass - 1;
            x0 = tx0 + (q % px) * sx;
            y0 = ty0 + ((q / px) % py) * sy;
            z0 = tz0 + (q / (px * py)) * sz;
            ex = min(sx, tx0 + td.tx - x0);
            ey = min(sy, ty0 + td.ty - y0);
            ez = min(sz, tz0 + td.tz - z0);
            if (x0 >= g.n[0] || y0 >= g.n[1] || z0 >= g.n[2]) continue;
        }
        const bool ok = kt_run<SITES, NK, OUTK>(sm, g, x0, y0, z0, ex, ey, ez, qrecs, qstart, srecs, scs, rec_id,
                                                divisor, out_idx, perm_out, pass > 0, td.cap);
        if (pass == 0 && ok) return;
    }
}

// ---- compact id tables: rec_id[t] = point id of record t; site_first[s] = first record of site s ------------
__global__ void __launch_bounds__(256)
site_first_kernel(int64_t nsites, const double4 *__restrict__ site_recs, int32_t *__restrict__ site_first)
{
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t <= nsites; t += (int64_t)gridDim.x * blockDim.x)
        site_first[t] = (int32_t)__double_as_longlong(site_recs[t].w);
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
        // 2. cell size.  Target: ~1.25 DISTINCT coordinates per cell of the OCCUPIED volume (measured optimum of the
        //    searches on lattice-like and random data, tools/exp_cell.py).  Neither quantity is known up front:
        //    * the occupied volume V is far below the bounding box for the meshes this path sees (a spherical shell
        //      fills half of its box, a 25 km crust layer a fraction of a per cent): V = (non-empty cells of a grid
        //      twice as coarse as the candidate) x (their volume), iterated to a fixed point of h = (1.25 V / D)^(1/d);
        //    * the number of distinct coordinates D is M unless the points repeat (the GLL-point form stores shared
        //      nodes 2-8 times, and duplicates never separate): when a grid twice as fine as the fixed point finds
        //      hardly any new non-empty cells although most cells hold several points, every non-empty cell of it IS
        //      one distinct coordinate, D = that count, and the fixed point is taken again.
        //    Multi-scale (clustered) data is then refined while cells stay crowded and halving still separates points.
        double h = 1.0;
        double emax = std::max(ext[0], std::max(ext[1], ext[2]));
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
        auto feasible = [&](double hh) {  // inside the per-axis and total cell caps
            if (!(hh > 0.0) || !std::isfinite(hh)) hh = emax;
            hh = std::min(std::max(hh, emax / 4096.0), emax);
            while (cells_at(hh) > MAX_CELLS) hh *= 1.1;
            return hh;
        };
        int64_t nonempty = 0;
        int rc = MM_OK;
        double distinct = (double)M;  // estimate of the number of distinct coordinates
        const double target = 1.25;
        if (nd > 0) {
            h = feasible(std::pow(vol * target / distinct, 1.0 / nd));
            auto fixed_point = [&]() -> int {
                for (int it = 0; it < 6; ++it) {
                    const double hc = feasible(2.0 * h);
                    int64_t ne = 0;
                    int r = evaluate(hc, &ne);
                    if (r != MM_OK) return r;
                    const double v_occ = (double)std::max<int64_t>(ne, 1) * std::pow(hc, nd);
                    const double hn = feasible(std::pow(target * v_occ / distinct, 1.0 / nd));
                    const bool converged = std::fabs(hn / h - 1.0) < 0.08;
                    h = hn;
                    if (converged) break;
                }
                return MM_OK;
            };
            rc = fixed_point();
            if (rc != MM_OK) return rc;
            rc = evaluate(h, &nonempty);
            if (rc != MM_OK) return rc;
            const double h2 = feasible(0.5 * h);  // as fine as the cell caps allow
            if (h2 <= 0.8 * h && cells_at(h2) > cells_at(h)) {
                int64_t ne2 = 0;
                rc = evaluate(h2, &ne2);
                if (rc != MM_OK) return rc;
                if ((double)ne2 < 0.6 * distinct && (double)ne2 <= 1.15 * (double)nonempty) {
                    distinct = (double)std::max<int64_t>(ne2, 1);
                    rc = fixed_point();
                    if (rc != MM_OK) return rc;
                }
            }
        }
        rc = evaluate(h, &nonempty);
        if (rc != MM_OK) return rc;
        for (int it = 0; it < 8 && nd > 0; ++it) {
            if (distinct / (double)std::max<int64_t>(nonempty, 1) <= 4.0) break;
            double h2 = 0.5 * h;
            if (h2 < emax / 4096.0) break;  // keep every axis below the per-axis cell cap
            if (cells_at(h2) > MAX_CELLS || cells_at(h2) == cells_at(h)) break;
            int64_t ne2 = 0;
            rc = evaluate(h2, &ne2);
            if (rc != MM_OK) return rc;
            // accept only if the finer grid separates points (clusters); duplicates never separate
            if ((double)ne2 >= 1.5 * (double)nonempty) {
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
    // compact id tables (K1's CTA-tile pass, K4)
    if (!ix->rec_id) {
        MM_CUDA(pool_alloc((void **)&ix->rec_id, sizeof(int32_t) * (size_t)ix->M, stream));
        ix->bytes += sizeof(int32_t) * (size_t)ix->M;
    }
    MM_CUDA(pool_alloc((void **)&ix->site_first, sizeof(int32_t) * (size_t)(ns + 1), stream));
    ix->bytes += sizeof(int32_t) * (size_t)(ns + 1);
    rec_id_kernel<<<launch_blocks(ix->M, 256, 8), 256, 0, stream>>>(ix->M, ix->recs, ix->rec_id);
    site_first_kernel<<<launch_blocks(ns + 1, 256, 8), 256, 0, stream>>>(ns, ix->site_recs, ix->site_first);
    MM_CUDA(cudaGetLastError());
    MM_CUDA(cudaStreamSynchronize(stream));  // the table is complete when this returns: any stream may use it
    return MM_OK;
}

bool mm_index_has_sites(const mm_index_t *ix) { return ix && ix->site_recs != nullptr; }

int64_t mm_index_size(const mm_index_t *ix) { return ix ? ix->M : 0; }

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

// first pass when the CTA-tile kernel does not apply (complete semantics, sparse query sets): thread per query
int mm_knn_first_pass(const mm_index_t *ix, int64_t N, const double *pts, int pts_stride, int k,
                      int32_t divisor, int32_t *idx, void *stream)
{
    MM_REQUIRE(ix, MM_ERR_INVALID, "mm_knn_first_pass: null index");
    if (N == 0) return MM_OK;
    return mm_knn_strided(ix, N, pts, pts_stride, k, divisor, idx, nullptr, stream);
}

int mm_knn_sites(const mm_index_t *ix, int64_t N, const double *pts, int pts_stride, int kout,
                 int32_t divisor, int32_t *idx, void *stream)
{
    MM_REQUIRE(ix && ix->site_recs, MM_ERR_INVALID, "mm_knn_sites: site table not built");
    if (N == 0) return MM_OK;
    knn_sites_kernel<<<launch_blocks(N, KNN_BLOCK, 8), KNN_BLOCK, 0, (cudaStream_t)stream>>>(
        grid_of(ix), N, pts, pts_stride, kout, fast_div(divisor), ix->site_recs, ix->site_cell_start, idx,
        ix->recs);
    MM_CUDA(cudaGetLastError());
    return MM_OK;
}

// CTA-tile first pass of the pipeline (knn_tile_kernel): PREFIX semantics only -- the caller re-runs the points
// whose row comes back (partly) empty.  `sorted`: the 32-byte query records and `sort_scratch`: the scratch of the
// mm_index_sort_queries call that produced them (its `starts` table locates the queries of every cell).
// *applied = false when the kernel does not apply (the caller then takes mm_knn_sites / mm_knn_first_pass).
int mm_knn_tile_first_pass(const mm_index_t *ix, int64_t N, const double *sorted, const void *sort_scratch, int kout,
                           int32_t divisor, bool sites, int32_t *idx, int32_t *perm_out, void *stream, bool *applied)
{
    MM_REQUIRE(ix && applied, MM_ERR_INVALID, "mm_knn_tile_first_pass: null");
    *applied = false;
    if (N == 0 || ix->M == 0) return MM_OK;
    if (const char *e = getenv("MM_KNN_TILE"))
        if (e[0] == '0') return MM_OK;
    if (!(ix->cell > 1e-15 && ix->cell < 1e15)) return MM_OK;  // fp32 frame and error bound need normal floats
    if (kout != 4 && kout != 8) return MM_OK;
    if (sites && !(ix->site_recs && ix->rec_id && kout == 8)) return MM_OK;
    // sparse query sets: staging a halo for a handful of queries costs more than the per-query search
    if ((double)N < 0.25 * (double)ix->ncells) return MM_OK;
    // tile: 16 x 4 x 4 cells (halo 18 x 6 x 6) when a halo of that many typically occupied cells fits the staging
    // buffers with 10 % to spare, else 8 x 4 x 4; crowded tiles split themselves once more inside the kernel
    kt_dims td{};
    const double per_cell = (double)(sites ? ix->nsites : ix->M) / (double)std::max<int64_t>(ix->nonempty, 1);
    td.tx = ix->dim == 3 ? (648.0 * per_cell * 1.1 + 36.0 <= (double)KT_CAP ? 16 : 8) : 16;
    td.ty = ix->dim == 3 ? 4 : 16;
    td.tz = ix->dim == 3 ? 4 : 1;
    if (const char *e = getenv("MM_KT_TILE")) {  // tuning experiments: "tx,ty,tz"
        int a = 0, b = 0, c = 0;
        if (sscanf(e, "%d,%d,%d", &a, &b, &c) == 3 && a >= 1 && b >= 1 && c >= 1) {
            td.tx = a;
            td.ty = b;
            td.tz = ix->dim == 3 ? c : 1;
        }
    }
    td.cap = KT_CAP;
    if (const char *e = getenv("MM_KT_CAP")) td.cap = std::max(1, std::min(KT_CAP, atoi(e)));  // tests
    MM_REQUIRE(td.tx + 3 <= KT_MAXCOLS && td.ty * td.tz <= KT_MAXQROWS &&
                   (td.ty + 2) * (ix->dim == 3 ? td.tz + 2 : 1) <= KT_MAXROWS,
               MM_ERR_INVALID, "mm_knn_tile_first_pass: tile %d x %d x %d too large", td.tx, td.ty, td.tz);
    td.ntx = (ix->n[0] + td.tx - 1) / td.tx;
    td.nty = (ix->n[1] + td.ty - 1) / td.ty;
    td.ntz = (ix->n[2] + td.tz - 1) / td.tz;
    const int64_t ntiles = (int64_t)td.ntx * td.nty * td.ntz;
    const int32_t *qstart = static_cast<const int32_t *>(sort_scratch) + (ix->ncells + 1);
    const double4 *qrecs = reinterpret_cast<const double4 *>(sorted);
    const size_t smem = sizeof(kt_smem);
    cudaStream_t st = (cudaStream_t)stream;
    const grid_t g = grid_of(ix);
    const fast_div dv(divisor);
#define MM_KT_LAUNCH(S, NK, OK, RECS, CS)                                                                     \
    {                                                                                                         \
        static mm_kernel_cfg kcfg;                                                                            \
        MM_CUDA(kcfg.prepare(knn_tile_kernel<S, NK, OK>, KT_THREADS, smem, nullptr));                         \
        knn_tile_kernel<S, NK, OK><<<(unsigned)ntiles, KT_THREADS, smem, st>>>(g, td, qrecs, qstart, RECS, CS, \
                                                                               ix->rec_id, dv, idx, perm_out); \
    }
    if (sites) MM_KT_LAUNCH(true, 4, 8, ix->site_recs, ix->site_cell_start)
    else if (kout == 4) MM_KT_LAUNCH(false, 5, 4, ix->recs, ix->cell_start)
    else MM_KT_LAUNCH(false, 5, 8, ix->recs, ix->cell_start)
#undef MM_KT_LAUNCH
    MM_CUDA(cudaGetLastError());
    *applied = true;
    return MM_OK;
}
