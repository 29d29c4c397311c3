// mm_interp_elem.cu -- K3 of the fused pipeline, ELEMENT-CENTRIC form.
//
//   out[n][f] = sum_a w_a(xi_n) * fields[elem_n][f][a]        (same arithmetic as mm_interp.cu: bit-identical)
//
// After K2 every point knows its source element.  The points are grouped by element with a counting sort (two light
// passes over the element ids; the value the histogram's atomic returns is the point's rank inside its element, so
// the placement pass needs no atomics), and K3 then walks the ELEMENTS in memory order:
//   * an element's F x P block of `fields` is read from HBM exactly once per call, sequentially over the array
//     (compulsory traffic only, perfectly streaming), straight into REGISTERS: lane (f, k) of a group keeps the
//     (order+1)^(d-1) values v_f[., ., k] of one field and one slab of the last tensor axis;
//   * the group then loops over the element's points: per point the lane reads the point's Lagrange values L0, L1
//     (shared memory, broadcast, 128-bit loads) and contracts its slab, u = sum_j L1[j] (sum_i L0[i] v[i, j]) -- the
//     values stay in registers for all points of the element, so the shared-memory traffic per point is 2 (order+1)
//     doubles instead of F (order+1)^d;
//   * a last phase finishes out = sum_k L2[k] u_k with lanes mapped to (point, field), so the F values of a point are
//     written by adjacent lanes.
// Order of the floating-point operations per output = the canonical nested contraction, each
// accumulation one fma; only WHICH lane performs an operation changes.
// A warp holds 32 / (Fc * (order+1)) groups (Fc = fields per pass = min(F, 32 / (order+1))): one group of 25 lanes
// at order 4 / F = 5, two groups of 15 at order 2, three of 10 at order 1; each group owns its own element.
#include <algorithm>
#include <cstdlib>

#include "mm_common.cuh"
#include "mm_scan.cuh"

namespace {

__device__ __forceinline__ int32_t valid_elem_e(int32_t e, int64_t E) { return (e >= 0 && e < E) ? e : -1; }

// pass 1: rank of every point inside its element (histogram); failed points get their zero row and their
// un-permuted location outputs here, because no element list will contain them
__global__ void __launch_bounds__(256)
elem_rank_kernel(int dim, int64_t N, int64_t E, int F, const int32_t *__restrict__ elem_s,
                 const double *__restrict__ xi_s, const uint8_t *__restrict__ status_s,
                 const int32_t *__restrict__ perm, int perm_stride, int32_t *__restrict__ counts,
                 int32_t *__restrict__ erank, double *__restrict__ out, int32_t *__restrict__ elem_u,
                 double *__restrict__ xi_u, uint8_t *__restrict__ status_u)
{
    for (int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; n < N; n += (int64_t)gridDim.x * blockDim.x) {
        const int32_t e = valid_elem_e(elem_s[n], E);
        if (e >= 0) {
            erank[n] = atomicAdd(&counts[e], 1);
            continue;
        }
        erank[n] = -1;
        const int64_t t = perm ? (int64_t)perm[n * perm_stride] : n;
        for (int f = 0; f < F; ++f) out[t * F + f] = 0.0;
        if (elem_u) {
            elem_u[t] = -1;
            if (status_u) status_u[t] = status_s[n];
            if (xi_u)
                for (int c = 0; c < dim; ++c) xi_u[t * dim + c] = xi_s[n * dim + c];
        }
    }
}

// pass 2: {sorted position, output row} of every located point, grouped by element
__global__ void __launch_bounds__(256)
elem_place_kernel(int64_t N, const int32_t *__restrict__ elem_s, const int32_t *__restrict__ perm, int perm_stride,
                  const int32_t *__restrict__ starts, const int32_t *__restrict__ erank, int2 *__restrict__ erec)
{
    for (int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; n < N; n += (int64_t)gridDim.x * blockDim.x) {
        const int32_t r = erank[n];
        if (r < 0) continue;
        erec[starts[elem_s[n]] + r] = make_int2((int)n, perm ? perm[n * perm_stride] : (int)n);
    }
}

// grouping of arbitrary points by an integer key (the pipeline groups K2's points by their first candidate element)
__global__ void __launch_bounds__(256)
key_rank_kernel(int64_t N, int64_t E, const int32_t *__restrict__ key, int key_stride, int32_t *__restrict__ counts,
                int32_t *__restrict__ rank)
{
    for (int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; n < N; n += (int64_t)gridDim.x * blockDim.x) {
        const int32_t e = valid_elem_e(key[n * key_stride], E);
        rank[n] = atomicAdd(&counts[e >= 0 ? e : E], 1);
    }
}

__global__ void __launch_bounds__(256)
key_place_kernel(int64_t N, int64_t E, const int32_t *__restrict__ key, int key_stride,
                 const int32_t *__restrict__ starts, const int32_t *__restrict__ rank, int32_t *__restrict__ order)
{
    for (int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; n < N; n += (int64_t)gridDim.x * blockDim.x) {
        const int32_t e = valid_elem_e(key[n * key_stride], E);
        order[starts[e >= 0 ? e : E] + rank[n]] = (int32_t)n;
    }
}

constexpr int IE_WARPS = 4;  // 4 CTAs / SM at order 4 (registers), 5 at lower orders (shared memory)

// per-warp shared-memory layout, computed identically on the host (size) and on the device
struct ie_cfg {
    int F, Fc;        // fields; fields per pass = min(F, 32 / (order+1))
    int G, ng;        // lanes per group = Fc (order+1); groups (elements) per warp
    int S, PCg;       // point slots per chunk (16 at order 4, else 32); slots per group = S / ng
    int ls_stride;    // doubles per slot of the Lagrange table = DIM * MP
    int fb_gstride;   // bytes of one group's field buffer (>= Fc P 8 + 16, multiple of 16)
    int off_part, off_llast, off_orow, off_fbuf, off_bar, per_warp;  // byte offsets inside the warp's region
};

template <int ORDER, int DIM>
__host__ inline ie_cfg ie_make_cfg(int F)
{
    constexpr int M = ORDER + 1;
    constexpr int P = DIM == 2 ? M * M : M * M * M;
    constexpr int MP = M + (M & 1);
    ie_cfg c{};
    c.F = F;
    c.Fc = std::min(F, 32 / M);
    c.G = c.Fc * M;
    c.ng = 32 / c.G;
    c.S = ORDER >= 4 ? 16 : 32;
    c.PCg = c.S / c.ng;
    c.ls_stride = DIM * MP;
    c.fb_gstride = ((c.Fc * P * 8 + 16 + 15) / 16) * 16;
    int o = c.S * c.ls_stride * 8 + c.ng * 32;  // + 32 bytes of bank skew per group
    c.off_part = o;
    o += c.S * c.G * 8;
    c.off_llast = o;
    o += c.S * M * 8;
    c.off_orow = o;
    o += c.S * 4;
    o = (o + 15) / 16 * 16;
    c.off_fbuf = o;
    o += c.ng * c.fb_gstride;
    c.off_bar = o;
    o += 16;
    c.per_warp = (o + 127) / 128 * 128;
    return c;
}

// Work of one warp: units u = 0, 1, ... = (element batch ebase_u = group0 + (u / npass) * groups_total, field pass
// f0_u = (u % npass) * Fc).  While the warp works on unit u, the field chunks of unit u + 1 are already on their way
// into the warp's field buffer (one bulk-async copy per group, UBLKCP, completion on the warp's mbarrier): the
// buffer is free again as soon as the lanes have moved their slabs into registers.
template <int ORDER, int DIM>
__global__ void __launch_bounds__(IE_WARPS * 32, ORDER >= 4 ? 4 : 5)
interp_elem_kernel(const mm_gll_table T, const ie_cfg cfg, int64_t E, const double *__restrict__ fields,
                   const int32_t *__restrict__ starts, const int2 *__restrict__ erec,
                   const double *__restrict__ xi_s, const uint8_t *__restrict__ status_s, double *__restrict__ out,
                   int32_t *__restrict__ elem_u, double *__restrict__ xi_u, uint8_t *__restrict__ status_u)
{
    constexpr int M = ORDER + 1;
    constexpr int P = DIM == 2 ? M * M : M * M * M;
    constexpr int R = DIM == 2 ? M : M * M;  // values a lane keeps: one slab of the last axis
    constexpr int MP = M + (M & 1);          // Lagrange rows padded to 16 bytes
    const int F = cfg.F, Fc = cfg.Fc, G = cfg.G, ng = cfg.ng, S = cfg.S, PCg = cfg.PCg;
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char *wb = smem + (size_t)warp * cfg.per_warp;
    double *Ls = reinterpret_cast<double *>(wb);                       // [S slots][DIM][MP] (+ 32 B skew per group)
    double *part = reinterpret_cast<double *>(wb + cfg.off_part);      // [S slots][G]
    int32_t *orow = reinterpret_cast<int32_t *>(wb + cfg.off_orow);    // [S slots] output row or -1
    double *Llast = reinterpret_cast<double *>(wb + cfg.off_llast);    // [S slots][M] Lagrange values of the last axis
    unsigned char *fbuf = wb + cfg.off_fbuf;                           // [ng][fb_gstride]
    uint64_t *bar = reinterpret_cast<uint64_t *>(wb + cfg.off_bar);
    if (lane == 0) mbar_init(bar, 1);
    fence_mbar_init();
    __syncwarp();
    // slot s of the Lagrange table: rows of group g start 32 bytes (4 doubles) further per group, so that the groups'
    // simultaneous 128-bit reads of "their" slot fall into different banks
    auto ls_at = [&](int slot, int g) { return Ls + (size_t)slot * cfg.ls_stride + g * 4; };

    // roles: phase A -- slot `lane` (< S) = (group ga, point qa of the chunk);  phase B -- (group gb, field fi, slab kk)
    const int ga = lane / PCg, qa = lane - ga * PCg;
    const int gb = lane / G, rb = lane - gb * G, fi = rb / M, kk = rb - fi * M;
    const bool role_a = lane < S && ga < ng, role_b = gb < ng;
    const int64_t groups_total = (int64_t)gridDim.x * IE_WARPS * ng;
    const int64_t group0 = ((int64_t)blockIdx.x * IE_WARPS + warp) * ng;
    const int npass = (F + Fc - 1) / Fc;
    const int64_t total_bytes = E * (int64_t)F * P * 8;
    if (group0 >= E) return;
    uint32_t phase = 0;

    // stage the field chunks of unit u (lane g < ng copies the chunk of group g, if its element has any points)
    auto issue = [&](int64_t ebase_u, int f0) {
        const int64_t e = ebase_u + lane;
        const int nf = min(Fc, F - f0);
        bool want = lane < ng && e < E;
        if (want) want = starts[e + 1] > starts[e];
        int64_t off = 0;
        int shift = 0, bytes = 0;
        bool use_tma = false;
        if (want) {
            off = (((int64_t)e * F + f0) * P) * 8;
            shift = (int)(off & 8);
            bytes = (nf * P * 8 + shift + 15) / 16 * 16;
            use_tma = off - shift + bytes <= total_bytes;
        }
        const unsigned tx = __reduce_add_sync(0xffffffffu, use_tma ? (unsigned)bytes : 0u);  // ng <= 16 copies
        // always arrive (tx may be 0), so that every unit completes exactly one phase of the barrier
        if (lane == 0) mbar_arrive_expect_tx(bar, tx);
        __syncwarp();
        unsigned char *dst = fbuf + (size_t)lane * cfg.fb_gstride;
        if (use_tma) {
            bulk_copy_g2s(dst, reinterpret_cast<const unsigned char *>(fields) + off - shift, (uint32_t)bytes, bar);
        } else if (want) {  // last bytes of the array: plain loads
            const double *src = reinterpret_cast<const double *>(reinterpret_cast<const unsigned char *>(fields) + off);
            double *d2 = reinterpret_cast<double *>(dst + shift);
            for (int q = 0; q < nf * P; ++q) d2[q] = src[q];
        }
    };

    // first batch at or after `from` in which at least one of the warp's groups has points (E when there is none):
    // batches without points cost two loads and a vote, not a trip through the pipeline (under strong scaling a
    // rank owns points in a fraction of the elements only)
    auto next_batch = [&](int64_t from) -> int64_t {
        for (int64_t b = from; b < E; b += groups_total) {
            const int64_t e = b + lane;
            const bool any = lane < ng && e < E && starts[e + 1] > starts[e];
            if (__ballot_sync(0xffffffffu, any)) return b;
        }
        return E;
    };
    int64_t ebase = next_batch(group0);
    if (ebase >= E) return;
    issue(ebase, 0);
    int64_t enext = -1;  // the batch after `ebase`, looked up when the last pass of `ebase` starts
    for (int pass = 0; ebase < E; ++pass) {
        if (pass == npass) {
            pass = 0;
            ebase = enext;
            if (ebase >= E) break;
        }
        const int f0 = pass * Fc;
        const int nf = min(Fc, F - f0);
        const int rcp_nf = (65536 + nf - 1) / nf;
        const int64_t ea = ebase + ga, eb = ebase + gb;
        int sa = 0, ca = 0, cb = 0;
        if (role_a && ea < E) {
            sa = starts[ea];
            ca = starts[ea + 1] - sa;
        }
        if (role_b && eb < E) cb = starts[eb + 1] - starts[eb];
        const int cmax = __reduce_max_sync(0xffffffffu, ca);
        // this lane's slab of its element's field chunk: registers for all points of the element
        mbar_wait(bar, phase);
        phase ^= 1;
        __syncwarp();  // (plain-load tail path: the copying lane's stores are visible)
        double v[R];
        const bool have_b = role_b && cb > 0 && f0 + fi < F;
        if (have_b) {
            const int shift = (int)(((((int64_t)eb * F + f0) * P) * 8) & 8);
            const double *src = reinterpret_cast<const double *>(fbuf + (size_t)gb * cfg.fb_gstride + shift) +
                                (size_t)fi * P + (size_t)kk * R;
#pragma unroll
            for (int a = 0; a < R; ++a) v[a] = src[a];
        } else {
#pragma unroll
            for (int a = 0; a < R; ++a) v[a] = 0.0;
        }
        // the buffer is free: order the generic-proxy reads before the async-proxy writes of the next unit's copies
        fence_proxy_async_smem();
        __syncwarp();
        if (pass + 1 < npass) {
            issue(ebase, f0 + Fc);
        } else {
            enext = next_batch(ebase + groups_total);
            if (enext < E) issue(enext, 0);
        }
        if (cmax == 0) continue;
        const int nchunk = (cmax + PCg - 1) / PCg;
        for (int c = 0; c < nchunk; ++c) {
            // ---- phase A: Lagrange values of the chunk's points ------------------------------------------------
            const int left_a = ca - c * PCg;
            if (pass == 0 || nchunk > 1) {  // (several chunks AND several field passes: the values are recomputed)
                __syncwarp();  // phase C of the previous chunk is done with Ls / orow
                int32_t row = -1;
                if (role_a && qa < left_a) {
                    const int2 rec = erec[sa + c * PCg + qa];
                    row = rec.y;
                    double x[DIM];
#pragma unroll
                    for (int ax = 0; ax < DIM; ++ax) x[ax] = xi_s[(int64_t)rec.x * DIM + ax];
                    if (elem_u && pass == 0) {  // fused un-permute of the location outputs
                        elem_u[row] = (int32_t)ea;
                        if (status_u) status_u[row] = status_s[rec.x];
                        if (xi_u) {
#pragma unroll
                            for (int ax = 0; ax < DIM; ++ax) xi_u[(int64_t)row * DIM + ax] = x[ax];
                        }
                    }
                    double *dstL = ls_at(lane, ga);
#pragma unroll
                    for (int ax = 0; ax < DIM; ++ax) {
                        double L[M];
                        lagrange_values<ORDER>(T, x[ax], L);
#pragma unroll
                        for (int m = 0; m < M; ++m) dstL[ax * MP + m] = L[m];
                        if (MP > M) dstL[ax * MP + M] = 0.0;
                        if (ax == DIM - 1) {  // the last axis once more, un-skewed: phase C addresses it by slot alone
#pragma unroll
                            for (int m = 0; m < M; ++m) Llast[lane * M + m] = L[m];
                        }
                    }
                }
                if (lane < S) orow[lane] = row;
            }
            __syncwarp();
            // ---- phase B: contract this lane's slab for every point of its group's chunk ---------------------
            const int left_b = min(cb - c * PCg, PCg);
            if (have_b) {
                for (int q = 0; q < left_b; ++q) {
                    const int slot = gb * PCg + q;
                    const double2 *l2 = reinterpret_cast<const double2 *>(ls_at(slot, gb));
                    double L0[MP];
#pragma unroll
                    for (int m = 0; m < MP / 2; ++m) {
                        const double2 t2 = l2[m];
                        L0[2 * m] = t2.x;
                        L0[2 * m + 1] = t2.y;
                    }
                    double uacc;
                    if constexpr (DIM == 2) {
                        uacc = 0.0;
#pragma unroll
                        for (int i = 0; i < M; ++i) uacc = __fma_rn(L0[i], v[i], uacc);
                    } else {
                        double L1[MP];
#pragma unroll
                        for (int m = 0; m < MP / 2; ++m) {
                            const double2 t2 = l2[MP / 2 + m];
                            L1[2 * m] = t2.x;
                            L1[2 * m + 1] = t2.y;
                        }
                        uacc = 0.0;
#pragma unroll
                        for (int j = 0; j < M; ++j) {
                            double t = 0.0;
#pragma unroll
                            for (int i = 0; i < M; ++i) t = __fma_rn(L0[i], v[i + M * j], t);
                            uacc = __fma_rn(L1[j], t, uacc);
                        }
                    }
                    part[(size_t)slot * G + rb] = uacc;
                }
            }
            __syncwarp();
            // ---- phase C: the last axis, lanes = (slot, field) ------------------------------------------------
            for (int o = lane; o < S * nf; o += 32) {
                const int slot = (o * rcp_nf) >> 16, f = o - slot * nf;  // o / nf (exact: o < 512, nf <= 16)
                const int32_t row = orow[slot];
                if (row < 0) continue;
                const double *Ll = Llast + slot * M;
                const double *pu = part + (size_t)slot * G + f * M;
                double acc = 0.0;
#pragma unroll
                for (int k = 0; k < M; ++k) acc = __fma_rn(Ll[k], pu[k], acc);
                out[(int64_t)row * F + f0 + f] = acc;
            }
            if (c + 1 < nchunk) __syncwarp();  // before `part` is overwritten (the next unit syncs on its own)
        }
    }
}

int ie_blocks(int64_t work)
{
    const int sms = mm_num_sms() > 0 ? mm_num_sms() : 148;
    const int64_t need = (work + 255) / 256, cap = (int64_t)sms * 16;
    return (int)(need < 1 ? 1 : (need > cap ? cap : need));
}

struct ie_layout {
    size_t counts, starts, tiles, erank, erec, total;
};

ie_layout ie_make_layout(int64_t E, int64_t N)
{
    ie_layout L{};
    size_t o = 0;
    auto take = [&](size_t b) { size_t at = o; o += (b + 255) & ~(size_t)255; return at; };
    L.counts = take(sizeof(int32_t) * (size_t)(E + 1));
    L.starts = take(sizeof(int32_t) * (size_t)(E + 1));
    L.tiles = take(sizeof(int32_t) * (size_t)(mm_scan_tiles(E) + 1));
    L.erank = take(sizeof(int32_t) * (size_t)N);
    L.erec = take(sizeof(int2) * (size_t)N);
    L.total = o;
    return L;
}

template <int ORDER, int DIM>
int launch_interp_elem(int64_t E, int F, const double *fields, const int32_t *starts, const int2 *erec,
                       const double *xi_s, const uint8_t *status_s, double *out, int32_t *elem_u, double *xi_u,
                       uint8_t *status_u, cudaStream_t stream)
{
    mm_gll_table T;
    mm_make_table(ORDER, &T);
    const ie_cfg cfg = ie_make_cfg<ORDER, DIM>(F);
    const size_t smem = (size_t)cfg.per_warp * IE_WARPS;
    MM_REQUIRE(smem <= 227 * 1024, MM_ERR_UNSUPPORTED, "mm_interpolate: K3 needs %zu B of shared memory", smem);
    auto kern = interp_elem_kernel<ORDER, DIM>;
    static mm_kernel_cfg kcfg;
    int per_sm = 1;
    MM_CUDA(kcfg.prepare(kern, IE_WARPS * 32, smem, &per_sm));
    const int sms = mm_num_sms() > 0 ? mm_num_sms() : 148;
    const int64_t need = (E + (int64_t)IE_WARPS * cfg.ng - 1) / ((int64_t)IE_WARPS * cfg.ng);
    int64_t grid = std::min<int64_t>((int64_t)sms * per_sm, need);
    if (grid < 1) grid = 1;
    kern<<<(int)grid, IE_WARPS * 32, smem, stream>>>(T, cfg, E, fields, starts, erec, xi_s, status_s, out, elem_u, xi_u,
                                                     status_u);
    MM_CUDA(cudaGetLastError());
    return MM_OK;
}

}  // namespace

size_t mm_interp_elem_scratch_bytes(int64_t E, int64_t N) { return ie_make_layout(E, N).total; }

// K3 of the fused pipeline, element-centric (see the header of this file).  `scratch`: mm_interp_elem_scratch_bytes.
int mm_interp_by_element(int order, int dim, int64_t E, int F, const double *fields, int64_t N,
                         const int32_t *elem_s, const double *xi_s, const uint8_t *status_s, const int32_t *perm,
                         int perm_stride, double *out, int32_t *elem_u, double *xi_u, uint8_t *status_u,
                         void *scratch, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    MM_REQUIRE(mm_valid_order(order) && (dim == 2 || dim == 3), MM_ERR_INVALID, "mm_interpolate: order/dim");
    MM_REQUIRE(E >= 0 && E < (int64_t)INT32_MAX && N >= 0 && N <= (int64_t)INT32_MAX && F >= 1, MM_ERR_INVALID,
               "mm_interpolate: sizes");
    if (N == 0) return MM_OK;
    MM_REQUIRE(((uintptr_t)fields & 15) == 0, MM_ERR_INVALID, "mm_interpolate: fields must be 16-byte aligned");
    MM_REQUIRE(scratch && ((uintptr_t)scratch & 255) == 0, MM_ERR_INVALID, "mm_interpolate: scratch");
    const ie_layout L = ie_make_layout(E, N);
    unsigned char *ws = static_cast<unsigned char *>(scratch);
    int32_t *counts = reinterpret_cast<int32_t *>(ws + L.counts);
    int32_t *starts = reinterpret_cast<int32_t *>(ws + L.starts);
    int32_t *tiles = reinterpret_cast<int32_t *>(ws + L.tiles);
    int32_t *erank = reinterpret_cast<int32_t *>(ws + L.erank);
    int2 *erec = reinterpret_cast<int2 *>(ws + L.erec);
    MM_CUDA(cudaMemsetAsync(counts, 0, sizeof(int32_t) * (size_t)(E + 1), stream));
    elem_rank_kernel<<<ie_blocks(N), 256, 0, stream>>>(dim, N, E, F, elem_s, xi_s, status_s, perm, perm_stride, counts,
                                                       erank, out, elem_u, xi_u, status_u);
    MM_CUDA(cudaGetLastError());
    if (E == 0) return MM_OK;  // nothing can be located: every point took the zero row above
    mm_exclusive_scan_i32(E, counts, starts, tiles, stream);  // starts[E] = located points
    elem_place_kernel<<<ie_blocks(N), 256, 0, stream>>>(N, elem_s, perm, perm_stride, starts, erank, erec);
    MM_CUDA(cudaGetLastError());
#define MM_IE(O, D)                                                                                              \
    if (order == O && dim == D)                                                                                  \
        return launch_interp_elem<O, D>(E, F, fields, starts, erec, xi_s, status_s, out, elem_u, xi_u, status_u, \
                                        stream);
    MM_IE(1, 2) MM_IE(2, 2) MM_IE(4, 2) MM_IE(1, 3) MM_IE(2, 3) MM_IE(4, 3)
#undef MM_IE
    mm_set_error("mm_interpolate: unsupported order/dim");
    return MM_ERR_UNSUPPORTED;
}

int mm_group_by_key(int64_t E, int64_t N, const int32_t *key, int key_stride, void *scratch, int32_t **order,
                    void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    MM_REQUIRE(E >= 0 && N >= 0 && key && scratch && order, MM_ERR_INVALID, "mm_group_by_key: arguments");
    const ie_layout L = ie_make_layout(E, N);  // counts [E + 1] | starts [E + 1] | tiles | erank [N] | erec [8 N bytes]
    unsigned char *ws = static_cast<unsigned char *>(scratch);
    int32_t *counts = reinterpret_cast<int32_t *>(ws + L.counts);
    int32_t *starts = reinterpret_cast<int32_t *>(ws + L.starts);
    int32_t *tiles = reinterpret_cast<int32_t *>(ws + L.tiles);
    int32_t *rank = reinterpret_cast<int32_t *>(ws + L.erank);
    *order = reinterpret_cast<int32_t *>(ws + L.erec);
    if (N == 0) return MM_OK;
    // E + 1 buckets (the last one: no usable key); the scan tables hold E + 1 counts and E + 1 starts -- the total
    // (starts[E + 1]) is not needed
    MM_CUDA(cudaMemsetAsync(counts, 0, sizeof(int32_t) * (size_t)(E + 1), stream));
    key_rank_kernel<<<ie_blocks(N), 256, 0, stream>>>(N, E, key, key_stride, counts, rank);
    // exclusive scan of E + 1 counts writes E + 2 values: the starts table has E + 1 entries, so scan E counts (their
    // total lands in starts[E], which is exactly the start of the last bucket)
    if (E > 0) {
        mm_exclusive_scan_i32(E, counts, starts, tiles, stream);
    } else {
        MM_CUDA(cudaMemsetAsync(starts, 0, sizeof(int32_t), stream));
    }
    key_place_kernel<<<ie_blocks(N), 256, 0, stream>>>(N, E, key, key_stride, starts, rank, *order);
    MM_CUDA(cudaGetLastError());
    return MM_OK;
}
