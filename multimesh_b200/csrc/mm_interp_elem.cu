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

constexpr int IE_WARPS = 8;

template <int ORDER, int DIM>
__global__ void __launch_bounds__(IE_WARPS * 32, 2)
interp_elem_kernel(const mm_gll_table T, int64_t E, int F, int Fc, const double *__restrict__ fields,
                   const int32_t *__restrict__ starts, const int2 *__restrict__ erec,
                   const double *__restrict__ xi_s, const uint8_t *__restrict__ status_s, double *__restrict__ out,
                   int32_t *__restrict__ elem_u, double *__restrict__ xi_u, uint8_t *__restrict__ status_u)
{
    constexpr int M = ORDER + 1;
    constexpr int P = DIM == 2 ? M * M : M * M * M;
    constexpr int R = DIM == 2 ? M : M * M;  // values a lane keeps: one slab of the last axis
    constexpr int MP = M + (M & 1);          // Lagrange rows padded to 16 bytes
    const int G = Fc * M;                    // lanes per group
    const int ng = 32 / G;                   // groups (elements) per warp
    const int PCg = 32 / ng;                 // points per group and chunk
    extern __shared__ __align__(16) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t ls_bytes = (size_t)32 * DIM * MP * 8, part_bytes = (size_t)32 * G * 8;
    unsigned char *wb = smem + (size_t)warp * (ls_bytes + part_bytes + 128);
    double *Ls = reinterpret_cast<double *>(wb);                 // [32 slots][DIM][MP]
    double *part = reinterpret_cast<double *>(wb + ls_bytes);    // [32 slots][G]
    int32_t *orow = reinterpret_cast<int32_t *>(wb + ls_bytes + part_bytes);  // [32 slots] output row or -1

    // roles: phase A / C -- slot `lane` = (group ga, point q of the chunk);  phase B -- (group gb, field fi, slab kk)
    const int ga = lane / PCg, qa = lane - ga * PCg;
    const int gb = lane / G, rb = lane - gb * G, fi = rb / M, kk = rb - fi * M;
    const bool role_a = ga < ng, role_b = gb < ng;
    const int64_t groups_total = (int64_t)gridDim.x * IE_WARPS * ng;
    const int64_t group0 = ((int64_t)blockIdx.x * IE_WARPS + warp) * ng;

    for (int64_t ebase = group0; ebase < E; ebase += groups_total) {
        const int64_t ea = ebase + ga, eb = ebase + gb;
        int sa = 0, ca = 0, cb = 0;
        if (role_a && ea < E) {
            sa = starts[ea];
            ca = starts[ea + 1] - sa;
        }
        if (role_b && eb < E) cb = starts[eb + 1] - starts[eb];
        const int cmax = __reduce_max_sync(0xffffffffu, ca);
        if (cmax == 0) continue;
        const int nchunk = (cmax + PCg - 1) / PCg;
        for (int f0 = 0; f0 < F; f0 += Fc) {
            // this lane's slab of its element's field block: registers for all points of the element
            double v[R];
            const bool have_b = role_b && cb > 0 && f0 + fi < F;
            if (have_b) {
                const double *src = fields + (((int64_t)eb * F + f0 + fi) * P + (int64_t)kk * R);
#pragma unroll
                for (int a = 0; a < R; ++a) v[a] = __ldg(src + a);
                if (f0 == 0 && eb + groups_total < E) {  // the group's next element: towards L2 now
                    const double *nx = fields + (((int64_t)(eb + groups_total) * F + fi) * P + (int64_t)kk * R);
                    prefetch_l2(nx);
                    if (R * 8 > 128) prefetch_l2(nx + 16);
                }
            } else {
#pragma unroll
                for (int a = 0; a < R; ++a) v[a] = 0.0;
            }
            for (int c = 0; c < nchunk; ++c) {
                // ---- phase A: Lagrange values of the chunk's points ------------------------------------------------
                const int left_a = ca - c * PCg;
                if (f0 == 0 || nchunk > 1) {  // (several chunks AND several field passes: the values are recomputed)
                    __syncwarp();  // phase C of the previous chunk is done with Ls / orow
                    int32_t row = -1;
                    if (role_a && qa < left_a) {
                        const int2 rec = erec[sa + c * PCg + qa];
                        row = rec.y;
                        double x[DIM];
#pragma unroll
                        for (int ax = 0; ax < DIM; ++ax) x[ax] = xi_s[(int64_t)rec.x * DIM + ax];
                        if (elem_u && f0 == 0) {  // fused un-permute of the location outputs
                            elem_u[row] = (int32_t)ea;
                            if (status_u) status_u[row] = status_s[rec.x];
                            if (xi_u) {
#pragma unroll
                                for (int ax = 0; ax < DIM; ++ax) xi_u[(int64_t)row * DIM + ax] = x[ax];
                            }
                        }
#pragma unroll
                        for (int ax = 0; ax < DIM; ++ax) {
                            double L[M];
                            lagrange_values<ORDER>(T, x[ax], L);
#pragma unroll
                            for (int m = 0; m < M; ++m) Ls[(lane * DIM + ax) * MP + m] = L[m];
                            if (MP > M) Ls[(lane * DIM + ax) * MP + M] = 0.0;
                        }
                    }
                    orow[lane] = row;
                }
                __syncwarp();
                // ---- phase B: contract this lane's slab for every point of its group's chunk ---------------------
                const int left_b = min(cb - c * PCg, PCg);
                if (have_b) {
                    for (int q = 0; q < left_b; ++q) {
                        const int slot = gb * PCg + q;
                        const double2 *l2 = reinterpret_cast<const double2 *>(Ls + (size_t)slot * DIM * MP);
                        double L0[MP];
#pragma unroll
                        for (int m = 0; m < MP / 2; ++m) {
                            const double2 t2 = l2[m];
                            L0[2 * m] = t2.x;
                            L0[2 * m + 1] = t2.y;
                        }
                        double u;
                        if constexpr (DIM == 2) {
                            u = 0.0;
#pragma unroll
                            for (int i = 0; i < M; ++i) u = __fma_rn(L0[i], v[i], u);
                        } else {
                            double L1[MP];
#pragma unroll
                            for (int m = 0; m < MP / 2; ++m) {
                                const double2 t2 = l2[MP / 2 + m];
                                L1[2 * m] = t2.x;
                                L1[2 * m + 1] = t2.y;
                            }
                            u = 0.0;
#pragma unroll
                            for (int j = 0; j < M; ++j) {
                                double t = 0.0;
#pragma unroll
                                for (int i = 0; i < M; ++i) t = __fma_rn(L0[i], v[i + M * j], t);
                                u = __fma_rn(L1[j], t, u);
                            }
                        }
                        part[(size_t)slot * G + rb] = u;
                    }
                }
                __syncwarp();
                // ---- phase C: the last axis, lanes = (slot, field) ------------------------------------------------
                const int nf = min(Fc, F - f0);
                for (int o = lane; o < 32 * nf; o += 32) {
                    const int slot = o / nf, f = o - slot * nf;
                    const int32_t row = orow[slot];
                    if (row < 0) continue;
                    const double *Ll = Ls + ((size_t)slot * DIM + (DIM - 1)) * MP;
                    const double *pu = part + (size_t)slot * G + f * M;
                    double acc = 0.0;
#pragma unroll
                    for (int k = 0; k < M; ++k) acc = __fma_rn(Ll[k], pu[k], acc);
                    out[(int64_t)row * F + f0 + f] = acc;
                }
                if (f0 + Fc < F || c + 1 < nchunk) __syncwarp();  // before `part` is overwritten
            }
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
    constexpr int M = ORDER + 1;
    constexpr int MP = M + (M & 1);
    mm_gll_table T;
    mm_make_table(ORDER, &T);
    const int Fc = std::min(F, 32 / M);
    const int G = Fc * M, ng = 32 / G;
    const size_t per_warp = (size_t)32 * DIM * MP * 8 + (size_t)32 * G * 8 + 128;
    const size_t smem = per_warp * IE_WARPS;
    auto kern = interp_elem_kernel<ORDER, DIM>;
    static mm_kernel_cfg kcfg;
    int per_sm = 1;
    MM_CUDA(kcfg.prepare(kern, IE_WARPS * 32, smem, &per_sm));
    const int sms = mm_num_sms() > 0 ? mm_num_sms() : 148;
    const int64_t need = (E + (int64_t)IE_WARPS * ng - 1) / ((int64_t)IE_WARPS * ng);
    int64_t grid = std::min<int64_t>((int64_t)sms * per_sm, need);
    if (grid < 1) grid = 1;
    kern<<<(int)grid, IE_WARPS * 32, smem, stream>>>(T, E, F, Fc, fields, starts, erec, xi_s, status_s, out, elem_u,
                                                     xi_u, status_u);
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
