// mm_locate.cu -- K2: batched point-in-element location (fp64 Newton inverse map + accept/fallback).
//
// One thread per target point, warp-synchronous rounds:
//   1. every unresolved lane advances through its candidate list (skipping negatives, repeats
//      and -- for V1 -- candidates whose node AABB excludes the point) to its next candidate;
//   2. the warp de-duplicates the candidate elements of its lanes (__match_any_sync): each
//      DISTINCT element's control nodes ([P][dim] f64, contiguous in HBM) are staged ONCE, by the
//      group's leader lane, with one bulk-async copy (cp.async.bulk -> UBLKCP) into one of SLOTS
//      shared-memory slots, completion tracked by a per-warp mbarrier; groups beyond SLOTS wait
//      for the next round.  With spatially sorted points (mm_interpolate) a warp needs 1-4 copies;
//   3. each lane runs Newton on the order-n map with the sum-factorised canonical evaluation
//      order, reading its element's nodes from the shared slot (lanes of one group read the same
//      addresses: broadcast) and shifting them by its own point on the fly (Y = X - p);
//   4. accept test / best tracking; lanes that run out of candidates take the variant's fallback.
// Rounds repeat until every lane is resolved (~1-2 rounds typically).
//
// Arithmetic = DESIGN.md section 3 (fixed order; the only fused operations are the explicit __fma_rn of the
// map contraction, mm_newton.cuh); restated independently by oracle/mm_oracle.c.
#include <cstdlib>

#include "mm_common.cuh"
#include "mm_newton.cuh"

namespace {

template <int ORDER, int DIM>
struct elem_traits {
    static constexpr int M = ORDER + 1;
    static constexpr int P = DIM == 2 ? M * M : M * M * M;
    static constexpr int DOUBLES = P * DIM;
    static constexpr int BYTES = DOUBLES * 8;
    static constexpr bool MISALIGNED = (BYTES % 16) != 0;  // odd elements start at 8 mod 16
    static constexpr int COPY_BYTES = MISALIGNED ? BYTES + 8 : BYTES;
    static constexpr int SLOT_BYTES = ((COPY_BYTES + 15) / 16) * 16;
};

template <int DIM>
__device__ __forceinline__ bool accept_xi(const mm_locate_params &prm, const double (&xi)[DIM])
{
    bool ok = true;
#pragma unroll
    for (int c = 0; c < DIM; ++c) {
        double a = fabs(xi[c]);
        if (prm.strict ? !(a < prm.tol) : !(a <= prm.tol)) ok = false;
    }
    return ok;
}

// STATS: also counts, over all points, the candidates that reached Newton and the map evaluations (Newton
// iterations) -- stats[0] += candidates, stats[1] += evaluations; a separate instantiation, used by the
// benchmarks only (mm_locate_set_stats)
// PREFIX: the first pass of the progressive search (mm_pipeline.cu) -- accept only, no fallback; a compile-time
// switch, so that none of the fallback bookkeeping (first AABB hit, nearest centre, best snap candidate) occupies
// registers in the pass that handles (almost) all points: the order-2 kernel sits at its 128-register cap.
// ORDERED: lane m works on point order[m] instead of point m (the pipeline groups the points by their first
// candidate element: the lanes of a warp then share a few elements whatever the point order is).
template <int ORDER, int DIM, int WARPS, int SLOTS, int MINB, bool STATS, bool PREFIX, bool ORDERED>
__global__ void __launch_bounds__(WARPS * 32, MINB)
locate_kernel(const mm_gll_table T, const mm_locate_params prm, int64_t E,
              const double *__restrict__ nodes, const double *__restrict__ centroid,
              const double *__restrict__ aabb, const double *__restrict__ presolve, int64_t N,
              const double *__restrict__ pts, int pstride, int k,
              const int32_t *__restrict__ cands, int32_t *__restrict__ elem_out,
              double *__restrict__ xi_out, uint8_t *__restrict__ status_out,
              unsigned long long *__restrict__ num_failed, int32_t *__restrict__ unresolved_list,
              unsigned long long *__restrict__ unresolved_count, const long long *__restrict__ n_dev,
              int64_t n_off, unsigned long long *__restrict__ stats, const int32_t *__restrict__ order)
{
    using tr = elem_traits<ORDER, DIM>;
    int st_cand = 0, st_eval = 0;
    if (n_dev) {  // point count known only on the device: [n_off, *n_dev)
        const long long have = *n_dev - n_off;
        N = have < 0 ? 0 : (have < N ? have : N);
    }
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char *wslots = smem + (size_t)warp * SLOTS * tr::SLOT_BYTES;
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + (size_t)WARPS * SLOTS * tr::SLOT_BYTES) + warp;
    if (lane == 0) mbar_init(bar, 1);
    fence_mbar_init();
    __syncthreads();
    uint32_t phase = 0;
    const int64_t total_bytes = E * (int64_t)tr::BYTES;
    unsigned long long failed_local = 0;

    const int64_t warps_total = (int64_t)gridDim.x * WARPS;
    for (int64_t batch = (int64_t)blockIdx.x * WARPS + warp; batch * 32 < N; batch += warps_total) {
        const int64_t m = batch * 32 + lane;
        const bool valid = m < N;
        bool done = !valid;
        const int64_t n = ORDERED ? (valid ? (int64_t)order[m] : 0) : m;
        if (!ORDERED) {  // the warp's next batch: pull its points and candidate rows towards L2 now
            const int64_t nn = n + warps_total * 32;
            if (nn < N) {
                prefetch_l2(pts + nn * pstride);
                prefetch_l2(cands + nn * (int64_t)k);
            }
        }
        double p[DIM];
#pragma unroll
        for (int c = 0; c < DIM; ++c) p[c] = done ? 0.0 : pts[n * pstride + c];
        const int32_t *cl = cands + (done ? 0 : n * (int64_t)k);
        int t = 0;
        int32_t r_elem = -1;
        uint8_t r_status = MM_ST_FAILED;
        double r_xi[DIM];
#pragma unroll
        for (int c = 0; c < DIM; ++c) r_xi[c] = 0.0;
        int32_t first_inside = -1, near_elem = -1, best_elem = -1;
        bool first_inside_nan = false;  // V1: Newton failed on the first AABB hit (the reference re-inverts it, :1460)
        double near_dist = INFINITY;
        double best_key = prm.fallback == MM_FB_SNAP ? 10e9 : INFINITY;
        double best_xi[DIM];
#pragma unroll
        for (int c = 0; c < DIM; ++c) best_xi[c] = 0.0;

        int32_t deferred_e = -1;  // candidate selected but not staged yet (more groups than SLOTS)
        bool deferred_fb = false;
        while (true) {
            int32_t e = -1;
            bool fb_newton = false;
            if (deferred_e >= 0) {
                e = deferred_e;
                fb_newton = deferred_fb;
                deferred_e = -1;
            } else if (!done) {
                while (t < k) {
                    int32_t c = cl[t];
                    ++t;
                    if (c < 0 || c >= E) continue;  // padding (-1) or an id the caller got wrong
                    bool dup = false;
                    for (int u = 0; u < t - 1; ++u) dup = dup || (cl[u] == c);
                    if (dup) continue;
                    if (prm.aabb_prefilter) {
                        const double *lo = aabb + ((int64_t)c * 2 + 0) * DIM;
                        const double *hi = lo + DIM;
                        bool inside = true;
#pragma unroll
                        for (int q = 0; q < DIM; ++q)
                            if (!(p[q] >= lo[q] && p[q] <= hi[q])) inside = false;
                        if (!inside) {
                            // prefix mode takes no fallback (unresolved points are re-run with the
                            // full list), so the nearest-centre bookkeeping is not needed there
                            if (PREFIX) continue;
                            const double *cc = centroid + (int64_t)c * DIM;
                            double dx = p[0] - cc[0], dy = p[1] - cc[1];
                            double s = dx * dx + dy * dy;
                            if constexpr (DIM == 3) {
                                double dz = p[2] - cc[2];
                                s = s + dz * dz;
                            }
                            double d = sqrt(s);
                            if (d < near_dist) {
                                near_dist = d;
                                near_elem = c;
                            }
                            continue;
                        }
                        if (!PREFIX && first_inside < 0) first_inside = c;
                    }
                    e = c;
                    break;
                }
                if (e < 0) {  // candidates exhausted without acceptance: fallback
                    done = true;
                    if (PREFIX) {
                        // first pass of the progressive search: the list is only a PREFIX of the
                        // k-NN list, so no fallback may be taken; the point is re-run with all k
                        r_status = MM_ST_UNRESOLVED;
                    } else if (prm.fallback == MM_FB_MAGIC) {
                        if (first_inside >= 0) {
                            r_elem = first_inside;
                            r_status = first_inside_nan ? MM_ST_FB_NAN_MAGIC : MM_ST_FB_INSIDE_MAGIC;
#pragma unroll
                            for (int c = 0; c < DIM; ++c) r_xi[c] = prm.magic_xi[c];
                        } else if (near_elem >= 0) {
                            e = near_elem;
                            fb_newton = true;
                            done = false;
                        }
                    } else if (prm.fallback == MM_FB_SNAP) {
                        if (best_elem >= 0) {
                            r_elem = best_elem;
                            r_status = MM_ST_SNAPPED;
#pragma unroll
                            for (int c = 0; c < DIM; ++c) {
                                double v = best_xi[c];
                                if (v < -prm.snap_clip) v = -prm.snap_clip;
                                if (v > prm.snap_clip) v = prm.snap_clip;
                                r_xi[c] = v;
                            }
                        } else {
                            r_elem = 0;
                            r_status = MM_ST_SNAP_NONE;
#pragma unroll
                            for (int c = 0; c < DIM; ++c) r_xi[c] = prm.snap_clip;
                        }
                    } else if (prm.fallback == MM_FB_MINL1) {
                        if (best_elem >= 0) {
                            r_elem = best_elem;
                            r_status = MM_ST_MINL1;
#pragma unroll
                            for (int c = 0; c < DIM; ++c) r_xi[c] = best_xi[c];
                        }
                    }
                }
            }
            const unsigned active = __ballot_sync(0xffffffffu, e >= 0);
            if (!active) break;

            // ---- stage each DISTINCT candidate element once: the group's leader lane copies ----
            const unsigned same = __match_any_sync(0xffffffffu, e);
            const int leader = __ffs(same) - 1;
            const unsigned leaders = __ballot_sync(0xffffffffu, e >= 0 && lane == leader);
            const int rank = e >= 0 ? __popc(leaders & ((1u << leader) - 1)) : -1;
            const bool served = e >= 0 && rank < SLOTS;
            if (e >= 0 && !served) {  // wait for a later round, keep the candidate
                deferred_e = e;
                deferred_fb = fb_newton;
            }
            int shift = 0;
            bool use_tma = false;
            const double *src = nodes;
            if (served) {
                const int64_t off = (int64_t)e * tr::BYTES;
                shift = tr::MISALIGNED ? (int)(off & 8) : 0;
                use_tma = lane == leader && (off - shift + tr::COPY_BYTES) <= total_bytes;
                src = nodes + (int64_t)e * tr::DOUBLES;
            }
            const unsigned tma_mask = __ballot_sync(0xffffffffu, use_tma);
            if (lane == 0 && tma_mask)
                mbar_arrive_expect_tx(bar, (uint32_t)__popc(tma_mask) * tr::COPY_BYTES);
            __syncwarp();
            unsigned char *slot = wslots + (size_t)(served ? rank : 0) * tr::SLOT_BYTES;
            const double *X = reinterpret_cast<const double *>(slot + shift);
            if (use_tma) {
                bulk_copy_g2s(slot, reinterpret_cast<const unsigned char *>(src) - shift,
                              tr::COPY_BYTES, bar);
            } else if (served && lane == leader) {  // array tail: plain loads
                double *dst = reinterpret_cast<double *>(slot + shift);
                for (int q = 0; q < tr::DOUBLES; ++q) dst[q] = src[q];
            }
            // Newton start value from the element's pre-solve row: needs no nodes, so its loads
            // overlap the bulk copy instead of queueing behind the barrier wait
            double x[DIM];
            if (served)
                newton_start<DIM>(p, presolve ? presolve + (int64_t)e * (2 * DIM + DIM * DIM) : nullptr, x);
            if (tma_mask) {
                mbar_wait(bar, phase);
                phase ^= 1;
            }
            __syncwarp();  // plain-load tail path: make the leader's stores visible to its group
            if (served) {
                if (STATS) ++st_cand;
                const bool ok = newton_iterate<ORDER, DIM>(T, X, p, x, STATS ? &st_eval : nullptr);
                if (!PREFIX && fb_newton) {  // V1: nearest-centre element, interpolator.py:1460-1473
                    bool big = false;
#pragma unroll
                    for (int c = 0; c < DIM; ++c)
                        if (ok && fabs(x[c]) >= prm.tol) big = true;
                    r_elem = e;
                    r_status = !ok ? MM_ST_FB_NAN_MAGIC : (big ? MM_ST_FB_NEAR_MAGIC : MM_ST_FB_NEAR_OK);
#pragma unroll
                    for (int c = 0; c < DIM; ++c)
                        r_xi[c] = (r_status == MM_ST_FB_NEAR_OK) ? x[c] : prm.magic_xi[c];
                    done = true;
                } else if (!ok) {
                    if (!PREFIX && e == first_inside) first_inside_nan = true;
                } else {
                    if (!PREFIX && (prm.fallback == MM_FB_SNAP || prm.fallback == MM_FB_MINL1)) {
                        double key = 0.0;
#pragma unroll
                        for (int c = 0; c < DIM; ++c) {
                            double a = fabs(x[c]);
                            if (prm.fallback == MM_FB_SNAP) key = a > key ? a : key;
                            else key = key + a;
                        }
                        if (key < best_key) {
                            best_key = key;
                            best_elem = e;
#pragma unroll
                            for (int c = 0; c < DIM; ++c) best_xi[c] = x[c];
                        }
                    }
                    if (accept_xi<DIM>(prm, x)) {
                        r_elem = e;
                        r_status = MM_ST_ACCEPTED;
#pragma unroll
                        for (int c = 0; c < DIM; ++c) r_xi[c] = x[c];
                        done = true;
                    }
                }
            }
            // generic-proxy reads/writes of the slots must be ordered before the next round's
            // async-proxy (bulk copy) writes into the same slots
            fence_proxy_async_smem();
            __syncwarp();
        }
        if (valid) {
            elem_out[n] = r_elem;
#pragma unroll
            for (int c = 0; c < DIM; ++c) xi_out[n * DIM + c] = r_elem < 0 ? 0.0 : r_xi[c];
            if (status_out) status_out[n] = r_status;
            if (r_elem < 0 && r_status != MM_ST_UNRESOLVED) failed_local += 1;
        }
        if (unresolved_list) {  // warp-aggregated append of the points that need the full search
            const unsigned um = __ballot_sync(0xffffffffu, valid && r_status == MM_ST_UNRESOLVED);
            if (um) {
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(unresolved_count, (unsigned long long)__popc(um));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (um & (1u << lane))
                    unresolved_list[base + __popc(um & ((1u << lane) - 1))] = (int32_t)n;
            }
        }
    }
    if (num_failed) {
        for (int o = 16; o > 0; o >>= 1) failed_local += __shfl_xor_sync(0xffffffffu, failed_local, o);
        if (lane == 0 && failed_local) atomicAdd(num_failed, failed_local);
    }
    if (STATS && stats) {
        for (int o = 16; o > 0; o >>= 1) {
            st_cand += __shfl_xor_sync(0xffffffffu, st_cand, o);
            st_eval += __shfl_xor_sync(0xffffffffu, st_eval, o);
        }
        if (lane == 0) {
            atomicAdd(stats, (unsigned long long)st_cand);
            atomicAdd(stats + 1, (unsigned long long)st_eval);
        }
    }
}

static thread_local unsigned long long *g_locate_stats = nullptr;

template <int ORDER, int DIM, int WARPS, int SLOTS, int MINB>
int launch_locate(const mm_locate_params &prm, int64_t E, const double *nodes,
                  const double *centroid, const double *aabb, const double *presolve, int64_t N,
                  const double *pts, int pstride, int k,
                  const int32_t *cands, int32_t *elem, double *xi, uint8_t *status,
                  int64_t *num_failed, int32_t *unresolved_list, int64_t *unresolved_count,
                  cudaStream_t stream, const int64_t *n_dev, int64_t n_off, const int32_t *order)
{
    using tr = elem_traits<ORDER, DIM>;
    mm_gll_table T;
    mm_make_table(ORDER, &T);
    unsigned long long *stats = g_locate_stats;
    const bool prefix = (prm.reserved & 1) != 0;
    MM_REQUIRE(!order || prefix, MM_ERR_INVALID, "mm_locate: an order array is a first-pass (prefix mode) option");
    auto kern = stats ? (prefix ? (order ? locate_kernel<ORDER, DIM, WARPS, SLOTS, MINB, true, true, true>
                                         : locate_kernel<ORDER, DIM, WARPS, SLOTS, MINB, true, true, false>)
                                : locate_kernel<ORDER, DIM, WARPS, SLOTS, MINB, true, false, false>)
                      : (prefix ? (order ? locate_kernel<ORDER, DIM, WARPS, SLOTS, MINB, false, true, true>
                                         : locate_kernel<ORDER, DIM, WARPS, SLOTS, MINB, false, true, false>)
                                : locate_kernel<ORDER, DIM, WARPS, SLOTS, MINB, false, false, false>);
    const size_t smem = (size_t)WARPS * SLOTS * tr::SLOT_BYTES + WARPS * sizeof(uint64_t);
    static mm_kernel_cfg kcfg[8];
    int per_sm = 1;
    MM_CUDA(kcfg[(stats ? 1 : 0) + (prefix ? 2 : 0) + (order ? 4 : 0)].prepare(kern, WARPS * 32, smem, &per_sm));
    const int sms = mm_num_sms() > 0 ? mm_num_sms() : 148;
    int64_t batches = (N + 32 * WARPS - 1) / (32 * WARPS);
    int64_t grid = (int64_t)sms * per_sm;  // persistent: resident CTAs loop over point batches
    if (grid > batches) grid = batches;
    if (grid < 1) grid = 1;
    kern<<<(int)grid, WARPS * 32, smem, stream>>>(T, prm, E, nodes, centroid, aabb, presolve, N, pts, pstride, k,
                                                  cands, elem, xi, status,
                                                  reinterpret_cast<unsigned long long *>(num_failed),
                                                  unresolved_list,
                                                  reinterpret_cast<unsigned long long *>(unresolved_count),
                                                  reinterpret_cast<const long long *>(n_dev), n_off, stats, order);
    MM_CUDA(cudaGetLastError());
    return MM_OK;
}

}  // namespace

int mm_locate_impl(int order, int dim, int64_t E, const double *nodes, const double *centroid,
                   const double *aabb, const double *presolve, int64_t N, const double *pts,
                   int pts_stride, int k,
                   const int32_t *cands,
                   const mm_locate_params *params, int32_t *elem, double *xi, uint8_t *status,
                   int64_t *num_failed, bool zero_num_failed, int32_t *unresolved_list,
                   int64_t *unresolved_count, void *stream_, const int64_t *n_dev, int64_t n_off,
                   const int32_t *point_order)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    MM_REQUIRE(mm_valid_order(order), MM_ERR_INVALID, "mm_locate: order %d (supported 1, 2, 4)", order);
    MM_REQUIRE(dim == 2 || dim == 3, MM_ERR_INVALID, "mm_locate: dim %d", dim);
    MM_REQUIRE(params, MM_ERR_INVALID, "mm_locate: null params");
    MM_REQUIRE(params->fallback >= MM_FB_FAIL && params->fallback <= MM_FB_MINL1, MM_ERR_INVALID,
               "mm_locate: fallback %d", (int)params->fallback);
    MM_REQUIRE(k >= 1, MM_ERR_INVALID, "mm_locate: k=%d", k);
    MM_REQUIRE(N >= 0 && E >= 0, MM_ERR_INVALID, "mm_locate: sizes");
    if (num_failed && zero_num_failed) MM_CUDA(cudaMemsetAsync(num_failed, 0, sizeof(int64_t), stream));
    if (N == 0) return MM_OK;
    MM_REQUIRE(nodes && pts && cands && elem && xi, MM_ERR_INVALID, "mm_locate: null buffer");
    MM_REQUIRE(((uintptr_t)nodes & 15) == 0, MM_ERR_INVALID,
               "mm_locate: nodes must be 16-byte aligned");
    MM_REQUIRE(!params->aabb_prefilter || (centroid && aabb), MM_ERR_INVALID,
               "mm_locate: aabb_prefilter needs centroid and aabb (mm_element_geometry)");
    // kernel configuration <order, dim, warps per CTA, shared slots per warp, min CTAs per SM>
#define MM_LOC(O, D, W, S, B)                                                                    \
    if (order == O && dim == D)                                                                  \
        return launch_locate<O, D, W, S, B>(*params, E, nodes, centroid, aabb, presolve, N, pts,     \
                                            pts_stride, k,                                        \
                                            cands,                                                \
                                            elem, xi, status, num_failed, unresolved_list,       \
                                            unresolved_count, stream, n_dev, n_off, point_order);
    MM_LOC(1, 2, 4, 8, 1)
    MM_LOC(2, 2, 4, 8, 1)
    MM_LOC(4, 2, 4, 8, 1)
    MM_LOC(1, 3, 4, 8, 1)
    MM_LOC(2, 3, 4, 8, 4)  // 128 registers, 16 warps per SM: measured best (profiles/)
    MM_LOC(4, 3, 2, 8, 1)
#undef MM_LOC
    mm_set_error("mm_locate: unsupported order/dim");
    return MM_ERR_UNSUPPORTED;
}

// benchmarks: make the calling thread's mm_locate / mm_interpolate launches accumulate {candidates that reached
// Newton, map evaluations} into a device array of two uint64 (NULL switches it off again)
extern "C" int mm_locate_set_stats(int64_t *device_counters)
{
    g_locate_stats = reinterpret_cast<unsigned long long *>(device_counters);
    return MM_OK;
}

extern "C" int mm_locate(int order, int dim, int64_t E, const double *nodes,
                         const double *centroid, const double *aabb, const double *presolve,
                         int64_t N, const double *pts, int k, const int32_t *cands,
                         const mm_locate_params *params,
                         int32_t *elem, double *xi, uint8_t *status, int64_t *num_failed,
                         void *stream)
{
    MM_REQUIRE(params, MM_ERR_INVALID, "mm_locate: null params");
    mm_locate_params prm = *params;
    prm.reserved = 0;  // the partial (progressive first pass) mode is internal to mm_interpolate
    return mm_locate_impl(order, dim, E, nodes, centroid, aabb, presolve, N, pts, dim, k, cands, &prm, elem, xi,
                          status, num_failed, true, nullptr, nullptr, stream);
}
