// mm_interp.cu -- K3, point-order kernels: fused Lagrange weights + multi-field gather for caller-supplied (elem, xi)
//                 (mm_interp / mm_interp_perm; the fused pipeline uses the element-centric mm_interp_elem.cu), plus
//                 coefficient write-out and the explicit-matrix gather for cached weights.
//
// out[n][f] = sum_a w_a(xi_n) * fields[elem_n][f][a]
//
// Data movement.  In the reference's MODEL/data layout [E][F][P] the F*P values a point needs are
// ONE contiguous block (1 080 B at order 2 / F=5, 5 000 B at order 4 / F=5).  Each lane owns one
// target point and fetches its block -- in chunks of FC fields -- with a single bulk-async copy
// (cp.async.bulk, SASS UBLKCP) into a private shared-memory slot.  A warp keeps S stages of 32
// slots in flight on per-stage mbarriers, so the copy engine streams whole contiguous blocks from
// HBM/L2 while the lanes contract the previous chunk out of shared memory.  No thread ever issues
// a scattered global load for field data.
//
// Arithmetic: nested tensor contraction, i innermost --
//   t[j,k] = sum_i Lx[i] v[i,j,k];  u[k] = sum_j Ly[j] t[j,k];  out = sum_k Lz[k] u[k]
// every accumulation one explicit fma; bit-identical to oracle/mm_oracle.c:mmo_interp.
#include <algorithm>
#include <cstdlib>

#include "mm_common.cuh"

namespace {

constexpr int INTERP_MAX_WARPS = 8;

struct interp_cfg {
    int F;           // fields
    int FC;          // fields per staged chunk
    int chunks;      // ceil(F / FC)
    int stages;      // pipeline depth per warp
    int slot_bytes;  // per-lane slot stride
    int odd_p;       // P odd: chunk starts alternate between 0 and 8 (mod 16)
    int warps;       // warps per CTA (each warp runs its own pipeline)
};

__host__ __device__ inline int chunk_copy_bytes(int nf, int P, int odd_p)
{
    int b = nf * P * 8;
    return odd_p ? ((b + 8 + 15) / 16) * 16 : b;
}

// element ids come from the caller: anything outside [0, E) is treated like -1 (failed point, zero row)
__device__ __forceinline__ int32_t valid_elem(int32_t e, int64_t E) { return (e >= 0 && e < E) ? e : -1; }

template <int ORDER, int DIM>
__device__ __forceinline__ double contract_field(const double *__restrict__ v,
                                                 const double (&L)[DIM][ORDER + 1])
{
    constexpr int M = ORDER + 1;
    if constexpr (DIM == 2) {
        double u = 0.0;
#pragma unroll
        for (int j = 0; j < M; ++j) {
            double t = 0.0;
#pragma unroll
            for (int i = 0; i < M; ++i) t = __fma_rn(L[0][i], v[i + M * j], t);
            u = __fma_rn(L[1][j], t, u);
        }
        return u;
    } else {
        double acc = 0.0;
#pragma unroll
        for (int k = 0; k < M; ++k) {
            double u = 0.0;
#pragma unroll
            for (int j = 0; j < M; ++j) {
                double t = 0.0;
#pragma unroll
                for (int i = 0; i < M; ++i) t = __fma_rn(L[0][i], v[i + M * j + M * M * k], t);
                u = __fma_rn(L[1][j], t, u);
            }
            acc = __fma_rn(L[2][k], u, acc);
        }
        return acc;
    }
}

template <int ORDER, int DIM>
__global__ void __launch_bounds__(INTERP_MAX_WARPS * 32)
interp_kernel(const mm_gll_table T, const interp_cfg cfg, int64_t E,
              const double *__restrict__ fields, int64_t N, const int32_t *__restrict__ elem,
              const double *__restrict__ xi, const int32_t *__restrict__ perm,
              double *__restrict__ out)
{
    constexpr int M = ORDER + 1;
    constexpr int P = DIM == 2 ? M * M : M * M * M;
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t stage_bytes = (size_t)32 * cfg.slot_bytes;
    unsigned char *wbase = smem + (size_t)warp * cfg.stages * stage_bytes;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)cfg.warps * cfg.stages * stage_bytes) +
                     warp * cfg.stages;
    if (lane < cfg.stages) mbar_init(&bars[lane], 1);
    fence_mbar_init();
    __syncthreads();

    const int64_t total_bytes = E * (int64_t)cfg.F * P * 8;
    const int64_t nbatch = (N + 31) / 32;
    const int64_t warps_total = (int64_t)gridDim.x * cfg.warps;
    const int64_t first = (int64_t)blockIdx.x * cfg.warps + warp;
    // this warp's batches: first, first + warps_total, ...
    const int64_t my_batches = first < nbatch ? (nbatch - first + warps_total - 1) / warps_total : 0;
    const int64_t items = my_batches * cfg.chunks;

    // producer side: stage item `q` (batch q / chunks, chunk q % chunks)
    auto issue = [&](int64_t q) {
        const int64_t b = first + (q / cfg.chunks) * warps_total;
        const int ch = (int)(q % cfg.chunks);
        const int f0 = ch * cfg.FC;
        const int nf = min(cfg.FC, cfg.F - f0);
        const int bytes = chunk_copy_bytes(nf, P, cfg.odd_p);
        const int s = (int)(q % cfg.stages);
        const int64_t n = b * 32 + lane;
        const int32_t e = n < N ? valid_elem(elem[n], E) : -1;
        int shift = 0;
        bool use_tma = false;
        int64_t off = 0;
        if (e >= 0) {
            off = (((int64_t)e * cfg.F + f0) * P) * 8;
            shift = cfg.odd_p ? (int)(off & 8) : 0;
            use_tma = (off - shift + bytes) <= total_bytes;
        }
        const unsigned mask = __ballot_sync(0xffffffffu, use_tma);
        // always arrive (tx may be 0) so that every use of a stage completes exactly one phase
        if (lane == 0) mbar_arrive_expect_tx(&bars[s], (uint32_t)__popc(mask) * bytes);
        __syncwarp();
        unsigned char *slot = wbase + s * stage_bytes + (size_t)lane * cfg.slot_bytes;
        if (use_tma) {
            bulk_copy_g2s(slot, reinterpret_cast<const unsigned char *>(fields) + off - shift,
                          (uint32_t)bytes, &bars[s]);
        } else if (e >= 0) {  // last bytes of the array: plain loads
            const double *src = reinterpret_cast<const double *>(
                reinterpret_cast<const unsigned char *>(fields) + off);
            double *dst = reinterpret_cast<double *>(slot + shift);
            for (int q2 = 0; q2 < nf * P; ++q2) dst[q2] = src[q2];
        }
    };

    const int prefetch = cfg.stages - 1;
    for (int64_t q = 0; q < prefetch && q < items; ++q) issue(q);

    double L[DIM][M];
    for (int64_t q = 0; q < items; ++q) {
        if (q + prefetch < items) {
            // the stage being refilled was consumed at item q-1; order those generic reads
            // before the async-proxy writes
            fence_proxy_async_smem();
            __syncwarp();
            issue(q + prefetch);
        }
        const int64_t b = first + (q / cfg.chunks) * warps_total;
        const int ch = (int)(q % cfg.chunks);
        const int f0 = ch * cfg.FC;
        const int nf = min(cfg.FC, cfg.F - f0);
        const int bytes = chunk_copy_bytes(nf, P, cfg.odd_p);
        const int s = (int)(q % cfg.stages);
        const int64_t n = b * 32 + lane;
        const int32_t e = n < N ? valid_elem(elem[n], E) : -1;
        int shift = 0;
        bool use_tma = false;
        if (e >= 0) {
            const int64_t off = (((int64_t)e * cfg.F + f0) * P) * 8;
            shift = cfg.odd_p ? (int)(off & 8) : 0;
            use_tma = (off - shift + bytes) <= total_bytes;
        }
        if (ch == 0 && e >= 0) {
#pragma unroll
            for (int ax = 0; ax < DIM; ++ax) lagrange_values<ORDER>(T, xi[n * DIM + ax], L[ax]);
        }
        mbar_wait(&bars[s], (uint32_t)((q / cfg.stages) & 1));
        if (n < N) {
            const double *v = reinterpret_cast<const double *>(
                wbase + s * stage_bytes + (size_t)lane * cfg.slot_bytes + shift);
            double *o = out + (perm ? (int64_t)perm[n] : n) * cfg.F + f0;
            if (e >= 0) {
                for (int f = 0; f < nf; ++f) o[f] = contract_field<ORDER, DIM>(v + f * P, L);
            } else {
                for (int f = 0; f < nf; ++f) o[f] = 0.0;
            }
        }
    }
}

template <int ORDER, int DIM>
int launch_interp(int64_t E, int F, const double *fields, int64_t N, const int32_t *elem,
                  const double *xi, const int32_t *perm, double *out, cudaStream_t stream)
{
    constexpr int M = ORDER + 1;
    constexpr int P = DIM == 2 ? M * M : M * M * M;
    mm_gll_table T;
    mm_make_table(ORDER, &T);
    interp_cfg cfg;
    cfg.F = F;
    cfg.odd_p = P % 2;
    // chunk size: keep a lane's slot near 512 B (a 32-lane stage near 16 KB) so that several
    // CTAs of 2 warps x 3 stages share an SM; one field is the minimum (1 008 B at order 4)
    int fc = std::max(1, std::min(F, 256 / (P * 8)));
    cfg.FC = fc;
    cfg.chunks = (F + fc - 1) / fc;
    cfg.slot_bytes = mm_slot_bytes(chunk_copy_bytes(fc, P, cfg.odd_p));
    cfg.stages = 3;
    cfg.warps = 4;
    // tuning overrides (profiling only)
    if (const char *e = getenv("MM_INTERP_FC")) fc = std::max(1, std::min(F, atoi(e)));
    if (const char *e = getenv("MM_INTERP_STAGES")) cfg.stages = std::max(2, std::min(8, atoi(e)));
    if (const char *e = getenv("MM_INTERP_WARPS")) cfg.warps = std::max(1, std::min(INTERP_MAX_WARPS, atoi(e)));
    cfg.FC = fc;
    cfg.chunks = (F + fc - 1) / fc;
    cfg.slot_bytes = mm_slot_bytes(chunk_copy_bytes(fc, P, cfg.odd_p));
    auto kern = interp_kernel<ORDER, DIM>;
    auto smem_of = [&](int warps) {
        return (size_t)warps * cfg.stages * 32 * cfg.slot_bytes + warps * cfg.stages * sizeof(uint64_t);
    };
    while (cfg.warps > 1 && smem_of(cfg.warps) > 200 * 1024) --cfg.warps;
    size_t smem = smem_of(cfg.warps);
    MM_REQUIRE(smem <= 227 * 1024, MM_ERR_UNSUPPORTED, "mm_interp: staging needs %zu B of shared memory", smem);
    static mm_kernel_cfg kcfg;
    int per_sm = 1;
    MM_CUDA(kcfg.prepare(kern, cfg.warps * 32, smem, &per_sm));
    const int sms = mm_num_sms() > 0 ? mm_num_sms() : 148;
    int64_t nbatch = (N + 31) / 32;
    int64_t grid = (int64_t)sms * per_sm;  // persistent CTAs, multiple of the SM count
    int64_t need = (nbatch + cfg.warps - 1) / cfg.warps;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    kern<<<(int)grid, cfg.warps * 32, smem, stream>>>(T, cfg, E, fields, N, elem, xi, perm, out);
    MM_CUDA(cudaGetLastError());
    return MM_OK;
}

// ================================================================================================
// K3, coherent variant (used by mm_interp_perm / the fused pipeline, where the points arrive
// spatially sorted): the 32 points of a warp batch usually fall into a handful of source elements,
// so the warp de-duplicates them (__match_any_sync), fetches each DISTINCT element block once with
// one bulk-async copy into one of C shared slots, and every lane contracts out of the slot of its
// element (lanes of one element read the same addresses: shared-memory broadcast).  Batches with
// more than C distinct elements take several rounds.  Correct for any point order; fast when the
// order is coherent.  Same arithmetic as interp_kernel (bit-identical results).
// ================================================================================================
struct coh_cfg {
    int F, FC, chunks, stages, slots, slot_bytes, odd_p, warps;
    int perm_stride;  // in int32 units: 1 for a plain permutation array, 8 when the permutation lives in
                      // the .w lane of the pipeline's 32-byte sorted-query records
};

template <int ORDER, int DIM>
struct coh_cursor {  // walks this warp's (batch, round, chunk) items; identical on both pipeline ends
    int64_t b;       // batch index
    int r, c;        // round (groups [r*C, r*C+C)), field chunk
    int D;           // distinct elements in the batch
    int rank;        // this lane's group rank (-1: no element)
    int32_t e;       // this lane's element
    int64_t n;       // this lane's point
};

template <int ORDER, int DIM>
__global__ void __launch_bounds__(INTERP_MAX_WARPS * 32)
interp_coherent_kernel(const mm_gll_table T, const coh_cfg cfg, int64_t E,
                       const double *__restrict__ fields, int64_t N,
                       const int32_t *__restrict__ elem, const double *__restrict__ xi,
                       const int32_t *__restrict__ perm, double *__restrict__ out,
                       const uint8_t *__restrict__ status, int32_t *__restrict__ elem_u,
                       double *__restrict__ xi_u, uint8_t *__restrict__ status_u)
{
    constexpr int M = ORDER + 1;
    constexpr int P = DIM == 2 ? M * M : M * M * M;
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t stage_bytes = (size_t)cfg.slots * cfg.slot_bytes;
    unsigned char *wbase = smem + (size_t)warp * cfg.stages * stage_bytes;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)cfg.warps * cfg.stages * stage_bytes) +
                     warp * cfg.stages;
    if (lane < cfg.stages) mbar_init(&bars[lane], 1);
    fence_mbar_init();
    __syncthreads();

    const int64_t total_bytes = E * (int64_t)cfg.F * P * 8;
    const int64_t nbatch = (N + 31) / 32;
    const int64_t stride = (int64_t)gridDim.x * cfg.warps;
    const int64_t first = (int64_t)blockIdx.x * cfg.warps + warp;

    using cur_t = coh_cursor<ORDER, DIM>;
    auto load_batch = [&](cur_t &k) {  // group the lanes of batch k.b by element
        k.n = k.b * 32 + lane;
        k.e = (k.b < nbatch && k.n < N) ? valid_elem(elem[k.n], E) : -1;
        const unsigned same = __match_any_sync(0xffffffffu, k.e);
        const int leader = __ffs(same) - 1;
        const unsigned leaders = __ballot_sync(0xffffffffu, lane == leader && k.e >= 0);
        k.rank = k.e >= 0 ? __popc(leaders & ((1u << leader) - 1)) : -1;
        k.D = __popc(leaders);
        k.r = 0;
        k.c = 0;
    };
    auto advance = [&](cur_t &k) {
        if (++k.c < cfg.chunks) return;
        k.c = 0;
        if ((++k.r) * cfg.slots < k.D) return;
        k.b += stride;
        load_batch(k);
    };
    auto valid = [&](const cur_t &k) { return k.b < nbatch; };

    auto issue = [&](const cur_t &k, int64_t q) {
        const int f0 = k.c * cfg.FC;
        const int nf = min(cfg.FC, cfg.F - f0);
        const int bytes = chunk_copy_bytes(nf, P, cfg.odd_p);
        const int s = (int)(q % cfg.stages);
        const int slot_id = k.rank - k.r * cfg.slots;
        // the group leader (lowest lane of the group) fetches the block for the whole group
        const unsigned same = __match_any_sync(0xffffffffu, k.e);
        const bool mine = k.e >= 0 && lane == __ffs(same) - 1 && slot_id >= 0 && slot_id < cfg.slots;
        int shift = 0;
        bool use_tma = false;
        int64_t off = 0;
        if (mine) {
            off = (((int64_t)k.e * cfg.F + f0) * P) * 8;
            shift = cfg.odd_p ? (int)(off & 8) : 0;
            use_tma = (off - shift + bytes) <= total_bytes;
        }
        const unsigned mask = __ballot_sync(0xffffffffu, use_tma);
        if (lane == 0) mbar_arrive_expect_tx(&bars[s], (uint32_t)__popc(mask) * bytes);
        __syncwarp();
        if (mine) {
            unsigned char *slot = wbase + s * stage_bytes + (size_t)slot_id * cfg.slot_bytes;
            if (use_tma) {
                bulk_copy_g2s(slot, reinterpret_cast<const unsigned char *>(fields) + off - shift,
                              (uint32_t)bytes, &bars[s]);
            } else {  // last bytes of the array: plain loads
                const double *src = reinterpret_cast<const double *>(
                    reinterpret_cast<const unsigned char *>(fields) + off);
                double *dst = reinterpret_cast<double *>(slot + shift);
                for (int q2 = 0; q2 < nf * P; ++q2) dst[q2] = src[q2];
            }
        }
    };

    cur_t ci, cc;  // issue cursor runs `prefetch` items ahead of the consume cursor
    ci.b = first;
    load_batch(ci);
    cc = ci;
    const int prefetch = cfg.stages - 1;
    int64_t qi = 0;
    for (; qi < prefetch && valid(ci); ++qi) {
        issue(ci, qi);
        advance(ci);
    }
    double L[DIM][M];
    int64_t last_b = -1;
    for (int64_t q = 0; valid(cc); ++q) {
        if (valid(ci)) {
            fence_proxy_async_smem();  // reads of the stage being refilled happened at item q-1
            __syncwarp();
            issue(ci, qi);
            advance(ci);
            ++qi;
        }
        const int f0 = cc.c * cfg.FC;
        const int nf = min(cfg.FC, cfg.F - f0);
        const int s = (int)(q % cfg.stages);
        if (cc.b != last_b) {  // new batch: Lagrange values of this lane's point
            last_b = cc.b;
            double x[DIM];
            if (cc.n < N) {
#pragma unroll
                for (int ax = 0; ax < DIM; ++ax) x[ax] = xi[cc.n * DIM + ax];
                if (elem_u) {  // fused un-permute of the location outputs (mm_interpolate)
                    const int64_t t = perm ? (int64_t)perm[cc.n * cfg.perm_stride] : cc.n;
                    elem_u[t] = cc.e;
                    if (status_u) status_u[t] = status[cc.n];
                    if (xi_u) {
#pragma unroll
                        for (int ax = 0; ax < DIM; ++ax) xi_u[t * DIM + ax] = x[ax];
                    }
                }
            }
            if (cc.e >= 0) {
#pragma unroll
                for (int ax = 0; ax < DIM; ++ax) lagrange_values<ORDER>(T, x[ax], L[ax]);
            }
        }
        mbar_wait(&bars[s], (uint32_t)((q / cfg.stages) & 1));
        const int slot_id = cc.rank - cc.r * cfg.slots;
        if (cc.e >= 0) {
            if (slot_id >= 0 && slot_id < cfg.slots) {
                const int64_t off = (((int64_t)cc.e * cfg.F + f0) * P) * 8;
                const int shift = cfg.odd_p ? (int)(off & 8) : 0;
                const double *v = reinterpret_cast<const double *>(
                    wbase + s * stage_bytes + (size_t)slot_id * cfg.slot_bytes + shift);
                double *o = out + (perm ? (int64_t)perm[cc.n * cfg.perm_stride] : cc.n) * cfg.F + f0;
                for (int f = 0; f < nf; ++f) o[f] = contract_field<ORDER, DIM>(v + f * P, L);
            }
        } else if (cc.n < N && cc.b < nbatch && cc.r == 0) {
            double *o = out + (perm ? (int64_t)perm[cc.n * cfg.perm_stride] : cc.n) * cfg.F + f0;
            for (int f = 0; f < nf; ++f) o[f] = 0.0;  // failed point: zero row
        }
        advance(cc);
    }
}

template <int ORDER, int DIM>
int launch_interp_coherent(int64_t E, int F, const double *fields, int64_t N, const int32_t *elem,
                           const double *xi, const int32_t *perm, double *out, cudaStream_t stream,
                           const uint8_t *status = nullptr, int32_t *elem_u = nullptr,
                           double *xi_u = nullptr, uint8_t *status_u = nullptr, int perm_stride = 1)
{
    constexpr int M = ORDER + 1;
    constexpr int P = DIM == 2 ? M * M : M * M * M;
    mm_gll_table T;
    mm_make_table(ORDER, &T);
    coh_cfg cfg;
    cfg.F = F;
    cfg.perm_stride = perm_stride;
    cfg.odd_p = P % 2;
    // slots of about 1 KB (whole block at order <= 2 / F = 5, one field at order 4), 4 slots x 2
    // stages per warp, 8 warps per CTA: measured best on B200 (profiles/r1_*), ~3 CTAs per SM
    int fc = std::max(1, std::min(F, 1100 / (P * 8)));
    cfg.stages = 2;
    cfg.slots = 4;
    cfg.warps = 8;
    if (const char *e = getenv("MM_COH_FC")) fc = std::max(1, std::min(F, atoi(e)));
    if (const char *e = getenv("MM_COH_STAGES")) cfg.stages = std::max(2, std::min(8, atoi(e)));
    if (const char *e = getenv("MM_COH_SLOTS")) cfg.slots = std::max(1, std::min(32, atoi(e)));
    if (const char *e = getenv("MM_COH_WARPS")) cfg.warps = std::max(1, std::min(INTERP_MAX_WARPS, atoi(e)));
    cfg.FC = fc;
    cfg.chunks = (F + fc - 1) / fc;
    cfg.slot_bytes = ((chunk_copy_bytes(fc, P, cfg.odd_p) + 15) / 16) * 16;
    auto kern = interp_coherent_kernel<ORDER, DIM>;
    auto smem_of = [&](int warps) {
        return (size_t)warps * cfg.stages * cfg.slots * cfg.slot_bytes + warps * cfg.stages * sizeof(uint64_t);
    };
    while (cfg.warps > 1 && smem_of(cfg.warps) > 110 * 1024) --cfg.warps;  // two CTAs per SM
    size_t smem = smem_of(cfg.warps);
    MM_REQUIRE(smem <= 227 * 1024, MM_ERR_UNSUPPORTED, "mm_interp: staging needs %zu B of shared memory", smem);
    static mm_kernel_cfg kcfg;
    int per_sm = 1;
    MM_CUDA(kcfg.prepare(kern, cfg.warps * 32, smem, &per_sm));
    const int sms = mm_num_sms() > 0 ? mm_num_sms() : 148;
    int64_t nbatch = (N + 31) / 32;
    int64_t grid = (int64_t)sms * per_sm;
    int64_t need = (nbatch + cfg.warps - 1) / cfg.warps;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    kern<<<(int)grid, cfg.warps * 32, smem, stream>>>(T, cfg, E, fields, N, elem, xi, perm, out, status,
                                                      elem_u, xi_u, status_u);
    MM_CUDA(cudaGetLastError());
    return MM_OK;
}

// ================================================================================================
// K3, CTA-tile variant (the fused pipeline's gather): a CTA owns a tile of 256 consecutive points of the
// spatially sorted order.  Their owning elements are de-duplicated for the WHOLE tile with a shared-memory
// hash table (a tile of 256 sorted points touches 10-40 source elements; a warp of 32 points alone would fetch
// 2-4 of them again in every warp), each distinct element block is fetched ONCE per tile with one bulk-async
// copy (UBLKCP) into a slot of the current stage, and every thread contracts out of the slot of its element
// (threads of one element read the same addresses: broadcast).  Two stages: while the threads contract item i
// (a round of up to `slots` elements x one chunk of FC fields), the copies of item i + 1 are in flight.  Every
// thread arrives once per item on the stage's mbarrier (element owners with their byte count), so the barrier
// needs no per-item re-initialisation.  Slot stride: a multiple of 16 B whose 8-byte count is = 2 (mod 4), so that
// consecutive slots start in different bank pairs (the warp-variant's stride of 136 x 8 B put every second slot on
// the same banks).  Same arithmetic as the other K3 kernels: bit-identical results.
// ================================================================================================
constexpr int TILE_THREADS = 256;
constexpr int TILE_HASH = 512;  // >= 2 x tile size: short probe sequences, never full

struct tile_cfg {
    int F, FC, chunks, slots, slot_bytes, odd_p, perm_stride;
};

__host__ __device__ inline int tile_slot_bytes(int bytes)
{
    int q = (bytes + 15) / 16 * 2;  // 8-byte words, even
    while ((q & 3) != 2) q += 2;
    return q * 8;
}

template <int ORDER, int DIM>
__global__ void __launch_bounds__(TILE_THREADS, 3)
interp_tile_kernel(const mm_gll_table T, const tile_cfg cfg, int64_t E, const double *__restrict__ fields, int64_t N,
                   const int32_t *__restrict__ elem, const double *__restrict__ xi,
                   const int32_t *__restrict__ perm, double *__restrict__ out,
                   const uint8_t *__restrict__ status, int32_t *__restrict__ elem_u, double *__restrict__ xi_u,
                   uint8_t *__restrict__ status_u)
{
    constexpr int M = ORDER + 1;
    constexpr int P = DIM == 2 ? M * M : M * M * M;
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t stage_bytes = (size_t)cfg.slots * cfg.slot_bytes;
    unsigned char *stage0 = smem;
    int32_t *keys = reinterpret_cast<int32_t *>(smem + 2 * stage_bytes);  // [TILE_HASH] element id or -1
    int32_t *dense = keys + TILE_HASH;                                   // [TILE_HASH] dense id of the entry
    int32_t *wsum = dense + TILE_HASH;                                   // [TILE_HASH / 32] + total
    uint64_t *bars = reinterpret_cast<uint64_t *>(wsum + TILE_HASH / 32 + 2);
    if (tid < 2) mbar_init(&bars[tid], TILE_THREADS);
    fence_mbar_init();
    __syncthreads();

    const int64_t total_bytes = E * (int64_t)cfg.F * P * 8;
    const int64_t ntiles = (N + TILE_THREADS - 1) / TILE_THREADS;
    uint32_t uses[2] = {0, 0};  // completed phases of each stage barrier

    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t n = tile * TILE_THREADS + tid;
        const bool valid = n < N;
        const int32_t e = valid ? valid_elem(elem[n], E) : -1;
        // ---- de-duplicate the tile's elements ---------------------------------------------------------
        __syncthreads();  // everybody is done with the previous tile's table and slots
        for (int i = tid; i < TILE_HASH; i += TILE_THREADS) keys[i] = -1;
        __syncthreads();
        int hpos = -1;
        if (e >= 0) {
            unsigned h = ((unsigned)e * 2654435761u) >> 23;  // top 9 bits
            while (true) {
                h &= TILE_HASH - 1;
                const int32_t old = atomicCAS(&keys[h], -1, e);
                if (old == -1 || old == e) break;
                ++h;
            }
            hpos = (int)h;
        }
        __syncthreads();
        // dense ids of the occupied entries (entry i is handled by thread i and thread i + 256)
        int occ[TILE_HASH / TILE_THREADS];
#pragma unroll
        for (int j = 0; j < TILE_HASH / TILE_THREADS; ++j) {
            const int i = tid + j * TILE_THREADS;
            const bool o = keys[i] >= 0;
            const unsigned b = __ballot_sync(0xffffffffu, o);
            occ[j] = o ? __popc(b & ((1u << lane) - 1)) : -1;
            if (lane == 0) wsum[warp + j * (TILE_THREADS / 32)] = __popc(b);
        }
        __syncthreads();
        if (tid == 0) {
            int run = 0;
            for (int w = 0; w < TILE_HASH / 32; ++w) {
                const int t = wsum[w];
                wsum[w] = run;
                run += t;
            }
            wsum[TILE_HASH / 32] = run;
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < TILE_HASH / TILE_THREADS; ++j)
            if (occ[j] >= 0) dense[tid + j * TILE_THREADS] = wsum[warp + j * (TILE_THREADS / 32)] + occ[j];
        __syncthreads();
        const int D = wsum[TILE_HASH / 32];
        const int my_id = hpos >= 0 ? dense[hpos] : -1;
        // this thread owns (fetches) the elements of table entries tid and tid + 256
        int32_t own_e[TILE_HASH / TILE_THREADS];
        int own_id[TILE_HASH / TILE_THREADS];
#pragma unroll
        for (int j = 0; j < TILE_HASH / TILE_THREADS; ++j) {
            own_e[j] = keys[tid + j * TILE_THREADS];
            own_id[j] = own_e[j] >= 0 ? dense[tid + j * TILE_THREADS] : -1;
        }
        // ---- per-point preparation: Lagrange values, un-permuted location outputs ----------------------
        double L[DIM][M];
        int64_t orow = 0;
        if (valid) {
            double x[DIM];
#pragma unroll
            for (int ax = 0; ax < DIM; ++ax) x[ax] = xi[n * DIM + ax];
            orow = perm ? (int64_t)perm[n * cfg.perm_stride] : n;
            if (elem_u) {
                elem_u[orow] = e;
                if (status_u) status_u[orow] = status[n];
                if (xi_u) {
#pragma unroll
                    for (int ax = 0; ax < DIM; ++ax) xi_u[orow * DIM + ax] = x[ax];
                }
            }
            if (e >= 0) {
#pragma unroll
                for (int ax = 0; ax < DIM; ++ax) lagrange_values<ORDER>(T, x[ax], L[ax]);
            } else {
                for (int f = 0; f < cfg.F; ++f) out[orow * cfg.F + f] = 0.0;  // failed point: zero row
            }
        }
        // ---- items: (round of `slots` elements) x (chunk of FC fields) -------------------------------
        const int rounds = (D + cfg.slots - 1) / cfg.slots;
        const int items = rounds * cfg.chunks;
        auto issue = [&](int item, int s) {
            const int r = item / cfg.chunks, ch = item % cfg.chunks;
            const int f0 = ch * cfg.FC, nf = min(cfg.FC, cfg.F - f0);
            const int bytes = chunk_copy_bytes(nf, P, cfg.odd_p);
            uint32_t tx = 0;
#pragma unroll
            for (int j = 0; j < TILE_HASH / TILE_THREADS; ++j) {
                const int slot = own_id[j] - r * cfg.slots;
                if (own_id[j] < 0 || slot < 0 || slot >= cfg.slots) continue;
                const int64_t off = (((int64_t)own_e[j] * cfg.F + f0) * P) * 8;
                const int shift = cfg.odd_p ? (int)(off & 8) : 0;
                unsigned char *dst = stage0 + s * stage_bytes + (size_t)slot * cfg.slot_bytes;
                if ((off - shift + bytes) <= total_bytes) {
                    tx += (uint32_t)bytes;
                } else {  // last bytes of the array: plain loads (visible to the others through the barrier's release)
                    const double *src = reinterpret_cast<const double *>(
                        reinterpret_cast<const unsigned char *>(fields) + off);
                    double *d2 = reinterpret_cast<double *>(dst + shift);
                    for (int q = 0; q < nf * P; ++q) d2[q] = src[q];
                }
            }
            if (tx) mbar_arrive_expect_tx(&bars[s], tx);
            else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bars[s])) : "memory");
#pragma unroll
            for (int j = 0; j < TILE_HASH / TILE_THREADS; ++j) {
                const int slot = own_id[j] - r * cfg.slots;
                if (own_id[j] < 0 || slot < 0 || slot >= cfg.slots) continue;
                const int64_t off = (((int64_t)own_e[j] * cfg.F + f0) * P) * 8;
                const int shift = cfg.odd_p ? (int)(off & 8) : 0;
                if ((off - shift + bytes) <= total_bytes)
                    bulk_copy_g2s(stage0 + s * stage_bytes + (size_t)slot * cfg.slot_bytes,
                                  reinterpret_cast<const unsigned char *>(fields) + off - shift, (uint32_t)bytes,
                                  &bars[s]);
            }
        };
        // generic-proxy reads of the previous tile's slots happened before the __syncthreads above
        fence_proxy_async_smem();
        if (items > 0) issue(0, 0);
        for (int item = 0; item < items; ++item) {
            const int s = item & 1;
            if (item + 1 < items) {
                // stage (item + 1) & 1 was read at item - 1: every thread must be past that before it is refilled
                fence_proxy_async_smem();
                __syncthreads();
                issue(item + 1, (item + 1) & 1);
            }
            mbar_wait(&bars[s], uses[s] & 1);
            ++uses[s];
            const int r = item / cfg.chunks, ch = item % cfg.chunks;
            const int slot = my_id - r * cfg.slots;
            if (my_id >= 0 && slot >= 0 && slot < cfg.slots) {
                const int f0 = ch * cfg.FC, nf = min(cfg.FC, cfg.F - f0);
                const int64_t off = (((int64_t)e * cfg.F + f0) * P) * 8;
                const int shift = cfg.odd_p ? (int)(off & 8) : 0;
                const double *v = reinterpret_cast<const double *>(stage0 + s * stage_bytes +
                                                                   (size_t)slot * cfg.slot_bytes + shift);
                double *o = out + orow * cfg.F + f0;
                for (int f = 0; f < nf; ++f) o[f] = contract_field<ORDER, DIM>(v + f * P, L);
            }
        }
    }
}

template <int ORDER, int DIM>
int launch_interp_tile(int64_t E, int F, const double *fields, int64_t N, const int32_t *elem, const double *xi,
                       const int32_t *perm, double *out, cudaStream_t stream, const uint8_t *status, int32_t *elem_u,
                       double *xi_u, uint8_t *status_u, int perm_stride)
{
    constexpr int M = ORDER + 1;
    constexpr int P = DIM == 2 ? M * M : M * M * M;
    mm_gll_table T;
    mm_make_table(ORDER, &T);
    tile_cfg cfg;
    cfg.F = F;
    cfg.perm_stride = perm_stride;
    cfg.odd_p = P % 2;
    int fc = std::max(1, std::min(F, 1100 / (P * 8)));  // ~1 KB slots: whole block at order <= 2, one field at order 4
    if (const char *e = getenv("MM_TILE_FC")) fc = std::max(1, std::min(F, atoi(e)));
    cfg.FC = fc;
    cfg.chunks = (F + fc - 1) / fc;
    cfg.slot_bytes = tile_slot_bytes(chunk_copy_bytes(fc, P, cfg.odd_p));
    cfg.slots = std::max(4, std::min(48, (32 * 1024) / cfg.slot_bytes));  // ~32 KB per stage
    if (const char *e = getenv("MM_TILE_SLOTS")) cfg.slots = std::max(1, std::min(256, atoi(e)));
    auto kern = interp_tile_kernel<ORDER, DIM>;
    size_t smem = 2 * (size_t)cfg.slots * cfg.slot_bytes + sizeof(int32_t) * (2 * TILE_HASH + TILE_HASH / 32 + 2) +
                  2 * sizeof(uint64_t);
    MM_REQUIRE(smem <= 227 * 1024, MM_ERR_UNSUPPORTED, "mm_interp: staging needs %zu B of shared memory", smem);
    static mm_kernel_cfg kcfg;
    int per_sm = 1;
    MM_CUDA(kcfg.prepare(kern, TILE_THREADS, smem, &per_sm));
    const int sms = mm_num_sms() > 0 ? mm_num_sms() : 148;
    const int64_t ntiles = (N + TILE_THREADS - 1) / TILE_THREADS;
    int64_t grid = std::min<int64_t>((int64_t)sms * per_sm, ntiles);
    if (grid < 1) grid = 1;
    kern<<<(int)grid, TILE_THREADS, smem, stream>>>(T, cfg, E, fields, N, elem, xi, perm, out, status, elem_u, xi_u,
                                                    status_u);
    MM_CUDA(cudaGetLastError());
    return MM_OK;
}

// ---- coefficient write-out: coeffs[n][a] = (Lx[i] * Ly[j]) * Lz[k] -----------------------------
template <int ORDER, int DIM>
__global__ void __launch_bounds__(256)
coeffs_kernel(const mm_gll_table T, int64_t N, const int32_t *__restrict__ elem,
              const double *__restrict__ xi, double *__restrict__ coeffs)
{
    constexpr int M = ORDER + 1;
    constexpr int P = DIM == 2 ? M * M : M * M * M;
    for (int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; n < N;
         n += (int64_t)gridDim.x * blockDim.x) {
        double *w = coeffs + n * P;
        if (elem && elem[n] < 0) {
            for (int a = 0; a < P; ++a) w[a] = 0.0;
            continue;
        }
        double L[DIM][M];
#pragma unroll
        for (int ax = 0; ax < DIM; ++ax) lagrange_values<ORDER>(T, xi[n * DIM + ax], L[ax]);
        if constexpr (DIM == 2) {
#pragma unroll
            for (int j = 0; j < M; ++j)
#pragma unroll
                for (int i = 0; i < M; ++i) w[i + M * j] = L[0][i] * L[1][j];
        } else {
#pragma unroll
            for (int k = 0; k < M; ++k)
#pragma unroll
                for (int j = 0; j < M; ++j)
#pragma unroll
                    for (int i = 0; i < M; ++i)
                        w[i + M * j + M * M * k] = (L[0][i] * L[1][j]) * L[2][k];
        }
    }
}

// ---- explicit-matrix gather (cached elements/coeffs): one warp per point -----------------------
// out[n][f] = sum_a fields[e][f][a] * coeffs[n][a], a ascending (sequential, lane 0 finishes).
__global__ void __launch_bounds__(256)
gather_coeffs_kernel(int P, int64_t E, int F, const double *__restrict__ fields, int64_t N,
                     const int32_t *__restrict__ elem, const double *__restrict__ coeffs,
                     double *__restrict__ out)
{
    // thread per (point, field): the P-long dot product walks one contiguous row
    int64_t total = N * F;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
         t += (int64_t)gridDim.x * blockDim.x) {
        int64_t n = t / F;
        int f = (int)(t - n * F);
        int32_t e = valid_elem(elem[n], E);
        double acc = 0.0;
        if (e >= 0) {
            const double *v = fields + ((int64_t)e * F + f) * P;
            const double *w = coeffs + n * P;
            for (int a = 0; a < P; ++a) acc = acc + v[a] * w[a];
        }
        out[t] = acc;
    }
}

int blocks_for(int64_t work, int block)
{
    int sms = mm_num_sms() > 0 ? mm_num_sms() : 148;
    int64_t need = (work + block - 1) / block;
    int64_t cap = (int64_t)sms * 16;
    return (int)(need < 1 ? 1 : (need > cap ? cap : need));
}

}  // namespace

static int interp_dispatch(bool coherent, int order, int dim, int64_t E, int F, const double *fields,
                           int64_t N, const int32_t *elem, const double *xi, const int32_t *perm,
                           double *out, void *stream_);

extern "C" int mm_interp(int order, int dim, int64_t E, int F, const double *fields, int64_t N,
                         const int32_t *elem, const double *xi, double *out, void *stream)
{
    return interp_dispatch(false, order, dim, E, F, fields, N, elem, xi, nullptr, out, stream);
}

extern "C" int mm_interp_perm(int order, int dim, int64_t E, int F, const double *fields, int64_t N,
                              const int32_t *elem, const double *xi, const int32_t *perm,
                              double *out, void *stream)
{
    const char *mode = getenv("MM_INTERP_MODE");
    const bool coherent = !(mode && mode[0] == 'l');
    return interp_dispatch(coherent, order, dim, E, F, fields, N, elem, xi, perm, out, stream);
}

static int interp_dispatch(bool coherent, int order, int dim, int64_t E, int F, const double *fields,
                           int64_t N, const int32_t *elem, const double *xi, const int32_t *perm,
                           double *out, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    MM_REQUIRE(mm_valid_order(order), MM_ERR_INVALID, "mm_interp: order %d (supported 1, 2, 4)", order);
    MM_REQUIRE(dim == 2 || dim == 3, MM_ERR_INVALID, "mm_interp: dim %d", dim);
    MM_REQUIRE(F >= 1 && N >= 0 && E >= 0, MM_ERR_INVALID, "mm_interp: sizes");
    if (N == 0) return MM_OK;
    MM_REQUIRE(fields && elem && xi && out, MM_ERR_INVALID, "mm_interp: null buffer");
    MM_REQUIRE(((uintptr_t)fields & 15) == 0, MM_ERR_INVALID,
               "mm_interp: fields must be 16-byte aligned");
#define MM_INT(O, D)                                                                              \
    if (order == O && dim == D)                                                                   \
        return coherent ? launch_interp_coherent<O, D>(E, F, fields, N, elem, xi, perm, out, stream) \
                        : launch_interp<O, D>(E, F, fields, N, elem, xi, perm, out, stream);
    MM_INT(1, 2) MM_INT(2, 2) MM_INT(4, 2) MM_INT(1, 3) MM_INT(2, 3) MM_INT(4, 3)
#undef MM_INT
    mm_set_error("mm_interp: unsupported order/dim");
    return MM_ERR_UNSUPPORTED;
}

extern "C" int mm_coeffs(int order, int dim, int64_t N, const int32_t *elem, const double *xi,
                         double *coeffs, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    MM_REQUIRE(mm_valid_order(order), MM_ERR_INVALID, "mm_coeffs: order %d", order);
    MM_REQUIRE(dim == 2 || dim == 3, MM_ERR_INVALID, "mm_coeffs: dim %d", dim);
    if (N == 0) return MM_OK;
    MM_REQUIRE(N > 0 && xi && coeffs, MM_ERR_INVALID, "mm_coeffs: arguments");
    mm_gll_table T;
    mm_make_table(order, &T);
#define MM_COE(O, D)                                                                       \
    if (order == O && dim == D) {                                                          \
        coeffs_kernel<O, D><<<blocks_for(N, 256), 256, 0, stream>>>(T, N, elem, xi, coeffs); \
        MM_CUDA(cudaGetLastError());                                                       \
        return MM_OK;                                                                      \
    }
    MM_COE(1, 2) MM_COE(2, 2) MM_COE(4, 2) MM_COE(1, 3) MM_COE(2, 3) MM_COE(4, 3)
#undef MM_COE
    mm_set_error("mm_coeffs: unsupported order/dim");
    return MM_ERR_UNSUPPORTED;
}

extern "C" int mm_gather_coeffs(int P, int64_t E, int F, const double *fields, int64_t N,
                                const int32_t *elem, const double *coeffs, double *out,
                                void *stream)
{
    MM_REQUIRE(P >= 1 && F >= 1 && N >= 0 && E >= 0, MM_ERR_INVALID, "mm_gather_coeffs: sizes");
    if (N == 0) return MM_OK;
    MM_REQUIRE(fields && elem && coeffs && out, MM_ERR_INVALID, "mm_gather_coeffs: null buffer");
    gather_coeffs_kernel<<<blocks_for(N * F, 256), 256, 0, (cudaStream_t)stream>>>(
        P, E, F, fields, N, elem, coeffs, out);
    MM_CUDA(cudaGetLastError());
    return MM_OK;
}

// internal: K3 of the fused pipeline -- gather in sorted order + un-permuted location outputs
int mm_interp_fused(int order, int dim, int64_t E, int F, const double *fields, int64_t N,
                    const int32_t *elem_s, const double *xi_s, const uint8_t *status_s,
                    const int32_t *perm, int perm_stride, double *out, int32_t *elem_u, double *xi_u,
                    uint8_t *status_u, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    MM_REQUIRE(((uintptr_t)fields & 15) == 0, MM_ERR_INVALID, "mm_interpolate: fields must be 16-byte aligned");
    const char *mode = getenv("MM_INTERP_MODE");  // "w": the warp-level variant (A/B comparisons)
    const bool tile = !(mode && mode[0] == 'w');
#define MM_INTF(O, D)                                                                             \
    if (order == O && dim == D)                                                                   \
        return tile ? launch_interp_tile<O, D>(E, F, fields, N, elem_s, xi_s, perm, out, stream,  \
                                               status_s, elem_u, xi_u, status_u, perm_stride)     \
                    : launch_interp_coherent<O, D>(E, F, fields, N, elem_s, xi_s, perm, out, stream, \
                                                   status_s, elem_u, xi_u, status_u, perm_stride);
    MM_INTF(1, 2) MM_INTF(2, 2) MM_INTF(4, 2) MM_INTF(1, 3) MM_INTF(2, 3) MM_INTF(4, 3)
#undef MM_INTF
    mm_set_error("mm_interpolate: unsupported order/dim");
    return MM_ERR_UNSUPPORTED;
}
