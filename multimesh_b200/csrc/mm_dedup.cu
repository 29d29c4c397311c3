// mm_dedup.cu -- K4: target-side de-duplication of GLL points and the write-back of the results.
//
// Replaces, on the device,
//   utils.get_unique_points            np.unique(points, axis=0, return_inverse=True)      utils.py:465-515
//   values[recon].reshape(...).swapaxes(1, 2)                                               interpolator.py:822-826
//   the fluid / solid repair of gll_2_gll                                                   interpolator.py:829-841
// so that a driver moves only the final [E_t][F][P_t] array to the host.
//
// De-duplication re-uses the spatial index: identical coordinates always fall into the same grid cell, so after
// the counting sort by cell (mm_index_create) and the in-cell sort by (x, y, z, id) (mm_index_prepare_sites) the
// "sites" of the index ARE the distinct points, each with the contiguous list of its copies.  What remains is to
// put the sites into numpy's order -- lexicographic by (x, y, z) -- with a stable LSD radix sort (8-bit digits)
// over order-preserving 64-bit keys, z first, then y, then x; digits that are equal for all keys are skipped.
// Only the N_u distinct sites are sorted, not the N raw points (N / N_u = 3.4 for an order-2 hex mesh).
//
// Differences from np.unique that cannot matter for coordinates: -0.0 and +0.0 are one key and come out as +0.0;
// NaN rows are not supported (the index build rejects non-finite coordinates).
#include <algorithm>
#include <vector>

#include "mm_common.cuh"
#include "mm_scan.cuh"

namespace {

#define MM_TRY(call)                  \
    do {                              \
        int _rc = (call);             \
        if (_rc != MM_OK) return _rc; \
    } while (0)

constexpr int RX_THREADS = 256;
constexpr int RX_ITEMS = 8;                      // per thread
constexpr int RX_TILE = RX_THREADS * RX_ITEMS;   // per block
constexpr int RX_WARPS = RX_THREADS / 32;

// monotone map double -> uint64 (a < b  <=>  key(a) < key(b)); -0.0 is folded onto +0.0
__device__ __forceinline__ unsigned long long order_key(double x)
{
    if (x == 0.0) x = 0.0;
    const unsigned long long b = (unsigned long long)__double_as_longlong(x);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

__global__ void __launch_bounds__(256)
iota_kernel(int64_t n, uint32_t *__restrict__ perm)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        perm[i] = (uint32_t)i;
}

// keys[i] = order_key(coordinate `axis` of site perm[i]); diff |= keys[i] ^ keys[0]
__global__ void __launch_bounds__(256)
gather_keys_kernel(int64_t n, const double4 *__restrict__ sites, const uint32_t *__restrict__ perm, int axis,
                   unsigned long long *__restrict__ keys, unsigned long long *__restrict__ diff)
{
    const double4 s0 = sites[perm[0]];
    const unsigned long long k0 = order_key(axis == 0 ? s0.x : (axis == 1 ? s0.y : s0.z));
    unsigned long long local = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double4 s = sites[perm[i]];
        const unsigned long long kk = order_key(axis == 0 ? s.x : (axis == 1 ? s.y : s.z));
        keys[i] = kk;
        local |= kk ^ k0;
    }
    for (int o = 16; o > 0; o >>= 1) local |= __shfl_xor_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0 && local) atomicOr(diff, local);
}

// per-block digit histogram, laid out digit-major: hist[digit * nblocks + block]
__global__ void __launch_bounds__(RX_THREADS)
radix_hist_kernel(int64_t n, const unsigned long long *__restrict__ keys, int shift, int nblocks,
                  int32_t *__restrict__ hist)
{
    __shared__ int32_t cnt[256];
    cnt[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * RX_TILE;
    for (int i = threadIdx.x; i < RX_TILE; i += RX_THREADS)
        if (base + i < n) atomicAdd(&cnt[(int)((keys[base + i] >> shift) & 0xff)], 1);
    __syncthreads();
    hist[(int64_t)threadIdx.x * nblocks + blockIdx.x] = cnt[threadIdx.x];
}

// stable scatter of one digit pass.  Warp w owns items [w * 256, (w + 1) * 256) of the tile, in rounds of 32
// consecutive items, so the order inside the tile is (warp, round, lane) = the input order.
__global__ void __launch_bounds__(RX_THREADS)
radix_scatter_kernel(int64_t n, const unsigned long long *__restrict__ keys_in, const uint32_t *__restrict__ perm_in,
                     unsigned long long *__restrict__ keys_out, uint32_t *__restrict__ perm_out, int shift,
                     int nblocks, const int32_t *__restrict__ offsets)
{
    __shared__ int32_t cnt[RX_WARPS][256];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < RX_WARPS * 256; i += RX_THREADS) (&cnt[0][0])[i] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * RX_TILE + warp * (RX_ITEMS * 32);
    unsigned long long key[RX_ITEMS];
    uint32_t val[RX_ITEMS];
    int32_t lrank[RX_ITEMS];
#pragma unroll
    for (int r = 0; r < RX_ITEMS; ++r) {
        const int64_t i = base + r * 32 + lane;
        const bool valid = i < n;
        key[r] = valid ? keys_in[i] : 0;
        val[r] = valid ? perm_in[i] : 0;
        const int d = valid ? (int)((key[r] >> shift) & 0xff) : -1 - lane;  // invalid lanes match nobody
        const unsigned same = __match_any_sync(0xffffffffu, d);
        const int leader = __ffs(same) - 1;
        int32_t old = 0;
        if (valid && lane == leader) {
            old = cnt[warp][d];
            cnt[warp][d] = old + __popc(same);
        }
        old = __shfl_sync(0xffffffffu, old, leader);
        lrank[r] = old + __popc(same & ((1u << lane) - 1));
        __syncwarp();
    }
    __syncthreads();
    {  // exclusive prefix over the warps, per digit (thread = digit)
        int32_t run = 0;
#pragma unroll
        for (int w = 0; w < RX_WARPS; ++w) {
            const int32_t t = cnt[w][threadIdx.x];
            cnt[w][threadIdx.x] = run;
            run += t;
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RX_ITEMS; ++r) {
        const int64_t i = base + r * 32 + lane;
        if (i < n) {
            const int d = (int)((key[r] >> shift) & 0xff);
            const int64_t pos = (int64_t)offsets[(int64_t)d * nblocks + blockIdx.x] + cnt[warp][d] + lrank[r];
            keys_out[pos] = key[r];
            perm_out[pos] = val[r];
        }
    }
}

// unique[i] = coordinates of site perm[i]; rank[perm[i]] = i
__global__ void __launch_bounds__(256)
emit_unique_kernel(int dim, int64_t n, const double4 *__restrict__ sites, const uint32_t *__restrict__ perm,
                   double *__restrict__ unique, int32_t *__restrict__ rank)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t s = perm[i];
        const double4 r = sites[s];
        double x = r.x, y = r.y, z = r.z;
        if (x == 0.0) x = 0.0;  // the key folds -0.0 onto +0.0; so does the output
        if (y == 0.0) y = 0.0;
        if (z == 0.0) z = 0.0;
        unique[i * dim + 0] = x;
        unique[i * dim + 1] = y;
        if (dim == 3) unique[i * dim + 2] = z;
        rank[s] = (int32_t)i;
    }
}

// inverse[point] = rank of the point's site; one thread per site walks its copies
__global__ void __launch_bounds__(256)
emit_inverse_kernel(int64_t nsites, const int32_t *__restrict__ site_first, const int32_t *__restrict__ rec_id,
                    const int32_t *__restrict__ rank, int32_t *__restrict__ inverse)
{
    for (int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s < nsites; s += (int64_t)gridDim.x * blockDim.x) {
        const int32_t r = rank[s];
        for (int32_t t = site_first[s]; t < site_first[s + 1]; ++t) inverse[rec_id[t]] = r;
    }
}

// out[e][f][p] = values[inverse[e * P + p]][f]   (inverse == nullptr: identity).  One warp per element: the P
// value rows an element needs are contiguous when the values are in element order (identity) and gathered
// row-wise otherwise; the [F][P] block of the element is written coalesced along p.
__global__ void __launch_bounds__(256)
scatter_back_kernel(int64_t E, int P, int F, const double *__restrict__ values, const int32_t *__restrict__ inverse,
                    double *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t e = warp; e < E; e += nwarps) {
        for (int p = lane; p < P; p += 32) {
            const int64_t row = inverse ? (int64_t)inverse[e * P + p] : e * P + p;
            const double *v = values + row * F;
            for (int f = 0; f < F; ++f) out[(e * F + f) * P + p] = v[f];
        }
    }
}

// gll_2_gll, interpolator.py:829-841: fluid elements keep their old values; a solid element that picked up a
// fluid value (any VS == 0) is restored as a whole.  One warp per element.
__global__ void __launch_bounds__(256)
fluid_fixup_kernel(int64_t E, int P, int F, double *__restrict__ values, const double *__restrict__ old_values,
                   const uint8_t *__restrict__ fluid, int vs_index)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t e = warp; e < E; e += nwarps) {
        bool restore = fluid[e] != 0;
        if (!restore && vs_index >= 0) {
            bool zero = false;
            for (int p = lane; p < P; p += 32) zero = zero || values[(e * F + vs_index) * P + p] == 0.0;
            restore = __any_sync(0xffffffffu, zero);
        }
        if (restore)
            for (int i = lane; i < F * P; i += 32) values[e * F * P + i] = old_values[e * F * P + i];
    }
}

int blocks_for(int64_t work, int block)
{
    int sms = mm_num_sms() > 0 ? mm_num_sms() : 148;
    int64_t need = (work + block - 1) / block;
    int64_t cap = (int64_t)sms * 16;
    return (int)(need < 1 ? 1 : (need > cap ? cap : need));
}

struct pool_buf {
    void *p = nullptr;
    cudaStream_t st;
    explicit pool_buf(cudaStream_t s) : st(s) {}
    cudaError_t alloc(size_t bytes) { return mm_pool_alloc(&p, bytes, st); }
    ~pool_buf() { mm_pool_free(p, st); }
    template <typename T> T *as() { return static_cast<T *>(p); }
};

}  // namespace

extern "C" int mm_unique_points(int dim, int64_t N, const double *pts, int64_t *n_unique, double *unique,
                                int32_t *inverse, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    MM_REQUIRE(dim == 2 || dim == 3, MM_ERR_INVALID, "mm_unique_points: dim %d", dim);
    MM_REQUIRE(N >= 0 && N <= (int64_t)INT32_MAX, MM_ERR_INVALID, "mm_unique_points: N=%lld outside [0, 2^31)",
               (long long)N);
    MM_REQUIRE(n_unique, MM_ERR_INVALID, "mm_unique_points: null n_unique");
    *n_unique = 0;
    if (N == 0) return MM_OK;
    MM_REQUIRE(pts && unique && inverse, MM_ERR_INVALID, "mm_unique_points: null buffer");

    // 1. cell sort + in-cell sort: sites = distinct coordinates, each with the list of its copies
    mm_index_t *ix = nullptr;
    MM_TRY(mm_index_create(&ix, dim, N, pts, stream));
    struct holder_t {
        mm_index_t *p;
        ~holder_t() { mm_index_destroy(p); }
    } holder{ix};
    MM_TRY(mm_index_prepare_sites(ix, stream));
    mm_index_sites_view sv;
    MM_REQUIRE(mm_index_sites_view_get(ix, &sv), MM_ERR_INVALID, "mm_unique_points: no site table");
    const int64_t ns = sv.nsites;
    *n_unique = ns;

    // 2. lexicographic order of the sites: stable LSD radix sort, last coordinate first
    const int nblocks = (int)((ns + RX_TILE - 1) / RX_TILE);
    const int64_t nh = (int64_t)256 * nblocks;
    pool_buf k0(stream), k1(stream), p0(stream), p1(stream), hist(stream), offs(stream), tiles(stream), diff(stream),
        rank(stream);
    MM_CUDA(k0.alloc(sizeof(unsigned long long) * ns));
    MM_CUDA(k1.alloc(sizeof(unsigned long long) * ns));
    MM_CUDA(p0.alloc(sizeof(uint32_t) * ns));
    MM_CUDA(p1.alloc(sizeof(uint32_t) * ns));
    MM_CUDA(hist.alloc(sizeof(int32_t) * nh));
    MM_CUDA(offs.alloc(sizeof(int32_t) * (nh + 1)));
    MM_CUDA(tiles.alloc(sizeof(int32_t) * (mm_scan_tiles(nh) + 1)));
    MM_CUDA(diff.alloc(sizeof(unsigned long long)));
    MM_CUDA(rank.alloc(sizeof(int32_t) * ns));
    unsigned long long *keys = k0.as<unsigned long long>(), *keys_alt = k1.as<unsigned long long>();
    uint32_t *perm = p0.as<uint32_t>(), *perm_alt = p1.as<uint32_t>();
    iota_kernel<<<blocks_for(ns, 256), 256, 0, stream>>>(ns, perm);
    for (int axis = dim - 1; axis >= 0; --axis) {
        MM_CUDA(cudaMemsetAsync(diff.p, 0, sizeof(unsigned long long), stream));
        gather_keys_kernel<<<blocks_for(ns, 256), 256, 0, stream>>>(ns, sv.site_recs, perm, axis, keys,
                                                                    diff.as<unsigned long long>());
        unsigned long long varying = 0;
        MM_CUDA(cudaMemcpyAsync(&varying, diff.p, sizeof varying, cudaMemcpyDeviceToHost, stream));
        MM_CUDA(cudaStreamSynchronize(stream));
        for (int digit = 0; digit < 8; ++digit) {
            if (((varying >> (8 * digit)) & 0xffull) == 0) continue;  // equal for all keys: the pass is the identity
            const int shift = 8 * digit;
            radix_hist_kernel<<<nblocks, RX_THREADS, 0, stream>>>(ns, keys, shift, nblocks, hist.as<int32_t>());
            mm_exclusive_scan_i32(nh, hist.as<int32_t>(), offs.as<int32_t>(), tiles.as<int32_t>(), stream);
            radix_scatter_kernel<<<nblocks, RX_THREADS, 0, stream>>>(ns, keys, perm, keys_alt, perm_alt, shift, nblocks,
                                                                     offs.as<int32_t>());
            MM_CUDA(cudaGetLastError());
            std::swap(keys, keys_alt);
            std::swap(perm, perm_alt);
        }
    }
    // 3. outputs
    emit_unique_kernel<<<blocks_for(ns, 256), 256, 0, stream>>>(dim, ns, sv.site_recs, perm, unique, rank.as<int32_t>());
    emit_inverse_kernel<<<blocks_for(ns, 256), 256, 0, stream>>>(ns, sv.site_first, sv.rec_id, rank.as<int32_t>(),
                                                                 inverse);
    MM_CUDA(cudaGetLastError());
    MM_CUDA(cudaStreamSynchronize(stream));  // the index (and its site table) is released on return
    return MM_OK;
}

extern "C" int mm_scatter_back(int64_t E, int P, int F, const double *values, const int32_t *inverse, double *out,
                               void *stream)
{
    MM_REQUIRE(E >= 0 && P >= 1 && F >= 1, MM_ERR_INVALID, "mm_scatter_back: sizes");
    if (E == 0) return MM_OK;
    MM_REQUIRE(values && out, MM_ERR_INVALID, "mm_scatter_back: null buffer");
    scatter_back_kernel<<<blocks_for(E * 32, 256), 256, 0, (cudaStream_t)stream>>>(E, P, F, values, inverse, out);
    MM_CUDA(cudaGetLastError());
    return MM_OK;
}

extern "C" int mm_fluid_fixup(int64_t E, int P, int F, double *values, const double *old_values,
                              const uint8_t *fluid, int vs_index, void *stream)
{
    MM_REQUIRE(E >= 0 && P >= 1 && F >= 1 && vs_index < F, MM_ERR_INVALID, "mm_fluid_fixup: sizes");
    if (E == 0) return MM_OK;
    MM_REQUIRE(values && old_values && fluid, MM_ERR_INVALID, "mm_fluid_fixup: null buffer");
    fluid_fixup_kernel<<<blocks_for(E * 32, 256), 256, 0, (cudaStream_t)stream>>>(E, P, F, values, old_values, fluid,
                                                                                 vs_index);
    MM_CUDA(cudaGetLastError());
    return MM_OK;
}
