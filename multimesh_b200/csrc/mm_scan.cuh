// mm_scan.cuh -- exclusive prefix sum of int32 counts on the device (three small kernels), shared by the
// index build / query sort (mm_index.cu) and the radix sort of K4 (mm_dedup.cu).
#pragma once

#include "mm_common.cuh"

namespace {

// ---- exclusive scan of int32 counts (n entries -> n + 1 starts), three small kernels -----------
constexpr int SCAN_BLOCK = 256;
constexpr int SCAN_ITEMS = 16;
constexpr int SCAN_TILE = SCAN_BLOCK * SCAN_ITEMS;

__global__ void __launch_bounds__(SCAN_BLOCK)
scan_tile_sums(int64_t n, const int32_t *__restrict__ in, int32_t *__restrict__ tile_sums)
{
    int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
    int32_t s = 0;
    for (int i = threadIdx.x; i < SCAN_TILE; i += SCAN_BLOCK)
        if (base + i < n) s += in[base + i];
    __shared__ int32_t sh[SCAN_BLOCK];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int w = SCAN_BLOCK / 2; w > 0; w >>= 1) {
        if (threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = sh[0];
}

__global__ void __launch_bounds__(1024)
scan_tile_offsets(int64_t ntiles, int32_t *__restrict__ tile_sums)  // in-place exclusive, 1 block
{
    __shared__ int32_t sh[1024];
    __shared__ int32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int64_t base = 0; base < ntiles; base += 1024) {
        int64_t i = base + threadIdx.x;
        int32_t v = i < ntiles ? tile_sums[i] : 0;
        sh[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            int32_t t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
            __syncthreads();
            sh[threadIdx.x] += t;
            __syncthreads();
        }
        if (i < ntiles) tile_sums[i] = carry + sh[threadIdx.x] - v;
        __syncthreads();
        if (threadIdx.x == 0) carry += sh[1023];
        __syncthreads();
    }
}

__global__ void __launch_bounds__(SCAN_BLOCK)
scan_apply(int64_t n, const int32_t *__restrict__ in, const int32_t *__restrict__ tile_offsets,
           int32_t *__restrict__ out /* n + 1 */)
{
    int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    int32_t v[SCAN_ITEMS];
    int32_t s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        v[i] = base + i < n ? in[base + i] : 0;
        s += v[i];
    }
    __shared__ int32_t sh[SCAN_BLOCK];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 1; o < SCAN_BLOCK; o <<= 1) {
        int32_t t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
        __syncthreads();
        sh[threadIdx.x] += t;
        __syncthreads();
    }
    int32_t run = tile_offsets[blockIdx.x] + sh[threadIdx.x] - s;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        if (base + i < n) out[base + i] = run;
        run += v[i];
        if (base + i == n - 1) out[n] = run;
    }
}


// in[n] -> out[n + 1] (out[n] = total); tile_sums: scratch of mm_scan_tiles(n) int32; in and out may not alias
inline int64_t mm_scan_tiles(int64_t n) { return (n + SCAN_TILE - 1) / SCAN_TILE; }

inline void mm_exclusive_scan_i32(int64_t n, const int32_t *in, int32_t *out, int32_t *tile_sums, cudaStream_t stream)
{
    const int64_t ntiles = mm_scan_tiles(n);
    scan_tile_sums<<<(int)ntiles, SCAN_BLOCK, 0, stream>>>(n, in, tile_sums);
    scan_tile_offsets<<<1, 1024, 0, stream>>>(ntiles, tile_sums);
    scan_apply<<<(int)ntiles, SCAN_BLOCK, 0, stream>>>(n, in, tile_sums, out);
}

}  // namespace
