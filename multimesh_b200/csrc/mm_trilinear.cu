// mm_trilinear.cu -- order-1 nodal (Exodus HEX8) path on the device, plus the two legacy
// host-pointer symbols the reference's ctypes loader binds (multi_mesh/helpers.py:43-81).
//
// The device arithmetic follows the expression trees of multi_mesh/src/trilinearinterpolator.c
// (vertex sign table :8-10, forward map :199-212, shape-function derivatives :214-227, Jacobian
// and cofactor inverse :230-257/:329-341, Newton loop incl. its residual test :260-305, weights
// :174-197, candidate loop :40-148) one IEEE operation at a time, so that weights, enclosing node
// ids and the failed count are bit-identical to the compiled reference.  One thread per point.
#include <cstdio>
#include <vector>

#include "mm_common.cuh"

namespace {

__device__ __constant__ double SR[8] = {-1, -1, +1, +1, -1, +1, +1, -1};
__device__ __constant__ double SS[8] = {-1, +1, +1, -1, -1, -1, +1, +1};
__device__ __constant__ double ST[8] = {-1, -1, -1, -1, +1, +1, +1, +1};

__device__ __forceinline__ double hex8_map1(const double (&v)[8][3], int c, double r, double s,
                                            double t)
{
    double hr = 0.5 * (r + 1.0), hs = 0.5 * (s + 1.0), ht = 0.5 * (t + 1.0);
    double v0 = v[0][c], v1 = v[1][c], v2 = v[2][c], v3 = v[3][c];
    double v4 = v[4][c], v5 = v[5][c], v6 = v[6][c], v7 = v[7][c];
    double e03 = hr * (-v0 + v3);
    double e12 = hr * (-v1 + v2);
    double e45 = hr * (-v4 + v5);
    double e76 = hr * (v6 - v7);
    double bot = -v0 + v1 - e03 + e12;
    double top = -v4 + v7 - e45 + e76;
    return v0 + e03 + hs * bot + ht * (-v0 + v4 - e03 + e45 - hs * bot + hs * top);
}

__device__ __forceinline__ void hex8_weights(const double (&q)[3], double (&w)[8])
{
    double r = q[0], s = q[1], t = q[2];
    double rst = 0.125 * r * s * t, rs = 0.125 * r * s, rt = 0.125 * r * t, st = 0.125 * s * t;
    double r8 = 0.125 * r, s8 = 0.125 * s, t8 = 0.125 * t;
    w[0] = -rst + rs + rt - r8 + st - s8 - t8 + 0.125;
    w[1] = +rst - rs + rt - r8 - st + s8 - t8 + 0.125;
    w[2] = -rst + rs - rt + r8 - st + s8 - t8 + 0.125;
    w[3] = +rst - rs - rt + r8 + st - s8 - t8 + 0.125;
    w[4] = +rst + rs - rt - r8 - st - s8 + t8 + 0.125;
    w[5] = -rst - rs + rt + r8 - st - s8 + t8 + 0.125;
    w[6] = +rst + rs + rt + r8 + st + s8 + t8 + 0.125;
    w[7] = -rst - rs - rt - r8 + st + s8 + t8 + 0.125;
}

__device__ bool hex8_inverse(const double (&pnt)[3], const double (&vtx)[8][3], double (&sol)[3])
{
    sol[0] = sol[1] = sol[2] = 0;
    double ax = fabs(vtx[1][0] - vtx[0][0]), ay = fabs(vtx[1][1] - vtx[0][1]);
    double az = fabs(vtx[1][2] - vtx[0][2]);
    double scalexy = ax > ay ? ax : ay;
    double scale = az > scalexy ? az : scalexy;
    double tol = 1e-8 * scale;
#pragma unroll 1
    for (int it = 0; it < 50; ++it) {
        double obj[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) obj[c] = pnt[c] - hex8_map1(vtx, c, sol[0], sol[1], sol[2]);
        // the reference tests component 0 twice and never component 2 (:290-291); kept on purpose
        if (fabs(obj[0]) < tol && fabs(obj[1]) < tol && fabs(obj[0]) < tol) return true;
        double jac[3][3];
#pragma unroll
        for (int q = 0; q < 3; ++q)
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                double sum = 0;
#pragma unroll
                for (int a = 0; a < 8; ++a) {
                    double dn;
                    if (q == 0) dn = 0.125 * SR[a] * (sol[1] * SS[a] + 1) * (sol[2] * ST[a] + 1);
                    else if (q == 1) dn = 0.125 * SS[a] * (sol[0] * SR[a] + 1) * (sol[2] * ST[a] + 1);
                    else dn = 0.125 * ST[a] * (sol[0] * SR[a] + 1) * (sol[1] * SS[a] + 1);
                    sum = sum + dn * vtx[a][j];
                }
                jac[q][j] = sum;
            }
        double det = jac[0][0] * (jac[1][1] * jac[2][2] - jac[2][1] * jac[1][2]) -
                     jac[0][1] * (jac[1][0] * jac[2][2] - jac[1][2] * jac[2][0]) +
                     jac[0][2] * (jac[1][0] * jac[2][1] - jac[1][1] * jac[2][0]);
        double id = 1 / det;
        double inv[3][3];
        inv[0][0] = (jac[1][1] * jac[2][2] - jac[2][1] * jac[1][2]) * id;
        inv[0][1] = (jac[0][2] * jac[2][1] - jac[0][1] * jac[2][2]) * id;
        inv[0][2] = (jac[0][1] * jac[1][2] - jac[0][2] * jac[1][1]) * id;
        inv[1][0] = (jac[1][2] * jac[2][0] - jac[1][0] * jac[2][2]) * id;
        inv[1][1] = (jac[0][0] * jac[2][2] - jac[0][2] * jac[2][0]) * id;
        inv[1][2] = (jac[1][0] * jac[0][2] - jac[0][0] * jac[1][2]) * id;
        inv[2][0] = (jac[1][0] * jac[2][1] - jac[2][0] * jac[1][1]) * id;
        inv[2][1] = (jac[2][0] * jac[0][1] - jac[0][0] * jac[2][1]) * id;
        inv[2][2] = (jac[0][0] * jac[1][1] - jac[1][0] * jac[0][1]) * id;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            double sum = 0;
#pragma unroll
            for (int j = 0; j < 3; ++j) sum = sum + inv[j][i] * obj[j];
            sol[i] = sol[i] + sum;
        }
    }
    return false;
}

__device__ __forceinline__ bool hex8_check_hull(const double (&pnt)[3], const double (&vtx)[8][3],
                                                double (&sol)[3])
{
    if (!hex8_inverse(pnt, vtx, sol)) return false;
#pragma unroll
    for (int c = 0; c < 3; ++c)
        if (fabs(sol[c]) > (1 + 1.0)) return false;
    return true;
}

__device__ __forceinline__ void load_vtx(const int64_t *__restrict__ conn,
                                         const double *__restrict__ nodes, int64_t e,
                                         double (&vtx)[8][3])
{
#pragma unroll
    for (int a = 0; a < 8; ++a) {
        int64_t id = conn[e * 8 + a];
#pragma unroll
        for (int c = 0; c < 3; ++c) vtx[a][c] = nodes[id * 3 + c];
    }
}

// Candidate loop of triLinearInterpolator (:40-148) for one point: the first candidate whose reference coordinates lie
// within 1.025 wins; otherwise the candidate with the smallest overshoot below 1.5 gets a second chance (:113-131).
// Returns the accepted element (sol = its reference coordinates) or -1.
// PREFIX: `cands` is only a prefix of the k-NN list -- an acceptance is what the full list would have decided (the
// loop stops at the first one), a miss decides nothing: no second chance, the caller re-runs the point with all k.
template <typename CandT, bool PREFIX>
__device__ __forceinline__ int64_t hex8_search(int64_t k, const CandT *__restrict__ cands,
                                               const int64_t *__restrict__ conn, const double *__restrict__ nodes,
                                               const double (&pnt)[3], double (&sol)[3])
{
    double vtx[8][3];
    double smallest = 99999999.9;
    int64_t best = -1, hit = -1;
    for (int64_t j = 0; j < k && hit < 0; ++j) {
        int64_t e = cands[j];
        if (e < 0) continue;  // -1 padding of a k-NN list longer than the mesh (the C twin has no such case)
        load_vtx(conn, nodes, e, vtx);
        if (hex8_check_hull(pnt, vtx, sol)) {
            double maxerr = 0.0;
#pragma unroll
            for (int c = 0; c < 3; ++c)
                if (fabs(sol[c]) > maxerr) maxerr = fabs(sol[c]);
            if (maxerr < (1 + 0.025)) hit = e;
            else if (!PREFIX && maxerr < smallest) {
                smallest = maxerr;
                best = e;
            }
        }
    }
    if (!PREFIX && hit < 0 && smallest < 1.5 && best >= 0) {  // :113-131
        load_vtx(conn, nodes, best, vtx);
        if (hex8_check_hull(pnt, vtx, sol)) hit = best;
    }
    return hit;
}

__device__ __forceinline__ void hex8_store(int64_t row, int64_t hit, const double (&sol)[3],
                                           const int64_t *__restrict__ conn, int64_t *__restrict__ enclosing,
                                           double *__restrict__ weights)
{
    double w[8];
    hex8_weights(sol, w);
#pragma unroll
    for (int a = 0; a < 8; ++a) {
        weights[row * 8 + a] = w[a];
        enclosing[row * 8 + a] = conn[hit * 8 + a];
    }
}

__global__ void __launch_bounds__(128)
trilinear_kernel(int64_t k, int64_t npoints, const int64_t *__restrict__ nearest,
                 const int64_t *__restrict__ conn, int64_t *__restrict__ enclosing,
                 const double *__restrict__ nodes, double *__restrict__ weights,
                 const double *__restrict__ points, unsigned long long *__restrict__ num_failed)
{
    unsigned long long failed = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < npoints;
         i += (int64_t)gridDim.x * blockDim.x) {
        double pnt[3] = {points[i * 3], points[i * 3 + 1], points[i * 3 + 2]};
        double sol[3];
        const int64_t hit = hex8_search<int64_t, false>(k, nearest + i * k, conn, nodes, pnt, sol);
        if (hit >= 0) hex8_store(i, hit, sol, conn, enclosing, weights);
        else failed += 1;  // outputs stay as the caller initialised them (zeros), :132-145
    }
    for (int o = 16; o > 0; o >>= 1) failed += __shfl_xor_sync(0xffffffffu, failed, o);
    if ((threadIdx.x & 31) == 0 && failed) atomicAdd(num_failed, failed);
}

// The same search inside the progressive pipeline (mm_trilinear_indexed): points come as the cell-sorted 32-byte
// query records {x, y, z, caller's row}, candidates as int32 rows of the first pass (PREFIX, k = 4) or of the re-run
// (all k).  list (optional): point i is record list[i] and owns candidate row i -- the re-run's work list, whose
// length *n_dev lives on the device.  PREFIX: a point without acceptance is appended to the work list instead of
// being counted as failed.
template <bool PREFIX>
__global__ void __launch_bounds__(128)
trilinear_rec_kernel(int k, int64_t npoints, const long long *__restrict__ n_dev, int64_t n_off,
                     const int32_t *__restrict__ list, const double4 *__restrict__ recs,
                     const int32_t *__restrict__ cands, const int64_t *__restrict__ conn,
                     const double *__restrict__ nodes, int64_t *__restrict__ enclosing, double *__restrict__ weights,
                     unsigned long long *__restrict__ num_failed, int32_t *__restrict__ unresolved_list,
                     unsigned long long *__restrict__ unresolved_count)
{
    if (n_dev) {
        const long long have = *n_dev - n_off;
        npoints = have < 0 ? 0 : (have < npoints ? have : npoints);
    }
    const int lane = threadIdx.x & 31;
    unsigned long long failed = 0;
    for (int64_t base = blockIdx.x * (int64_t)blockDim.x + threadIdx.x - lane; base < npoints;
         base += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = base + lane;
        const bool valid = i < npoints;
        int64_t n = 0, hit = -1;
        if (valid) {
            n = list ? (int64_t)list[i] : i;
            const double4 r = recs[n];
            const double pnt[3] = {r.x, r.y, r.z};
            double sol[3];
            hit = hex8_search<int32_t, PREFIX>(k, cands + i * (int64_t)k, conn, nodes, pnt, sol);
            if (hit >= 0) hex8_store((int64_t)(int32_t)__double_as_longlong(r.w), hit, sol, conn, enclosing, weights);
            else if (!PREFIX) failed += 1;
        }
        if (PREFIX) {
            const bool open = valid && hit < 0;
            const unsigned mask = __ballot_sync(0xffffffffu, open);
            if (mask) {
                const int leader = __ffs(mask) - 1;
                unsigned long long at = 0;
                if (lane == leader) at = atomicAdd(unresolved_count, (unsigned long long)__popc(mask));
                at = __shfl_sync(0xffffffffu, at, leader);
                if (open) unresolved_list[at + __popc(mask & ((1u << lane) - 1u))] = (int32_t)n;
            }
        }
    }
    if (!PREFIX) {
        for (int o = 16; o > 0; o >>= 1) failed += __shfl_xor_sync(0xffffffffu, failed, o);
        if (lane == 0 && failed) atomicAdd(num_failed, failed);
    }
}

// values[f][n] = sum_a param[f][enc[n][a]] * w[n][a], a ascending
__global__ void __launch_bounds__(256)
gather_nodal_kernel(int F, int64_t npm, const double *__restrict__ param, int64_t N,
                    const int64_t *__restrict__ enc, const double *__restrict__ w,
                    double *__restrict__ values)
{
    for (int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; n < N;
         n += (int64_t)gridDim.x * blockDim.x) {
        int64_t id[8];
        double ww[8];
#pragma unroll
        for (int a = 0; a < 8; ++a) {
            id[a] = enc[n * 8 + a];
            ww[a] = w[n * 8 + a];
        }
        // four fields at a time: 32 independent loads in flight per thread (the gathers hit L2 at random; with one
        // field per trip the kernel waited on 8 loads at a time -- ncu: 36 cycles of long-scoreboard stall per issue).
        // Every output is still the same sequential sum over a = 0..7.
        int f = 0;
        for (; f + 4 <= F; f += 4) {
            const double *pf = param + (int64_t)f * npm;
            double acc[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
            for (int a = 0; a < 8; ++a)
#pragma unroll
                for (int u = 0; u < 4; ++u) acc[u] = acc[u] + pf[(int64_t)u * npm + id[a]] * ww[a];
#pragma unroll
            for (int u = 0; u < 4; ++u) values[(int64_t)(f + u) * N + n] = acc[u];
        }
        for (; f + 2 <= F; f += 2) {
            const double *pf = param + (int64_t)f * npm;
            double acc[2] = {0.0, 0.0};
#pragma unroll
            for (int a = 0; a < 8; ++a)
#pragma unroll
                for (int u = 0; u < 2; ++u) acc[u] = acc[u] + pf[(int64_t)u * npm + id[a]] * ww[a];
#pragma unroll
            for (int u = 0; u < 2; ++u) values[(int64_t)(f + u) * N + n] = acc[u];
        }
        for (; f < F; ++f) {
            const double *pf = param + (int64_t)f * npm;
            double acc = 0.0;
#pragma unroll
            for (int a = 0; a < 8; ++a) acc = acc + pf[id[a]] * ww[a];
            values[(int64_t)f * N + n] = acc;
        }
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

extern "C" int mm_trilinear(int64_t k, int64_t npoints, const int64_t *nearest,
                            const int64_t *connectivity, int64_t *enclosing, const double *nodes,
                            double *weights, const double *points, int64_t *num_failed,
                            void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    MM_REQUIRE(k >= 1 && npoints >= 0, MM_ERR_INVALID, "mm_trilinear: sizes");
    MM_REQUIRE(num_failed, MM_ERR_INVALID, "mm_trilinear: null num_failed");
    MM_CUDA(cudaMemsetAsync(num_failed, 0, sizeof(int64_t), stream));
    if (npoints == 0) return MM_OK;
    MM_REQUIRE(nearest && connectivity && enclosing && nodes && weights && points, MM_ERR_INVALID,
               "mm_trilinear: null buffer");
    trilinear_kernel<<<blocks_for(npoints, 128), 128, 0, stream>>>(
        k, npoints, nearest, connectivity, enclosing, nodes, weights, points,
        reinterpret_cast<unsigned long long *>(num_failed));
    MM_CUDA(cudaGetLastError());
    return MM_OK;
}

// launcher of trilinear_rec_kernel for mm_trilinear_indexed (mm_pipeline.cu)
int mm_trilinear_records(bool prefix, int k, int64_t npoints, const int64_t *n_dev, int64_t n_off, const int32_t *list,
                         const double *recs, const int32_t *cands, const int64_t *connectivity, const double *nodes,
                         int64_t *enclosing, double *weights, int64_t *num_failed, int32_t *unresolved_list,
                         int64_t *unresolved_count, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (npoints == 0) return MM_OK;
    const int grid = blocks_for(npoints, 128);
    auto nd = reinterpret_cast<const long long *>(n_dev);
    auto r4 = reinterpret_cast<const double4 *>(recs);
    auto nf = reinterpret_cast<unsigned long long *>(num_failed);
    auto uc = reinterpret_cast<unsigned long long *>(unresolved_count);
    if (prefix)
        trilinear_rec_kernel<true><<<grid, 128, 0, stream>>>(k, npoints, nd, n_off, list, r4, cands, connectivity, nodes,
                                                             enclosing, weights, nf, unresolved_list, uc);
    else
        trilinear_rec_kernel<false><<<grid, 128, 0, stream>>>(k, npoints, nd, n_off, list, r4, cands, connectivity,
                                                              nodes, enclosing, weights, nf, unresolved_list, uc);
    MM_CUDA(cudaGetLastError());
    return MM_OK;
}

extern "C" int mm_gather_nodal(int F, int64_t npoints_mesh, const double *param, int64_t N,
                               const int64_t *enclosing, const double *weights, double *values,
                               void *stream)
{
    MM_REQUIRE(F >= 1 && N >= 0 && npoints_mesh >= 0, MM_ERR_INVALID, "mm_gather_nodal: sizes");
    if (N == 0) return MM_OK;
    MM_REQUIRE(param && enclosing && weights && values, MM_ERR_INVALID, "mm_gather_nodal: null");
    gather_nodal_kernel<<<blocks_for(N, 256), 256, 0, (cudaStream_t)stream>>>(
        F, npoints_mesh, param, N, enclosing, weights, values);
    MM_CUDA(cudaGetLastError());
    return MM_OK;
}

// ------------------------------------------------------------------------------------------------
// Legacy host-pointer symbols (exact reference signatures).  They stage through device memory on
// the current CUDA device; on any CUDA failure they report on stderr -- `centroid` leaves its
// output untouched, `triLinearInterpolator` returns npoints (= everything failed).  There is no
// CPU implementation behind them.
// ------------------------------------------------------------------------------------------------
namespace {
struct dev_buf {
    void *p = nullptr;
    ~dev_buf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1); }
};

int64_t max_index(const long long *a, int64_t n)
{
    long long m = -1;
    for (int64_t i = 0; i < n; ++i) m = a[i] > m ? a[i] : m;
    return m;
}
}  // namespace

extern "C" void centroid(long long int ndim, long long int nelem, long long int npe,
                         long long int *connectivity, double *points, double *cent)
{
    if (nelem <= 0 || ndim <= 0 || npe <= 0) return;
    int64_t npts = max_index(connectivity, nelem * npe) + 1;
    dev_buf dc, dp, dout;
    bool ok = dc.alloc(sizeof(int64_t) * nelem * npe) == cudaSuccess &&
              dp.alloc(sizeof(double) * npts * ndim) == cudaSuccess &&
              dout.alloc(sizeof(double) * nelem * ndim) == cudaSuccess;
    ok = ok && cudaMemcpy(dc.p, connectivity, sizeof(int64_t) * nelem * npe, cudaMemcpyHostToDevice) == cudaSuccess;
    ok = ok && cudaMemcpy(dp.p, points, sizeof(double) * npts * ndim, cudaMemcpyHostToDevice) == cudaSuccess;
    ok = ok && mm_centroid_conn(ndim, nelem, npe, (const int64_t *)dc.p, (const double *)dp.p,
                                (double *)dout.p, nullptr) == MM_OK;
    ok = ok && cudaMemcpy(cent, dout.p, sizeof(double) * nelem * ndim, cudaMemcpyDeviceToHost) == cudaSuccess;
    if (!ok) fprintf(stderr, "multimesh_b200: centroid() failed on the CUDA device: %s\n", mm_last_error());
}

extern "C" long long int triLinearInterpolator(long long int k, long long int npoints,
                                               long long int *nearest, long long int *connectivity,
                                               long long int *enclosing, double *nodes,
                                               double *weights, double *points)
{
    if (npoints <= 0 || k <= 0) return 0;
    // the C interface carries no array lengths: derive them from the indices that will be read
    int64_t nelem = max_index(nearest, npoints * k) + 1;
    int64_t nnode = max_index(connectivity, nelem * 8) + 1;
    dev_buf dn, dc, de, dx, dw, dp, df;
    bool ok = dn.alloc(sizeof(int64_t) * npoints * k) == cudaSuccess &&
              dc.alloc(sizeof(int64_t) * nelem * 8) == cudaSuccess &&
              de.alloc(sizeof(int64_t) * npoints * 8) == cudaSuccess &&
              dx.alloc(sizeof(double) * nnode * 3) == cudaSuccess &&
              dw.alloc(sizeof(double) * npoints * 8) == cudaSuccess &&
              dp.alloc(sizeof(double) * npoints * 3) == cudaSuccess &&
              df.alloc(sizeof(int64_t)) == cudaSuccess;
    auto h2d = [&](void *d, const void *h, size_t b) {
        return cudaMemcpy(d, h, b, cudaMemcpyHostToDevice) == cudaSuccess;
    };
    ok = ok && h2d(dn.p, nearest, sizeof(int64_t) * npoints * k) &&
         h2d(dc.p, connectivity, sizeof(int64_t) * nelem * 8) &&
         h2d(de.p, enclosing, sizeof(int64_t) * npoints * 8) &&
         h2d(dx.p, nodes, sizeof(double) * nnode * 3) &&
         h2d(dw.p, weights, sizeof(double) * npoints * 8) &&
         h2d(dp.p, points, sizeof(double) * npoints * 3);
    ok = ok && mm_trilinear(k, npoints, (const int64_t *)dn.p, (const int64_t *)dc.p,
                            (int64_t *)de.p, (const double *)dx.p, (double *)dw.p,
                            (const double *)dp.p, (int64_t *)df.p, nullptr) == MM_OK;
    int64_t nfailed = npoints;
    ok = ok && cudaMemcpy(enclosing, de.p, sizeof(int64_t) * npoints * 8, cudaMemcpyDeviceToHost) == cudaSuccess &&
         cudaMemcpy(weights, dw.p, sizeof(double) * npoints * 8, cudaMemcpyDeviceToHost) == cudaSuccess &&
         cudaMemcpy(&nfailed, df.p, sizeof(int64_t), cudaMemcpyDeviceToHost) == cudaSuccess;
    if (!ok) {
        fprintf(stderr, "multimesh_b200: triLinearInterpolator() failed on the CUDA device: %s\n",
                mm_last_error());
        return npoints;
    }
    return nfailed;
}
