// mm_geometry.cu -- K0: element centroids + AABBs, connectivity-gathered centroids, sphere map.
//
// Centroid = (sequential sum over the P control nodes, a ascending) / P.  The sequential order is
// what makes it bit-equal to np.mean(points, axis=1) (salvus_mesh_reader.py:99-100) and to the
// reference's centroid.c:15-24; a tree reduction would not be.
#include "mm_common.cuh"
#include "mm_newton.cuh"

namespace {

// one thread per (element, coordinate); the P reads of a thread walk its element block.
__global__ void __launch_bounds__(256)
element_geometry_kernel(int P, int dim, int64_t E, const double *__restrict__ nodes,
                        double *__restrict__ centroid, double *__restrict__ aabb)
{
    int64_t total = E * dim;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
         t += (int64_t)gridDim.x * blockDim.x) {
        int64_t e = t / dim;
        int c = (int)(t - e * dim);
        const double *x = nodes + (e * P) * dim + c;
        double v = x[0];
        double s = 0.0 + v, lo = v, hi = v;
        for (int a = 1; a < P; ++a) {
            v = x[(int64_t)a * dim];
            s = s + v;
            lo = v < lo ? v : lo;
            hi = v > hi ? v : hi;
        }
        if (centroid) centroid[t] = s / (double)P;
        if (aabb) {
            aabb[(e * 2 + 0) * dim + c] = lo;
            aabb[(e * 2 + 1) * dim + c] = hi;
        }
    }
}

// centroid.c:3-25 on the device: sum of connectivity-gathered nodes, then one divide.
__global__ void __launch_bounds__(256)
centroid_conn_kernel(int64_t ndim, int64_t nelem, int64_t npe, const int64_t *__restrict__ conn,
                     const double *__restrict__ points, double *__restrict__ cent)
{
    int64_t total = nelem * ndim;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
         t += (int64_t)gridDim.x * blockDim.x) {
        int64_t e = t / ndim;
        int64_t c = t - e * ndim;
        double s = 0.;
        for (int64_t a = 0; a < npe; ++a) s = s + points[conn[e * npe + a] * ndim + c];
        cent[t] = s / (double)npe;
    }
}

// map_to_sphere (interpolator.py:1136-1144): x <- ((x * r_earth) * rad_1D) / r, per component,
// r = sqrt((x^2 + y^2) + z^2), only where r > 0.
__global__ void __launch_bounds__(256)
map_to_sphere_kernel(int64_t n, double *__restrict__ nodes, const double *__restrict__ rad,
                     double r_earth)
{
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n;
         t += (int64_t)gridDim.x * blockDim.x) {
        double x = nodes[t * 3 + 0], y = nodes[t * 3 + 1], z = nodes[t * 3 + 2];
        double r = sqrt((x * x + y * y) + z * z);
        if (r > 0) {
            double q = rad[t];
            nodes[t * 3 + 0] = ((x * r_earth) * q) / r;
            nodes[t * 3 + 1] = ((y * r_earth) * q) / r;
            nodes[t * 3 + 2] = ((z * r_earth) * q) / r;
        }
    }
}

// affine pre-solve per element: pre[e] = {ref, x(0) - ref, Jinv(0)}, ref = first node, evaluated on
// the ref-shifted nodes (thread per element)
template <int ORDER, int DIM>
__global__ void __launch_bounds__(128)
presolve_kernel(const mm_gll_table T, int64_t E, const double *__restrict__ nodes,
                double *__restrict__ pre)
{
    constexpr int M = ORDER + 1;
    constexpr int P = DIM == 2 ? M * M : M * M * M;
    constexpr int W = 2 * DIM + DIM * DIM;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < E;
         e += (int64_t)gridDim.x * blockDim.x) {
        const double *X = nodes + e * (int64_t)(P * DIM);
        double ref[DIM], xi0[DIM], x[DIM], J[DIM][DIM];
#pragma unroll
        for (int c = 0; c < DIM; ++c) {
            ref[c] = X[c];
            xi0[c] = 0.0;
        }
        eval_map<ORDER, DIM>(T, X, ref, xi0, x, J);
        double *o = pre + e * W;
#pragma unroll
        for (int c = 0; c < DIM; ++c) {
            o[c] = ref[c];
            o[DIM + c] = x[c];
        }
        double *I = o + 2 * DIM;
        if constexpr (DIM == 2) {
            double det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
            I[0] = J[1][1] / det;
            I[1] = (-J[0][1]) / det;
            I[2] = (-J[1][0]) / det;
            I[3] = J[0][0] / det;
        } else {
            double C00 = J[1][1] * J[2][2] - J[1][2] * J[2][1];
            double C01 = J[1][2] * J[2][0] - J[1][0] * J[2][2];
            double C02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
            double C10 = J[0][2] * J[2][1] - J[0][1] * J[2][2];
            double C11 = J[0][0] * J[2][2] - J[0][2] * J[2][0];
            double C12 = J[0][1] * J[2][0] - J[0][0] * J[2][1];
            double C20 = J[0][1] * J[1][2] - J[0][2] * J[1][1];
            double C21 = J[0][2] * J[1][0] - J[0][0] * J[1][2];
            double C22 = J[0][0] * J[1][1] - J[0][1] * J[1][0];
            double det = (J[0][0] * C00 + J[0][1] * C01) + J[0][2] * C02;
            I[0] = C00 / det; I[1] = C10 / det; I[2] = C20 / det;
            I[3] = C01 / det; I[4] = C11 / det; I[5] = C21 / det;
            I[6] = C02 / det; I[7] = C12 / det; I[8] = C22 / det;
        }
    }
}

int grid_for(int64_t work, int block)
{
    int sms = mm_num_sms();
    int64_t need = (work + block - 1) / block;
    int64_t cap = (int64_t)(sms > 0 ? sms : 148) * 16;
    return (int)(need < 1 ? 1 : (need > cap ? cap : need));
}

}  // namespace

extern "C" int mm_element_geometry(int order, int dim, int64_t E, const double *nodes,
                                   double *centroid, double *aabb, void *stream)
{
    MM_REQUIRE(mm_valid_order(order), MM_ERR_INVALID, "mm_element_geometry: order %d", order);
    MM_REQUIRE(dim == 2 || dim == 3, MM_ERR_INVALID, "mm_element_geometry: dim %d", dim);
    MM_REQUIRE(E >= 0 && (E == 0 || nodes), MM_ERR_INVALID, "mm_element_geometry: null nodes");
    if (E == 0) return MM_OK;
    int P = mm_pow(order + 1, dim);
    element_geometry_kernel<<<grid_for(E * dim, 256), 256, 0, (cudaStream_t)stream>>>(
        P, dim, E, nodes, centroid, aabb);
    MM_CUDA(cudaGetLastError());
    return MM_OK;
}

extern "C" int mm_centroid_conn(int64_t ndim, int64_t nelem, int64_t npe,
                                const int64_t *connectivity, const double *points,
                                double *centroid, void *stream)
{
    MM_REQUIRE(ndim > 0 && npe > 0 && nelem >= 0, MM_ERR_INVALID, "mm_centroid_conn: sizes");
    if (nelem == 0) return MM_OK;
    MM_REQUIRE(connectivity && points && centroid, MM_ERR_INVALID, "mm_centroid_conn: null");
    centroid_conn_kernel<<<grid_for(nelem * ndim, 256), 256, 0, (cudaStream_t)stream>>>(
        ndim, nelem, npe, connectivity, points, centroid);
    MM_CUDA(cudaGetLastError());
    return MM_OK;
}

extern "C" int mm_map_to_sphere(int64_t n, double *nodes, const double *radius_1d,
                                double r_earth, void *stream)
{
    if (n == 0) return MM_OK;
    MM_REQUIRE(n > 0 && nodes && radius_1d, MM_ERR_INVALID, "mm_map_to_sphere: arguments");
    map_to_sphere_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(n, nodes, radius_1d,
                                                                             r_earth);
    MM_CUDA(cudaGetLastError());
    return MM_OK;
}

extern "C" int mm_element_presolve(int order, int dim, int64_t E, const double *nodes,
                                   double *presolve, void *stream)
{
    MM_REQUIRE(mm_valid_order(order), MM_ERR_INVALID, "mm_element_presolve: order %d", order);
    MM_REQUIRE(dim == 2 || dim == 3, MM_ERR_INVALID, "mm_element_presolve: dim %d", dim);
    if (E == 0) return MM_OK;
    MM_REQUIRE(E > 0 && nodes && presolve, MM_ERR_INVALID, "mm_element_presolve: arguments");
    mm_gll_table T;
    mm_make_table(order, &T);
#define MM_PRE(O, D)                                                                          \
    if (order == O && dim == D) {                                                             \
        presolve_kernel<O, D><<<grid_for(E, 128), 128, 0, (cudaStream_t)stream>>>(T, E, nodes, \
                                                                                 presolve);   \
        MM_CUDA(cudaGetLastError());                                                          \
        return MM_OK;                                                                         \
    }
    MM_PRE(1, 2) MM_PRE(2, 2) MM_PRE(4, 2) MM_PRE(1, 3) MM_PRE(2, 3) MM_PRE(4, 3)
#undef MM_PRE
    mm_set_error("mm_element_presolve: unsupported order/dim");
    return MM_ERR_UNSUPPORTED;
}
