// mm_host.cu -- one-shot end-to-end entry point over HOST buffers (what a ctypes/cffi caller without
// torch uses for a single interpolation): upload the source mesh, build geometry + index, run the
// chunked three-stream pipeline of mm_source.cu, release.
//
// Copy order on the PCIe link: nodes -> fields -> target-point chunks.  The index is built from the
// nodes while the fields are still streaming in; K1/K2 of the first chunks overlap the tail of the field
// upload (only K3 waits for it), and the values of chunk i go back to the host while chunk i+1 is
// uploaded and computed (full duplex).  All device memory comes from the library's own stream-ordered
// pool, so repeated calls re-use their blocks; mm_host_release() returns them to the driver.
#include "mm_common.cuh"

int mm_source_create_host_impl(mm_source_t **out, int order, int dim, int64_t E, const double *nodes, int F,
                               const double *fields, int gll_points_form, bool defer_fields);

extern "C" int mm_host_release(void) { return mm_pool_trim(); }

extern "C" int mm_interpolate_host(int order, int dim, int64_t E, const double *nodes, int F,
                                   const double *fields, int64_t N, const double *pts, int k,
                                   int gll_points_form, const mm_locate_params *params,
                                   double *values, int32_t *elem, double *xi, int64_t *num_failed)
{
    MM_REQUIRE(mm_valid_order(order), MM_ERR_INVALID, "mm_interpolate_host: order %d", order);
    MM_REQUIRE(dim == 2 || dim == 3, MM_ERR_INVALID, "mm_interpolate_host: dim %d", dim);
    MM_REQUIRE(E > 0 && F >= 1 && N >= 0 && k >= 1 && k <= 64, MM_ERR_INVALID,
               "mm_interpolate_host: sizes");
    MM_REQUIRE(nodes && fields && params && values && (N == 0 || pts), MM_ERR_INVALID,
               "mm_interpolate_host: null buffer");
    if (num_failed) *num_failed = 0;
    if (N == 0) return MM_OK;
    mm_source_t *src = nullptr;
    int rc = mm_source_create_host_impl(&src, order, dim, E, nodes, F, fields, gll_points_form, true);
    if (rc != MM_OK) return rc;
    rc = mm_source_interpolate_host(src, N, pts, k, params, values, elem, xi, num_failed);
    mm_source_destroy(src);
    return rc;
}
