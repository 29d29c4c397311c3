/*
 * multimesh_b200.h -- C-ABI of the B200-native mesh-to-mesh interpolation path.
 *
 * This is the drop-in boundary for the hot path of solvithrastar/MultiMesh: every entry point
 * below names the reference interface it replaces (paths relative to the reference root).
 * Conventions:
 *   - extern "C", plain pointers and sizes, no torch / C++ types;
 *   - mm_* functions take DEVICE pointers (sm_100a, CUDA 12.9), caller-owned buffers,
 *     a cudaStream_t passed as void*, and return 0 (MM_OK) or a negative error code;
 *     mm_last_error() returns a thread-local message for the last failure;
 *   - base pointers of double arrays must be 16-byte aligned (bulk-async copies);
 *   - no hidden allocation except the opaque mm_index handle (create/destroy);
 *   - no global mutable state besides per-device constant tables; thread-safe per handle+stream;
 *   - the two legacy symbols `centroid` and `triLinearInterpolator` keep the reference's exact
 *     signatures and operate on HOST pointers, so multi_mesh/helpers.py:load_lib() keeps working.
 * There is NO CPU fallback: without a CUDA device every mm_* call fails with MM_ERR_CUDA.
 *
 * All floating-point work is IEEE binary64 in a documented operation order and
 * matches the CPU oracle bit for bit.
 */
#ifndef MULTIMESH_B200_H
#define MULTIMESH_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MM_OK 0
#define MM_ERR_INVALID (-1)     /* bad argument (order/dim/k/alignment/null pointer) */
#define MM_ERR_CUDA (-2)        /* CUDA runtime error or no device */
#define MM_ERR_UNSUPPORTED (-3) /* valid request outside the implemented range */
#define MM_ERR_NOMEM (-4)

#define MM_VERSION 100

int mm_version(void);
const char *mm_last_error(void);

/* ------------------------------------------------------------------------------------------
 * K0  element geometry: centroid (sequential mean of the P control nodes) and inclusive AABB.
 * Replaces: SalvusMesh.get_element_centroids (components/salvus_mesh_reader.py:99-100),
 *           _find_gll_centroids (components/interpolator.py:1389-1406),
 *           the min/max of boundary_box_check (components/interpolator.py:1360).
 *   nodes    [E][P][dim] f64, P = (order+1)^dim
 *   centroid [E][dim]    f64 (out, may be NULL)
 *   aabb     [E][2][dim] f64 (out, may be NULL): row 0 = min, row 1 = max
 * ---------------------------------------------------------------------------------------- */
int mm_element_geometry(int order, int dim, int64_t E, const double *nodes, double *centroid,
                        double *aabb, void *stream);

/* Affine pre-solve per element (optional accelerator of K2): presolve [E][2*dim + dim*dim] f64 =
 * { ref[dim] = first control node, x(xi=0) - ref [dim], inverse Jacobian at xi=0 [dim][dim] }, the
 * last two evaluated on ref-shifted nodes.  K2 then starts Newton at xi0 = Jinv0 ((p - ref) - x0)
 * -- the first Newton step with point-independent quantities -- so an exactly affine element needs
 * one map evaluation instead of two.  Converged xi agree with the xi0 = 0
 * start to roundoff (1e-16); the CPU oracle implements the same start. */
int mm_element_presolve(int order, int dim, int64_t E, const double *nodes, double *presolve,
                        void *stream);

/* In-place x <- x * r_earth * z_node_1D / |x| for |x| > 0.
 * Replaces: map_to_sphere (components/interpolator.py:1125-1144). nodes [n][3], radius_1d [n]. */
int mm_map_to_sphere(int64_t n, double *nodes, const double *radius_1d, double r_earth,
                     void *stream);

/* ------------------------------------------------------------------------------------------
 * K1  spatial index (uniform grid, counting-sorted) + exact k-nearest-neighbour query.
 * Replaces: pykdtree KDTree(data) and .query(pts, k)  (components/interpolator.py:101-105,
 *           255-264, 363-373, 515-522, 678, 751-756, 898-902, 949-951, 1053, 1176-1178;
 *           utils.py:199-200).
 * Neighbour order is the canonical total order (d2, index) with
 *   d2 = (dx*dx + dy*dy) + dz*dz evaluated in binary64 without FMA,
 * which fixes the tie order a KD-tree leaves implementation-defined.
 * ---------------------------------------------------------------------------------------- */
typedef struct mm_index mm_index_t;

/* points [M][dim] f64 (device); the index keeps its own sorted copy. Synchronises the stream. */
int mm_index_create(mm_index_t **out, int dim, int64_t M, const double *points, void *stream);
int mm_index_destroy(mm_index_t *index);
/* info[0]=M, info[1]=dim, info[2..4]=cells per axis, info[5]=non-empty cells, info[6]=bytes held */
int mm_index_info(const mm_index_t *index, int64_t info[8], double *cell_size);

/* Optional: build the index's SITE TABLE (one record per distinct coordinate).  The GLL-point form stores
 * every node shared between elements up to 8 times; with the table the first pass of mm_interpolate
 * searches distinct coordinates only (3-4x fewer distance evaluations).  Mutates the index: call it once
 * after mm_index_create, before the index is shared between threads / streams.  Synchronises the stream. */
int mm_index_prepare_sites(mm_index_t *index, void *stream);

/* idx [N][k] int32 (out): neighbour ids divided (integer division) by `divisor`
 *   divisor = 1: plain ids;  divisor = P: the reference's GLL-point form `idx // P`
 *   (components/interpolator.py:116-118, 754-756).  Missing neighbours (k > M) are -1.
 * d2  [N][k] f64  (out, may be NULL): squared distances. 1 <= k <= 64. */
int mm_knn(const mm_index_t *index, int64_t N, const double *pts, int k, int32_t divisor,
           int32_t *idx, double *d2, void *stream);

/* ------------------------------------------------------------------------------------------
 * K2  point-in-element location: candidate iteration, optional AABB prefilter, fp64 Newton
 *     inversion of the order-n isoparametric map, accept test and fallback.
 * Replaces: inverse_transform -> salvus.fem InverseCoordinateTransformWrapper
 *             (components/interpolator.py:1370-1386),
 *           boundary_box_check (components/interpolator.py:1350-1367),
 *           _check_if_inside_element  V1 (components/interpolator.py:1409-1473),
 *           get_element_weights.check_inside V2 (components/interpolator.py:1181-1233),
 *           get_element_weights_layered.check_inside V3 (components/interpolator.py:1271-1297),
 *           v2_interpolation_tools.get_element_weights V4 (v2_interpolation_tools.py:71-164),
 *           cli._check_if_inside_element V5 (scripts/cli.py:401-430),
 *           and the per-point loops find_gll_coeffs / fill_value_array
 *             (components/interpolator.py:1476-1597).
 * ---------------------------------------------------------------------------------------- */
enum { MM_FB_FAIL = 0, MM_FB_MAGIC = 1, MM_FB_SNAP = 2, MM_FB_MINL1 = 3 };
enum {
    MM_ST_ACCEPTED = 0,
    MM_ST_FB_INSIDE_MAGIC = 1,
    MM_ST_FB_NEAR_OK = 2,
    MM_ST_FB_NEAR_MAGIC = 3,
    MM_ST_FB_NAN_MAGIC = 4, /* reference raises ValueError unless ignore_hard_elements */
    MM_ST_SNAPPED = 5,
    MM_ST_FAILED = 6, /* elem = -1, zero weights */
    MM_ST_MINL1 = 7,
    MM_ST_SNAP_NONE = 8,
    MM_ST_UNRESOLVED = 9 /* internal to the progressive search; never returned */
};

typedef struct {
    int32_t aabb_prefilter; /* V1 */
    int32_t strict;         /* 1: all |xi| < tol, 0: all |xi| <= tol */
    int32_t fallback;       /* MM_FB_* */
    int32_t reserved;
    double tol;
    double snap_clip;
    double magic_xi[3];
} mm_locate_params;

/*   nodes    [E][P][dim]   source control nodes
 *   centroid [E][dim]      from mm_element_geometry (needed when aabb_prefilter)
 *   aabb     [E][2][dim]   from mm_element_geometry (needed when aabb_prefilter)
 *   presolve [E][2*dim+dim*dim] from mm_element_presolve, or NULL (Newton starts at xi = 0)
 *   pts      [N][dim]
 *   cands    [N][k] int32  candidate element ids in neighbour order; negatives are skipped
 *   elem     [N] int32 (out), xi [N][dim] f64 (out), status [N] u8 (out, may be NULL)
 *   num_failed : device int64 (out, may be NULL) = number of points with elem = -1 */
int mm_locate(int order, int dim, int64_t E, const double *nodes, const double *centroid,
              const double *aabb, const double *presolve, int64_t N, const double *pts, int k,
              const int32_t *cands,
              const mm_locate_params *params, int32_t *elem, double *xi, uint8_t *status,
              int64_t *num_failed, void *stream);

/* Statistics for the benchmarks: while a device array of two int64 is set, the calling thread's mm_locate /
 * mm_interpolate launches add {candidates that reached Newton, map evaluations (Newton iterations)} to it
 * (a separate kernel instantiation; the regular kernels carry no counters).  NULL switches it off. */
int mm_locate_set_stats(int64_t *device_counters);

/* ------------------------------------------------------------------------------------------
 * K3  fused Lagrange weights + multi-field gather.
 *   out[n][f] = sum_a w_a(xi_n) * fields[elem_n][f][a];  rows with elem < 0 are zero.
 * Replaces: get_coefficients -> salvus.fem GetInterpolationCoefficients
 *             (components/interpolator.py:1337-1347) and the gathers
 *             np.sum(data[element] * coeffs, axis=...) (components/interpolator.py:136-138,
 *             279, 406-414, 598-605, 814-826, 974-976, 1072-1081).
 *   fields [E][F][P] f64 (MODEL/data layout), out [N][F] f64.
 * ---------------------------------------------------------------------------------------- */
int mm_interp(int order, int dim, int64_t E, int F, const double *fields, int64_t N,
              const int32_t *elem, const double *xi, double *out, void *stream);

/* Same with an output permutation: row n is written to out[perm[n]] (perm may be NULL). Used by
 * mm_interpolate, which processes the points in spatially sorted order. */
int mm_interp_perm(int order, int dim, int64_t E, int F, const double *fields, int64_t N,
                   const int32_t *elem, const double *xi, const int32_t *perm, double *out,
                   void *stream);

/* coeffs [N][P] f64 (out) = w_a(xi_n), zero rows where elem < 0 (elem may be NULL).
 * For the reference's stored interpolation matrices (components/interpolator.py:391-398,
 * 797-810). */
int mm_coeffs(int order, int dim, int64_t N, const int32_t *elem, const double *xi,
              double *coeffs, void *stream);

/* Explicit-matrix gather for cached (elements, coeffs):
 *   out[n][f] = sum_a fields[elem_n][f][a] * coeffs[n][a]   (sequential in a). */
int mm_gather_coeffs(int P, int64_t E, int F, const double *fields, int64_t N,
                     const int32_t *elem, const double *coeffs, double *out, void *stream);

/* ------------------------------------------------------------------------------------------
 * K4  target-side de-duplication and write-back (device pointers).
 * mm_unique_points replaces utils.get_unique_points = np.unique(points, axis=0, return_inverse=True)
 *   (utils.py:465-515): unique [n_unique][dim] (caller provides room for N rows) in numpy's lexicographic order,
 *   inverse [N] int32 with points[i] == unique[inverse[i]]; *n_unique is written on the HOST (the call
 *   synchronises).  -0.0 is folded onto +0.0; non-finite coordinates are rejected.
 * mm_scatter_back replaces values[recon].reshape(E, P, F).swapaxes(1, 2) (components/interpolator.py:822-826,
 *   412-427): out[e][f][p] = values[inverse[e * P + p]][f]; inverse == NULL means the identity (every GLL node was
 *   interpolated itself).
 * mm_fluid_fixup replaces the fluid / solid repair of gll_2_gll (components/interpolator.py:829-841), in place:
 *   elements with fluid[e] != 0 keep old_values; a solid element with any values[e][vs_index][p] == 0.0 is
 *   restored from old_values as a whole (vs_index < 0 skips that test).
 * ---------------------------------------------------------------------------------------- */
int mm_unique_points(int dim, int64_t N, const double *pts, int64_t *n_unique, double *unique, int32_t *inverse,
                     void *stream);
int mm_scatter_back(int64_t E, int P, int F, const double *values, const int32_t *inverse, double *out,
                    void *stream);
int mm_fluid_fixup(int64_t E, int P, int F, double *values, const double *old_values, const uint8_t *fluid,
                   int vs_index, void *stream);

/* ------------------------------------------------------------------------------------------
 * Order-1 nodal (Exodus HEX8) path, device pointers.  Arithmetic is bit-identical to
 * multi_mesh/src/trilinearinterpolator.c:40-305 (same argument meaning as
 * triLinearInterpolator) and multi_mesh/src/centroid.c:3-25.
 *   num_failed: device int64 (out).
 * ---------------------------------------------------------------------------------------- */
int mm_trilinear(int64_t nelem_to_search, int64_t npoints, const int64_t *nearest,
                 const int64_t *connectivity, int64_t *enclosing, const double *nodes,
                 double *weights, const double *points, int64_t *num_failed, void *stream);
int mm_centroid_conn(int64_t ndim, int64_t nelem, int64_t npe, const int64_t *connectivity,
                     const double *points, double *centroid, void *stream);
/* values[f][n] = sum_a param[f][enclosing[n][a]] * weights[n][a]  (a = 0..7, sequential);
 * replaces np.sum(param_exodus[:, enc] * w, axis=2) (components/interpolator.py:219-222). */
int mm_gather_nodal(int F, int64_t npoints_mesh, const double *param, int64_t N,
                    const int64_t *enclosing, const double *weights, double *values,
                    void *stream);

/* The two calls every exodus driver makes back to back -- centroid_tree.query(points, k=nelem_to_search) and
 * lib.triLinearInterpolator (components/interpolator.py:193-224; scripts/cli.py:74-99) -- as ONE stream-ordered call
 * with the progressive search of mm_interpolate: points sorted by index cell, certified 4-prefix of the k-NN list,
 * trilinear candidate loop on the prefix, re-run of the points it did not accept with all k candidates.  enclosing /
 * weights / num_failed are those of mm_knn(k) + mm_trilinear, bit for bit (a point accepted by a candidate of the
 * prefix is accepted by the same candidate of the full list: the loop of trilinearinterpolator.c:63-111 stops there).
 *   index: over the nelem element centroids (mm_centroid_conn + mm_index_create);
 *   connectivity [nelem, 8] in the C routine's vertex order; enclosing [N, 8] / weights [N, 8]: rows of points that
 *   fail keep what the caller put there (the drivers pass zeros); num_failed: device int64 (out);
 *   workspace: mm_trilinear_indexed_workspace_bytes(index, N, k) bytes, 256-byte aligned. */
size_t mm_trilinear_indexed_workspace_bytes(const mm_index_t *index, int64_t N, int k);
int mm_trilinear_indexed(const mm_index_t *index, int64_t nelem, const int64_t *connectivity, const double *nodes,
                         int64_t N, const double *pts, int k, int64_t *enclosing, double *weights,
                         int64_t *num_failed, void *workspace, size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------------------------
 * Fused device pipeline K1 -> K2 -> K3 over one batch of target points (the bench's "step").
 * Replaces the body of the reference's drivers between "arrays loaded" and "values computed":
 *   KDTree.query + find_gll_coeffs / fill_value_array / get_element_weights + the gather
 *   (components/interpolator.py:744-826, 363-427, 949-976).
 * Same results as mm_knn -> mm_locate -> mm_interp, but
 *   - the points are first counting-sorted by index cell (coherent warps, L2 locality) and the
 *     results are written back through the permutation;
 *   - the search is progressive: a first pass over a certified prefix of the canonical k-NN list (up to
 *     min(k, 8) entries) resolves most points (any prefix of the canonical list is the list of the nearest);
 *     only points whose prefix is exhausted are re-run with all k candidates and the variant's fallback;
 *   - the gather walks the source elements in memory order (points grouped by element): every field block is
 *     read once per call;
 *   - the call is STREAM-ORDERED: it never synchronises with the host (the number of points to re-run
 *     stays on the device), so it can be captured in a CUDA graph.  0 <= N < 2^31 per call.
 *   index    : over element centroids (divisor = 1) or over all GLL points (divisor = P)
 *   fields   : [E][F][P], may be NULL (locate only; out ignored)
 *   out      : [N][F];  elem [N], xi [N][dim], status [N]: optional (NULL to skip)
 *   num_failed: device int64 (may be NULL)
 *   workspace: device scratch of at least mm_interpolate_workspace_bytes(...) bytes
 * ---------------------------------------------------------------------------------------- */
size_t mm_interpolate_workspace_bytes(const mm_index_t *index, int dim, int64_t N, int k);
int mm_interpolate(const mm_index_t *index, int32_t divisor, int order, int dim, int64_t E,
                   const double *nodes, const double *centroid, const double *aabb,
                   const double *presolve /* may be NULL */, int F,
                   const double *fields, int64_t N, const double *pts, int k,
                   const mm_locate_params *params, double *out, int32_t *elem, double *xi,
                   uint8_t *status, int64_t *num_failed, void *workspace, size_t workspace_bytes,
                   void *stream);

/* Stage timing of mm_interpolate with CUDA events recorded on the caller's stream (no sync inside
 * mm_interpolate).  Stages: 0 query sort, 1 k-NN first pass, 2 locate first pass, 3 re-run of
 * unresolved points (k-NN + locate with all k), 4 gather (K3), 5 un-permute of elem/xi/status.
 * mm_profile_begin makes the CALLING THREAD's subsequent mm_interpolate calls record into `p`
 * (up to max_calls); mm_profile_end stops it; mm_profile_read synchronises the events and returns
 * the per-call stage durations in milliseconds, stage_ms[call][MM_N_STAGES]. */
#define MM_N_STAGES 6
typedef struct mm_profile mm_profile_t;
int mm_profile_create(mm_profile_t **out, int max_calls);
int mm_profile_destroy(mm_profile_t *p);
int mm_profile_begin(mm_profile_t *p);
int mm_profile_end(void);
int mm_profile_read(mm_profile_t *p, int *n_calls, float *stage_ms);

/* ------------------------------------------------------------------------------------------
 * Legacy symbols, HOST pointers, exact reference signatures (helpers.py:43-81).
 * They copy to the current CUDA device, run the kernels above and copy back.
 * ---------------------------------------------------------------------------------------- */
void centroid(long long int ndim, long long int nelem, long long int npointsperelem,
              long long int *connectivity, double *points, double *centroid);
long long int triLinearInterpolator(long long int nelem_to_search, long long int npoints,
                                    long long int *nearest_element_indices,
                                    long long int *connectivity,
                                    long long int *enclosing_elem_indices, double *nodes,
                                    double *weights, double *points);

/* ------------------------------------------------------------------------------------------
 * End-to-end convenience over HOST buffers (the call the e2e benchmark times):
 * H2D of the source mesh and targets, index build over centroids (gll_points_form = 0) or over
 * all GLL points (gll_points_form = 1, the gll_2_gll form), k-NN, locate, gather, D2H
 * (= mm_source_create_host + mm_source_interpolate_host + mm_source_destroy, with the field upload
 * overlapping the index build; any N).
 *   values [N][F] f64 (out, host), elem [N] int32 (out, host, may be NULL),
 *   xi [N][dim] (out, host, may be NULL).  Returns MM_OK or an error; *num_failed (host).
 * ---------------------------------------------------------------------------------------- */
int mm_interpolate_host(int order, int dim, int64_t E, const double *nodes, int F,
                        const double *fields, int64_t N, const double *pts, int k,
                        int gll_points_form, const mm_locate_params *params, double *values,
                        int32_t *elem, double *xi, int64_t *num_failed);
/* All device memory behind mm_interpolate_host / mm_source_* / mm_index_* comes from a stream-ordered pool
 * PRIVATE to this library (the device's default pool is never reconfigured); freed blocks stay cached for
 * the next call.  mm_pool_trim() (alias mm_host_release) returns the cached blocks to the driver. */
int mm_host_release(void);
int mm_pool_trim(void);

/* ------------------------------------------------------------------------------------------
 * Resident source mesh.  The reference re-reads the source model and rebuilds its KD-tree in every call
 * (components/interpolator.py:660-760, 288-373); its users interpolate between a fixed mesh pair many
 * times (sum of gradients, model updates).  The handle keeps nodes, fields, centroid/AABB, affine
 * pre-solve and the spatial index (+ site table in the GLL-point form) in HBM; a call then moves only
 * target points in and values out.
 *   mm_source_create_host   : nodes [E][P][dim], fields [E][F][P] (F may be 0 / fields NULL) on the HOST
 *                             (pinned memory makes the upload asynchronous); copies them to the device
 *   mm_source_create_device : the same from DEVICE arrays, which are BORROWED (caller keeps them alive,
 *                             16-byte aligned); work already enqueued on `stream` (e.g. an NCCL broadcast
 *                             of the mesh from the loading rank) is ordered before the build
 *   mm_source_set_fields_host: replace the fields (new model on the same geometry), host-owned sources
 *   mm_source_interpolate   : device pointers, one mm_interpolate on the caller's stream (stream-ordered,
 *                             no host synchronisation unless the internal workspace has to grow)
 *   mm_source_interpolate_host: HOST pointers; the points are cut into chunks (default 1 M points after a
 *                             short ramp of smaller ones, MM_HOST_CHUNK / MM_HOST_RAMP=0 override) and pipelined over three streams -- H2D of chunk
 *                             i+1, K1-K3 of chunk i, D2H of chunk i-1 -- so both PCIe directions and the SMs
 *                             work concurrently.  Any N (chunks are < 2^31 points each).  values/elem/xi:
 *                             any may be NULL (not all).  Returns after the last byte has arrived.
 * Results are identical to mm_interpolate for every chunking (each point is a pure function of the source).
 * gll_points_form: 0 = index over element centroids, 1 = over all GLL points (`idx // P`, gll_2_gll).
 * info: E, P, dim, F, order, gll_points_form, bytes resident, device.
 * A handle is bound to the device that was current at creation; not thread-safe (one call at a time).
 * ---------------------------------------------------------------------------------------- */
typedef struct mm_source mm_source_t;
int mm_source_create_host(mm_source_t **out, int order, int dim, int64_t E, const double *nodes, int F,
                          const double *fields, int gll_points_form);
int mm_source_create_device(mm_source_t **out, int order, int dim, int64_t E, const double *nodes, int F,
                            const double *fields, int gll_points_form, void *stream);
int mm_source_set_fields_host(mm_source_t *src, int F, const double *fields);
int mm_source_destroy(mm_source_t *src);
int mm_source_info(const mm_source_t *src, int64_t info[8]);
const mm_index_t *mm_source_index(const mm_source_t *src);
int mm_source_interpolate(mm_source_t *src, int64_t N, const double *pts, int k,
                          const mm_locate_params *params, double *out, int32_t *elem, double *xi,
                          uint8_t *status, int64_t *num_failed, void *stream);
int mm_source_interpolate_host(mm_source_t *src, int64_t N, const double *pts, int k,
                               const mm_locate_params *params, double *values, int32_t *elem, double *xi,
                               int64_t *num_failed);

#ifdef __cplusplus
}
#endif
#endif /* MULTIMESH_B200_H */
