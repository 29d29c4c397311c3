/*
 * oracle/mm_oracle.c -- CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of the mesh-to-mesh interpolation hot path of
 * solvithrastar/MultiMesh.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this library.
 * The product (multimesh_b200/) never links, imports or calls it.
 *
 * PARITY STATUS
 *   * order-1 HEX8 path (mmo_trilinear_interpolator, mmo_centroid): PINNED against
 *     the reference's own C sources compiled into oracle/_ref (see oracle/build.py
 *     and tests/test_oracle.py) -- bit-for-bit.
 *   * GLL order-n path (Newton inverse map, Lagrange weights): PARITY UNPINNED.
 *     The reference delegates this arithmetic to the closed-source salvus.fem
 *     module (multi_mesh/components/interpolator.py:12,22-57,1337-1347,1370-1386),
 *     which is not vendored, not pinned to a version and not installable here;
 *     the reference ships no tests or golden vectors.  What is restated below is
 *     the published mathematics (tensor-product GLL Lagrange basis, Newton on the
 *     order-n isoparametric map) anchored on the reference's call sites, with a
 *     documented canonical operation order so that CPU and GPU agree bit for bit.
 *   * location logic around that arithmetic (candidate loops V1-V5, fall-backs, layer
 *     masks, de-duplication, gathers, fix-ups): PINNED against the reference's own Python,
 *     run in the build container with salvus.fem's two functions served by this file
 *     (tests/golden/make_golden_glue.py -> tests/golden/glue_*.npz, checked by
 *     tests/test_golden_glue.py).
 *
 * CANONICAL ARITHMETIC (shared spec with the CUDA kernels)
 *   - all arithmetic IEEE-754 binary64; the accumulations of the two tensor contractions
 *     (eval_map: Newton's x and J; mmo_interp: the gather) are explicit fused multiply-adds,
 *     acc <- fma(w, v, acc), one rounding each (C fma() here, __fma_rn on the device); every
 *     other operation is a separate IEEE operation with its own rounding -- the compilers are
 *     kept from contracting anything themselves (-ffp-contract=off; nvcc -fmad=false);
 *   - node index a = i + m*j + m*m*k, m = order+1, i along xi (fastest);
 *   - sums are evaluated in the loop order written below, never re-associated.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define MMO_MAXM 5            /* order <= 4 */
#define MMO_NEWTON_MAXIT 50   /* iteration cap, as trilinearinterpolator.c:264 */
#define MMO_NEWTON_TOL 1e-13  /* on max|delta xi|; see DESIGN.md 3.3 */
#define MMO_NEWTON_TOL_FAST 1e-7    /* accepted when the iteration is visibly quadratic (newton_inverse) */
#define MMO_NEWTON_FAST_RATIO 1e-3
#define MMO_NEWTON_DIVERGE 1e10

/* ------------------------------------------------------------------------------------------
 * 1-D GLL nodes.  Orders fixed by the reference's template list
 * (interpolator.py:26-57: n = 1, 2, 4).
 * ---------------------------------------------------------------------------------------- */
static int gll_nodes(int order, double *z)
{
    switch (order) {
    case 1: z[0] = -1.0; z[1] = 1.0; return 2;
    case 2: z[0] = -1.0; z[1] = 0.0; z[2] = 1.0; return 3;
    case 4:
        z[0] = -1.0; z[1] = -0x1.4f2ec413cb52ap-1; z[2] = 0.0;
        z[3] = 0x1.4f2ec413cb52ap-1; z[4] = 1.0;   /* sqrt(3/7) */
        return 5;
    default: return 0;
    }
}

/* c_i = 1 / prod_{j != i, j ascending} (z_i - z_j) */
static void lagrange_denominators(int m, const double *z, double *c)
{
    for (int i = 0; i < m; ++i) {
        double prod = 1.0;
        for (int j = 0; j < m; ++j)
            if (j != i) prod = prod * (z[i] - z[j]);
        c[i] = 1.0 / prod;
    }
}

typedef struct {
    int m;
    double z[MMO_MAXM];
    double c[MMO_MAXM];
} basis_t;

static int basis_init(basis_t *b, int order)
{
    b->m = gll_nodes(order, b->z);
    if (!b->m) return 0;
    lagrange_denominators(b->m, b->z, b->c);
    return 1;
}

/* L_i(x) = c_i * prod_{j != i} (x - z_j), factors applied in ascending j. */
static void lagrange_values(const basis_t *b, double x, double *L)
{
    double d[MMO_MAXM];
    for (int j = 0; j < b->m; ++j) d[j] = x - b->z[j];
    for (int i = 0; i < b->m; ++i) {
        double prod = b->c[i];
        for (int j = 0; j < b->m; ++j)
            if (j != i) prod = prod * d[j];
        L[i] = prod;
    }
}

/* L_i'(x) = sum_{q != i, ascending} ( c_i * prod_{j != i, j != q, ascending} (x - z_j) ) */
static void lagrange_derivs(const basis_t *b, double x, double *dL)
{
    double d[MMO_MAXM];
    for (int j = 0; j < b->m; ++j) d[j] = x - b->z[j];
    for (int i = 0; i < b->m; ++i) {
        double sum = 0.0;
        for (int q = 0; q < b->m; ++q) {
            if (q == i) continue;
            double term = b->c[i];
            for (int j = 0; j < b->m; ++j)
                if (j != i && j != q) term = term * d[j];
            sum = sum + term;
        }
        dL[i] = sum;
    }
}

/* exported helpers (tests) */
int mmo_gll_nodes(int order, double *z) { return gll_nodes(order, z); }

int mmo_lagrange(int order, double x, double *L, double *dL)
{
    basis_t b;
    if (!basis_init(&b, order)) return -1;
    lagrange_values(&b, x, L);
    lagrange_derivs(&b, x, dL);
    return b.m;
}

/* ------------------------------------------------------------------------------------------
 * Tensor-product interpolation coefficients w_a(xi) = (L_i(xi) * L_j(eta)) * L_k(zeta).
 * Restates salvus.fem GetInterpolationCoefficients<n,n,n> / <4,4,0>
 * (interpolator.py:1337-1347); output length P = m^dim.
 * ---------------------------------------------------------------------------------------- */
static void weights_from_xi(const basis_t *b, int dim, const double *xi, double *w)
{
    double L[3][MMO_MAXM];
    int m = b->m;
    for (int ax = 0; ax < dim; ++ax) lagrange_values(b, xi[ax], L[ax]);
    if (dim == 2) {
        for (int j = 0; j < m; ++j)
            for (int i = 0; i < m; ++i) w[i + m * j] = L[0][i] * L[1][j];
    } else {
        for (int k = 0; k < m; ++k)
            for (int j = 0; j < m; ++j)
                for (int i = 0; i < m; ++i)
                    w[i + m * j + m * m * k] = (L[0][i] * L[1][j]) * L[2][k];
    }
}

int mmo_weights(int order, int dim, const double *xi, double *w)
{
    basis_t b;
    if (!basis_init(&b, order) || (dim != 2 && dim != 3)) return -1;
    weights_from_xi(&b, dim, xi, w);
    return 0;
}

/* coeffs[N,P]; zero rows where elem < 0 (interpolator.py:1233,1297,1583). */
int mmo_coeffs(int order, int dim, long long N, const int32_t *elem, const double *xi,
               double *coeffs)
{
    basis_t b;
    if (!basis_init(&b, order) || (dim != 2 && dim != 3)) return -1;
    int P = dim == 2 ? b.m * b.m : b.m * b.m * b.m;
#pragma omp parallel for schedule(static)
    for (long long n = 0; n < N; ++n) {
        double *w = coeffs + n * P;
        if (elem && elem[n] < 0) {
            for (int a = 0; a < P; ++a) w[a] = 0.0;
        } else {
            weights_from_xi(&b, dim, xi + n * dim, w);
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Isoparametric map evaluated on point-shifted control nodes Y_a = X_a - p.
 *   x[c]      = sum_a w_a Y_a[c]            ( = x(xi) - p )
 *   J[c][s]   = sum_a (dw_a/dxi_s) Y_a[c]
 * in the canonical nested (sum-factorised) order:  i innermost, then j, then k.
 * Generalises dNdR/dNdS/dNdT + dot_product_matrix_matrix (trilinearinterpolator.c:214-257)
 * to the order-n Lagrange basis over all P control nodes (interpolator.py:1373-1384).
 * ---------------------------------------------------------------------------------------- */
static void eval_map(const basis_t *b, int dim, const double *Y /*[P][dim]*/, const double *xi,
                     double *x /*[dim]*/, double J[3][3])
{
    int m = b->m;
    double L[3][MMO_MAXM], dL[3][MMO_MAXM];
    for (int ax = 0; ax < dim; ++ax) {
        lagrange_values(b, xi[ax], L[ax]);
        lagrange_derivs(b, xi[ax], dL[ax]);
    }
    if (dim == 2) {
        double V[2] = {0, 0}, Dxi[2] = {0, 0}, Deta[2] = {0, 0};
        for (int j = 0; j < m; ++j) {
            double a[2] = {0, 0}, bb[2] = {0, 0};
            for (int i = 0; i < m; ++i) {
                const double *y = Y + (size_t)(i + m * j) * 2;
                for (int c = 0; c < 2; ++c) {
                    a[c] = fma(L[0][i], y[c], a[c]);
                    bb[c] = fma(dL[0][i], y[c], bb[c]);
                }
            }
            for (int c = 0; c < 2; ++c) {
                V[c] = fma(L[1][j], a[c], V[c]);
                Deta[c] = fma(dL[1][j], a[c], Deta[c]);
                Dxi[c] = fma(L[1][j], bb[c], Dxi[c]);
            }
        }
        for (int c = 0; c < 2; ++c) {
            x[c] = V[c];
            J[c][0] = Dxi[c];
            J[c][1] = Deta[c];
        }
        return;
    }
    double X[3] = {0, 0, 0}, Jx[3] = {0, 0, 0}, Jy[3] = {0, 0, 0}, Jz[3] = {0, 0, 0};
    for (int k = 0; k < m; ++k) {
        double V[3] = {0, 0, 0}, Dxi[3] = {0, 0, 0}, Deta[3] = {0, 0, 0};
        for (int j = 0; j < m; ++j) {
            double a[3] = {0, 0, 0}, bb[3] = {0, 0, 0};
            for (int i = 0; i < m; ++i) {
                const double *y = Y + (size_t)(i + m * j + m * m * k) * 3;
                for (int c = 0; c < 3; ++c) {
                    a[c] = fma(L[0][i], y[c], a[c]);
                    bb[c] = fma(dL[0][i], y[c], bb[c]);
                }
            }
            for (int c = 0; c < 3; ++c) {
                V[c] = fma(L[1][j], a[c], V[c]);
                Deta[c] = fma(dL[1][j], a[c], Deta[c]);
                Dxi[c] = fma(L[1][j], bb[c], Dxi[c]);
            }
        }
        for (int c = 0; c < 3; ++c) {
            X[c] = fma(L[2][k], V[c], X[c]);
            Jz[c] = fma(dL[2][k], V[c], Jz[c]);
            Jx[c] = fma(L[2][k], Dxi[c], Jx[c]);
            Jy[c] = fma(L[2][k], Deta[c], Jy[c]);
        }
    }
    for (int c = 0; c < 3; ++c) {
        x[c] = X[c];
        J[c][0] = Jx[c];
        J[c][1] = Jy[c];
        J[c][2] = Jz[c];
    }
}

/* Forward map x(xi) on unshifted nodes (tests / mesh generators). */
int mmo_forward_map(int order, int dim, const double *nodes, const double *xi, double *x)
{
    basis_t b;
    double J[3][3];
    if (!basis_init(&b, order) || (dim != 2 && dim != 3)) return -1;
    eval_map(&b, dim, nodes, xi, x, J);
    return 0;
}

/*
 * Affine pre-solve per element (p-independent, computed once per source mesh, K0):
 *   pre[e] = { ref[dim], x0[dim], Jinv[dim][dim] }
 *   ref  = the element's first control node (an exact reference point),
 *   x0   = x(xi = 0) - ref and Jinv = inverse of J(xi = 0), both evaluated on the nodes shifted
 *          by ref with the same nested order (so roundoff scales with the element size, not with
 *          the coordinate magnitude); Jinv[s][c] = cofactor / det.
 * Newton then starts at xi0 = Jinv ((p - ref) - x0).
 */
void mmo_presolve(int order, int dim, long long E, const double *nodes, double *pre)
{
    basis_t b;
    if (!basis_init(&b, order)) return;
    int P = dim == 2 ? b.m * b.m : b.m * b.m * b.m;
    int W = 2 * dim + dim * dim;
#pragma omp parallel for schedule(static)
    for (long long e = 0; e < E; ++e) {
        double xi0[3] = {0, 0, 0}, x[3], J[3][3];
        double Y[MMO_MAXM * MMO_MAXM * MMO_MAXM * 3];
        const double *X = nodes + (size_t)e * P * dim;
        for (int a = 0; a < P; ++a)
            for (int c = 0; c < dim; ++c) Y[a * dim + c] = X[a * dim + c] - X[c];
        eval_map(&b, dim, Y, xi0, x, J);
        double *o = pre + (size_t)e * W;
        for (int c = 0; c < dim; ++c) {
            o[c] = X[c];
            o[dim + c] = x[c];
        }
        double *I = o + 2 * dim;
        if (dim == 2) {
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
            /* delta_s = sum_c Jinv[s][c] r_c,  Jinv[s][c] = C[c][s] / det */
            I[0] = C00 / det; I[1] = C10 / det; I[2] = C20 / det;
            I[3] = C01 / det; I[4] = C11 / det; I[5] = C21 / det;
            I[6] = C02 / det; I[7] = C12 / det; I[8] = C22 / det;
        }
    }
}

/*
 * Newton inverse of the isoparametric map.  Structure follows
 * inverseCoordinateTransform (trilinearinterpolator.c:260-305): start at xi = 0,
 * <= 50 iterations, update = J^-1 (p - x(xi)) through an explicit cofactor inverse
 * (trilinearinterpolator.c:329-341), non-convergence => reject (the Python drivers'
 * "NaN" branch, interpolator.py:1200,1286,1436).  Differences, all deliberate:
 *   - the map is the order-n Lagrange map over all P control nodes;
 *   - nodes are shifted by p first (Y = X - p) so roundoff scales with the element,
 *     not with |x| ~ 6.4e6 m on global meshes;
 *   - convergence test is on the update, max|delta| <= 1e-13 (the C twin's
 *     1e-8*scale residual test cannot deliver 1e-12 on xi and checks component 0
 *     twice, trilinearinterpolator.c:290-291 -- not replicated here).
 * Returns 1 if converged (xi valid), 0 otherwise.
 */
static int newton_inverse(const basis_t *b, int dim, const double *nodes, const double *p,
                          const double *pre, double *xi, int *iters)
{
    double Y[MMO_MAXM * MMO_MAXM * MMO_MAXM * 3];
    int m = b->m;
    int P = dim == 2 ? m * m : m * m * m;
    for (int a = 0; a < P; ++a)
        for (int c = 0; c < dim; ++c) Y[a * dim + c] = nodes[a * dim + c] - p[c];
    for (int c = 0; c < dim; ++c) xi[c] = 0.0;
    if (pre) {
        /* affine pre-solve (mmo_presolve): xi0 = Jinv0 (p - x0), the first Newton step from the
           element centre with p-independent quantities; an exactly affine element then needs one
           evaluation instead of two.  Falls back to xi0 = 0 when the pre-solve is unusable. */
        double r[3], g[3];
        int ok = 1;
        for (int c = 0; c < dim; ++c) r[c] = (p[c] - pre[c]) - pre[dim + c];
        for (int s = 0; s < dim; ++s) {
            const double *row = pre + 2 * dim + s * dim;
            double v = row[0] * r[0] + row[1] * r[1];
            if (dim == 3) v = v + row[2] * r[2];
            g[s] = v;
            if (!(fabs(v) <= MMO_NEWTON_DIVERGE)) ok = 0;
        }
        if (ok)
            for (int c = 0; c < dim; ++c) xi[c] = g[c];
    }
    double dprev = INFINITY; /* max|delta| of the previous iteration */
    for (int it = 0; it < MMO_NEWTON_MAXIT; ++it) {
        double x[3], J[3][3], delta[3];
        eval_map(b, dim, Y, xi, x, J);
        if (dim == 2) {
            double r0 = -x[0], r1 = -x[1];
            double det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
            delta[0] = (J[1][1] * r0 - J[0][1] * r1) / det;
            delta[1] = (J[0][0] * r1 - J[1][0] * r0) / det;
        } else {
            double r0 = -x[0], r1 = -x[1], r2 = -x[2];
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
            delta[0] = ((C00 * r0 + C10 * r1) + C20 * r2) / det;
            delta[1] = ((C01 * r0 + C11 * r1) + C21 * r2) / det;
            delta[2] = ((C02 * r0 + C12 * r1) + C22 * r2) / det;
        }
        double dmax = 0.0;
        int bad = 0;
        for (int c = 0; c < dim; ++c) {
            double ad = fabs(delta[c]);
            if (!(ad <= MMO_NEWTON_DIVERGE)) bad = 1; /* NaN, inf or runaway */
            if (ad > dmax) dmax = ad;
            xi[c] = xi[c] + delta[c];
        }
        if (iters) *iters = it + 1;
        if (bad) return 0;
        /* converged: update below 1e-13, or below 1e-7 and a thousand times smaller than the previous one (quadratic
           regime: the update just applied leaves an error of O(dmax^2) <= 1e-14; the confirming evaluation is saved) */
        if (dmax <= MMO_NEWTON_TOL || (dmax <= MMO_NEWTON_TOL_FAST && dmax <= MMO_NEWTON_FAST_RATIO * dprev)) return 1;
        dprev = dmax;
    }
    return 0;
}

int mmo_inverse_map(int order, int dim, const double *nodes, const double *p, double *xi,
                    int *iters)
{
    basis_t b;
    if (!basis_init(&b, order) || (dim != 2 && dim != 3)) return -1;
    return newton_inverse(&b, dim, nodes, p, NULL, xi, iters);
}

int mmo_inverse_map_pre(int order, int dim, const double *nodes, const double *p, const double *pre,
                        double *xi, int *iters)
{
    basis_t b;
    if (!basis_init(&b, order) || (dim != 2 && dim != 3)) return -1;
    return newton_inverse(&b, dim, nodes, p, pre, xi, iters);
}

/* ------------------------------------------------------------------------------------------
 * Element centroids and AABBs.
 *   centroid = (sequential sum over nodes a = 0..P-1) / P  -- bit-equal to
 *   np.mean(points, axis=1) (salvus_mesh_reader.py:99-100) and to centroid.c:15-24.
 *   AABB = min/max over all P nodes (boundary_box_check, interpolator.py:1360).
 * ---------------------------------------------------------------------------------------- */
void mmo_centroids(long long E, int P, int dim, const double *nodes, double *cent)
{
#pragma omp parallel for schedule(static)
    for (long long e = 0; e < E; ++e)
        for (int c = 0; c < dim; ++c) {
            double s = 0.0;
            for (int a = 0; a < P; ++a) s = s + nodes[((size_t)e * P + a) * dim + c];
            cent[e * dim + c] = s / (double)P;
        }
}

void mmo_aabb(long long E, int P, int dim, const double *nodes, double *box /*[E][2][dim]*/)
{
#pragma omp parallel for schedule(static)
    for (long long e = 0; e < E; ++e)
        for (int c = 0; c < dim; ++c) {
            double lo = nodes[((size_t)e * P) * dim + c], hi = lo;
            for (int a = 1; a < P; ++a) {
                double v = nodes[((size_t)e * P + a) * dim + c];
                if (v < lo) lo = v;
                if (v > hi) hi = v;
            }
            box[(e * 2 + 0) * dim + c] = lo;
            box[(e * 2 + 1) * dim + c] = hi;
        }
}

/* Connectivity-gathered centroid; restates centroid.c:3-25 (same signature). */
void mmo_centroid(long long ndim, long long nelem, long long npe, const long long *conn,
                  const double *points, double *cent)
{
#pragma omp parallel for schedule(static)
    for (long long e = 0; e < nelem; ++e)
        for (long long c = 0; c < ndim; ++c) {
            double s = 0.;
            for (long long a = 0; a < npe; ++a) s = s + points[conn[e * npe + a] * ndim + c];
            cent[e * ndim + c] = s / npe;
        }
}

/* ------------------------------------------------------------------------------------------
 * Canonical exact k-NN (brute force).  Replaces KDTree(data).query(pts, k)
 * (pykdtree; interpolator.py:101-105,363-373,751-753,949-951,1176-1178) with the
 * documented total order (d2, index):  d2 = (dx*dx + dy*dy) + dz*dz.
 * idx is padded with -1 when k > M.
 * ---------------------------------------------------------------------------------------- */
static inline double dist2(int dim, const double *a, const double *b)
{
    double dx = a[0] - b[0], dy = a[1] - b[1];
    double s = dx * dx + dy * dy;
    if (dim == 3) {
        double dz = a[2] - b[2];
        s = s + dz * dz;
    }
    return s;
}

void mmo_knn_bruteforce(long long M, int dim, const double *data, long long N, const double *pts,
                        int k, int32_t *idx_out, double *d2_out /* may be NULL */)
{
#pragma omp parallel
    {
        double *bd = (double *)malloc(sizeof(double) * (size_t)k);
        int32_t *bi = (int32_t *)malloc(sizeof(int32_t) * (size_t)k);
#pragma omp for schedule(dynamic, 64)
        for (long long n = 0; n < N; ++n) {
            int cnt = 0;
            const double *q = pts + n * dim;
            for (long long j = 0; j < M; ++j) {
                double d2 = dist2(dim, q, data + j * dim);
                if (cnt == k && !(d2 < bd[k - 1])) continue; /* idx ascending => ties lose */
                int pos = cnt < k ? cnt : k - 1;
                while (pos > 0 && (d2 < bd[pos - 1])) { /* equal d2: earlier index stays first */
                    bd[pos] = bd[pos - 1];
                    bi[pos] = bi[pos - 1];
                    --pos;
                }
                bd[pos] = d2;
                bi[pos] = (int32_t)j;
                if (cnt < k) ++cnt;
            }
            for (int t = 0; t < k; ++t) {
                idx_out[n * k + t] = t < cnt ? bi[t] : -1;
                if (d2_out) d2_out[n * k + t] = t < cnt ? bd[t] : INFINITY;
            }
        }
        free(bd);
        free(bi);
    }
}

/* ------------------------------------------------------------------------------------------
 * Point-in-element location: candidate iteration, accept test, fallback.
 * One parameterised routine for the reference's variants:
 *   V1 _check_if_inside_element            interpolator.py:1409-1473
 *   V2 get_element_weights.check_inside    interpolator.py:1181-1233
 *   V3 get_element_weights_layered         interpolator.py:1271-1297
 *   V4 v2_interpolation_tools              v2_interpolation_tools.py:71-164
 *   V5 cli._check_if_inside_element        scripts/cli.py:401-430
 * ---------------------------------------------------------------------------------------- */
enum { MMO_FB_FAIL = 0, MMO_FB_MAGIC = 1, MMO_FB_SNAP = 2, MMO_FB_MINL1 = 3 };
enum {
    MMO_ST_ACCEPTED = 0,      /* accepted inside the candidate loop */
    MMO_ST_FB_INSIDE_MAGIC = 1, /* V1: first AABB-containing candidate, magic xi */
    MMO_ST_FB_NEAR_OK = 2,    /* V1: nearest-centre candidate, |xi| < 1.04 */
    MMO_ST_FB_NEAR_MAGIC = 3, /* V1: nearest-centre candidate, some |xi| >= 1.04 -> magic */
    MMO_ST_FB_NAN_MAGIC = 4,  /* V1: nearest-centre candidate did not converge -> magic
                                 (ValueError unless ignore_hard_elements, :1465-1468) */
    MMO_ST_SNAPPED = 5,       /* V2 snap_to_nearest */
    MMO_ST_FAILED = 6,        /* elem = -1, zero weights */
    MMO_ST_MINL1 = 7,         /* V5 fallback */
    MMO_ST_SNAP_NONE = 8      /* V2 snap with no convergent candidate: elem 0, xi = +clip */
};

typedef struct {
    int32_t aabb_prefilter; /* V1: test only candidates whose node AABB contains the point */
    int32_t strict;         /* 1: all |xi| <  tol ; 0: all |xi| <= tol */
    int32_t fallback;       /* MMO_FB_* */
    int32_t reserved;
    double tol;             /* 1.04 (V1) 1.05 (V2,V4) 1.03 (V3) 1.02 (V5) */
    double snap_clip;       /* 1.02 (interpolator.py:1219) */
    double magic_xi[3];     /* 0.645, -0.5, 0.22 (interpolator.py:1468-1471) */
} mmo_locate_params;

static int accept_xi(const mmo_locate_params *prm, int dim, const double *xi)
{
    for (int c = 0; c < dim; ++c) {
        double a = fabs(xi[c]);
        if (prm->strict ? !(a < prm->tol) : !(a <= prm->tol)) return 0;
    }
    return 1;
}

/*
 * nodes  [E][P][dim]     source element control nodes
 * cent   [E][dim]        canonical centroids (mmo_centroids)       -- V1 fallback distance
 * box    [E][2][dim]     AABBs (mmo_aabb)                          -- V1 prefilter
 * cands  [N][k] int32    candidate element ids in k-NN order; negative entries are skipped;
 *                        repeated ids are tested once (first occurrence) -- equivalent to the
 *                        reference, which would recompute the same result.
 * Returns the number of points with elem = -1.
 */
long long mmo_locate(int order, int dim, long long E, const double *nodes, const double *cent,
                     const double *box, const double *pre /* may be NULL */, long long N,
                     const double *pts, int k,
                     const int32_t *cands, const mmo_locate_params *prm, int32_t *elem_out,
                     double *xi_out, uint8_t *status_out)
{
    basis_t b;
    if (!basis_init(&b, order) || (dim != 2 && dim != 3)) return -1;
    int P = dim == 2 ? b.m * b.m : b.m * b.m * b.m;
    int W = 2 * dim + dim * dim;
    long long nfailed = 0;
    (void)E;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : nfailed)
    for (long long n = 0; n < N; ++n) {
        const double *p = pts + n * dim;
        const int32_t *cl = cands + n * (long long)k;
        double xi[3] = {0, 0, 0};
        int32_t elem = -1;
        uint8_t status = MMO_ST_FAILED;
        int accepted = 0;
        /* fallback trackers */
        int32_t first_inside = -1;            /* V1 */
        int first_inside_nan = 0;             /* V1: Newton failed on it (:1460-1467 re-inverts and raises) */
        int32_t near_elem = -1;               /* V1: min distance to centroid, first occurrence */
        double near_dist = INFINITY;
        int32_t best_elem = -1;               /* V2: smallest max|xi|; V5: smallest sum|xi| */
        double best_key = prm->fallback == MMO_FB_SNAP ? 10e9 : INFINITY;
        double best_xi[3] = {0, 0, 0};

        for (int t = 0; t < k && !accepted; ++t) {
            int32_t e = cl[t];
            if (e < 0) continue;
            int dup = 0;
            for (int u = 0; u < t; ++u)
                if (cl[u] == e) { dup = 1; break; }
            if (dup) continue;
            if (prm->aabb_prefilter) {
                const double *lo = box + ((size_t)e * 2 + 0) * dim;
                const double *hi = box + ((size_t)e * 2 + 1) * dim;
                int inside = 1;
                for (int c = 0; c < dim; ++c)
                    if (!(p[c] >= lo[c] && p[c] <= hi[c])) inside = 0;
                if (!inside) {
                    /* dist = || p - centre ||  (boundary_box_check, :1365-1366) */
                    double d = sqrt(dist2(dim, p, cent + (size_t)e * dim));
                    if (d < near_dist) { near_dist = d; near_elem = e; }
                    continue;
                }
                if (first_inside < 0) first_inside = e;
            }
            double x[3];
            int ok = newton_inverse(&b, dim, nodes + (size_t)e * P * dim, p,
                                    pre ? pre + (size_t)e * W : NULL, x, NULL);
            if (!ok) { /* "NaN" branch */
                if (e == first_inside) first_inside_nan = 1;
                continue;
            }
            if (prm->fallback == MMO_FB_SNAP || prm->fallback == MMO_FB_MINL1) {
                double key = 0.0;
                for (int c = 0; c < dim; ++c) {
                    double a = fabs(x[c]);
                    if (prm->fallback == MMO_FB_SNAP) { if (a > key) key = a; }
                    else key = key + a;
                }
                if (key < best_key) {
                    best_key = key; best_elem = e;
                    for (int c = 0; c < dim; ++c) best_xi[c] = x[c];
                }
            }
            if (accept_xi(prm, dim, x)) {
                accepted = 1; elem = e; status = MMO_ST_ACCEPTED;
                for (int c = 0; c < dim; ++c) xi[c] = x[c];
            }
        }
        if (!accepted) {
            switch (prm->fallback) {
            case MMO_FB_MAGIC: /* interpolator.py:1448-1473 */
                if (first_inside >= 0) {
                    /* the re-inversion (:1460) repeats a result already rejected above: either NaN -- the
                     * reference raises unless ignore_hard_elements -- or some |xi| > 1.04; magic xi either way */
                    elem = first_inside;
                    status = first_inside_nan ? MMO_ST_FB_NAN_MAGIC : MMO_ST_FB_INSIDE_MAGIC;
                    for (int c = 0; c < dim; ++c) xi[c] = prm->magic_xi[c];
                } else if (near_elem >= 0) {
                    double x[3];
                    elem = near_elem;
                    int ok = newton_inverse(&b, dim, nodes + (size_t)elem * P * dim, p,
                                            pre ? pre + (size_t)elem * W : NULL, x, NULL);
                    int big = 0;
                    if (ok)
                        for (int c = 0; c < dim; ++c)
                            if (fabs(x[c]) >= prm->tol) big = 1;
                    if (!ok) status = MMO_ST_FB_NAN_MAGIC;
                    else if (big) status = MMO_ST_FB_NEAR_MAGIC;
                    else status = MMO_ST_FB_NEAR_OK;
                    for (int c = 0; c < dim; ++c)
                        xi[c] = (status == MMO_ST_FB_NEAR_OK) ? x[c] : prm->magic_xi[c];
                }
                break;
            case MMO_FB_SNAP: /* interpolator.py:1217-1230 */
                if (best_elem >= 0) {
                    elem = best_elem; status = MMO_ST_SNAPPED;
                    for (int c = 0; c < dim; ++c) {
                        double v = best_xi[c];
                        if (v < -prm->snap_clip) v = -prm->snap_clip;
                        if (v > prm->snap_clip) v = prm->snap_clip;
                        xi[c] = v;
                    }
                } else {
                    elem = 0; status = MMO_ST_SNAP_NONE;
                    for (int c = 0; c < dim; ++c) xi[c] = prm->snap_clip;
                }
                break;
            case MMO_FB_MINL1: /* scripts/cli.py:424-428 */
                if (best_elem >= 0) {
                    elem = best_elem; status = MMO_ST_MINL1;
                    for (int c = 0; c < dim; ++c) xi[c] = best_xi[c];
                }
                break;
            default: break;
            }
        }
        if (elem < 0) nfailed += 1;
        elem_out[n] = elem;
        for (int c = 0; c < dim; ++c) xi_out[n * dim + c] = elem < 0 ? 0.0 : xi[c];
        if (status_out) status_out[n] = status;
    }
    return nfailed;
}

/* ------------------------------------------------------------------------------------------
 * Gather-and-weight:  out[n,f] = sum_a w_a(xi_n) * fields[elem_n, f, a]
 * (interpolator.py:136-138,814-826,974-976; layout MODEL/data [E,F,P]).
 * Canonical order = nested tensor contraction, i innermost:
 *   t[j,k] = sum_i Lx[i] v[i,j,k];  u[k] = sum_j Ly[j] t[j,k];  out = sum_k Lz[k] u[k]
 * (differs from numpy's pairwise sum over explicit coeffs by O(1e-16) relative).
 * Rows with elem < 0 are zero (interpolator.py:963-976).
 * ---------------------------------------------------------------------------------------- */
int mmo_interp(int order, int dim, long long E, int F, const double *fields, long long N,
               const int32_t *elem, const double *xi, double *out)
{
    basis_t b;
    if (!basis_init(&b, order) || (dim != 2 && dim != 3)) return -1;
    int m = b.m;
    int P = dim == 2 ? m * m : m * m * m;
    (void)E;
#pragma omp parallel for schedule(static)
    for (long long n = 0; n < N; ++n) {
        double *o = out + n * F;
        if (elem[n] < 0) {
            for (int f = 0; f < F; ++f) o[f] = 0.0;
            continue;
        }
        double L[3][MMO_MAXM];
        for (int ax = 0; ax < dim; ++ax) lagrange_values(&b, xi[n * dim + ax], L[ax]);
        const double *blk = fields + (size_t)elem[n] * F * P;
        for (int f = 0; f < F; ++f) {
            const double *v = blk + (size_t)f * P;
            double acc = 0.0;
            int nk = dim == 3 ? m : 1;
            for (int k = 0; k < nk; ++k) {
                double u = 0.0;
                for (int j = 0; j < m; ++j) {
                    double t = 0.0;
                    for (int i = 0; i < m; ++i) t = fma(L[0][i], v[i + m * j + m * m * k], t);
                    u = fma(L[1][j], t, u);
                }
                if (dim == 3) acc = fma(L[2][k], u, acc);
                else acc = u;
            }
            o[f] = acc;
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Order-1 HEX8 "C-compat" path: restates triLinearInterpolator and its helpers
 * (trilinearinterpolator.c:40-305) with the same expression trees so results are
 * bit-identical to the compiled reference (oracle/_ref), including the convergence
 * test that looks at residual component 0 twice (trilinearinterpolator.c:290-291).
 * Vertex order R,S,T signs: trilinearinterpolator.c:8-10.
 * ---------------------------------------------------------------------------------------- */
static const double SR[8] = {-1, -1, +1, +1, -1, +1, +1, -1};
static const double SS[8] = {-1, +1, +1, -1, -1, -1, +1, +1};
static const double ST[8] = {-1, -1, -1, -1, +1, +1, +1, +1};

/* trilinear forward map of one coordinate, expression tree of referenceToElementMapping (:199-212) */
static double hex8_map1(const double v[8], double r, double s, double t)
{
    double hr = 0.5 * (r + 1.0), hs = 0.5 * (s + 1.0), ht = 0.5 * (t + 1.0);
    double e03 = hr * (-v[0] + v[3]);
    double e12 = hr * (-v[1] + v[2]);
    double e45 = hr * (-v[4] + v[5]);
    double e76 = hr * (v[6] - v[7]);
    double bot = -v[0] + v[1] - e03 + e12; /* ((-v0 + v1) - e03) + e12 */
    double top = -v[4] + v[7] - e45 + e76;
    return v[0] + e03 + hs * bot + ht * (-v[0] + v[4] - e03 + e45 - hs * bot + hs * top);
}

/* weights at (r,s,t); expression tree of interpolateAtPoint (:174-197) */
static void hex8_weights(const double q[3], double w[8])
{
    double r = q[0], s = q[1], t = q[2];
    double rst = 0.125 * r * s * t, rs = 0.125 * r * s, rt = 0.125 * r * t, st = 0.125 * s * t;
    double r8 = 0.125 * r, s8 = 0.125 * s, t8 = 0.125 * t;
    /* signs per vertex: (rst, rs, rt, r, st, s, t) */
    w[0] = -rst + rs + rt - r8 + st - s8 - t8 + 0.125;
    w[1] = +rst - rs + rt - r8 - st + s8 - t8 + 0.125;
    w[2] = -rst + rs - rt + r8 - st + s8 - t8 + 0.125;
    w[3] = +rst - rs - rt + r8 + st - s8 - t8 + 0.125;
    w[4] = +rst + rs - rt - r8 - st - s8 + t8 + 0.125;
    w[5] = -rst - rs + rt + r8 - st - s8 + t8 + 0.125;
    w[6] = +rst + rs + rt + r8 + st + s8 + t8 + 0.125;
    w[7] = -rst - rs - rt - r8 + st + s8 + t8 + 0.125;
}

static int hex8_inverse(const double pnt[3], double vtx[8][3], double sol[3])
{
    sol[0] = sol[1] = sol[2] = 0;
    double ax = fabs(vtx[1][0] - vtx[0][0]), ay = fabs(vtx[1][1] - vtx[0][1]);
    double az = fabs(vtx[1][2] - vtx[0][2]);
    double scalexy = ax > ay ? ax : ay;
    double scale = az > scalexy ? az : scalexy;
    double tol = 1e-8 * scale;
    for (int it = 0; it < 50; ++it) {
        double obj[3];
        for (int c = 0; c < 3; ++c) {
            double v[8];
            for (int a = 0; a < 8; ++a) v[a] = vtx[a][c];
            obj[c] = pnt[c] - hex8_map1(v, sol[0], sol[1], sol[2]);
        }
        if (fabs(obj[0]) < tol && fabs(obj[1]) < tol && fabs(obj[0]) < tol) return 1;
        /* jac[q][j] = sum_a dN_a/dq * vtx[a][j]  (:230-257, :343-360) */
        double jac[3][3];
        for (int q = 0; q < 3; ++q)
            for (int j = 0; j < 3; ++j) {
                double sum = 0;
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
        /* update = inv^T obj  (:298-299, :363-375) */
        for (int i = 0; i < 3; ++i) {
            double sum = 0;
            for (int j = 0; j < 3; ++j) sum = sum + inv[j][i] * obj[j];
            sol[i] = sol[i] + sum;
        }
    }
    return 0;
}

static int hex8_check_hull(const double pnt[3], double vtx[8][3], double sol[3])
{
    if (!hex8_inverse(pnt, vtx, sol)) return 0;
    for (int c = 0; c < 3; ++c)
        if (fabs(sol[c]) > (1 + 1.0)) return 0;
    return 1;
}

long long mmo_trilinear_interpolator(long long nelem_to_search, long long npoints,
                                     const long long *nearest, const long long *conn,
                                     long long *enclosing, const double *nodes, double *weights,
                                     const double *points)
{
    long long nfailed = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : nfailed)
    for (long long i = 0; i < npoints; ++i) {
        double pnt[3] = {points[i * 3], points[i * 3 + 1], points[i * 3 + 2]};
        double vtx[8][3], sol[3], w[8];
        double smallest = 99999999.9;
        long long best = -1;
        for (long long j = 0; j < nelem_to_search; ++j) {
            long long e = nearest[i * nelem_to_search + j];
            int done = 0;
            /* e < 0: -1 padding when the k-NN list is longer than the mesh (no such case in the reference, whose
             * KD-tree would hand out an out-of-range id): the candidate is skipped, the end-of-list logic below
             * still runs */
            if (e >= 0)
                for (int a = 0; a < 8; ++a)
                    for (int c = 0; c < 3; ++c) vtx[a][c] = nodes[conn[e * 8 + a] * 3 + c];
            if (e >= 0 && hex8_check_hull(pnt, vtx, sol)) {
                double maxerr = 0.0;
                for (int c = 0; c < 3; ++c)
                    if (fabs(sol[c]) > maxerr) maxerr = fabs(sol[c]);
                if (maxerr < (1 + 0.025)) {
                    hex8_weights(sol, w);
                    for (int a = 0; a < 8; ++a) {
                        weights[i * 8 + a] = w[a];
                        enclosing[i * 8 + a] = conn[e * 8 + a];
                    }
                    done = 1;
                } else if (maxerr < smallest) {
                    smallest = maxerr;
                    best = e;
                }
            }
            if (done) break;
            if (j == nelem_to_search - 1) {
                int ok = 0;
                if (smallest < 1.5 && best >= 0) {
                    for (int a = 0; a < 8; ++a)
                        for (int c = 0; c < 3; ++c) vtx[a][c] = nodes[conn[best * 8 + a] * 3 + c];
                    if (hex8_check_hull(pnt, vtx, sol)) {
                        hex8_weights(sol, w);
                        for (int a = 0; a < 8; ++a) {
                            weights[i * 8 + a] = w[a];
                            enclosing[i * 8 + a] = conn[best * 8 + a];
                        }
                        ok = 1;
                    }
                }
                if (!ok) nfailed += 1;
            }
        }
    }
    return nfailed;
}

/* hex8 helpers exported for unit tests */
void mmo_hex8_weights(const double q[3], double w[8]) { hex8_weights(q, w); }
int mmo_hex8_inverse(const double pnt[3], const double *vtx_flat, double sol[3])
{
    double vtx[8][3];
    memcpy(vtx, vtx_flat, sizeof vtx);
    return hex8_inverse(pnt, vtx, sol);
}

void mmo_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int mmo_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
