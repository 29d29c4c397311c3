// mm_newton.cuh -- order-n isoparametric map evaluation and Newton inversion (device),
// shared by K0 (affine pre-solve) and K2 (locate).  Arithmetic = DESIGN.md section 3.
#pragma once

#include "mm_common.cuh"

// x[c] = sum_a w_a Y_a[c],  J[c][s] = sum_a dw_a/dxi_s Y_a[c]; i innermost, then j, then k.
// Every accumulation of the contraction is ONE fused multiply-add, acc <- fma(w, v, acc) (explicit
// __fma_rn; the build keeps -fmad=false, so nothing else is contracted) -- the oracle uses C fma()
// at the same places, which is what keeps CPU and GPU bit-identical.
template <int ORDER, int DIM>
__device__ __forceinline__ void eval_map(const mm_gll_table &T, const double *__restrict__ Xn,
                                         const double (&p)[DIM], const double (&xi)[DIM],
                                         double (&x)[DIM], double (&J)[DIM][DIM])
{
    constexpr int M = ORDER + 1;
    double L[DIM][M], dL[DIM][M];
#pragma unroll
    for (int ax = 0; ax < DIM; ++ax) lagrange_values_derivs<ORDER>(T, xi[ax], L[ax], dL[ax]);

    if constexpr (DIM == 2) {
        double V[2] = {0, 0}, Dxi[2] = {0, 0}, Deta[2] = {0, 0};
#pragma unroll
        for (int j = 0; j < M; ++j) {
            double a[2] = {0, 0}, b[2] = {0, 0};
#pragma unroll
            for (int i = 0; i < M; ++i) {
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    double y = Xn[(i + M * j) * 2 + c] - p[c];
                    a[c] = __fma_rn(L[0][i], y, a[c]);
                    b[c] = __fma_rn(dL[0][i], y, b[c]);
                }
            }
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                V[c] = __fma_rn(L[1][j], a[c], V[c]);
                Deta[c] = __fma_rn(dL[1][j], a[c], Deta[c]);
                Dxi[c] = __fma_rn(L[1][j], b[c], Dxi[c]);
            }
        }
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            x[c] = V[c];
            J[c][0] = Dxi[c];
            J[c][1] = Deta[c];
        }
    } else {
        // the z-direction values are indexed by the k loop; for order 4 that loop stays rolled
        // (keeps the live register set and the code size down), so they live in their own arrays
        double Lz[M], dLz[M];
#pragma unroll
        for (int k = 0; k < M; ++k) {
            Lz[k] = L[2][k];
            dLz[k] = dL[2][k];
        }
        double X[3] = {0, 0, 0}, Jx[3] = {0, 0, 0}, Jy[3] = {0, 0, 0}, Jz[3] = {0, 0, 0};
#pragma unroll(M >= 3 ? 1 : M)
        for (int k = 0; k < M; ++k) {
            double V[3] = {0, 0, 0}, Dxi[3] = {0, 0, 0}, Deta[3] = {0, 0, 0};
            const double *Xk = Xn + (M * M * k) * 3;
#pragma unroll
            for (int j = 0; j < M; ++j) {
                double a[3] = {0, 0, 0}, b[3] = {0, 0, 0};
#pragma unroll
                for (int i = 0; i < M; ++i) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        double y = Xk[(i + M * j) * 3 + c] - p[c];
                        a[c] = __fma_rn(L[0][i], y, a[c]);
                        b[c] = __fma_rn(dL[0][i], y, b[c]);
                    }
                }
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    V[c] = __fma_rn(L[1][j], a[c], V[c]);
                    Deta[c] = __fma_rn(dL[1][j], a[c], Deta[c]);
                    Dxi[c] = __fma_rn(L[1][j], b[c], Dxi[c]);
                }
            }
            const double lz = Lz[k], dlz = dLz[k];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                X[c] = __fma_rn(lz, V[c], X[c]);
                Jz[c] = __fma_rn(dlz, V[c], Jz[c]);
                Jx[c] = __fma_rn(lz, Dxi[c], Jx[c]);
                Jy[c] = __fma_rn(lz, Deta[c], Jy[c]);
            }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            x[c] = X[c];
            J[c][0] = Jx[c];
            J[c][1] = Jy[c];
            J[c][2] = Jz[c];
        }
    }
}

// Newton on the point-shifted nodes Y = X - p; true when max|delta| <= 1e-13.
// Start: xi0 = Jinv0 ((p - ref) - x0) from the element's affine pre-solve
// `pre` = {ref[DIM], x0[DIM], Jinv[DIM][DIM]} (K0, mm_element_presolve; ref = first control node,
// x0 and Jinv evaluated on ref-shifted nodes) -- an exactly affine element then needs one
// evaluation instead of two -- or xi0 = 0 when `pre` is null or unusable.
template <int DIM>
__device__ __forceinline__ void newton_start(const double (&p)[DIM], const double *__restrict__ pre,
                                             double (&xi)[DIM])
{
#pragma unroll
    for (int c = 0; c < DIM; ++c) xi[c] = 0.0;
    if (pre) {
        double r[DIM], g[DIM];
        bool ok = true;
#pragma unroll
        for (int c = 0; c < DIM; ++c) r[c] = (p[c] - pre[c]) - pre[DIM + c];
#pragma unroll
        for (int s = 0; s < DIM; ++s) {
            const double *row = pre + 2 * DIM + s * DIM;
            double v = row[0] * r[0] + row[1] * r[1];
            if constexpr (DIM == 3) v = v + row[2] * r[2];
            g[s] = v;
            if (!(fabs(v) <= MM_NEWTON_DIVERGE)) ok = false;
        }
        if (ok) {
#pragma unroll
            for (int c = 0; c < DIM; ++c) xi[c] = g[c];
        }
    }
}

// Newton iterations from the start value already in xi (newton_start needs only the pre-solve row
// and the point, so K2 computes it while the element block is still in flight).
template <int ORDER, int DIM>
__device__ __forceinline__ bool newton_iterate(const mm_gll_table &T,
                                               const double *__restrict__ X,
                                               const double (&p)[DIM], double (&xi)[DIM],
                                               int *evaluations = nullptr)
{
    double dprev = INFINITY;  // max|delta| of the previous iteration
#pragma unroll 1
    for (int it = 0; it < MM_NEWTON_MAXIT; ++it) {
        if (evaluations) ++*evaluations;  // statistics build only (compile-time null otherwise)
        double x[DIM], J[DIM][DIM], delta[DIM];
        eval_map<ORDER, DIM>(T, X, p, xi, x, J);
        if constexpr (DIM == 2) {
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
        bool bad = false;
#pragma unroll
        for (int c = 0; c < DIM; ++c) {
            double ad = fabs(delta[c]);
            if (!(ad <= MM_NEWTON_DIVERGE)) bad = true;
            if (ad > dmax) dmax = ad;
            xi[c] = xi[c] + delta[c];
        }
        if (bad) return false;
        // converged: the update is below 1e-13 -- or it is below 1e-7 AND a thousand times smaller than the previous
        // one: Newton is then in its quadratic regime, the update just applied leaves an error of O(dmax^2) <= 1e-14,
        // and the evaluation that would only confirm it is not spent (curved elements: 2 evaluations instead of 3)
        if (dmax <= MM_NEWTON_TOL || (dmax <= MM_NEWTON_TOL_FAST && dmax <= MM_NEWTON_FAST_RATIO * dprev)) return true;
        dprev = dmax;
    }
    return false;
}


template <int ORDER, int DIM>
__device__ __forceinline__ bool newton_inverse(const mm_gll_table &T,
                                               const double *__restrict__ X,
                                               const double (&p)[DIM],
                                               const double *__restrict__ pre, double (&xi)[DIM])
{
    newton_start<DIM>(p, pre, xi);
    return newton_iterate<ORDER, DIM>(T, X, p, xi);
}
