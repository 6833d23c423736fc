// ggp_linalg.cuh — small dense algebra of the backward pass, the forward/backward combination and the
// 8-dim joints, in registers: N x N inverse by partially pivoted LU, products, Gaussian product and
// division by the stationary prior.
//
// Replaces, from the reference (paths under src/):
//   multiply_gaussian()   predictions.h:183-188
//   divide_by_prior()     predictions.h:446-463
//   reverse_mean/cov()    predictions.h:278-301
//   every Eigen MatrixXd::inverse() / product on 4x4 and 8x8 dynamic matrices
//
// The reference lets Eigen 3.3 pick the algorithm; what Eigen does for these sizes is part of the
// rounding, so it is restated here: dynamic-size inverse = unblocked partial-pivot LU followed by a unit
// lower and an upper column-oriented triangular solve of P*I (diagonal applied as a multiplication by
// 1/u_ii); dynamic products = coefficient-wise with the inner sum left to right; 4x4 matrix * vector =
// column-major GEMV kernel, four columns at a time.  All loops are compile-time unrolled and row swaps
// are predicated element swaps, so the matrices never leave registers on the device.
// Host+device; compile with FMA contraction off.
#pragma once
#include "ggp_libm.cuh"

template <int N>
GGP_HD void ggp_cswap_rows(double* __restrict__ A, int k, int row) {
    // swap row k (compile-time after unrolling) with row `row` (run-time, > k) without dynamic indexing
#pragma unroll
    for (int i = 0; i < N; ++i) {
        if (i > k) {
            const bool sw = (row == i);
#pragma unroll
            for (int j = 0; j < N; ++j) {
                double a = A[k * N + j], b = A[i * N + j];
                A[k * N + j] = sw ? b : a;
                A[i * N + j] = sw ? a : b;
            }
        }
    }
}

// R = A^-1 (row-major).  Returns false if a zero pivot was met (result then holds inf/nan like Eigen's).
template <int N>
GGP_HD bool ggp_inv_lu(const double* __restrict__ A, double* __restrict__ R) {
    double lu[N * N];
#pragma unroll
    for (int i = 0; i < N * N; ++i) lu[i] = A[i];
    int tr[N];
    bool ok = true;
#pragma unroll
    for (int k = 0; k < N; ++k) {
        int row = k;
        double big = fabs(lu[k * N + k]);
#pragma unroll
        for (int i = k + 1; i < N; ++i) {
            double v = fabs(lu[i * N + k]);
            if (v > big) { big = v; row = i; }
        }
        tr[k] = row;
        if (big != 0.0) {
            ggp_cswap_rows<N>(lu, k, row);
#pragma unroll
            for (int i = k + 1; i < N; ++i) lu[i * N + k] = lu[i * N + k] / lu[k * N + k];
        } else {
            ok = false;
        }
#pragma unroll
        for (int i = k + 1; i < N; ++i)
#pragma unroll
            for (int j = k + 1; j < N; ++j) lu[i * N + j] = lu[i * N + j] - lu[i * N + k] * lu[k * N + j];
    }
#pragma unroll
    for (int i = 0; i < N * N; ++i) R[i] = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) R[i * N + i] = 1.0;
#pragma unroll
    for (int k = 0; k < N; ++k) ggp_cswap_rows<N>(R, k, tr[k]);
    double dinv[N];
#pragma unroll
    for (int i = 0; i < N; ++i) dinv[i] = 1.0 / lu[i * N + i];
#pragma unroll
    for (int j = 0; j < N; ++j) {
#pragma unroll
        for (int k = 0; k < N; ++k) {   // unit lower
            double b = R[k * N + j];
#pragma unroll
            for (int i = k + 1; i < N; ++i) R[i * N + j] = R[i * N + j] - b * lu[i * N + k];
        }
#pragma unroll
        for (int i = N - 1; i >= 0; --i) {   // upper
            double b = R[i * N + j] * dinv[i];
            R[i * N + j] = b;
#pragma unroll
            for (int s = 0; s < i; ++s) R[s * N + j] = R[s * N + j] - b * lu[s * N + i];
        }
    }
    return ok;
}

// C = A * B, coefficient-wise, inner sum left to right.  C may alias neither input.
template <int N>
GGP_HD void ggp_matmul(const double* __restrict__ A, const double* __restrict__ B, double* __restrict__ C) {
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = 0; j < N; ++j) {
            double s = A[i * N] * B[j];
#pragma unroll
            for (int k = 1; k < N; ++k) s = s + A[i * N + k] * B[k * N + j];
            C[i * N + j] = s;
        }
}

// y = A x for 4x4 A: Eigen's column-major GEMV processes four columns at once
GGP_HD void ggp_gemv4(const double* __restrict__ A, const double* __restrict__ x, double* __restrict__ y) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
        y[i] = 0.0 + ((A[i * 4] * x[0] + A[i * 4 + 1] * x[1]) + (A[i * 4 + 2] * x[2] + A[i * 4 + 3] * x[3]));
}

// N(m1, c1) * N(m2, c2) -> (m1, c1), predictions.h:183-188
GGP_HD void ggp_multiply_gaussian(double* __restrict__ m1, double* __restrict__ c1, const double* __restrict__ m2,
                                  const double* __restrict__ c2) {
    double i1[16], i2[16], s[16], nc[16];
    ggp_inv_lu<4>(c1, i1);
    ggp_inv_lu<4>(c2, i2);
#pragma unroll
    for (int i = 0; i < 16; ++i) s[i] = i1[i] + i2[i];
    ggp_inv_lu<4>(s, nc);
    double A[16], a[4], b[4];
    ggp_matmul<4>(nc, i1, A);
    ggp_gemv4(A, m1, a);
    ggp_matmul<4>(nc, i2, A);
    ggp_gemv4(A, m2, b);
#pragma unroll
    for (int i = 0; i < 4; ++i) m1[i] = a[i] + b[i];
#pragma unroll
    for (int i = 0; i < 16; ++i) c1[i] = nc[i];
}

// division by the stationary OU prior on (lambda, q), predictions.h:446-463; p = 11 parameters
GGP_HD void ggp_divide_by_prior(double* __restrict__ m, double* __restrict__ c, const double* __restrict__ p) {
    const double mean_prior[4] = {0, 0, p[0], p[3]};
    double P[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) P[i] = 0.0;
    P[10] = (2. * p[1]) / p[2];
    P[15] = (2. * p[4]) / p[5];
    double ci[16], d[16], nc[16];
    ggp_inv_lu<4>(c, ci);
#pragma unroll
    for (int i = 0; i < 16; ++i) d[i] = ci[i] - P[i];
    ggp_inv_lu<4>(d, nc);
    double a[4], b[4], r[4], nm[4];
    ggp_gemv4(ci, m, a);
    ggp_gemv4(P, mean_prior, b);
#pragma unroll
    for (int i = 0; i < 4; ++i) r[i] = a[i] - b[i];
    ggp_gemv4(nc, r, nm);
#pragma unroll
    for (int i = 0; i < 4; ++i) m[i] = nm[i];
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = nc[i];
}

// sign flip of the (lambda, q) frame, predictions.h:278-301
GGP_HD void ggp_reverse_mean(const double* __restrict__ m, double* __restrict__ out) {
    out[0] = m[0]; out[1] = m[1]; out[2] = -m[2]; out[3] = -m[3];
}
GGP_HD void ggp_reverse_cov(const double* __restrict__ c, double* __restrict__ out) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const bool flip = ((i < 2) != (j < 2));
            out[4 * i + j] = flip ? -c[4 * i + j] : c[4 * i + j];
        }
}
