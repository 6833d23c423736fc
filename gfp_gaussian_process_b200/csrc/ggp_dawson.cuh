// ggp_dawson.cuh — Dawson's integral D(x) = sqrt(pi)/2 * exp(-x^2) * erfi(x) for real x,
// bit-compatible with the Faddeeva package the reference vendors
// (reference src/Faddeeva.cc:467-471 Dawson, :1864-1888 w_im, :1458-1862 w_im_y100).
//
// Algorithm (S. G. Johnson, 2012): D(x) = (sqrt(pi)/2) * w_im(x) with w_im odd and
//   |x| > 5e7        : 1/(sqrt(pi) x)
//   45 < |x| <= 5e7  : 5-term continued fraction collapsed to a rational function
//   |x| <= 45        : y100 = 100/(1+|x|); piece i = (int)y100; Horner polynomial in
//                      t = 2*y100 - (2i+1) for i <= 96, Taylor series for i >= 97 (|x| <= 0.0309).
// The reference evaluates these in plain double without FMA (g++ -O3, no -march);
// this file must be compiled with contraction off to keep the same roundings.
//
// Instead of the reference's 100-way switch (divergent on a GPU) the 97 coefficient
// rows live in one table (ggp_dawson_tables.h, zero padded to degree 8, which is
// bit-neutral) and every lane runs the same 8-step Horner loop on its own row.
#pragma once
#include "ggp_libm.cuh"

GGP_HD double ggp_w_im_pos(double x, const double* __restrict__ tab) {
    // x >= 0 and not NaN
    if (x > 45.0) {
        const double ispi = 0.56418958354775628694807945156;
        if (x > 5e7) return ispi / x;
        double xx = x * x;
        return ispi * (xx * (xx - 4.5) + 2) / (x * (xx * (xx - 5) + 3.75));
    }
    double y100 = 100 / (1 + x);
    int i = (int)y100;
    if (i >= 97) {
        double x2 = x * x;
        return x * (1.1283791670955125739
                    - x2 * (0.75225277806367504925
                            - x2 * (0.30090111122547001970
                                    - x2 * (0.085971746064420005629
                                            - x2 * 0.016931216931216931217))));
    }
    double t = 2 * y100 - (double)(2 * i + 1);
    const double* __restrict__ c = tab + 9 * i;
    double p = GGP_LDG(c + 8);
#pragma unroll
    for (int k = 7; k >= 0; --k) p = GGP_LDG(c + k) + p * t;
    return p;
}

GGP_HD_NOINLINE double ggp_dawson(double x, const GgpMathTables* __restrict__ M) {
    const double spi2 = 0.8862269254527580136490837416705725913990;   // sqrt(pi)/2
    if (x != x) return x;
    M = GGP_TABLES(M);
    double w = (x >= 0) ? ggp_w_im_pos(x, M->dawson_tab) : -ggp_w_im_pos(-x, M->dawson_tab);
    return spi2 * w;
}

// N independent Dawson values with the common path (Chebyshev piece, 0.0309 < |x| <= 45) as straight-line
// code; the rare branches go through the full routine.  Same operations, same bits as ggp_dawson.
template <int N>
GGP_HD void ggp_dawson_n(const double* __restrict__ x, double* __restrict__ y, const GgpMathTables* __restrict__ M) {
    const double spi2 = 0.8862269254527580136490837416705725913990;
    bool slow = false;
    double ax[N], y100[N];
    int idx[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        ax[i] = fabs(x[i]);
        y100[i] = 100 / (1 + ax[i]);
        const int pc = (int)y100[i];
        const bool ok = (ax[i] <= 45.0) && (pc < 97);   // false for NaN as well
        slow = slow || !ok;
        idx[i] = ok ? pc : 0;
    }
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const double t = 2 * y100[i] - (double)(2 * idx[i] + 1);
        const double* __restrict__ c = M->dawson_tab + 9 * idx[i];
        double p = GGP_LDG(c + 8);
#pragma unroll
        for (int k = 7; k >= 0; --k) p = GGP_LDG(c + k) + p * t;
        y[i] = spi2 * ((x[i] >= 0) ? p : -p);
    }
    if (slow) {
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const int pc = (int)y100[i];
            if (!((ax[i] <= 45.0) && (pc < 97))) y[i] = ggp_dawson(x[i], M);
        }
    }
}
