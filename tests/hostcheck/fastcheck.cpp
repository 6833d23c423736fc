// tests/hostcheck/fastcheck.cpp — the FAST likelihood step (gfp_gaussian_process_b200/csrc/ggp_fast.cuh) compiled for the
// host, in double (the arithmetic of the device kernel up to FMA contraction, any quadrature order) and in binary128 with a
// 16-node rule (the "truth" both the reference and the fast path are measured against in the gate report).
// Test infrastructure (tests/test_fast_host.py, tools/fast_gate.py); not part of the product.
#include <quadmath.h>

#include <vector>

#include "../../gfp_gaussian_process_b200/csrc/ggp_tables_data.h"
#include "../../gfp_gaussian_process_b200/csrc/ggp_dawson.cuh"
#include "../../gfp_gaussian_process_b200/csrc/ggp_layout.hpp"
#include "../../gfp_gaussian_process_b200/csrc/ggp_fast.cuh"

typedef __float128 quad;
template <> struct GgpFx<quad> {
    static quad exp_(quad x) { return expq(x); }
    static quad exp_mid_(quad x) { return expq(x); }
    static quad exp_small_(quad x) { return expq(x); }
    static quad exp_small8_(quad x) { return expq(x); }
    static quad exp_tiny_(quad x) { return expq(x); }
    static quad exp_tiny4_(quad x) { return expq(x); }
    static quad log_(quad x) { return logq(x); }
    static quad expm1_(quad x) { return expm1q(x); }
    static quad abs_(quad x) { return fabsq(x); }
    static bool finite_(quad x) { return x - x == 0; }
    static quad ln2() { return logq((quad)2); }
};

// Gauss-Legendre rule on [0, 1] in binary128 (Newton on P_n)
struct QuadGL {
    int n;
    std::vector<quad> x, w;
    explicit QuadGL(int n_) : n(n_), x(n_), w(n_) {
        const quad pi = M_PIq;
        for (int i = 0; i < n; ++i) {
            quad z = cosq(pi * (i + (quad)0.75) / (n + (quad)0.5)), dp = 0;
            for (int it = 0; it < 100; ++it) {
                quad p0 = 1, p1 = z;
                for (int k = 2; k <= n; ++k) { const quad p2 = ((2 * k - 1) * z * p1 - (k - 1) * p0) / k; p0 = p1; p1 = p2; }
                dp = n * (z * p1 - p0) / (z * z - 1);
                const quad dz = p1 / dp;
                z -= dz;
                if (fabsq(dz) < (quad)1e-33L) break;
            }
            quad p0 = 1, p1 = z;
            for (int k = 2; k <= n; ++k) { const quad p2 = ((2 * k - 1) * z * p1 - (k - 1) * p0) / k; p0 = p1; p1 = p2; }
            dp = n * (z * p1 - p0) / (z * z - 1);
            x[n - 1 - i] = (z + 1) / 2;
            w[n - 1 - i] = 1 / ((1 - z * z) * dp * dp);
        }
    }
    quad xi(int j) const { return x[j]; }
    quad om(int j) const { return w[j]; }
};
template <int N> struct DoubleGL {
    double xi(int j) const { return GgpGL<N>::xi(j); }
    double om(int j) const { return GgpGL<N>::om(j); }
};

static GgpDevForest fc_dev(const GgpLayout& L, const ggp_forest_desc* d) {
    GgpDevForest F;
    F.n_cells = L.n_cells; F.n_ctp = L.n_ctp;
    F.time = d->time; F.x = d->log_length; F.g = d->fp; F.seg = L.seg.data();
    F.s_off = L.s_off.data(); F.s_n = L.s_n.data(); F.s_parent = L.s_parent.data(); F.s_d1 = L.s_d1.data();
    F.s_d2 = L.s_d2.data(); F.s_root = L.s_root.data(); F.s_cell = L.s_cell.data(); F.s_dfs0 = L.s_dfs0.data();
    F.model.noise_scaled = d->noise_model == GGP_NOISE_SCALED;
    F.model.division_binomial = d->division_model == GGP_DIVISION_BINOMIAL;
    F.model.fp_auto = d->fp_auto;
    for (int i = 0; i < 4; ++i) { F.init_f[i] = L.init_f[i]; F.init_r[i] = L.init_r[i]; }
    return F;
}

template <class T, int N, class GL>
static int run(const ggp_forest_desc* d, const double* params, int n_vec, double* out_cell_ll, double* out_total, int* out_valid,
               double* out_state14, const GL& gln) {
    GgpLayout L;
    if (!L.build(d).empty()) return -1;
    const GgpDevForest F = fc_dev(L, d);
    std::vector<GgpFastState<T>> state((size_t)L.n_cells);
    for (int v = 0; v < n_vec; ++v) {
        T p[11];
        for (int i = 0; i < 11; ++i) p[i] = (T)params[11 * v + i];
        GgpFastConstsOnDemand<T, N, GL> kp(F, p, gln);
        bool valid = true;
        T total = 0;
        for (int64_t slot = 0; slot < L.n_cells; ++slot) {
            GgpFastState<T> s;
            if (F.s_parent[slot] >= 0) s = state[(size_t)F.s_parent[slot]];
            const T own = ggp_fast_cell<T, N>(F, (int)slot, p, s, kp, valid);
            state[(size_t)slot] = s;
            if (out_cell_ll) out_cell_ll[(size_t)v * L.n_cells + L.cell_of_slot[slot]] = (double)own;
            total += own;
        }
        out_total[v] = (double)total;
        out_valid[v] = valid ? 1 : 0;
        if (out_state14 && v == n_vec - 1)
            for (int64_t slot = 0; slot < L.n_cells; ++slot) {
                double* o = out_state14 + 14 * (size_t)L.cell_of_slot[slot];
                for (int i = 0; i < 4; ++i) o[i] = (double)state[(size_t)slot].m[i];
                for (int i = 0; i < 10; ++i) o[4 + i] = (double)state[(size_t)slot].c[i];
            }
    }
    return 0;
}

extern "C" {
// likelihood of the forest with the fast step in double, n_nodes in {3,4,5,6,8,10,12,16}; out_cell_ll NULL or [n_vec][n_cells];
// out_total [n_vec]; out_valid [n_vec]; out_state14 NULL or [n_cells][14] end-of-cell posteriors of the last vector
int fc_loglik(const ggp_forest_desc* d, const double* params, int n_vec, int n_nodes, double* out_cell_ll, double* out_total,
              int* out_valid, double* out_state14) {
    switch (n_nodes) {
        case 3: return run<double, 3>(d, params, n_vec, out_cell_ll, out_total, out_valid, out_state14, DoubleGL<3>());
        case 4: return run<double, 4>(d, params, n_vec, out_cell_ll, out_total, out_valid, out_state14, DoubleGL<4>());
        case 5: return run<double, 5>(d, params, n_vec, out_cell_ll, out_total, out_valid, out_state14, DoubleGL<5>());
        case 6: return run<double, 6>(d, params, n_vec, out_cell_ll, out_total, out_valid, out_state14, DoubleGL<6>());
        case 8: return run<double, 8>(d, params, n_vec, out_cell_ll, out_total, out_valid, out_state14, DoubleGL<8>());
        case 10: return run<double, 10>(d, params, n_vec, out_cell_ll, out_total, out_valid, out_state14, DoubleGL<10>());
        case 12: return run<double, 12>(d, params, n_vec, out_cell_ll, out_total, out_valid, out_state14, DoubleGL<12>());
        case 16: return run<double, 16>(d, params, n_vec, out_cell_ll, out_total, out_valid, out_state14, DoubleGL<16>());
    }
    return -2;
}
// the same in binary128 with the 16-node rule: the reference value of the gate report
int fc_loglik_quad(const ggp_forest_desc* d, const double* params, int n_vec, double* out_cell_ll, double* out_total, int* out_valid,
                   double* out_state14) {
    return run<quad, 16>(d, params, n_vec, out_cell_ll, out_total, out_valid, out_state14, QuadGL(16));
}
}
