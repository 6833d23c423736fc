// oracle/ggp_oracle.cpp — CPU oracle: restatement of the reference's likelihood / forward-backward /
// combine passes over lineage trees.  TEST INFRASTRUCTURE ONLY (see ggp_oracle.h): used by tests/,
// __graft_entry__.smoke() and bench.py's CPU baseline as the checker of the CUDA path; the product
// never includes, links or calls it.
//
// Each function cites the reference lines it follows (paths relative to /root/reference/src).
// Eigen is not available in this environment; the Eigen expressions are restated with the algorithm
// choices of Eigen 3.3:
//   * Matrix2d::inverse()            -> 1/det, cofactors (LU/InverseImpl.h, size-2 specialisation)
//   * MatrixXd::determinant()        -> partial-pivot LU, product of the diagonal times the sign
//   * MatrixXd::inverse() (4x4, 8x8) -> unblocked partial-pivot LU, then P*I solved with a unit-lower
//                                       and an upper column-oriented triangular solve (diagonal applied
//                                       as a multiplication by 1/u_ii)
//   * MatrixXd * MatrixXd            -> coefficient-wise lazy product, inner sum left to right
//   * MatrixXd(4x4) * VectorXd       -> column-major GEMV, 4 columns at once: (a0x0+a1x1)+(a2x2+a3x3)
//   * chained products               -> evaluated left to right with temporaries
// The likelihood path only meets inner dimensions of 2, where none of these choices can change a bit.
#include "ggp_oracle.h"

#include <cmath>
#include <cstring>
#include <numeric>
#include <vector>

#ifdef GGP_ORACLE_USE_REF
// variant built into oracle/_ref/libggp_oracle_ref.so: the arithmetic core is the REFERENCE'S OWN
// mean_cov_model.h + Faddeeva.cc (compiled unmodified by ref_shim.cpp); only the Eigen-dependent wrappers
// below are the restatement.  Used to time the reference's CPU math in bench.py and to cross-check.
extern "C" {
void ggp_ref_mean_cov_model(const double*, const double*, double, const double*, double*, double*);
void ggp_ref_cross_cov_model(const double*, const double*, double, const double*, double*);
double ggp_ref_dawson(double);
double ggp_ref_tauint(int, double, double, double, double, double);
}
namespace ggp_oracle_math {
static void mean_cov_model(double* mean, double* cov, double t, const double* p7, bool flip = false) {
    double q[7] = {flip ? -p7[0] : p7[0], p7[1], p7[2], flip ? -p7[3] : p7[3], p7[4], p7[5], flip ? -p7[6] : p7[6]};
    double m[4], c[16];
    ggp_ref_mean_cov_model(mean, cov, t, q, m, c);
    std::memcpy(mean, m, sizeof m);
    std::memcpy(cov, c, sizeof c);
}
static void cross_cov_model(const double* mean, const double* cov, double t, const double* p7, double* out) {
    ggp_ref_cross_cov_model(mean, cov, t, p7, out);
}
static double dawson(double x) { return ggp_ref_dawson(x); }
static double I0(double a, double b, double c, double t1, double t0) { return ggp_ref_tauint(0, a, b, c, t1, t0); }
static double I1(double a, double b, double c, double t1, double t0) { return ggp_ref_tauint(1, a, b, c, t1, t0); }
static double I2(double a, double b, double c, double t1, double t0) { return ggp_ref_tauint(2, a, b, c, t1, t0); }
static double I3(double a, double b, double c, double t1, double t0) { return ggp_ref_tauint(3, a, b, c, t1, t0); }
}  // namespace ggp_oracle_math
#else
#include "oracle_math.inc"
#endif

namespace {
namespace M = ggp_oracle_math;

// ------------------------------------------------------------------------------------------------
// small dense helpers (row-major)
// ------------------------------------------------------------------------------------------------
template <int N>
bool inv_lu(const double* A, double* R) {
    // Eigen 3.3 PartialPivLU (unblocked_lu for size <= 16) + solve(Identity)
    double lu[N * N];
    std::memcpy(lu, A, sizeof lu);
    int tr[N];
    bool ok = true;
    for (int k = 0; k < N; ++k) {
        int row = k;
        double big = std::fabs(lu[k * N + k]);
        for (int i = k + 1; i < N; ++i) {
            double v = std::fabs(lu[i * N + k]);
            if (v > big) { big = v; row = i; }
        }
        tr[k] = row;
        if (big != 0.0) {
            if (row != k)
                for (int j = 0; j < N; ++j) std::swap(lu[k * N + j], lu[row * N + j]);
            for (int i = k + 1; i < N; ++i) lu[i * N + k] /= lu[k * N + k];
        } else {
            ok = false;
        }
        for (int i = k + 1; i < N; ++i)
            for (int j = k + 1; j < N; ++j) lu[i * N + j] -= lu[i * N + k] * lu[k * N + j];
    }
    for (int i = 0; i < N * N; ++i) R[i] = 0.0;
    for (int i = 0; i < N; ++i) R[i * N + i] = 1.0;
    for (int k = 0; k < N; ++k)
        if (tr[k] != k)
            for (int j = 0; j < N; ++j) std::swap(R[k * N + j], R[tr[k] * N + j]);
    for (int j = 0; j < N; ++j) {
        for (int k = 0; k < N; ++k) {   // unit lower
            double b = R[k * N + j];
            for (int i = k + 1; i < N; ++i) R[i * N + j] -= b * lu[i * N + k];
        }
        for (int i = N - 1; i >= 0; --i) {   // upper
            double a = 1.0 / lu[i * N + i];
            double b = (R[i * N + j] *= a);
            for (int s = 0; s < i; ++s) R[s * N + j] -= b * lu[s * N + i];
        }
    }
    return ok;
}

double det2_lu(double s00, double s01, double s10, double s11) {
    // MatrixXd::determinant() on a 2x2: partial-pivot LU (likelihood.h:26,31)
    double sign = 1.0;
    if (std::fabs(s10) > std::fabs(s00)) {
        std::swap(s00, s10);
        std::swap(s01, s11);
        sign = -1.0;
    }
    if (s00 != 0.0) s10 /= s00;
    s11 -= s10 * s01;
    return sign * (s00 * s11);
}

template <int N>
void matmul(const double* A, const double* B, double* C) {   // lazy coefficient product, sequential inner sum
    double T[N * N];
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) {
            double s = A[i * N] * B[j];
            for (int k = 1; k < N; ++k) s += A[i * N + k] * B[k * N + j];
            T[i * N + j] = s;
        }
    std::memcpy(C, T, sizeof T);
}

void gemv4(const double* A, const double* x, double* y) {   // column-major GEMV kernel, 4 columns at once
    double t[4];
    for (int i = 0; i < 4; ++i)
        t[i] = 0.0 + ((A[i * 4] * x[0] + A[i * 4 + 1] * x[1]) + (A[i * 4 + 2] * x[2] + A[i * 4 + 3] * x[3]));
    std::memcpy(y, t, sizeof t);
}

inline const double* pv(const double* params, int seg) { return params + 11 * seg; }

// ------------------------------------------------------------------------------------------------
// filter pieces
// ------------------------------------------------------------------------------------------------
struct Meas {   // S and Si of one measurement update
    double xg0, xg1, S00, S01, S10, S11, Si00, Si01, Si10, Si11;
};

Meas measurement(const ggp_oracle_forest* f, const double* p, const double* mean, const double* cov, double x, double g) {
    // likelihood.h:54-67 (= predictions.h:123-135, :394-409)
    Meas m;
    m.xg0 = x - mean[0];
    m.xg1 = g - mean[1];
    double D11 = f->noise_model == 1 ? p[8] * (mean[1] + f->fp_auto) : p[8];
    m.S00 = cov[0] + p[7];
    m.S01 = cov[1] + 0;
    m.S10 = cov[4] + 0;
    m.S11 = cov[5] + D11;
    double det = m.S00 * m.S11 - m.S10 * m.S01;   // Matrix2d::inverse()
    double invdet = 1.0 / det;
    m.Si00 = m.S11 * invdet;
    m.Si10 = -m.S10 * invdet;
    m.Si01 = -m.S01 * invdet;
    m.Si11 = m.S00 * invdet;
    return m;
}

double log_likelihood(const Meas& m) {
    // likelihood.h:26-32
    double r0 = (-0.5 * m.xg0) * m.Si00 + (-0.5 * m.xg1) * m.Si10;
    double r1 = (-0.5 * m.xg0) * m.Si01 + (-0.5 * m.xg1) * m.Si11;
    double a = r0 * m.xg0 + r1 * m.xg1;
    return a - 0.5 * std::log(det2_lu(m.S00, m.S01, m.S10, m.S11)) - 2 * std::log(2 * M_PI);
}

void posterior(const Meas& m, double* mean, double* cov) {
    // predictions.h:84-89
    double K[2][4];
    for (int j = 0; j < 4; ++j) { K[0][j] = cov[j]; K[1][j] = cov[4 + j]; }
    double T[4][2];
    for (int i = 0; i < 4; ++i) {
        T[i][0] = K[0][i] * m.Si00 + K[1][i] * m.Si10;
        T[i][1] = K[0][i] * m.Si01 + K[1][i] * m.Si11;
    }
    for (int i = 0; i < 4; ++i) mean[i] = mean[i] + ((0.0 + T[i][0] * m.xg0) + T[i][1] * m.xg1);
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) cov[4 * i + j] = cov[4 * i + j] - (T[i][0] * K[0][j] + T[i][1] * K[1][j]);
}

void mean_cov_after_division(const ggp_oracle_forest* f, long c, const double* p, double* cell_mean, double* cell_cov) {
    // predictions.h:18-61
    double* mean = cell_mean + 4 * c;
    double* cov = cell_cov + 16 * c;
    long par = f->parent[c];
    std::memcpy(mean, cell_mean + 4 * par, 4 * sizeof(double));
    std::memcpy(cov, cell_cov + 16 * par, 16 * sizeof(double));
    double dt = f->time[f->cell_offset[c]] - f->time[f->cell_offset[par + 1] - 1];
    M::mean_cov_model(mean, cov, dt, p);
    double var_dx = p[9], var_dg = p[10];
    if (f->division_model == 1) {   // binomial
        cov[0] += var_dx;
        cov[1] = cov[4] = mean[1] / 2. * var_dx + cov[1];
        cov[5] = var_dx * (mean[1] * mean[1] + cov[5]) / 2. + var_dg * mean[1] / 4. * (1 - var_dx) + cov[5] / 4.;
        cov[9] /= 2; cov[6] /= 2;
        cov[13] /= 2; cov[7] /= 2;
        mean[0] = mean[0] + -std::log(2.);   // F*mean + f
        mean[1] = 0.5 * mean[1] + 0.0;
    } else {                        // gauss: D + F C F^T, F = diag(1, 1/2, 1, 1)
        mean[0] = mean[0] + -std::log(2.);
        mean[1] = 0.5 * mean[1] + 0.0;
        const double F[4] = {1, 0.5, 1, 1};
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) {
                double d = (i == j) ? (i == 0 ? var_dx : (i == 1 ? var_dg : 0.0)) : 0.0;
                cov[4 * i + j] = d + (F[i] * cov[4 * i + j]) * F[j];
            }
    }
}

void init_sc_distribution(const ggp_oracle_forest* f, long c, const double* p, double* cell_mean, double* cell_cov) {
    // predictions.h:63-82 — NOTE only the diagonal of a root's covariance is reset
    double* mean = cell_mean + 4 * c;
    double* cov = cell_cov + 16 * c;
    if (f->parent[c] < 0) {
        mean[0] = f->init_f[0];
        mean[1] = f->init_f[1];
        cov[0] = f->init_f[2];
        cov[5] = f->init_f[3];
        mean[2] = p[0];
        mean[3] = p[3];
        cov[10] = p[2] / (2. * p[1]);
        cov[15] = p[5] / (2. * p[4]);
    } else {
        mean_cov_after_division(f, c, p, cell_mean, cell_cov);
    }
}

struct LikCtx {
    const ggp_oracle_forest* f;
    const double* p;
    double* cell_mean;
    double* cell_cov;
    double* per_cell;
    double tl;
    long nan_cell, nan_t;
};

void sc_likelihood(LikCtx& L, long c) {
    // likelihood.h:36-103
    const ggp_oracle_forest* f = L.f;
    init_sc_distribution(f, c, L.p, L.cell_mean, L.cell_cov);
    double* mean = L.cell_mean + 4 * c;
    double* cov = L.cell_cov + 16 * c;
    long o = f->cell_offset[c], n = f->cell_offset[c + 1] - o;
    double own = 0.0;
    for (long t = 0; t < n; ++t) {
        Meas m = measurement(f, L.p, mean, cov, f->log_length[o + t], f->fp[o + t]);
        double ll = log_likelihood(m);
        L.tl += ll;
        own += ll;
        if (std::isnan(L.tl) && L.nan_cell < 0) {   // the reference throws here (likelihood.h:71-93)
            L.nan_cell = c;
            L.nan_t = t;
        }
        posterior(m, mean, cov);
        if (t < n - 1) M::mean_cov_model(mean, cov, f->time[o + t + 1] - f->time[o + t], L.p);
    }
    if (L.per_cell) L.per_cell[c] = own;
}

void likelihood_recr(LikCtx& L, long c) {
    // likelihood.h:110-122 (depth first: cell, daughter1 subtree, daughter2 subtree); explicit stack
    std::vector<long> st;
    st.push_back(c);
    while (!st.empty()) {
        long u = st.back();
        st.pop_back();
        sc_likelihood(L, u);
        if (L.f->daughter2[u] >= 0) st.push_back(L.f->daughter2[u]);
        if (L.f->daughter1[u] >= 0) st.push_back(L.f->daughter1[u]);
    }
}

// ------------------------------------------------------------------------------------------------
// backward pieces (predictions.h:183-444)
// ------------------------------------------------------------------------------------------------
void multiply_gaussian(double* m1, double* c1, const double* m2, const double* c2) {
    // predictions.h:183-188
    double i1[16], i2[16], s[16], nc[16];
    inv_lu<4>(c1, i1);
    inv_lu<4>(c2, i2);
    for (int i = 0; i < 16; ++i) s[i] = i1[i] + i2[i];
    inv_lu<4>(s, nc);
    double A[16], B[16], a[4], b[4];
    matmul<4>(nc, i1, A);
    matmul<4>(nc, i2, B);
    gemv4(A, m1, a);
    gemv4(B, m2, b);
    for (int i = 0; i < 4; ++i) m1[i] = a[i] + b[i];
    std::memcpy(c1, nc, sizeof nc);
}

void division_r_one(const ggp_oracle_forest* f, const double* p, const double* dmean, const double* dcov, double* mean, double* cov) {
    // predictions.h:213-238 (daughter 1) / :243-266 (daughter 2)
    double var_dx = p[9], var_dg = p[10];
    std::memcpy(mean, dmean, 4 * sizeof(double));
    std::memcpy(cov, dcov, 16 * sizeof(double));
    if (f->division_model == 1) {
        cov[0] += var_dx;
        cov[5] = 8. * var_dx * (mean[1] * mean[1] + cov[5]) + 2. * var_dg * mean[1] + 8. * cov[5];
        cov[1] = cov[4] = 2. * mean[1] * var_dx + 4. * cov[1];
        cov[9] *= 2; cov[6] *= 2;
        cov[13] *= 2; cov[7] *= 2;
        mean[0] += std::log(2.);
        mean[1] *= 2;
    } else {
        mean[0] = mean[0] + std::log(2.);
        mean[1] = 2 * mean[1] + 0.0;
        const double F[4] = {1, 2, 1, 1};
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) {
                double d = (i == j) ? (i == 0 ? var_dx : (i == 1 ? var_dg : 0.0)) : 0.0;
                cov[4 * i + j] = d + (F[i] * cov[4 * i + j]) * F[j];
            }
    }
}

void mean_cov_after_division_r(const ggp_oracle_forest* f, long c, const double* p, double* cell_mean, double* cell_cov) {
    // predictions.h:201-275
    double* mean = cell_mean + 4 * c;
    double* cov = cell_cov + 16 * c;
    long d1 = f->daughter1[c], d2 = f->daughter2[c];
    division_r_one(f, p, cell_mean + 4 * d1, cell_cov + 16 * d1, mean, cov);
    if (d2 >= 0) {
        double mean2[4], cov2[16];
        division_r_one(f, p, cell_mean + 4 * d2, cell_cov + 16 * d2, mean2, cov2);
        multiply_gaussian(mean, cov, mean2, cov2);
    }
    double dt = f->time[f->cell_offset[d1]] - f->time[f->cell_offset[c + 1] - 1];
    M::mean_cov_model(mean, cov, dt, p, /*flip=*/true);
}

void reverse_mean(const double* m, double* out) {   // predictions.h:278-285
    out[0] = m[0]; out[1] = m[1]; out[2] = -m[2]; out[3] = -m[3];
}
void reverse_cov(const double* c, double* out) {    // predictions.h:287-301
    for (int i = 0; i < 16; ++i) out[i] = c[i];
    const int e[4][2] = {{0, 2}, {0, 3}, {1, 2}, {1, 3}};
    for (auto& k : e) {
        out[4 * k[0] + k[1]] = -c[4 * k[0] + k[1]];
        out[4 * k[1] + k[0]] = -c[4 * k[1] + k[0]];
    }
}

void divide_by_prior(double* m, double* c, const double* p) {
    // predictions.h:446-463
    double mean_prior[4] = {0, 0, p[0], p[3]};
    double P[16] = {0};
    P[10] = (2. * p[1]) / p[2];
    P[15] = (2. * p[4]) / p[5];
    double ci[16], d[16], nc[16];
    inv_lu<4>(c, ci);
    for (int i = 0; i < 16; ++i) d[i] = ci[i] - P[i];
    inv_lu<4>(d, nc);
    double a[4], b[4], r[4], nm[4];
    gemv4(ci, m, a);
    gemv4(P, mean_prior, b);
    for (int i = 0; i < 4; ++i) r[i] = a[i] - b[i];
    gemv4(nc, r, nm);
    std::memcpy(m, nm, sizeof nm);
    std::memcpy(c, nc, sizeof nc);
}

}  // namespace

// ================================================================================================
// C interface
// ================================================================================================
extern "C" {

void ggp_oracle_init_stats(ggp_oracle_forest* f) {
    // moma_input.h:663-735: vec_mean = accumulate/size; vec_var = inner_product/size - mean^2
    for (int dir = 0; dir < 2; ++dir) {
        std::vector<double> x0, g0;
        for (long c = 0; c < f->n_cells; ++c) {
            long o = f->cell_offset[c], n = f->cell_offset[c + 1] - o;
            if (n > 1) {
                long k = dir == 0 ? o : o + n - 1;
                x0.push_back(f->log_length[k]);
                g0.push_back(f->fp[k]);
            }
        }
        auto vmean = [](const std::vector<double>& v) { return std::accumulate(v.begin(), v.end(), 0.0) / v.size(); };
        auto vvar = [&](const std::vector<double>& v) {
            double sq_sum = std::inner_product(v.begin(), v.end(), v.begin(), 0.0);
            double m = vmean(v);
            return sq_sum / v.size() - m * m;
        };
        double* out = dir == 0 ? f->init_f : f->init_r;
        out[0] = vmean(x0);
        out[1] = vmean(g0);
        out[2] = vvar(x0);
        out[3] = vvar(g0);
    }
}

void ggp_oracle_mean_cov_model(const double* mean, const double* cov, double t, const double* p7, double* mean_out, double* cov_out) {
    double m[4], c[16];
    std::memcpy(m, mean, sizeof m);
    std::memcpy(c, cov, sizeof c);
    M::mean_cov_model(m, c, t, p7);
    std::memcpy(mean_out, m, sizeof m);
    std::memcpy(cov_out, c, sizeof c);
}

void ggp_oracle_cross_cov_model(const double* mean, const double* cov, double t, const double* p7, double* out) {
    M::cross_cov_model(mean, cov, t, p7, out);
}

double ggp_oracle_dawson(double x) { return M::dawson(x); }

double ggp_oracle_tauint(int k, double a, double b, double c, double t1, double t0) {
    switch (k) {
        case 0: return M::I0(a, b, c, t1, t0);
        case 1: return M::I1(a, b, c, t1, t0);
        case 2: return M::I2(a, b, c, t1, t0);
        default: return M::I3(a, b, c, t1, t0);
    }
}

double ggp_oracle_total_loglik(const ggp_oracle_forest* f, const double* params11, double* cell_mean, double* cell_cov,
                               double* per_cell_ll, long* nan_cell, long* nan_t) {
    // likelihood.h:125-138 and :170-174; roots in file order (moma_input.h:177-189)
    LikCtx L{f, params11, cell_mean, cell_cov, per_cell_ll, 0.0, -1, -1};
    for (long c = 0; c < f->n_cells; ++c)
        if (f->parent[c] < 0) likelihood_recr(L, c);
    if (nan_cell) *nan_cell = L.nan_cell;
    if (nan_t) *nan_t = L.nan_t;
    return L.tl;
}

void ggp_oracle_prediction_forward(const ggp_oracle_forest* f, const double* params, int n_seg, double* cell_mean, double* cell_cov,
                                   double* mean_f, double* cov_f) {
    // predictions.h:93-173 (pre-order)
    (void)n_seg;
    std::vector<long> st;
    for (long root = 0; root < f->n_cells; ++root) {
        if (f->parent[root] >= 0) continue;
        st.push_back(root);
        while (!st.empty()) {
            long c = st.back();
            st.pop_back();
            long o = f->cell_offset[c], n = f->cell_offset[c + 1] - o;
            int seg = f->parent[c] < 0 ? f->segment[o] : f->segment[f->cell_offset[f->parent[c] + 1] - 1];
            init_sc_distribution(f, c, pv(params, seg), cell_mean, cell_cov);
            double* mean = cell_mean + 4 * c;
            double* cov = cell_cov + 16 * c;
            for (long t = 0; t < n; ++t) {
                const double* p = pv(params, f->segment[o + t]);
                Meas m = measurement(f, p, mean, cov, f->log_length[o + t], f->fp[o + t]);
                posterior(m, mean, cov);
                std::memcpy(mean_f + 4 * (o + t), mean, 4 * sizeof(double));
                std::memcpy(cov_f + 16 * (o + t), cov, 16 * sizeof(double));
                if (t < n - 1) M::mean_cov_model(mean, cov, f->time[o + t + 1] - f->time[o + t], p);
            }
            if (f->daughter2[c] >= 0) st.push_back(f->daughter2[c]);
            if (f->daughter1[c] >= 0) st.push_back(f->daughter1[c]);
        }
    }
}

void ggp_oracle_prediction_backward(const ggp_oracle_forest* f, const double* params, int n_seg, double* cell_mean, double* cell_cov,
                                    double* mean_b, double* cov_b) {
    // predictions.h:368-444 (post-order: daughter1 subtree, daughter2 subtree, cell)
    (void)n_seg;
    std::vector<long> order;
    {
        std::vector<std::pair<long, int>> st;
        for (long root = 0; root < f->n_cells; ++root) {
            if (f->parent[root] >= 0) continue;
            st.push_back({root, 0});
            while (!st.empty()) {
                auto& top = st.back();
                long c = top.first;
                if (top.second == 0) { top.second = 1; if (f->daughter1[c] >= 0) { st.push_back({f->daughter1[c], 0}); continue; } }
                if (top.second == 1) { top.second = 2; if (f->daughter2[c] >= 0) { st.push_back({f->daughter2[c], 0}); continue; } }
                order.push_back(c);
                st.pop_back();
            }
        }
    }
    for (long c : order) {
        long o = f->cell_offset[c], n = f->cell_offset[c + 1] - o;
        int seg = f->segment[o + n - 1];
        const double* p0 = pv(params, seg);
        double* mean = cell_mean + 4 * c;
        double* cov = cell_cov + 16 * c;
        if (f->daughter1[c] < 0 && f->daughter2[c] < 0) {
            // init_sc_distribution_r, leaf branch (predictions.h:317-331): diagonal only
            mean[0] = f->init_r[0];
            mean[1] = f->init_r[1];
            cov[0] = f->init_r[2];
            cov[5] = f->init_r[3];
            mean[2] = -p0[0];
            mean[3] = -p0[3];
            cov[10] = p0[2] / (2. * p0[1]);
            cov[15] = p0[5] / (2. * p0[4]);
        } else {
            mean_cov_after_division_r(f, c, p0, cell_mean, cell_cov);
        }
        for (long t = n - 1; t > -1; --t) {
            reverse_mean(mean, mean_b + 4 * (o + t));
            reverse_cov(cov, cov_b + 16 * (o + t));
            const double* p = pv(params, f->segment[o + t]);
            Meas m = measurement(f, p, mean, cov, f->log_length[o + t], f->fp[o + t]);
            posterior(m, mean, cov);
            if (t > 0) {
                const double* pp = pv(params, f->segment[o + t - 1]);
                M::mean_cov_model(mean, cov, f->time[o + t] - f->time[o + t - 1], pp, /*flip=*/true);
            }
        }
    }
}

void ggp_oracle_combine_predictions(const ggp_oracle_forest* f, const double* params, int n_seg, const double* mean_f, const double* cov_f,
                                    const double* mean_b, const double* cov_b, double* mean_p, double* cov_p) {
    // predictions.h:466-499
    (void)n_seg;
    for (long c = 0; c < f->n_cells; ++c) {
        long o = f->cell_offset[c], n = f->cell_offset[c + 1] - o;
        for (long j = 0; j < n; ++j) {
            double tm[4], tc[16];
            std::memcpy(tm, mean_f + 4 * (o + j), sizeof tm);
            std::memcpy(tc, cov_f + 16 * (o + j), sizeof tc);
            multiply_gaussian(tm, tc, mean_b + 4 * (o + j), cov_b + 16 * (o + j));
            int seg;
            if (j == 0)
                seg = f->parent[c] < 0 ? f->segment[o] : f->segment[f->cell_offset[f->parent[c] + 1] - 1];
            else
                seg = f->segment[o + j];
            divide_by_prior(tm, tc, pv(params, seg));
            std::memcpy(mean_p + 4 * (o + j), tm, sizeof tm);
            std::memcpy(cov_p + 16 * (o + j), tc, sizeof tc);
        }
    }
}

}  // extern "C"

// ================================================================================================
// joints (correlation_tree.h, Gaussians.h) — restated with small dynamic matrices, like the reference
// ================================================================================================
namespace {

struct Mat {   // row-major dynamic matrix, at most 8x8
    int r = 0, c = 0;
    double v[64];
    Mat() {}
    Mat(int r_, int c_) : r(r_), c(c_) { for (int i = 0; i < r * c; ++i) v[i] = 0.0; }
    double& operator()(int i, int j) { return v[i * c + j]; }
    double operator()(int i, int j) const { return v[i * c + j]; }
};
struct Vec {
    int n = 0;
    double v[8];
    Vec() {}
    explicit Vec(int n_) : n(n_) { for (int i = 0; i < n; ++i) v[i] = 0.0; }
    double& operator()(int i) { return v[i]; }
    double operator()(int i) const { return v[i]; }
};

Mat mat4(const double* p) { Mat m(4, 4); for (int i = 0; i < 16; ++i) m.v[i] = p[i]; return m; }
Vec vec4(const double* p) { Vec x(4); for (int i = 0; i < 4; ++i) x.v[i] = p[i]; return x; }

Mat T(const Mat& a) { Mat t(a.c, a.r); for (int i = 0; i < a.r; ++i) for (int j = 0; j < a.c; ++j) t(j, i) = a(i, j); return t; }
Mat add(const Mat& a, const Mat& b) { Mat m(a.r, a.c); for (int i = 0; i < a.r * a.c; ++i) m.v[i] = a.v[i] + b.v[i]; return m; }
Mat sub(const Mat& a, const Mat& b) { Mat m(a.r, a.c); for (int i = 0; i < a.r * a.c; ++i) m.v[i] = a.v[i] - b.v[i]; return m; }
Vec add(const Vec& a, const Vec& b) { Vec m(a.n); for (int i = 0; i < a.n; ++i) m.v[i] = a.v[i] + b.v[i]; return m; }
Vec sub(const Vec& a, const Vec& b) { Vec m(a.n); for (int i = 0; i < a.n; ++i) m.v[i] = a.v[i] - b.v[i]; return m; }
Vec neg(const Vec& a) { Vec m(a.n); for (int i = 0; i < a.n; ++i) m.v[i] = -a.v[i]; return m; }

// MatrixXd * MatrixXd for these sizes: lazy coefficient product, inner sum left to right
Mat mul(const Mat& a, const Mat& b) {
    Mat m(a.r, b.c);
    for (int i = 0; i < a.r; ++i)
        for (int j = 0; j < b.c; ++j) {
            double s = a(i, 0) * b(0, j);
            for (int k = 1; k < a.c; ++k) s += a(i, k) * b(k, j);
            m(i, j) = s;
        }
    return m;
}
// MatrixXd * VectorXd: column-major GEMV, four columns at a time, the remaining ones one by one
Vec mul(const Mat& a, const Vec& x) {
    Vec y(a.r);
    int j = 0;
    for (; j + 4 <= a.c; j += 4)
        for (int i = 0; i < a.r; ++i)
            y(i) = y(i) + ((a(i, j) * x(j) + a(i, j + 1) * x(j + 1)) + (a(i, j + 2) * x(j + 2) + a(i, j + 3) * x(j + 3)));
    for (; j < a.c; ++j)
        for (int i = 0; i < a.r; ++i) y(i) = y(i) + a(i, j) * x(j);
    return y;
}
Mat inverse(const Mat& a) {
    Mat m(a.r, a.c);
    if (a.r == 2) inv_lu<2>(a.v, m.v);
    else inv_lu<4>(a.v, m.v);
    return m;
}
Mat block(const Mat& a, int i0, int j0, int r, int c) {
    Mat m(r, c);
    for (int i = 0; i < r; ++i) for (int j = 0; j < c; ++j) m(i, j) = a(i0 + i, j0 + j);
    return m;
}
Vec head(const Vec& a, int n) { Vec x(n); for (int i = 0; i < n; ++i) x(i) = a(i); return x; }
Vec tail(const Vec& a, int n) { Vec x(n); for (int i = 0; i < n; ++i) x(i) = a(a.n - n + i); return x; }
Vec cat(const Vec& a, const Vec& b) { Vec x(a.n + b.n); for (int i = 0; i < a.n; ++i) x(i) = a(i); for (int i = 0; i < b.n; ++i) x(a.n + i) = b(i); return x; }
Mat blocks(const Mat& a, const Mat& b, const Mat& c, const Mat& d) {   // vstack(hstack(a, b), hstack(c, d))
    Mat m(a.r + c.r, a.c + b.c);
    for (int i = 0; i < a.r; ++i) { for (int j = 0; j < a.c; ++j) m(i, j) = a(i, j); for (int j = 0; j < b.c; ++j) m(i, a.c + j) = b(i, j); }
    for (int i = 0; i < c.r; ++i) { for (int j = 0; j < c.c; ++j) m(a.r + i, j) = c(i, j); for (int j = 0; j < d.c; ++j) m(a.r + i, c.c + j) = d(i, j); }
    return m;
}

struct Gaussian { Vec m; Mat C; };                 // Gaussians.h:24-40
struct Affine { Vec a; Mat F; Mat A; };            // Gaussians.h:52-69
struct Separated { Gaussian marginal; Affine conditional; };   // Gaussians.h:91-107

Gaussian gaussian_multiply(const Gaussian& n1, const Gaussian& n2) {   // Gaussians.h:42-49
    Mat Si = inverse(add(n1.C, n2.C));
    Gaussian n3;
    n3.m = add(mul(mul(n2.C, Si), n1.m), mul(mul(n1.C, Si), n2.m));
    n3.C = mul(mul(n1.C, Si), n2.C);
    return n3;
}
Affine affine_transform(const Affine& g) {   // Gaussians.h:71-81
    Mat Fi = inverse(g.F);
    Affine n;
    n.a = neg(mul(Fi, g.a));
    n.F = Fi;
    n.A = mul(mul(Fi, g.A), T(Fi));
    return n;
}
Gaussian affine_transform_at(const Affine& g, const Vec& y) {   // Gaussians.h:83-87
    Mat Fi = inverse(g.F);
    Gaussian n;
    n.m = mul(Fi, sub(y, g.a));
    n.C = mul(mul(Fi, g.A), T(Fi));
    return n;
}
Gaussian to_joint(const Separated& s) {   // Gaussians.h:109-124
    Gaussian j;
    j.m = cat(s.marginal.m, add(s.conditional.a, mul(s.conditional.F, s.marginal.m)));
    j.C = blocks(s.marginal.C, mul(T(s.marginal.C), T(s.conditional.F)), mul(s.conditional.F, s.marginal.C),
                 add(s.conditional.A, mul(mul(s.conditional.F, T(s.marginal.C)), T(s.conditional.F))));
    return j;
}
Separated separate_gaussian(const Gaussian& joint) {   // Gaussians.h:127-145
    const int n = 4;
    Mat B = block(joint.C, n, n, n, n), K = block(joint.C, 0, n, n, n), A = block(joint.C, 0, 0, n, n);
    Vec a = head(joint.m, n), b = tail(joint.m, n);
    Separated s;
    s.marginal.m = a;
    s.marginal.C = A;
    Mat KtAi = mul(T(K), inverse(A));
    s.conditional.a = sub(b, mul(KtAi, a));
    s.conditional.F = KtAi;
    s.conditional.A = sub(B, mul(KtAi, K));
    return s;
}
Gaussian flip_xy(const Gaussian& g) {   // Gaussians.h:147-158
    const int n = 4;
    Gaussian r;
    r.m = cat(tail(g.m, n), head(g.m, n));
    r.C = blocks(block(g.C, n, n, n, n), T(block(g.C, 0, n, n, n)), T(block(g.C, n, 0, n, n)), block(g.C, 0, 0, n, n));
    return r;
}

Gaussian include_measurement(const Gaussian& joint, double D00, double D11, double x, double g) {   // correlation_tree.h:132-154
    Mat S = block(joint.C, 0, 0, 2, 2);
    S(0, 0) = S(0, 0) + D00; S(0, 1) = S(0, 1) + 0.0; S(1, 0) = S(1, 0) + 0.0; S(1, 1) = S(1, 1) + D11;
    Vec xg(2);
    xg(0) = x - joint.m(0);
    xg(1) = g - joint.m(1);
    Mat Si = inverse(S);
    Mat K = block(joint.C, 0, 0, 2, joint.C.c);
    Mat KtSi = mul(T(K), Si);
    Gaussian n;
    n.m = add(joint.m, mul(KtSi, xg));
    n.C = sub(joint.C, mul(KtSi, K));
    return n;
}

struct StepOut { Vec mean1, mean2; Mat cov1, cov2, cross; };

// the common part of consecutive_joint / consecutive_conditional (correlation_tree.h:325-345, :360-382)
StepOut step_from_forward(const double* mf, const double* cf, double dt, const double* p) {
    StepOut o;
    o.mean1 = vec4(mf);
    o.cov1 = mat4(cf);
    double cross[16], m[4], c[16];
    M::cross_cov_model(mf, cf, dt, p, cross);
    std::memcpy(m, mf, sizeof m);
    std::memcpy(c, cf, sizeof c);
    M::mean_cov_model(m, c, dt, p);
    o.cross = mat4(cross);
    o.mean2 = vec4(m);
    o.cov2 = mat4(c);
    return o;
}
// the binomial branch shared by consecutive_joint_cell_division / consecutive_conditional_cell_division
StepOut division_binomial(const double* mf, const double* cf, double dt, const double* p) {   // correlation_tree.h:165-203
    StepOut o;
    o.mean1 = vec4(mf);
    o.cov1 = mat4(cf);
    double mean[4], cov[16];
    std::memcpy(mean, mf, sizeof mean);
    std::memcpy(cov, cf, sizeof cov);
    M::mean_cov_model(mean, cov, dt, p);
    double var_dx = p[9], var_dg = p[10];
    cov[0] += var_dx;
    cov[1] = cov[4] = mean[1] / 2. * var_dx + cov[1];
    cov[5] = var_dx * (mean[1] * mean[1] + cov[5]) / 2. + var_dg * mean[1] / 4. * (1 - var_dx) + cov[5] / 4.;
    cov[9] /= 2; cov[6] /= 2;
    cov[13] /= 2; cov[7] /= 2;
    mean[0] = mean[0] + -std::log(2.);
    mean[1] = 0.5 * mean[1] + 0.0;
    o.cross = mat4(cf);
    for (int j = 0; j < 4; ++j) o.cross(1, j) /= 2.;
    o.mean2 = vec4(mean);
    o.cov2 = mat4(cov);
    return o;
}
Affine division_gauss_conditional(const double* p) {   // correlation_tree.h:176-179, :224-228, :302-312
    Affine c;
    c.F = Mat(4, 4);
    for (int i = 0; i < 4; ++i) c.F(i, i) = 1.0;
    c.F(1, 1) = 0.5;
    c.a = Vec(4);
    c.a(0) = -std::log(2.);
    c.A = Mat(4, 4);
    c.A(0, 0) = p[9];
    c.A(1, 1) = p[10];
    return c;
}

struct JointCtx {
    const ggp_oracle_forest* f;
    const double* params;
    double tol;
    const double* cell_mean;   // after the backward pass
    const double *mean_f, *cov_f, *mean_b, *cov_b;
    long cap, count;
    long *row, *col;
    double* rec;
    long row_ctp;
    std::vector<Gaussian> cell_joint;   // MOMAdata::joint
};

Gaussian consecutive_joint(const JointCtx& X, long c, long t, const double* p) {   // correlation_tree.h:325-357
    long k = X.f->cell_offset[c] + t;
    StepOut o = step_from_forward(X.mean_f + 4 * k, X.cov_f + 16 * k, X.f->time[k + 1] - X.f->time[k], p);
    Gaussian j;
    j.m = cat(o.mean2, o.mean1);
    j.C = blocks(o.cov2, o.cross, T(o.cross), o.cov1);
    return j;
}
Affine consecutive_conditional(const JointCtx& X, long c, long t, const double* p) {   // correlation_tree.h:360-396
    long k = X.f->cell_offset[c] + t;
    StepOut o = step_from_forward(X.mean_f + 4 * k, X.cov_f + 16 * k, X.f->time[k + 1] - X.f->time[k], p);
    Gaussian j;
    j.m = cat(o.mean1, o.mean2);
    j.C = blocks(o.cov1, T(o.cross), o.cross, o.cov2);
    return affine_transform(separate_gaussian(j).conditional);
}
Gaussian consecutive_joint_cell_division(const JointCtx& X, long c, long t, const double* p) {   // correlation_tree.h:160-238
    const ggp_oracle_forest* f = X.f;
    long k = f->cell_offset[c] + t;
    double dt = f->time[f->cell_offset[f->daughter1[c]]] - f->time[k];
    if (f->division_model == 1) {
        StepOut o = division_binomial(X.mean_f + 4 * k, X.cov_f + 16 * k, dt, p);
        Gaussian j;
        j.m = cat(o.mean2, o.mean1);
        j.C = blocks(o.cov2, o.cross, T(o.cross), o.cov1);
        return j;
    }
    Separated s;
    s.marginal.m = vec4(X.mean_f + 4 * k);
    s.marginal.C = mat4(X.cov_f + 16 * k);
    s.conditional = division_gauss_conditional(p);
    return flip_xy(to_joint(s));
}
Affine consecutive_conditional_cell_division(const JointCtx& X, long c, long t, const double* p) {   // correlation_tree.h:241-319
    const ggp_oracle_forest* f = X.f;
    long k = f->cell_offset[c] + t;
    double dt = f->time[f->cell_offset[f->daughter1[c]]] - f->time[k];
    if (f->division_model == 1) {
        StepOut o = division_binomial(X.mean_f + 4 * k, X.cov_f + 16 * k, dt, p);
        Gaussian j;
        j.m = cat(o.mean1, o.mean2);
        j.C = blocks(o.cov1, T(o.cross), o.cross, o.cov2);
        return affine_transform(separate_gaussian(j).conditional);
    }
    return affine_transform(division_gauss_conditional(p));
}

Gaussian next_joint(const Gaussian& joint, const Affine& conditional) {   // correlation_tree.h:403-454
    Separated sep = separate_gaussian(joint);
    const Gaussian& n1 = sep.marginal;
    Mat Si = inverse(add(n1.C, conditional.A));
    Affine NX;
    NX.a = add(mul(mul(conditional.A, Si), n1.m), mul(mul(n1.C, Si), conditional.a));   // calc_x
    NX.F = mul(mul(n1.C, Si), conditional.F);                                            // calc_X
    NX.A = mul(mul(n1.C, Si), conditional.A);                                            // calc_Y
    Affine G;
    G.a = conditional.a;
    G.F = conditional.F;
    G.A = add(n1.C, conditional.A);
    Gaussian next_marginal = affine_transform_at(G, n1.m);
    Affine nc;   // propagation(sep.conditional, NX), correlation_tree.h:418-423
    nc.a = add(sep.conditional.a, mul(sep.conditional.F, NX.a));
    nc.F = mul(sep.conditional.F, NX.F);
    nc.A = add(sep.conditional.A, mul(mul(sep.conditional.F, NX.A), T(sep.conditional.F)));
    Separated nj;
    nj.marginal = next_marginal;
    nj.conditional = nc;
    return to_joint(nj);
}

Gaussian incorporate_backward_prob(const Separated& joint, const double* mb, const double* cb, const double* p) {   // correlation_tree.h:457-482
    double m[4], c[16];
    std::memcpy(m, mb, sizeof m);
    std::memcpy(c, cb, sizeof c);
    divide_by_prior(m, c, p);   // same expressions as predictions.h:446-463
    Gaussian backward;
    backward.m = vec4(m);
    backward.C = mat4(c);
    Separated nj;
    nj.marginal = gaussian_multiply(joint.marginal, backward);
    nj.conditional = joint.conditional;
    return to_joint(nj);
}

bool crosscovariance_is_small(const Gaussian& joint, double tolerance) {   // correlation_tree.h:484-493
    for (int i = 0; i < 4; ++i)
        for (int j = 4; j < 8; ++j)
            if (std::abs(joint.C(i, j) / (joint.m(i) * joint.m(j))) > tolerance) return false;
    return true;
}

void emit(JointCtx& X, long col, const Gaussian& g) {   // Joint_vector::add, correlation_tree.h:65-75
    if (X.count < X.cap) {
        X.row[X.count] = X.row_ctp;
        X.col[X.count] = col;
        double* r = X.rec + 44 * X.count;
        for (int i = 0; i < 8; ++i) r[i] = g.m(i);
        int q = 8;
        for (int i = 0; i < 8; ++i) for (int j = i; j < 8; ++j) r[q++] = g.C(i, j);
    }
    ++X.count;
}

bool calc_joint_distributions(JointCtx& X, long c, long n) {   // correlation_tree.h:499-558
    const ggp_oracle_forest* f = X.f;
    long o = f->cell_offset[c], T_ = f->cell_offset[c + 1] - o;
    for (long m = 1; n + m < T_; ++m) {
        long idx = n + m;
        const double* p = pv(X.params, f->segment[o + idx]);
        double D11 = f->noise_model == 1 ? p[8] * (X.cell_mean[4 * c + 1] + f->fp_auto) : p[8];
        X.cell_joint[c] = include_measurement(X.cell_joint[c], p[7], D11, f->log_length[o + idx], f->fp[o + idx]);
        Gaussian combined = incorporate_backward_prob(separate_gaussian(X.cell_joint[c]), X.mean_b + 4 * (o + idx), X.cov_b + 16 * (o + idx), p);
        if (crosscovariance_is_small(combined, X.tol)) return true;
        emit(X, o + idx, combined);
        if (idx < T_ - 1) {
            Affine cond = consecutive_conditional(X, c, idx, p);
            X.cell_joint[c] = next_joint(X.cell_joint[c], cond);
        } else {
            if (f->daughter1[c] >= 0) {
                Affine cond = consecutive_conditional_cell_division(X, c, idx, pv(X.params, f->segment[o + T_ - 1]));
                X.cell_joint[c] = next_joint(X.cell_joint[c], cond);
                X.cell_joint[f->daughter1[c]] = X.cell_joint[c];
            }
            if (f->daughter2[c] >= 0) X.cell_joint[f->daughter2[c]] = X.cell_joint[c];
        }
    }
    return false;
}

void joint_distributions_recr(JointCtx& X, long c, long n, bool is_joint_at_division) {   // correlation_tree.h:566-585
    if (c < 0) return;
    if (!is_joint_at_division) {
        if (calc_joint_distributions(X, c, n)) return;
    }
    joint_distributions_recr(X, X.f->daughter1[c], -1, false);
    joint_distributions_recr(X, X.f->daughter2[c], -1, false);
}

}  // namespace

extern "C" long ggp_oracle_joints(const ggp_oracle_forest* f, const double* params, int n_seg, double tol,
                                  const double* cell_mean_after_backward, const double* mean_f, const double* cov_f,
                                  const double* mean_b, const double* cov_b, long cap, long* row_ctp, long* col_ctp, double* rec44) {
    // collect_joint_distributions / sc_joint_distributions (correlation_tree.h:588-648)
    (void)n_seg;
    JointCtx X{f, params, tol, cell_mean_after_backward, mean_f, cov_f, mean_b, cov_b, cap, 0, row_ctp, col_ctp, rec44, 0, {}};
    X.cell_joint.resize(f->n_cells);
    for (long c = 0; c < f->n_cells; ++c) {
        long o = f->cell_offset[c], T_ = f->cell_offset[c + 1] - o;
        for (long n = 0; n < T_; ++n) {
            X.row_ctp = o + n;
            const double* p = pv(params, f->segment[o + n]);
            if (n < T_ - 1) {
                X.cell_joint[c] = consecutive_joint(X, c, n, p);
                joint_distributions_recr(X, c, n, false);
            } else {
                if (f->daughter1[c] >= 0) {
                    X.cell_joint[c] = consecutive_joint_cell_division(X, c, n, p);
                    X.cell_joint[f->daughter1[c]] = X.cell_joint[c];
                }
                if (f->daughter2[c] >= 0) X.cell_joint[f->daughter2[c]] = X.cell_joint[c];
                joint_distributions_recr(X, c, n, true);
            }
        }
    }
    return X.count;
}
