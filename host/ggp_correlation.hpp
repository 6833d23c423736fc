// host/ggp_correlation.hpp — correlation functions from the joint posteriors: the reference's post-processing
// script python_src/correlation_from_joint.py (Gaussian :69-123, log_likelihood_function/_error :131-176, Correlation
// :179-406, cell_lineage_lookup :408-440, files2correlation_function :443-560, corr_to_csv :571-643, process_file
// :662-703) restated in C++ so that `-j` no longer has to go through the dense joints file, whose size is quadratic in
// the data set.
//
// For every lag dt in np.arange(0, dt * n_data, dt) the script sums, over all pairs of points (t, t + dt) on one
// lineage, the first and second moments of the 8-dim joint P(z(t+dt), z(t) | D):  n, sum m, sum (m m^T + C), and the same
// for the concentration c = g / exp(x).  Pairs the joints pass did not emit (their cross-covariance fell below the
// tolerance) enter with the product of their marginals; dt = 0 comes from the prediction file.  Then per lag: the
// covariance, the naive correlation and a grid-search maximum-likelihood correlation with its error bar.
// The sums are kept in long double like the script's np.longfloat, additions happen in the script's order (rows in file
// order, columns in file order), so the file front end reproduces the script bit for bit up to numpy's vectorised log.
//
// Two front ends:
//   CorrelationFromJoints  sparse records straight from ggp_joints (full precision, no file in between)
//   correlation_from_files the reference's <prefix>_joints.csv + <prefix>_prediction.csv (drop-in for the script)
#pragma once
#include <cmath>
#include <fstream>
#include <functional>
#include <unordered_map>

#include "ggp_util.hpp"

namespace ggp {

struct Gauss8 {
    double m[8];
    double C[8][8];
};

inline void tri10_to_sym4(const double* u, double C[4][4]) {   // xx xg xl xq gg gl gq ll lq qq
    int k = 0;
    for (int i = 0; i < 4; ++i)
        for (int j = i; j < 4; ++j) { C[i][j] = u[k]; C[j][i] = u[k]; ++k; }
}

// full joint: 8 means + 36 upper-triangular covariances (Gaussian n=8)
inline Gauss8 gauss_from_joint44(const double* v) {
    Gauss8 G;
    for (int i = 0; i < 8; ++i) G.m[i] = v[i];
    int k = 8;
    for (int i = 0; i < 8; ++i)
        for (int j = i; j < 8; ++j) { G.C[i][j] = v[k]; G.C[j][i] = v[k]; ++k; }
    return G;
}
// "joint" of a point with itself from its marginal (Gaussian n=4): every 4x4 block is the marginal covariance
inline Gauss8 gauss_from_marginal14(const double* v) {
    Gauss8 G;
    double C[4][4];
    tri10_to_sym4(v + 4, C);
    for (int i = 0; i < 8; ++i) {
        G.m[i] = v[i & 3];
        for (int j = 0; j < 8; ++j) G.C[i][j] = C[i & 3][j & 3];
    }
    return G;
}
// approximate joint of two points from their marginals (Gaussian n=2): block diagonal, (first, second)
inline Gauss8 gauss_from_two_marginals(const double* first14, const double* second14) {
    Gauss8 G;
    double C1[4][4], C2[4][4];
    tri10_to_sym4(first14 + 4, C1);
    tri10_to_sym4(second14 + 4, C2);
    for (int i = 0; i < 8; ++i) {
        G.m[i] = i < 4 ? first14[i] : second14[i - 4];
        for (int j = 0; j < 8; ++j) G.C[i][j] = (i < 4 && j < 4) ? C1[i][j] : ((i >= 4 && j >= 4) ? C2[i - 4][j - 4] : 0.0);
    }
    return G;
}

inline double corr_loglik(double Vyy, double Vyx, double Vxx, double sy, double sx, double r, double n) {
    const double r2 = r * r, q = sy / sx;
    return -n / 2 * (std::log(1 - r2) + (Vyy - 2 * r * sy / sx * Vyx + r2 * (q * q) * Vxx) / (sy * sy * (1 - r2)));
}

inline double corr_loglik_error(double Vyy, double Vyx, double Vxx, double sy, double sx, double r, double n) {
    // (scalar ** in the script is C pow)
    const double r2 = std::pow(r, 2.0), q = sy / sx, om = 1 - r2;
    const double log_term = n * (1 + r2) / std::pow(om, 2.0);
    const double q2 = std::pow(q, 2.0);
    const double v_term = -n / 2 * 1 / std::pow(sy, 2.0) *
                          ((2 * q2 * Vxx) / om + (8 * r * (r * q2 * Vxx - sy / sx * Vyx)) / std::pow(om, 2.0) +
                           ((8 * r2) / std::pow(om, 3.0) + 2 / std::pow(om, 2.0)) * (Vyy - 2 * r * sy / sx * Vyx + r2 * q2 * Vxx));
    const double dd = log_term + v_term;
    return -1 / dd > 0 ? std::sqrt(-1 / dd) : 0.0;
}

struct Correlation {
    double dt = 0;
    long n = 0;
    long double m[8] = {}, mm[8][8] = {}, c[2] = {}, cc[2][2] = {};
    double cov[8][8], cov_c[2][2], corr_naive[8][8], corr_c_naive[2][2];
    double corr_mle[8][8], corr_mle_err[8][8], corr_c_mle[2][2], corr_c_mle_err[2][2];
    double cov_mle[8][8], cov_mle_err[8][8], cov_c_mle[2][2], cov_c_mle_err[2][2];

    void add(const Gauss8& G) {
        ++n;
        for (int i = 0; i < 8; ++i) m[i] += G.m[i];
        for (int i = 0; i < 8; ++i)
            for (int j = 0; j < 8; ++j) mm[i][j] += G.m[j] * G.m[i] + G.C[i][j];
        const long double cv[2] = {G.m[1] / expl((long double)G.m[0]), G.m[5] / expl((long double)G.m[4])};
        for (int i = 0; i < 2; ++i) c[i] += cv[i];
        for (int i = 0; i < 2; ++i)
            for (int j = 0; j < 2; ++j) cc[i][j] += cv[j] * cv[i];
    }
    void average() {
        for (int i = 0; i < 8; ++i)
            for (int j = 0; j < 8; ++j) cov[i][j] = n > 0 ? (double)(mm[i][j] / n - m[i] / n * m[j] / n) : NAN;
        for (int i = 0; i < 2; ++i)
            for (int j = 0; j < 2; ++j) cov_c[i][j] = n > 0 ? (double)(cc[i][j] / n - c[i] / n * c[j] / n) : NAN;
    }
    void naive() {
        for (int i = 0; i < 8; ++i)
            for (int j = 0; j < 8; ++j) corr_naive[i][j] = n > 0 ? cov[i][j] / std::sqrt(cov[i][i] * cov[j][j]) : 0.0;
        for (int i = 0; i < 2; ++i)
            for (int j = 0; j < 2; ++j) corr_c_naive[i][j] = n > 0 ? cov_c[i][j] / std::sqrt(cov_c[i][i] * cov_c[j][j]) : 0.0;
    }
    // grid-search MLE of the correlation given the lag-0 variances (norm: correlation, else covariance)
    template <int N>
    void mle_block(const double (*V)[N], const double (*V0)[N], double (*out)[N], double (*err)[N], bool norm) const {
        const int G = 10000;
        const double lo = -1 + 1e-12, hi = 1 - 1e-12, step = (hi - lo) / (G - 1);
        for (int i = 0; i < N; ++i)
            for (int j = 0; j < N; ++j) {
                out[i][j] = 0; err[i][j] = 0;
                if (n <= 0) continue;
                const double Vyx = V[j][i], Vxx = V[i][i], Vyy = V[j][j], sy = std::sqrt(V0[j][j]), sx = std::sqrt(V0[i][i]);
                double best = -HUGE_VAL, r_max = lo;
                bool any = false;
                for (int k = 0; k < G; ++k) {
                    const double r = k == G - 1 ? hi : lo + k * step;
                    const double ll = corr_loglik(Vyy, Vyx, Vxx, sy, sx, r, (double)n);
                    // np.argmax: first maximum; NaN wins (numpy treats NaN as the maximum)
                    if (ll != ll) { if (!any || best == best) { best = ll; r_max = r; } any = true; break; }
                    if (!any || ll > best) { best = ll; r_max = r; any = true; }
                }
                double e = corr_loglik_error(Vyy, Vyx, Vxx, sy, sx, r_max, (double)n);
                if (!norm) { r_max = r_max * sy * sx; e *= sy * sx; }
                out[i][j] = r_max; err[i][j] = e;
            }
    }
    void mle(const Correlation& zero) {
        mle_block<8>(cov, zero.cov, corr_mle, corr_mle_err, true);
        mle_block<2>(cov_c, zero.cov_c, corr_c_mle, corr_c_mle_err, true);
        mle_block<8>(cov, zero.cov, cov_mle, cov_mle_err, false);
        mle_block<2>(cov_c, zero.cov_c, cov_c_mle, cov_c_mle_err, false);
    }
};

class CorrelationSet {
public:
    std::vector<Correlation> bins;
    double tol;
    // dts = np.arange(0, dt_max, dt) (start + i * step), tol as np.isclose's atol (rtol = 1e-5 like numpy's default)
    CorrelationSet(double dt_step, double dt_max, double tol_) : tol(tol_) {
        const long len = (long)std::ceil(dt_max / dt_step);
        for (long i = 0; i < len; ++i) { bins.emplace_back(); bins.back().dt = 0 + i * dt_step; }
    }
    void add(double dt, const Gauss8& G) {
        if (!(dt == dt) || std::isinf(dt)) return;
        for (Correlation& b : bins)
            if (std::fabs(b.dt - dt) <= tol + 1e-5 * std::fabs(dt)) { b.add(G); return; }
    }
    // moment sums accumulated elsewhere (ggp_correlation_sums): [n_bins][50] = n, m[8], upper(mm)[36], c[2], upper(cc)[3], each as
    // the leading double of the sum and (lo, may be null) its remainder
    void set_sums(const double* hi, const double* lo) {
        for (size_t b = 0; b < bins.size(); ++b) {
            auto at = [&](int k) { return (long double)hi[50 * b + k] + (lo ? (long double)lo[50 * b + k] : 0.0L); };
            Correlation& B = bins[b];
            B.n = (long)llroundl(at(0));
            for (int i = 0; i < 8; ++i) B.m[i] = at(1 + i);
            int q = 9;
            for (int i = 0; i < 8; ++i)
                for (int j = i; j < 8; ++j) { B.mm[i][j] = at(q); B.mm[j][i] = at(q); ++q; }
            B.c[0] = at(45); B.c[1] = at(46);
            B.cc[0][0] = at(47); B.cc[0][1] = at(48); B.cc[1][0] = at(48); B.cc[1][1] = at(49);
            // a single pair: its second moment IS the product of its first moments (the script's long-double sums give an exact
            // zero covariance there, hence 0 / 0 in the naive correlation; a product rounded to double would leave a residue)
            if (B.n == 1) for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) B.cc[i][j] = B.c[j] * B.c[i];
        }
    }
    void finalize() {
        for (Correlation& b : bins) { b.average(); b.naive(); }
        for (Correlation& b : bins) b.mle(bins[0]);
    }
    void to_csv(const std::string& file) const {
        std::ofstream f(file);
        f << "dt,cov_l(t+dt)l(t),cov_l(t+dt)l(t)_err,cov_l(t+dt)q(t),cov_l(t+dt)q(t)_err,cov_q(t+dt)l(t),cov_q(t+dt)l(t)_err,"
             "cov_q(t+dt)q(t),cov_q(t+dt)q(t)_err,cov_c(t+dt)c(t),cov_c(t+dt)c(t)_err,corr_l(t+dt)l(t),corr_l(t+dt)l(t)_err,"
             "corr_l(t+dt)q(t),corr_l(t+dt)q(t)_err,corr_q(t+dt)l(t),corr_q(t+dt)l(t)_err,corr_q(t+dt)q(t),corr_q(t+dt)q(t)_err,"
             "corr_c(t+dt)c(t),corr_c(t+dt)c(t)_err,corr_naive_l(t+dt)l(t),corr_naive_l(t+dt)q(t),corr_naive_q(t+dt)l(t),"
             "corr_naive_q(t+dt)q(t),corr_naive_c(t+dt)c(t),n_pairs\n";
        f.precision(17);
        const int idx[4][2] = {{2, 6}, {2, 7}, {3, 6}, {3, 7}};
        for (const Correlation& b : bins) {
            f << b.dt << ",";
            for (auto& ij : idx) f << b.cov_mle[ij[0]][ij[1]] << "," << b.cov_mle_err[ij[0]][ij[1]] << ",";
            f << b.cov_c_mle[0][1] << "," << b.cov_c_mle_err[0][1] << ",";
            for (auto& ij : idx) f << b.corr_mle[ij[0]][ij[1]] << "," << b.corr_mle_err[ij[0]][ij[1]] << ",";
            f << b.corr_c_mle[0][1] << "," << b.corr_c_mle_err[0][1] << ",";
            for (auto& ij : idx) f << b.corr_naive[ij[0]][ij[1]] << ",";
            f << b.corr_c_naive[0][1] << "," << b.n << "\n";
        }
    }
};

// cells on one lineage: a is an ancestor of b, b of a, or a == b (cell_lineage_lookup, :408-440)
class LineageLookup {
public:
    LineageLookup(const std::vector<std::string>& cell_ids, const std::vector<std::string>& parent_ids) {
        const size_t N = cell_ids.size();
        std::unordered_map<std::string, int> index;
        for (size_t c = 0; c < N; ++c) index[cell_ids[c]] = (int)c;   // a dict: the last occurrence wins
        parent.assign(N, -1);
        for (size_t c = 0; c < N; ++c) {
            const auto it = index.find(parent_ids[c]);
            if (it != index.end()) parent[c] = it->second;
        }
        children.assign(N, {});
        for (size_t c = 0; c < N; ++c) if (parent[c] >= 0) children[parent[c]].push_back((int)c);
    }
    // all cells on a lineage with c, sorted by index (= file order)
    std::vector<int> related(int c) const {
        std::vector<int> out, stack{c};
        while (!stack.empty()) {
            const int u = stack.back(); stack.pop_back();
            out.push_back(u);
            for (int k : children[u]) stack.push_back(k);
        }
        for (int u = parent[c]; u >= 0; u = parent[u]) out.push_back(u);
        std::sort(out.begin(), out.end());
        return out;
    }
    std::vector<int> parent;
    std::vector<std::vector<int>> children;
};

// Front end on the library's sparse joints.  `marginal14(k)`: 4 means + 10 upper-triangular covariances of the combined
// prediction at cell-timepoint k; `joints_of_rows(r0, r1, row, col, rec44)`: the records of the start points [r0, r1)
// sorted by (row, col).  cell_offset/time in file order.
struct JointsSource {
    std::function<const double*(int64_t)> marginal14;
    std::function<void(int64_t, int64_t, std::vector<int64_t>&, std::vector<int64_t>&, std::vector<double>&)> joints_of_rows;
};

inline void correlation_from_joints(CorrelationSet& CS, const std::vector<std::string>& cell_ids, const std::vector<std::string>& parent_ids,
                                    const std::vector<int64_t>& cell_offset, const std::vector<double>& time, const JointsSource& src,
                                    bool normalize_time, int64_t row_block = 4096) {
    const int64_t N = (int64_t)cell_ids.size(), M = cell_offset[N];
    const LineageLookup L(cell_ids, parent_ids);
    // lag 0 from the marginals (prediction-file loop, :477-499)
    for (int64_t k = 0; k < M; ++k) CS.add(0.0, gauss_from_marginal14(src.marginal14(k)));
    std::vector<int64_t> row, col;
    std::vector<double> rec;
    std::vector<int64_t> cell_of(M);
    for (int64_t c = 0; c < N; ++c) for (int64_t k = cell_offset[c]; k < cell_offset[c + 1]; ++k) cell_of[k] = c;
    for (int64_t r0 = 0; r0 < M; r0 += row_block) {
        const int64_t r1 = std::min(M, r0 + row_block);
        src.joints_of_rows(r0, r1, row, col, rec);
        size_t at = 0;
        int last_cell = -1;
        std::vector<int> rel;
        for (int64_t i = r0; i < r1; ++i) {
            const int c = (int)cell_of[i];
            if (c != last_cell) { rel = L.related(c); last_cell = c; }
            const double cycle = time[cell_offset[c + 1] - 1] - time[cell_offset[c]];
            // columns in file order: joints where emitted, else (same lineage and j > i) the product of the marginals
            for (int b : rel)
                for (int64_t j = cell_offset[b]; j < cell_offset[b + 1]; ++j) {
                    while (at < row.size() && (row[at] < i || (row[at] == i && col[at] < j))) {
                        if (row[at] == i) {   // an emitted joint of a column that is not on this lineage list (cannot happen) or before j
                            double dt = time[col[at]] - time[i];
                            if (normalize_time) dt /= cycle;
                            CS.add(dt, gauss_from_joint44(&rec[44 * at]));
                        }
                        ++at;
                    }
                    double dt = time[j] - time[i];
                    if (normalize_time) dt /= cycle;
                    if (at < row.size() && row[at] == i && col[at] == j) {
                        CS.add(dt, gauss_from_joint44(&rec[44 * at]));
                        ++at;
                    } else if (j > i) {
                        CS.add(dt, gauss_from_two_marginals(src.marginal14(j), src.marginal14(i)));
                    }
                }
            while (at < row.size() && row[at] == i) {
                double dt = time[col[at]] - time[i];
                if (normalize_time) dt /= cycle;
                CS.add(dt, gauss_from_joint44(&rec[44 * at]));
                ++at;
            }
        }
    }
}

// Front end on the reference's files (files2correlation_function, :443-560)
inline void correlation_from_files(CorrelationSet& CS, const std::string& joint_file, const std::string& prediction_file, bool normalize_time) {
    std::vector<std::string> cell_ids, parent_ids, point_cell;
    std::vector<std::vector<double>> marginals;
    std::vector<double> point_time;
    std::unordered_map<std::string, std::pair<double, double>> span;   // first / last time of a cell
    {
        std::ifstream fin(prediction_file);
        if (!fin) throw std::invalid_argument("prediction file not found: " + prediction_file);
        std::string line, last;
        bool skip = true;
        while (std::getline(fin, line)) {
            if (!skip) {
                const auto p = split(line, ",");
                if (p.size() < 19) continue;
                std::vector<double> m14;
                for (int k = 5; k < 19; ++k) m14.push_back(std::stod(p[k]));
                marginals.push_back(m14);
                point_cell.push_back(p[0]);
                const double t = std::stod(p[2]);
                point_time.push_back(t);
                CS.add(0.0, gauss_from_marginal14(m14.data()));
                if (p[0] != last) { cell_ids.push_back(p[0]); parent_ids.push_back(p[1]); span[p[0]] = {t, t}; }
                span[p[0]].second = t;
                last = p[0];
            }
            if (line.rfind("cell_id", 0) == 0) skip = false;
        }
    }
    const LineageLookup L(cell_ids, parent_ids);
    std::unordered_map<std::string, int> cidx;
    for (size_t c = 0; c < cell_ids.size(); ++c) cidx[cell_ids[c]] = (int)c;
    std::ifstream fin(joint_file);
    if (!fin) throw std::invalid_argument("joint file not found: " + joint_file);
    std::vector<std::string> col_cell;
    std::vector<double> col_time;
    std::string line;
    bool skip = true;
    int64_t i = 0;
    std::vector<char> related;
    std::string related_for;
    while (std::getline(fin, line)) {
        if (!skip) {
            const auto p = split(line, ",");
            if (p.size() < 3 + 44 * col_cell.size()) continue;
            const std::string& row_cell = p[0];
            const double row_time = std::stod(p[2]);
            if (row_cell != related_for) {
                related.assign(cell_ids.size(), 0);
                const auto it = cidx.find(row_cell);
                if (it != cidx.end()) for (int u : L.related(it->second)) related[u] = 1;
                related_for = row_cell;
            }
            for (size_t j = 0; j < col_cell.size(); ++j) {
                double dt = col_time[j] - row_time;
                if (normalize_time) { const auto& s = span[row_cell]; dt /= (s.second - s.first); }
                const size_t base = 3 + 44 * j;
                if (!p[base].empty()) {
                    double v[44];
                    for (int k = 0; k < 44; ++k) v[k] = std::stod(p[base + k]);
                    CS.add(dt, gauss_from_joint44(v));
                } else {
                    const auto it = cidx.find(col_cell[j]);
                    if (it != cidx.end() && related[it->second] && (int64_t)j > i)
                        CS.add(dt, gauss_from_two_marginals(marginals[j].data(), marginals[i].data()));
                }
            }
            ++i;
        }
        if (line.rfind("cell_id", 0) == 0) {
            skip = false;
            const auto p = split(line, ",");
            for (size_t k = 3; k < p.size(); ++k)
                if (!p[k].empty()) {
                    const auto e = split(p[k], "_");
                    col_cell.push_back(e[0]);
                    col_time.push_back(std::stod(e[1]));
                }
        }
    }
}

}  // namespace ggp
