// oracle/ref_wrappers.cpp — builds oracle/_ref/libggp_ref_wrappers.so from the REFERENCE'S OWN WRAPPER SOURCES, in place and
// unmodified:  /root/reference/src/{likelihood.h, correlation_tree.h, predictions.h, moma_input.h, Gaussians.h,
// mean_cov_model.h, Parameters.h, CSVconfig.h, utils.h} (pulled in by the one #include below) + Faddeeva.cc (oracle/Makefile).
// Nothing of the reference is copied into this repository.  Eigen, which those headers need, does not exist in this image:
// oracle/eigen_shim/ supplies the subset they use (its header states exactly which Eigen evaluation rules it restates).
//
// TEST INFRASTRUCTURE ONLY: tests/test_ref_wrappers.py holds oracle/ggp_oracle.cpp (the restatement the GPU path is checked
// against) to this build bit for bit - log-likelihood in fresh and carry mode, forward / backward / combined predictions,
// joints - and tools/make_wrapper_golden.py stores its outputs as fixtures for machines without /root/reference.
// Never linked into the product.
#include <cmath>
#include <cstring>
#include <fstream>
#include <limits>
#include <sstream>
#include <string>
#include <vector>

std::ofstream _file_log;   // main.cpp:9: the global every header of the reference logs to

#include "/root/reference/src/likelihood.h"

namespace {

struct RefForest {
    std::vector<MOMAdata> cells;
    std::vector<long> offset;
};

std::vector<double> pvec(const double* p) { return std::vector<double>(p, p + 11); }
std::vector<std::vector<double>> pvecs(const double* p, int n_seg) {
    std::vector<std::vector<double>> v;
    for (int s = 0; s < n_seg; ++s) v.push_back(pvec(p + 11 * s));
    return v;
}

void per_cell_recr(const std::vector<double>& p, MOMAdata* cell, const MOMAdata* base, double* out) {
    // the traversal of likelihood_recr (likelihood.h:110-122) with a running sum per cell instead of one for the forest
    if (cell == nullptr) return;
    double tl = 0;
    sc_likelihood(p, *cell, tl);
    out[cell - base] = tl;
    per_cell_recr(p, cell->daughter1, base, out);
    per_cell_recr(p, cell->daughter2, base, out);
}

void store4(double* dst, const Eigen::Vector4d& v) { for (int i = 0; i < 4; ++i) dst[i] = v(i); }
void store16(double* dst, const Eigen::Matrix4d& m) {
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) dst[4 * i + j] = m(i, j);
}

}  // namespace

extern "C" {

// std::vector<MOMAdata> as read_data would leave it (moma_input.h:401-527), then build_cell_genealogy (:125) and init_cells (:738)
void* ggp_refw_create(long n_cells, const long* cell_offset, const int* parent, const double* time, const double* log_length,
                      const double* fp, const int* segment, int noise_scaled, int division_binomial, double fp_auto) {
    if (!_file_log.is_open()) _file_log.open("/dev/null");
    _save_ll = false;
    _print_level = 0;
    RefForest* f = new RefForest();
    f->offset.assign(cell_offset, cell_offset + n_cells + 1);
    f->cells.resize((size_t)n_cells);
    for (long c = 0; c < n_cells; ++c) {
        MOMAdata& m = f->cells[(size_t)c];
        m.cell_id = "c" + std::to_string(c);
        m.parent_id = parent[c] >= 0 ? "c" + std::to_string(parent[c]) : std::string("none");
        m.noise_model = noise_scaled ? "scaled" : "const";
        m.cell_division_model = division_binomial ? "binomial" : "gauss";
        m.fp_auto = fp_auto;
        for (long k = cell_offset[c]; k < cell_offset[c + 1]; ++k) {
            append_vec(m.time, time[k]);
            append_vec(m.log_length, log_length[k]);
            append_vec(m.fp, fp[k]);
            append_vec(m.segment, segment ? segment[k] : 0);
        }
    }
    build_cell_genealogy(f->cells);
    init_cells(f->cells);
    return f;
}

void ggp_refw_destroy(void* h) { delete (RefForest*)h; }

// daughter1 / daughter2 as build_cell_genealogy assigned them (cell indices or -1)
void ggp_refw_get_daughters(void* h, int* d1, int* d2) {
    RefForest* f = (RefForest*)h;
    const MOMAdata* base = f->cells.data();
    for (size_t c = 0; c < f->cells.size(); ++c) {
        d1[c] = f->cells[c].daughter1 ? (int)(f->cells[c].daughter1 - base) : -1;
        d2[c] = f->cells[c].daughter2 ? (int)(f->cells[c].daughter2 - base) : -1;
    }
}

// the statistics init_cells_f / init_cells_r put into the roots / leafs (moma_input.h:675-735); NaN where the forest has none
void ggp_refw_get_init(void* h, double* init_f4, double* init_r4) {
    RefForest* f = (RefForest*)h;
    for (int i = 0; i < 4; ++i) init_f4[i] = init_r4[i] = std::numeric_limits<double>::quiet_NaN();
    for (MOMAdata& c : f->cells) {
        if (c.is_root()) {
            init_f4[0] = c.mean_init_forward(0); init_f4[1] = c.mean_init_forward(1);
            init_f4[2] = c.cov_init_forward(0, 0); init_f4[3] = c.cov_init_forward(1, 1);
        }
        if (c.is_leaf()) {
            init_r4[0] = c.mean_init_backward(0); init_r4[1] = c.mean_init_backward(1);
            init_r4[2] = c.cov_init_backward(0, 0); init_r4[3] = c.cov_init_backward(1, 1);
        }
    }
}

// MOMAdata::mean / cov of every cell: the persistent state that makes evaluations history dependent
void ggp_refw_get_state(void* h, double* mean, double* cov) {
    RefForest* f = (RefForest*)h;
    for (size_t c = 0; c < f->cells.size(); ++c)
        for (int i = 0; i < 4; ++i) {
            mean[4 * c + i] = f->cells[c].mean(i);
            for (int j = 0; j < 4; ++j) cov[16 * c + 4 * i + j] = f->cells[c].cov(i, j);
        }
}
void ggp_refw_set_state(void* h, const double* mean, const double* cov) {
    RefForest* f = (RefForest*)h;
    for (size_t c = 0; c < f->cells.size(); ++c)
        for (int i = 0; i < 4; ++i) {
            f->cells[c].mean(i) = mean[4 * c + i];
            for (int j = 0; j < 4; ++j) f->cells[c].cov(i, j) = cov[16 * c + 4 * i + j];
        }
}

// total_likelihood(params_vec, cells) (likelihood.h:170-174): +log-likelihood; NaN if the reference threw "Likelihood is Nan"
double ggp_refw_total_loglik(void* h, const double* p11) {
    RefForest* f = (RefForest*)h;
    try {
        return total_likelihood(pvec(p11), f->cells);
    } catch (const std::domain_error&) {
        return std::numeric_limits<double>::quiet_NaN();
    }
}

// sc_likelihood (likelihood.h:36-103) over the forest in the order of likelihood_recr, one sum per cell
int ggp_refw_per_cell_loglik(void* h, const double* p11, double* per_cell) {
    RefForest* f = (RefForest*)h;
    const std::vector<double> p = pvec(p11);
    try {
        for (MOMAdata& c : f->cells)
            if (c.is_root()) per_cell_recr(p, &c, f->cells.data(), per_cell);
    } catch (const std::domain_error&) {
        return 1;
    }
    return 0;
}

// prediction_forward, prediction_backward, combine_predictions as run_prediction_segments calls them (main.cpp:132-140)
void ggp_refw_predict(void* h, const double* params, int n_seg, double* mf, double* cf, double* mb, double* cb, double* mp,
                      double* cp) {
    RefForest* f = (RefForest*)h;
    const std::vector<std::vector<double>> pv = pvecs(params, n_seg);
    for (MOMAdata& c : f->cells) {
        c.mean_forward.clear(); c.cov_forward.clear();
        c.mean_backward.clear(); c.cov_backward.clear();
        c.mean_prediction.clear(); c.cov_prediction.clear();
    }
    prediction_forward(pv, f->cells);
    prediction_backward(pv, f->cells);
    combine_predictions(f->cells, pv);
    for (size_t c = 0; c < f->cells.size(); ++c) {
        const MOMAdata& m = f->cells[c];
        for (size_t t = 0; t < m.mean_forward.size(); ++t) {
            const long k = f->offset[c] + (long)t;
            store4(mf + 4 * k, m.mean_forward[t]); store16(cf + 16 * k, m.cov_forward[t]);
            store4(mb + 4 * k, m.mean_backward[t]); store16(cb + 16 * k, m.cov_backward[t]);
            store4(mp + 4 * k, m.mean_prediction[t]); store16(cp + 16 * k, m.cov_prediction[t]);
        }
    }
}

// collect_joint_distributions (correlation_tree.h:629-648) into a string stream at 17 significant digits (the reference
// streams at the stream's precision; 17 digits round-trip a double), parsed back into sparse records
// (row ctp, col ctp, 8 means + 36 upper-triangular covariances).  Requires ggp_refw_predict with the same parameters.
long ggp_refw_joints(void* h, const double* params, int n_seg, double tol, long cap, long* row, long* col, double* rec44) {
    RefForest* f = (RefForest*)h;
    std::ostringstream out;
    out.precision(17);
    collect_joint_distributions(pvecs(params, n_seg), f->cells, out, tol);
    std::istringstream in(out.str());
    std::string line;
    std::getline(in, line);   // column indices
    long n = 0, r = 0;
    while (std::getline(in, line)) {
        size_t pos = 0;
        for (int k = 0; k < 3; ++k) pos = line.find(',', pos) + 1;   // cell_id, parent_id, time
        --pos;                                                       // at the comma in front of column 0
        for (long c = 0; pos < line.size(); ++c) {
            // a column is 44 comma-prefixed fields, all empty or all numbers
            if (pos + 1 < line.size() && line[pos + 1] != ',') {
                double v[44];
                for (int k = 0; k < 44; ++k) {
                    ++pos;   // the comma
                    size_t end = line.find(',', pos);
                    if (end == std::string::npos) end = line.size();
                    v[k] = std::stod(line.substr(pos, end - pos));
                    pos = end;
                }
                if (n < cap) {
                    row[n] = r; col[n] = c;
                    std::memcpy(rec44 + 44 * n, v, sizeof v);
                }
                ++n;
            } else {
                pos += 44;
            }
        }
        ++r;
    }
    return n;
}

}  // extern "C"
