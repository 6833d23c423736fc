// ggp_bridge.h — the reference-side binding of libggp_b200.so: what a maintainer of bjks/gfp_gaussian_process adds to
// keep the reference's own host code (moma_input.h parsing, genealogy, nlopt, output writers, Eigen types) and move
// only the tree recursion to the GPU.  C++17, header only; include it AFTER the reference's likelihood.h (it uses
// MOMAdata, Joint_vector, Gaussian and the logging globals _iteration / _save_ll / _file_iteration / _print_level /
// _file_log declared there and in main.cpp:9).  INTEGRATION.md walks through the four call sites.
//
// Built and run for real: oracle/ref_bridge.cpp compiles this header against the reference's unmodified headers
// (oracle/_ref/libggp_ref_bridge.so, linked to libggp_b200.so) and tests/test_gpu_bridge.py drives it on the GPU next to
// the reference's own CPU passes on the same std::vector<MOMAdata>.
#ifndef GGP_BRIDGE_H
#define GGP_BRIDGE_H
#include <cstdint>
#include <iomanip>
#include <stdexcept>
#include <string>
#include <vector>

#include "ggp_b200.h"

// std::vector<MOMAdata> -> SoA forest handle; runs once after build_cell_genealogy + init_cells (main.cpp:408-411, :434-435)
struct GgpBridge {
    ggp_forest* h = nullptr;
    std::vector<MOMAdata>* cells = nullptr;
    std::vector<int64_t> off;
    std::vector<int32_t> parent, d1, d2, seg;
    std::vector<double> t, x, g;
    std::vector<double> carry;   // [n_roots][16]: the roots' persistent MOMAdata::cov (predictions.h:64-78);
                                 // zero after get_segment (moma_input.h:587), then chained through every evaluation
    int64_t n_ctp = 0;

    explicit GgpBridge(std::vector<MOMAdata>& cells_, int device = 0) : cells(&cells_) {
        const MOMAdata* base = cells_.data();   // genealogy pointers point into this vector (moma_input.h:133-139)
        off.push_back(0);
        const MOMAdata *root = nullptr, *leaf = nullptr;
        for (const MOMAdata& c : cells_) {
            for (long i = 0; i < c.time.size(); ++i) {
                t.push_back(c.time(i)); x.push_back(c.log_length(i)); g.push_back(c.fp(i)); seg.push_back(c.segment[i]);
            }
            off.push_back((int64_t)t.size());
            parent.push_back(c.parent ? int32_t(c.parent - base) : -1);
            d1.push_back(c.daughter1 ? int32_t(c.daughter1 - base) : -1);
            d2.push_back(c.daughter2 ? int32_t(c.daughter2 - base) : -1);
            if (c.is_root() && !root) root = &c;
            if (c.is_leaf() && !leaf) leaf = &c;
        }
        n_ctp = (int64_t)t.size();
        ggp_forest_desc d{};
        d.n_cells = (int64_t)cells_.size(); d.n_ctp = n_ctp;
        d.cell_offset = off.data(); d.parent = parent.data(); d.daughter1 = d1.data(); d.daughter2 = d2.data();
        d.time = t.data(); d.log_length = x.data(); d.fp = g.data(); d.segment = seg.data();
        d.noise_model = cells_[0].noise_model == "scaled" ? GGP_NOISE_SCALED : GGP_NOISE_CONST;
        d.division_model = cells_[0].cell_division_model == "binomial" ? GGP_DIVISION_BINOMIAL : GGP_DIVISION_GAUSS;
        d.fp_auto = cells_[0].fp_auto;
        // init_cells(cells) has already run: its statistics live in the roots (forward) and leafs (backward), moma_input.h:697-703, :728-734
        d.init_f[0] = root->mean_init_forward(0); d.init_f[1] = root->mean_init_forward(1);
        d.init_f[2] = root->cov_init_forward(0, 0); d.init_f[3] = root->cov_init_forward(1, 1);
        d.init_r[0] = leaf->mean_init_backward(0); d.init_r[1] = leaf->mean_init_backward(1);
        d.init_r[2] = leaf->cov_init_backward(0, 0); d.init_r[3] = leaf->cov_init_backward(1, 1);
        d.compute_init = 0;
        d.device = device;
        if (ggp_forest_create(&d, &h) != GGP_OK) throw std::invalid_argument(ggp_last_error());
        carry.assign((size_t)ggp_forest_n_roots(h) * 16, 0.0);
    }
    GgpBridge(const GgpBridge&) = delete;
    GgpBridge& operator=(const GgpBridge&) = delete;
    ~GgpBridge() { ggp_forest_destroy(h); }
};

// total_likelihood (likelihood.h:125-159), nlopt vfunc signature; `c` carries the bridge instead of the root pointers.
// Logging, _iteration, the exp() of log-space parameters (:161-167) and the sign flip stay on the host as they are.
inline double ggp_bridge_total_likelihood(const std::vector<double>& params_vec, std::vector<double>& /*grad*/, void* c) {
    GgpBridge* B = static_cast<GgpBridge*>(c);
    double tl = 0;
    ggp_nan_info nan{-1, -1};
    const int rc = ggp_loglik(B->h, params_vec.data(), 1, B->carry.data(), &tl, nullptr, &nan);
    if (rc == GGP_ERR_NAN) {   // the diagnostics of likelihood.h:71-93 for the first failing point in depth-first order
        const MOMAdata& cell = (*B->cells)[(size_t)nan.cell];
        if (_save_ll) {
            _file_iteration << _iteration + 1 << ",";
            for (size_t i = 0; i < params_vec.size(); ++i) _file_iteration << std::setprecision(20) << params_vec[i] << ",";
            _file_iteration << std::setprecision(30) << tl << std::setprecision(15) << "\n";
        }
        _file_log << "\n(sc_likelihood) ERROR: Log likelihood is Nan\n";
        _file_log << "____________________________________________\n";
        _file_log << "cell_id: " << cell.cell_id << ", at time " << cell.time(nan.t_index) << "\n";
        _file_log << _iteration + 1 << ": ";
        for (size_t i = 0; i < params_vec.size(); ++i) _file_log << params_vec[i] << ", ";
        _file_log << "ll=" << std::setprecision(10) << tl << "\n";
        _file_iteration.close();
        throw std::domain_error("Likelihood is Nan");
    }
    if (rc != GGP_OK) throw std::runtime_error(ggp_last_error());
    ++_iteration;
    if (_save_ll) {   // likelihood.h:140-148
        _file_iteration << _iteration << ",";
        for (size_t i = 0; i < params_vec.size(); ++i) _file_iteration << std::setprecision(20) << params_vec[i] << ",";
        _file_iteration << std::setprecision(30) << tl << std::setprecision(15) << "\n";
    }
    if (_print_level > 0) {   // likelihood.h:150-157
        std::cout << _iteration << ": ";
        for (size_t i = 0; i < params_vec.size(); ++i) std::cout << std::setprecision(20) << params_vec[i] << ", ";
        std::cout << "ll=" << std::setprecision(30) << tl << std::setprecision(15) << "\n";
    }
    return -tl;
}

// the +log-likelihood overload the scan and the Hessian use (likelihood.h:170-174)
inline double ggp_bridge_total_likelihood(const std::vector<double>& params_vec, GgpBridge& B) {
    std::vector<double> g;
    return -ggp_bridge_total_likelihood(params_vec, g, &B);
}

// run_bound_1dscan's loop (main.cpp:102-108) / num_hessian_ll's stencil (likelihood.h:211-258) as ONE call: the vectors are
// evaluated as if one after the other (the root chain is sequential on the device), so the numbers equal the loop's
inline std::vector<double> ggp_bridge_total_likelihood_batch(const std::vector<std::vector<double>>& vecs, GgpBridge& B) {
    std::vector<double> P, out(vecs.size());
    for (const std::vector<double>& v : vecs) P.insert(P.end(), v.begin(), v.end());
    std::vector<ggp_nan_info> nan(vecs.size());
    const int rc = ggp_loglik(B.h, P.data(), (int32_t)vecs.size(), B.carry.data(), out.data(), nullptr, nan.data());
    if (rc == GGP_ERR_NAN) throw std::domain_error("Likelihood is Nan");
    if (rc != GGP_OK) throw std::runtime_error(ggp_last_error());
    _iteration += (int)vecs.size();
    return out;
}

inline std::vector<double> ggp_bridge_flat_params(const std::vector<std::vector<double>>& params_vecs) {
    std::vector<double> P;
    for (const std::vector<double>& v : params_vecs) P.insert(P.end(), v.begin(), v.end());
    return P;
}

// prediction_forward + prediction_backward + combine_predictions (main.cpp:132-140): fills the cells' prediction vectors
// so that write_predictions_to_file (predictions.h:563-601) runs unchanged
inline void ggp_bridge_predictions(GgpBridge& B, const std::vector<std::vector<double>>& params_vecs) {
    const std::vector<double> P = ggp_bridge_flat_params(params_vecs);
    const size_t n = (size_t)B.n_ctp;
    std::vector<double> fwd(20 * n), bwd(20 * n), comb(20 * n);
    if (ggp_predict(B.h, P.data(), (int32_t)params_vecs.size(), fwd.data(), bwd.data(), comb.data()) != GGP_OK)
        throw std::runtime_error(ggp_last_error());
    auto mean_of = [](const double* r) { return Eigen::Vector4d(r[0], r[1], r[2], r[3]); };
    auto cov_of = [](const double* r) {
        Eigen::Matrix4d C;
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) C(i, j) = r[4 + 4 * i + j];
        return C;
    };
    size_t k = 0;
    for (MOMAdata& c : *B.cells) {
        c.mean_forward.clear(); c.cov_forward.clear();
        c.mean_backward.clear(); c.cov_backward.clear();
        c.mean_prediction.clear(); c.cov_prediction.clear();
        for (long i = 0; i < c.time.size(); ++i, ++k) {
            c.mean_forward.push_back(mean_of(&fwd[20 * k])); c.cov_forward.push_back(cov_of(&fwd[20 * k]));
            c.mean_backward.push_back(mean_of(&bwd[20 * k])); c.cov_backward.push_back(cov_of(&bwd[20 * k]));
            c.mean_prediction.push_back(mean_of(&comb[20 * k])); c.cov_prediction.push_back(cov_of(&comb[20 * k]));
        }
    }
}

// collect_joint_distributions (correlation_tree.h:629-648): same dense text, rendered by the reference's own Joint_vector
// writer from the sparse records, one block of rows at a time.  Requires ggp_bridge_predictions with the same parameters.
inline void ggp_bridge_collect_joint_distributions(GgpBridge& B, const std::vector<std::vector<double>>& params_vecs,
                                                   std::ostream& out, double tolerance_joint, int64_t rows_per_call = 4096) {
    const std::vector<double> P = ggp_bridge_flat_params(params_vecs);
    std::vector<MOMAdata>& cells = *B.cells;
    Joint_vector joint_vector(cells);
    out << ",";
    joint_vector.write_column_indices(out);
    out << "\n";
    std::vector<int64_t> row, col;
    std::vector<double> rec;
    size_t cell = 0;
    long in_cell = 0;
    for (int64_t r0 = 0; r0 < B.n_ctp; r0 += rows_per_call) {
        const int64_t r1 = r0 + rows_per_call < B.n_ctp ? r0 + rows_per_call : B.n_ctp;
        int64_t n = 0;
        if (ggp_joints(B.h, P.data(), (int32_t)params_vecs.size(), tolerance_joint, r0, r1, 0, &n, nullptr, nullptr, nullptr) != GGP_OK)
            throw std::runtime_error(ggp_last_error());
        row.resize((size_t)n); col.resize((size_t)n); rec.resize((size_t)n * 44);
        if (n && ggp_joints(B.h, P.data(), (int32_t)params_vecs.size(), tolerance_joint, r0, r1, n, &n, row.data(), col.data(), rec.data()) != GGP_OK)
            throw std::runtime_error(ggp_last_error());
        size_t k = 0;   // records are sorted by (row, col)
        for (int64_t r = r0; r < r1; ++r) {
            while (in_cell >= cells[cell].time.size()) { ++cell; in_cell = 0; }
            joint_vector.clear();
            for (; k < (size_t)n && row[k] == r; ++k) {
                const double* v = &rec[44 * k];
                Eigen::VectorXd m(8);
                Eigen::MatrixXd C(8, 8);
                for (int i = 0; i < 8; ++i) m(i) = v[i];
                for (int i = 0, q = 8; i < 8; ++i)
                    for (int j = i; j < 8; ++j, ++q) C(i, j) = C(j, i) = v[q];
                joint_vector.joints[(size_t)col[k]] = Gaussian(m, C);
                joint_vector.is_set[(size_t)col[k]] = true;
            }
            out << cells[cell].cell_id << "," << cells[cell].parent_id << "," << cells[cell].time[in_cell];
            joint_vector.write(out);
            out << "\n";
            ++in_cell;
        }
    }
}

#endif
