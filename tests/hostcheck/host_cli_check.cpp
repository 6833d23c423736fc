// tests/hostcheck/host_cli_check.cpp — exposes the host-side pieces of the gfp_gaussian command line (host/*.hpp:
// readers, segment slicing, genealogy, parameter tables, Nelder-Mead) to the CPU tests.  Test infrastructure.
#include <cstring>
#include <fstream>
#include <sstream>

#include "../../host/ggp_data.hpp"
#include "../../host/ggp_neldermead.hpp"
#include "../../host/ggp_params.hpp"

using namespace ggp;

static LineageTable g_table;
static std::string g_text;
static double g_nm_nan_above = HUGE_VAL;   // hcli_neldermead's objective is NaN where a coordinate exceeds this

extern "C" {

// reads infile with csv_config (may be ""), optionally keeps one segment (-1: all), builds the genealogy;
// returns n_cells (or -1 and the message in hcli_text())
long hcli_load(const char* infile, const char* config, int segment) {
    std::ostringstream log;
    try {
        CsvConfig cfg(config, &log);
        LineageTable T = is_binary_forest(infile) ? read_binary_forest(infile, cfg, "scaled", "binomial", log)
                                                  : read_data(infile, cfg, "scaled", "binomial", log);
        segment_indices(T, log);
        if (segment >= 0) T = get_segment(T, segment);
        build_genealogy(T, log);
        g_table = T;
        g_text = log.str();
        return (long)g_table.n_cells();
    } catch (std::exception& e) {
        g_text = log.str() + e.what();
        return -1;
    }
}
long hcli_n_ctp() { return (long)g_table.n_ctp(); }
void hcli_copy(long long* offset, int* parent, int* d1, int* d2, double* time, double* x, double* g, int* seg) {
    std::copy(g_table.offset.begin(), g_table.offset.end(), offset);
    std::copy(g_table.parent.begin(), g_table.parent.end(), parent);
    std::copy(g_table.daughter1.begin(), g_table.daughter1.end(), d1);
    std::copy(g_table.daughter2.begin(), g_table.daughter2.end(), d2);
    std::copy(g_table.time.begin(), g_table.time.end(), time);
    std::copy(g_table.log_length.begin(), g_table.log_length.end(), x);
    std::copy(g_table.fp.begin(), g_table.fp.end(), g);
    std::copy(g_table.segment.begin(), g_table.segment.end(), seg);
}
const char* hcli_cell_id(long c) { return g_table.cell_id[c].c_str(); }
const char* hcli_text() { return g_text.c_str(); }

// parameter file -> the csv table + log table + file-name code; returns 0 or -1
int hcli_params(const char* file, double* final_or_null) {
    std::ostringstream out, log;
    try {
        ParameterSet ps(file, &log);
        ps.check_if_complete(log);
        if (final_or_null) ps.set_final(std::vector<double>(final_or_null, final_or_null + 11));
        ps.to_csv(out);
        out << "CODE " << ps.code() << "\n" << ps;
        g_text = out.str();
        return 0;
    } catch (std::exception& e) {
        g_text = log.str() + e.what();
        return -1;
    }
}

// Nelder-Mead on f(x) = sum_i w_i (x_i - c_i)^2 + rosenbrock coupling; logs every recorded evaluation into hcli_text()
// as "x0 x1 ... f" lines.  Returns the number of recorded evaluations; launches and the optimum through the out arrays.
// speculate: 0 = sequential, 1 = this iteration's candidates in one launch, k > 1 = up to k points per speculative launch
int hcli_neldermead(int n, const double* x0, const double* lb, const double* ub, const double* step, double ftol, int speculate,
                    double* x_out, double* f_out, int* launches_out) {
    std::ostringstream rec;
    rec.precision(17);
    auto f = [&](const std::vector<double>& x) {
        for (double xi : x) if (xi > g_nm_nan_above) return std::nan("");
        double s = 0;
        for (int i = 0; i + 1 < n; ++i) s += 100 * (x[i + 1] - x[i] * x[i]) * (x[i + 1] - x[i] * x[i]) + (1 - x[i]) * (1 - x[i]);
        return s;
    };
    BatchObjective obj;
    auto note = [&](const std::vector<double>& x, double v) {
        for (double xi : x) rec << xi << " ";
        rec << v << "\n";
    };
    obj.evaluate = [&](const std::vector<std::vector<double>>& X, bool record) {
        std::vector<double> v;
        for (const auto& x : X) {
            v.push_back(f(x));
            if (record) note(x, v.back());
        }
        return v;
    };
    obj.commit = note;
    const NelderMeadResult R = nelder_mead(obj, std::vector<double>(x0, x0 + n), std::vector<double>(lb, lb + n), std::vector<double>(ub, ub + n),
                                           std::vector<double>(step, step + n), ftol, speculate != 0, 20000, speculate <= 1 ? 4 : speculate);
    std::copy(R.x.begin(), R.x.end(), x_out);
    *f_out = R.f;
    *launches_out = R.launches;
    g_text = rec.str() + R.reason;
    return R.evaluations;
}

void hcli_nm_nan_above(double v) { g_nm_nan_above = v; }

void hcli_arange(double a, double b, double s, double* out, int* n) {
    const auto v = arange(a, b, s);
    *n = (int)v.size();
    if (out) std::copy(v.begin(), v.end(), out);
}
}
