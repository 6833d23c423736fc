// host/ggp_params.hpp — the parameter-bounds file and the parameter tables written to every output file.
// Format and semantics follow the reference's Parameters.h (Parameter :8-86, Parameter_set :89-350) and
// predictions.h:505-519 (file-name code); the code is new.
//   name = init                      fixed
//   name = init, step                free
//   name = init, step, lower, upper  bound
#pragma once
#include <fstream>
#include <numeric>
#include <ostream>

#include "ggp_util.hpp"

namespace ggp {

struct Parameter {
    std::string name;
    bool fixed = false, bound = false, free = false, set = false, minimized = false;
    double init = 0, step = 0, lower = 0, upper = HUGE_VAL, final_value = 0;   // reference defaults: lower 0, upper +inf
};

inline const char* const kParamNames[11] = {"mean_lambda", "gamma_lambda", "var_lambda", "mean_q", "gamma_q", "var_q",
                                            "beta", "var_x", "var_g", "var_dx", "var_dg"};

class ParameterSet {
public:
    std::vector<Parameter> all;

    explicit ParameterSet(const std::string& filename, std::ostream* log = nullptr) {
        all.resize(11);
        for (int i = 0; i < 11; ++i) all[i].name = kParamNames[i];
        std::ifstream fin(filename);
        std::string line;
        while (std::getline(fin, line)) {
            if (line.empty() || line[0] == '#') continue;
            const auto parts = split(line, "=");
            const std::string key = trim(parts[0]);
            for (Parameter& p : all) {
                if (p.name != key) continue;
                try {
                    if (parts.size() < 2) throw std::invalid_argument("Invalide number of arguments");
                    auto vals = split(parts[1], ",");
                    for (auto& v : vals) v = trim(v);
                    if (vals.size() == 4) {
                        p.init = to_double_no_nan(vals[0]); p.step = to_double_no_nan(vals[1]);
                        p.lower = to_double_no_nan(vals[2]); p.upper = to_double_no_nan(vals[3]);
                        p.bound = true;
                    } else if (vals.size() == 1) {
                        p.init = to_double_no_nan(vals[0]);
                        p.fixed = true;
                    } else if (vals.size() == 2) {
                        p.init = to_double_no_nan(vals[0]); p.step = to_double_no_nan(vals[1]);
                        p.free = true;
                    } else {
                        throw std::invalid_argument("Invalide number of arguments");
                    }
                    p.set = true;
                } catch (std::exception& e) {
                    if (log) *log << "(set_paramter) ERROR: Parameter settings of '" << key << "' cannot be processed (" << e.what() << ")" << std::endl;
                    throw;
                }
            }
        }
    }

    void check_if_complete(std::ostream& log) const {
        for (const Parameter& p : all)
            if (!p.set) {
                log << "(check_if_complete) ERROR: Parameter " << p.name << " not found in parameter file\n";
                throw std::invalid_argument("Invalide argument");
            }
    }
    bool has_nonfixed() const {
        for (const Parameter& p : all) if (!p.fixed) return true;
        return false;
    }
    void set_final(const std::vector<double>& v) {
        for (size_t i = 0; i < all.size(); ++i) { all[i].final_value = v[i]; all[i].minimized = true; }
    }
    std::vector<double> get_final() const {
        std::vector<double> v;
        for (const Parameter& p : all) v.push_back(p.minimized ? p.final_value : p.init);
        return v;
    }
    std::vector<int> non_fixed() const {
        std::vector<int> idx;
        for (size_t i = 0; i < all.size(); ++i) if (!all[i].fixed) idx.push_back((int)i);
        return idx;
    }
    // "_f<indices of free>_b<indices of bound>" used in every output file name
    std::string code() const {
        std::string c = "_f";
        for (size_t i = 0; i < all.size(); ++i) if (!all[i].bound && !all[i].fixed) c += std::to_string(i);
        c += "_b";
        for (size_t i = 0; i < all.size(); ++i) if (all[i].bound) c += std::to_string(i);
        return c;
    }
    // the table at the top of every csv output
    void to_csv(std::ostream& f) const {
        f << "no,name,type,init,step,lower_bound,upper_bound,final\n";
        for (size_t i = 0; i < all.size(); ++i) {
            const Parameter& p = all[i];
            f << i << ",";
            if (p.fixed) f << p.name << ",fixed," << p.init << ", , , ,";
            else if (p.bound) f << p.name << ",bound," << p.init << "," << p.step << "," << p.lower << "," << p.upper << ",";
            else f << p.name << ",free," << p.init << "," << p.step << ", , ,";
            if (p.minimized) f << p.final_value;
            f << "\n";
        }
    }
    void to_csv(const std::string& file, std::ios_base::openmode mode = std::ios_base::out) const {
        std::ofstream f(file, mode);
        to_csv(f);
    }
};

// log-file table (Parameters.h:305-348)
inline std::ostream& operator<<(std::ostream& os, const ParameterSet& ps) {
    const int w[7] = {4, 15, 8, 20, 20, 15, 15};
    os << pad("No", w[0]) << pad("Name", w[1]) << pad("Type", w[2]) << pad("Init", w[3]) << pad("Step", w[4]) << pad("Bounds", w[5]) << "\n";
    os << std::string(std::accumulate(w, w + 7, 0), '_') << "\n";
    for (size_t i = 0; i < ps.all.size(); ++i) {
        const Parameter& p = ps.all[i];
        os << pad(std::to_string(i) + ":", w[0]) << pad(p.name, w[1]);
        if (p.fixed) os << pad("(fixed)", w[2]) << pad(p.init, w[3]) << pad("", w[4] + w[5] + w[6]);
        else if (p.bound) os << pad("(bound)", w[2]) << pad(p.init, w[3]) << pad(p.step, w[4]) << pad(p.lower, w[5]) << pad(p.upper, w[6]);
        else os << pad("(free)", w[2]) << pad(p.init, w[3]) << pad(p.step, w[4]) << pad("", w[5] + w[6]);
        if (p.minimized && !p.fixed) os << " -> " << p.final_value;
        os << "\n";
    }
    return os;
}

}  // namespace ggp
