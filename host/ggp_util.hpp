// host/ggp_util.hpp — small string / path helpers of the gfp_gaussian command line.
// Behaviour follows the reference's utils.h (split_string_at :77, trim :88, pad_str :30-44, arange :96-103,
// default_out_dir :106, out_dir :125, add_segment_to_filename :139, file_base :149) because file names, log
// layout and the sampling grid of -s are part of the drop-in contract; the code is new.
#pragma once
#include <cmath>
#include <filesystem>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace ggp {

using Args = std::map<std::string, std::string>;

inline std::vector<std::string> split(const std::string& s, const std::string& delim = ",") {
    std::vector<std::string> out;
    size_t from = 0;
    for (;;) {
        const size_t at = s.find(delim, from);
        if (at == std::string::npos) break;
        out.push_back(s.substr(from, at - from));
        from = at + delim.size();
    }
    out.push_back(s.substr(from));
    return out;
}

// strips `c` on the left and blanks on the right (the reference trims the right end by ' ' whatever `c` is)
inline std::string trim(const std::string& s, char c = ' ') {
    const size_t first = s.find_first_not_of(c);
    if (first == std::string::npos) return s;
    const size_t last = s.find_last_not_of(' ');
    return s.substr(first, last - first + 1);
}

inline std::string trim_all(std::string s) {
    for (char c : {' ', '\t', '\n', '\v', '\f', '\r'}) s = trim(s, c);
    return s;
}

inline std::string pad(std::string s, size_t width) {
    if (s.size() < width) s.append(width - s.size(), ' ');
    return s;
}

inline std::string pad(double d, size_t width) {
    std::ostringstream b;
    b << d;
    return pad(b.str(), width);
}

inline double to_double_no_nan(const std::string& s) {
    const double d = std::stod(s);
    if (std::isnan(d)) throw std::invalid_argument("String is Nan");
    return d;
}

inline bool to_bool(const std::string& s) {
    if (s == "True" || s == "true" || s == "TRUE" || s == "1") return true;
    if (s == "False" || s == "false" || s == "FALSE" || s == "0") return false;
    throw std::invalid_argument("Invalid argument");
}

// numpy-like arange by repeated addition (NOT start + i*step): the rounding of the grid is part of -s
inline std::vector<double> arange(double start, double stop, double step) {
    std::vector<double> v;
    for (double x = start; x < stop; x += step) v.push_back(x);
    return v;
}

inline std::string file_base(const std::string& infile) {
    const auto path = split(infile, "/");
    const auto parts = split(path.back(), ".");
    std::string b;
    for (size_t i = 0; i + 1 < parts.size(); ++i) {
        if (i) b += '.';
        b += parts[i];
    }
    return b;
}

inline std::string out_dir(Args& a) {
    std::string dir;
    if (!a.count("outdir")) {
        const auto path = split(a["infile"], "/");
        for (size_t i = 0; i + 1 < path.size(); ++i) dir += path[i] + "/";
        dir += file_base(a["infile"]) + "_out/";
    } else {
        dir = a["outdir"];
        if (dir.back() != '/') dir += "/";
    }
    std::filesystem::create_directory(dir);
    return dir;
}

inline std::string with_segment(const std::string& name, int segment) {
    return segment == -1 ? name : name + "_segment" + std::to_string(segment);
}

}  // namespace ggp
