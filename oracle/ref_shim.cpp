// oracle/ref_shim.cpp — builds oracle/_ref/libggp_ref.so from the REFERENCE'S OWN SOURCES, in place:
//   /root/reference/src/mean_cov_model.h  (included unmodified below)
//   /root/reference/src/Faddeeva.cc       (compiled unmodified by oracle/Makefile)
// Nothing from the reference is copied into this repository.  Eigen is not installed in this
// image, so the two Eigen types mean_cov_model.h touches (element access, copy,
// comma initialiser) get a minimal stand-in here; all arithmetic is the reference's.
// TEST INFRASTRUCTURE ONLY: used by tests/ and bench.py's cpu_baseline/reference arm to pin
// oracle/ggp_oracle.cpp and to time the reference's CPU math.  Never linked into the product.
#include <cmath>
#include <cstring>

namespace Eigen {
struct VectorXd {
    double v[4];
    VectorXd() { std::memset(v, 0, sizeof v); }
    explicit VectorXd(int) { std::memset(v, 0, sizeof v); }
    double& operator()(int i) { return v[i]; }
    double operator()(int i) const { return v[i]; }
};
struct MatrixXd {
    double m[16];   // row-major 4x4
    struct Comma {
        MatrixXd* M; int n;
        Comma& operator,(double x) { M->m[n++] = x; return *this; }
    };
    MatrixXd() { std::memset(m, 0, sizeof m); }
    MatrixXd(int, int) { std::memset(m, 0, sizeof m); }
    double& operator()(int i, int j) { return m[4 * i + j]; }
    double operator()(int i, int j) const { return m[4 * i + j]; }
    Comma operator<<(double x) { m[0] = x; return Comma{this, 1}; }
};
}  // namespace Eigen

struct MOMAdata {
    Eigen::VectorXd mean;
    Eigen::MatrixXd cov;
};

#include "/root/reference/src/mean_cov_model.h"

extern "C" {

// one call of the reference's mean_cov_model (mean_cov_model.h:211); cov is row-major 4x4
void ggp_ref_mean_cov_model(const double* mean, const double* cov, double t, const double* p7,
                            double* mean_out, double* cov_out) {
    MOMAdata c;
    std::memcpy(c.mean.v, mean, sizeof c.mean.v);
    std::memcpy(c.cov.m, cov, sizeof c.cov.m);
    mean_cov_model(c, t, p7[0], p7[1], p7[2], p7[3], p7[4], p7[5], p7[6]);
    std::memcpy(mean_out, c.mean.v, sizeof c.mean.v);
    std::memcpy(cov_out, c.cov.m, sizeof c.cov.m);
}

// reference cross_cov_model (mean_cov_model.h:380); returns row-major 4x4
void ggp_ref_cross_cov_model(const double* mean, const double* cov, double t, const double* p7,
                             double* cross_out) {
    MOMAdata c;
    std::memcpy(c.mean.v, mean, sizeof c.mean.v);
    std::memcpy(c.cov.m, cov, sizeof c.cov.m);
    Eigen::MatrixXd r = cross_cov_model(c, t, p7[0], p7[1], p7[2], p7[3], p7[4], p7[5], p7[6]);
    std::memcpy(cross_out, r.m, sizeof r.m);
}

double ggp_ref_dawson(double x) { return Faddeeva::Dawson(x); }

// the four integral primitives (mean_cov_model.h:9-67)
double ggp_ref_tauint(int k, double a, double b, double c, double t1, double t0) {
    switch (k) {
        case 0: return zerotauint(a, b, c, t1, t0);
        case 1: return onetauint(a, b, c, t1, t0);
        case 2: return twotauint(a, b, c, t1, t0);
        default: return treetauint(a, b, c, t1, t0);
    }
}

// host libm as the reference binary sees it (for tests/test_libm_bits.py)
double ggp_ref_exp(double x) { return exp(x); }
double ggp_ref_log(double x) { return log(x); }
double ggp_ref_pow(double x, double y) { return pow(x, y); }

// many steps of the reference propagation, for CPU timing (bench.py --impl reference)
void ggp_ref_mean_cov_model_batch(long n, const double* mean, const double* cov, const double* t,
                                  const double* p7, double* mean_out, double* cov_out) {
    for (long i = 0; i < n; ++i)
        ggp_ref_mean_cov_model(mean + 4 * i, cov + 16 * i, t[i], p7, mean_out + 4 * i, cov_out + 16 * i);
}
}
