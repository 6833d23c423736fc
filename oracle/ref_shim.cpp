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

#ifdef GGP_REF_ULP_PERTURB
// +-1 ulp envelope build (tools/ulp_envelope.py): every exp, true pow and Dawson RESULT inside the
// reference's mean_cov_model.h is moved one ulp up or down at random (counter-based generator, seed set by
// ggp_ref_set_ulp_seed; seed 0 = untouched).  pow(x, 2) stays x * x: g++ -O3 folds it to a multiplication in the
// reference binary, so it is not a libm result there.  The macros below only rename the calls; the header's text is
// the reference's, included in place.
#include <cstdint>
#include "/root/reference/src/Faddeeva.hh"
static uint64_t g_ulp_seed = 0, g_ulp_counter = 0;
static inline double ulp_nudge(double v) {
    if (g_ulp_seed == 0 || !(v == v) || v == 0.0 || std::isinf(v)) return v;
    uint64_t z = g_ulp_seed + 0x9e3779b97f4a7c15ull * ++g_ulp_counter;   // splitmix64
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    z ^= z >> 31;
    return std::nextafter(v, (z & 1) ? HUGE_VAL : -HUGE_VAL);
}
static inline double ulp_exp(double x) { return ulp_nudge(std::exp(x)); }
static inline double ulp_pow(double x, double y) { return y == 2.0 ? x * x : ulp_nudge(std::pow(x, y)); }
namespace FaddeevaUlp {
static inline double Dawson(double x) { return ulp_nudge(Faddeeva::Dawson(x)); }
}
extern "C" void ggp_ref_set_ulp_seed(unsigned long long seed) { g_ulp_seed = seed; g_ulp_counter = 0; }
#define exp ulp_exp
#define pow ulp_pow
#define Faddeeva FaddeevaUlp
#include "/root/reference/src/mean_cov_model.h"
#undef exp
#undef pow
#undef Faddeeva
#else
#include "/root/reference/src/mean_cov_model.h"
#endif

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
