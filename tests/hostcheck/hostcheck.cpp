// tests/hostcheck/hostcheck.cpp — compiles the product's host+device math headers FOR THE HOST
// (g++ -ffp-contract=off) so that CPU-only tests can compare them bit-for-bit against the
// reference build (oracle/_ref) and the host libm.  Test infrastructure; not part of the product.
#include "../../gfp_gaussian_process_b200/csrc/ggp_tables_data.h"
#include "../../gfp_gaussian_process_b200/csrc/ggp_dawson.cuh"
#ifdef GGP_HOSTCHECK_STEP
#include "../../gfp_gaussian_process_b200/csrc/ggp_step.cuh"
#endif

static const GgpMathTables g_tables = GGP_MATH_TABLES_INIT;

extern "C" {
void hc_exp(long n, const double* x, double* y) { for (long i = 0; i < n; ++i) y[i] = ggp_exp(x[i], &g_tables); }
void hc_log(long n, const double* x, double* y) { for (long i = 0; i < n; ++i) y[i] = ggp_log(x[i], &g_tables); }
void hc_pow(long n, const double* x, const double* e, double* y) { for (long i = 0; i < n; ++i) y[i] = ggp_pow(x[i], e[i], &g_tables); }
void hc_dawson(long n, const double* x, double* y) { for (long i = 0; i < n; ++i) y[i] = ggp_dawson(x[i], &g_tables); }
#ifdef GGP_HOSTCHECK_STEP
// state = 4 means + 10 upper-triangular covariances (xx,xg,xl,xq,gg,gl,gq,ll,lq,qq)
void hc_propagate(long n, const double* state14, const double* dt, const double* p7, double* out14) {
    for (long i = 0; i < n; ++i) {
        GgpState s;
        for (int k = 0; k < 4; ++k) s.m[k] = state14[14 * i + k];
        for (int k = 0; k < 10; ++k) s.c[k] = state14[14 * i + 4 + k];
        GgpOuParams p = {p7[0], p7[1], p7[2], p7[3], p7[4], p7[5], p7[6]};
        ggp_propagate(s, dt[i], p, &g_tables);
        for (int k = 0; k < 4; ++k) out14[14 * i + k] = s.m[k];
        for (int k = 0; k < 10; ++k) out14[14 * i + 4 + k] = s.c[k];
    }
}
#endif
}
