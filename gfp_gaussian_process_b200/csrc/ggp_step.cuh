// ggp_step.cuh — one propagation step of the 4-dim Gaussian belief (x, g, lambda, q) over dt
// without cell division: the B200 path's replacement for the reference's mean_cov_model()
// (reference src/mean_cov_model.h:211-274) and cross_cov_model() (:380-432).
//
// Same mathematics, different evaluation plan.  The reference calls four integral primitives
//   I_k(a,B,c;t0,t1) = int_{t0}^{t1} s^k exp(a s^2 + B s + c) ds,  k = 0..3   (mean_cov_model.h:9-67)
// 59 times per step (82 with the cross covariances), each call re-evaluating two Dawson functions,
// 2-9 exp, a sqrt and a pow.  Here the step is organised around what is actually distinct:
//   * a = C_ll/2 is common to all calls: sqrt(a), a^1.5, a^2.5, a^3.5 are computed once;
//   * the linear coefficient takes 6 values B in {b+bl+Cxl, b+bl+2Cxl} x {., -gq, +gq}; per (B, t'),
//     t' in {0, t, 2t}, one Dawson argument u, one Dawson value D, one exp G (14 pairs in all);
//   * the constant takes 10 values c; per (B, c) pair the exp(c), the two "zeroth" exponentials E
//     and the two "completed square" exponentials H are computed once and shared by k = 0..3;
//   * 17 (B, c, range) groups yield the 39 distinct integrals the five g-moments consume.
// STRICT ROUNDING: every floating-point operation below is the one the reference performs, in the
// reference's order (its results are ill-conditioned enough that re-association is visible at the
// 1e-7 level); only *repeated* evaluations of bit-identical subexpressions were
// removed.  No FMA contraction may be applied to this file (nvcc -fmad=false); exp/pow are the
// glibc-exact routines of ggp_libm.cuh; pow(x,2) is x*x as g++ -O3 folds it.
// Checked bit-for-bit against the reference's own code on the CPU (tests/test_step_bits.py) and
// against the oracle on the GPU (tests/test_gpu_parity.py).
#pragma once
#include "ggp_dawson.cuh"

struct GgpState {
    double m[4];    // mean: x, g, lambda, q
    double c[10];   // covariance, upper triangle row-major: xx xg xl xq gg gl gq ll lq qq
};

struct GgpOuParams {   // the seven dynamic parameters, reference order (likelihood.h:40-42)
    double ml, gl, sl2, mq, gq, sq2, b;
};

// per linear-coefficient quantities at one time t'
struct GgpBT {
    double D;    // Dawson((B + 2 a t')/(2 sqrt a))
    double u2;   // ((B + 2 a t')/(2 sqrt a))^2
    double G;    // exp(t' (B + a t'))
};

struct GgpStepCommon {
    double a, sqa, twoa, two_sqa, m2sqa, p2sqa, foura;
    double den0, den1, den2, den3;   // 2 sqrt a, 4 a^1.5, 8 a^2.5, 16 a^3.5
    double foura2;                   // 4 a^2
};

GGP_HD GgpBT ggp_bt(const GgpStepCommon& k, double B, double tp, const GgpMathTables* __restrict__ M) {
    GgpBT r;
    double u = (B + k.twoa * tp) / k.two_sqa;
    r.D = ggp_dawson(u, M);
    r.u2 = u * u;
    r.G = ggp_exp(tp * (B + k.a * tp), M);
    return r;
}

// integrals of one (B, c) pair over one range [t0, t1]; NK = highest order needed (0..3)
template <int NK>
struct GgpInts { double I[NK + 1]; };

// E(t') = exp(a t'^2 + B t' + c) ; H(t') = exp(-B^2/(4a) + c + u(t')^2)
template <int NK>
GGP_HD GgpInts<NK> ggp_integrals(const GgpStepCommon& k, double B, double Ec, double E0, double E1,
                                 double H0, double H1, const GgpBT& b0, const GgpBT& b1, double t0, double t1) {
    GgpInts<NK> r;
    {   // k = 0, mean_cov_model.h:9-21
        double x = 2. * (-E0 * b0.D + E1 * b1.D);
        r.I[0] = x / k.den0;
    }
    if (NK >= 1) {   // mean_cov_model.h:23-34
        double x = (k.m2sqa * Ec * (b0.G - b1.G) + B * 2. * (H0 * b0.D - H1 * b1.D));
        r.I[NK >= 1 ? 1 : 0] = x / k.den1;
    }
    if (NK >= 2) {   // mean_cov_model.h:36-49
        double B2 = B * B;
        double x = (k.p2sqa * Ec * (b0.G * (B - k.twoa * t0) - b1.G * (B - k.twoa * t1))
                    + (H0 * (k.twoa - B2) * 2. * b0.D + H1 * (-k.twoa + B2) * 2. * b1.D));
        r.I[NK >= 2 ? 2 : 0] = x / k.den2;
    }
    if (NK >= 3) {   // mean_cov_model.h:51-67
        double B2 = B * B;
        double x = (k.m2sqa * Ec *
                    (B2 * (b0.G - b1.G) - k.twoa * b0.G * (2. + B * t0) + k.twoa * b1.G * (2 + B * t1)
                     + k.foura2 * (b0.G * (t0 * t0) - b1.G * (t1 * t1))))
                   + H0 * B * (-6. * k.a + B2) * 2. * b0.D
                   - H1 * B * (-6 * k.a + B2) * 2. * b1.D;
        r.I[NK >= 3 ? 3 : 0] = x / k.den3;
    }
    return r;
}

// a (B, c) pair over [0, t]
template <int NK>
GGP_HD GgpInts<NK> ggp_group_0t(const GgpStepCommon& k, double B, double nb, double c, double t, double at2,
                                const GgpBT& b0, const GgpBT& bt, const GgpMathTables* __restrict__ M) {
    double Ec = ggp_exp(c, M);                       // = exp(a*0 + B*0 + c) as well
    double E1 = ggp_exp(at2 + B * t + c, M);
    double H0 = 0, H1 = 0;
    if (NK >= 1) {
        H0 = ggp_exp(nb + c + b0.u2, M);
        H1 = ggp_exp(nb + c + bt.u2, M);
    }
    return ggp_integrals<NK>(k, B, Ec, Ec, E1, H0, H1, b0, bt, 0.0, t);
}

// a (B, c) pair over [t, 2t]
template <int NK>
GGP_HD GgpInts<NK> ggp_group_t2t(const GgpStepCommon& k, double B, double nb, double c, double t, double at2, double a4t2,
                                 const GgpBT& bt, const GgpBT& b2t, const GgpMathTables* __restrict__ M) {
    double t2 = 2 * t;
    double Ec = (NK >= 1) ? ggp_exp(c, M) : 0.0;
    double E0 = ggp_exp(at2 + B * t + c, M);
    double E1 = ggp_exp(a4t2 + B * t2 + c, M);
    double H0 = 0, H1 = 0;
    if (NK >= 1) {
        H0 = ggp_exp(nb + c + bt.u2, M);
        H1 = ggp_exp(nb + c + b2t.u2, M);
    }
    return ggp_integrals<NK>(k, B, Ec, E0, E1, H0, H1, bt, b2t, t, t2);
}

// a (B, c) pair over both [0, t] and [t, 2t]: the t' = t exponentials are shared
template <int NK>
GGP_HD void ggp_group_both(const GgpStepCommon& k, double B, double nb, double c, double t, double at2, double a4t2,
                           const GgpBT& b0, const GgpBT& bt, const GgpBT& b2t, const GgpMathTables* __restrict__ M,
                           GgpInts<NK>& lo, GgpInts<NK>& hi) {
    double t2 = 2 * t;
    double Ec = ggp_exp(c, M);
    double Et = ggp_exp(at2 + B * t + c, M);
    double E2 = ggp_exp(a4t2 + B * t2 + c, M);
    double nbc = nb + c;
    double H0 = ggp_exp(nbc + b0.u2, M);
    double Ht = ggp_exp(nbc + bt.u2, M);
    double H2 = ggp_exp(nbc + b2t.u2, M);
    lo = ggp_integrals<NK>(k, B, Ec, Ec, Et, H0, Ht, b0, bt, 0.0, t);
    hi = ggp_integrals<NK>(k, B, Ec, Et, E2, Ht, H2, bt, b2t, t, t2);
}

// The step.  If `cross` is non-null it receives Cov(z_{n+1}, z_n) row-major 4x4 (mean_cov_model.h:380-432).
GGP_HD void ggp_propagate_impl(GgpState& s, double t, const GgpOuParams& p, const GgpMathTables* __restrict__ M,
                               double* __restrict__ cross) {
    const double bx = s.m[0], bg = s.m[1], bl = s.m[2], bq = s.m[3];
    const double Cxx = s.c[0], Cxg = s.c[1], Cxl = s.c[2], Cxq = s.c[3], Cgg = s.c[4], Cgl = s.c[5], Cgq = s.c[6],
                 Cll = s.c[7], Clq = s.c[8], Cqq = s.c[9];
    const double ml = p.ml, gl = p.gl, sl2 = p.sl2, mq = p.mq, gq = p.gq, sq2 = p.sq2, b = p.b;

    // ---- elementary exponentials of the x, lambda, q block ----
    const double egl = ggp_exp(-gl * t, M);        // exp(-gl t)
    const double egq = ggp_exp(-gq * t, M);        // exp(-gq t)
    const double ebt = ggp_exp(b * t, M);          // exp(b t)
    const double ebgl = ggp_exp((b + gl) * t, M);  // exp((b+gl) t)
    const double ebgq = ggp_exp((b + gq) * t, M);  // exp((b+gq) t)
    const double e2bt = ggp_exp(2 * b * t, M);     // exp(2 b t)
    const double omegl = 1 - egl;

    // ---- quantities common to all integrals ----
    GgpStepCommon k;
    k.a = Cll / 2.;
    k.sqa = GGP_SQRT(k.a);
    k.twoa = 2. * k.a;
    k.two_sqa = 2. * k.sqa;
    k.m2sqa = -2. * k.sqa;
    k.p2sqa = 2. * k.sqa;
    k.foura = 4. * k.a;
    k.foura2 = 4 * (k.a * k.a);
    k.den0 = 2. * k.sqa;
    k.den1 = 4. * ggp_pow(k.a, 1.5, M);
    k.den2 = 8. * ggp_pow(k.a, 2.5, M);
    k.den3 = 16. * ggp_pow(k.a, 3.5, M);
    const double t2 = 2 * t;
    const double at2 = k.a * (t * t);
    const double a4t2 = k.a * (t2 * t2);

    // ---- the six linear coefficients ----
    const double B = b + bl + Cxl, Bm = b + bl + Cxl - gq, Bp = b + bl + Cxl + gq;
    const double W = b + bl + 2 * Cxl, Wm = b + bl + 2 * Cxl - gq, Wp = b + bl + 2 * Cxl + gq;
    const double nB = -(B * B) / k.foura, nBm = -(Bm * Bm) / k.foura;
    const double nW = -(W * W) / k.foura, nWm = -(Wm * Wm) / k.foura;

    const GgpBT B_0 = ggp_bt(k, B, 0.0, M), B_t = ggp_bt(k, B, t, M);
    const GgpBT Bm_0 = ggp_bt(k, Bm, 0.0, M), Bm_t = ggp_bt(k, Bm, t, M);
    const GgpBT Bp_0 = ggp_bt(k, Bp, 0.0, M), Bp_t = ggp_bt(k, Bp, t, M);
    const GgpBT W_0 = ggp_bt(k, W, 0.0, M), W_t = ggp_bt(k, W, t, M), W_2t = ggp_bt(k, W, t2, M);
    const GgpBT Wm_0 = ggp_bt(k, Wm, 0.0, M), Wm_t = ggp_bt(k, Wm, t, M), Wm_2t = ggp_bt(k, Wm, t2, M);
    const GgpBT Wp_t = ggp_bt(k, Wp, t, M), Wp_2t = ggp_bt(k, Wp, t2, M);

    // ---- the constants ----
    const double c1 = bx + Cxx / 2. - b * t;
    const double c1l = bx + Cxx / 2. - b * t - gl * t;
    const double c1q = bx + Cxx / 2. - b * t - gq * t;
    const double c1qw = -b * t + bx + Cxx / 2. - gq * t;   // the reference's second spelling (mean_cov_model.h:184,186)
    const double c2 = bx + Cxx / 2. - 2 * b * t;
    const double d1 = 2 * (bx + Cxx - b * t);              // == 2*bx + 2*Cxx - 2*b*t bit for bit (scaling by 2 is exact)
    const double d2 = 2 * bx + 2 * Cxx - (2 * b + gq) * t;
    const double d3 = 2 * bx + 2 * Cxx - 2 * b * t + gq * t;
    const double d4 = 2 * bx + 2 * Cxx - 2 * b * t - 2 * gq * t;

    // ---- the 17 groups, 39 integrals ----
    const GgpInts<1> jB_c1 = ggp_group_0t<1>(k, B, nB, c1, t, at2, B_0, B_t, M);
    const GgpInts<2> jBm_c1 = ggp_group_0t<2>(k, Bm, nBm, c1, t, at2, Bm_0, Bm_t, M);

    // new mean (mean_cov_model.h:73-87); needed by the covariance terms below
    double nm0 = bx + ml * t + (bl - ml) * omegl / gl;
    double nm1 = bg / ebt + Clq * jBm_c1.I[1] + mq * jB_c1.I[0] + (bq + Cxq - mq) * jBm_c1.I[0];
    double nm2 = ml + (bl - ml) * egl;
    double nm3 = mq + (bq - mq) * egq;

    if (cross) {   // mean_cov_model.h:282-377: uses only the two (., c1) groups over [0, t]
        cross[0] = Cxx + Cxl * omegl / gl;
        cross[1] = Cxg + Cgl * omegl / gl;
        cross[2] = Cxl + Cll * omegl / gl;
        cross[3] = Cxq + Clq * omegl / gl;
        cross[4] = (bg * bx) / ebt + Cxg / ebt + Cxl * mq * jB_c1.I[1]
                   + (bx * Clq + bq * Cxl + Cxl * Cxq + Clq * Cxx - Cxl * mq) * jBm_c1.I[1]
                   + Clq * Cxl * jBm_c1.I[2] + (bx * mq + Cxx * mq) * jB_c1.I[0]
                   + (bq * bx + Cxq + bx * Cxq + bq * Cxx + Cxq * Cxx - bx * mq - Cxx * mq) * jBm_c1.I[0] - nm1 * bx;
        cross[5] = (bg * bg) / ebt + Cgg / ebt + Cgl * mq * jB_c1.I[1]
                   + (bq * Cgl + bg * Clq + Clq * Cxg + Cgl * Cxq - Cgl * mq) * jBm_c1.I[1]
                   + Cgl * Clq * jBm_c1.I[2] + (bg * mq + Cxg * mq) * jB_c1.I[0]
                   + (bg * bq + Cgq + bq * Cxg + bg * Cxq + Cxg * Cxq - bg * mq - Cxg * mq) * jBm_c1.I[0] - nm1 * bg;
        cross[6] = (bg * bl) / ebt + Cgl / ebt + Cll * mq * jB_c1.I[1]
                   + (bq * Cll + bl * Clq + Clq * Cxl + Cll * Cxq - Cll * mq) * jBm_c1.I[1]
                   + Cll * Clq * jBm_c1.I[2] + (bl * mq + Cxl * mq) * jB_c1.I[0]
                   + (bl * bq + Clq + bq * Cxl + bl * Cxq + Cxl * Cxq - bl * mq - Cxl * mq) * jBm_c1.I[0] - nm1 * bl;
        cross[7] = (bg * bq) / ebt + Cgq / ebt + Clq * mq * jB_c1.I[1]
                   + (2 * bq * Clq + 2 * Clq * Cxq - Clq * mq) * jBm_c1.I[1]
                   + (Clq * Clq) * jBm_c1.I[2] + (bq * mq + Cxq * mq) * jB_c1.I[0]
                   + ((bq * bq) + Cqq + 2 * bq * Cxq + (Cxq * Cxq) - bq * mq - Cxq * mq) * jBm_c1.I[0] - nm1 * bq;
        cross[8] = Cxl * egl; cross[9] = Cgl * egl; cross[10] = Cll * egl; cross[11] = Clq * egl;
        cross[12] = Cxq * egq; cross[13] = Cgq * egq; cross[14] = Clq * egq; cross[15] = Cqq * egq;
    }

    const GgpInts<1> jB_c1l = ggp_group_0t<1>(k, B, nB, c1l, t, at2, B_0, B_t, M);
    const GgpInts<2> jBm_c1l = ggp_group_0t<2>(k, Bm, nBm, c1l, t, at2, Bm_0, Bm_t, M);

    // ---- cov_xg, mean_cov_model.h:97-115 ----
    double n_xg =
        (bg * bx) / ebt + Cxg / ebt + (bg * bl) / (ebt * gl) + Cgl / (ebt * gl) - (bg * bl) / (ebgl * gl)
        - Cgl / (ebgl * gl) - (bg * ml) / (ebt * gl) + (bg * ml) / (ebgl * gl) + (bg * ml * t) / ebt
        + (Cxl * mq + (Cll * mq) / gl) * jB_c1.I[1]
        - (Cll * mq * jB_c1l.I[1]) / gl
        + (bx * Clq + bq * Cxl + Cxl * Cxq + Clq * Cxx + (bq * Cll) / gl + (bl * Clq) / gl + (Clq * Cxl) / gl
           + (Cll * Cxq) / gl - (Clq * ml) / gl - Cxl * mq - (Cll * mq) / gl + Clq * ml * t) * jBm_c1.I[1]
        + (-((bq * Cll) / gl) - (bl * Clq) / gl - (Clq * Cxl) / gl - (Cll * Cxq) / gl + (Clq * ml) / gl
           + (Cll * mq) / gl) * jBm_c1l.I[1]
        + (Clq * Cxl + (Cll * Clq) / gl) * jBm_c1.I[2]
        - (Cll * Clq * jBm_c1l.I[2]) / gl
        + (bx * mq + Cxx * mq + (bl * mq) / gl + (Cxl * mq) / gl - (ml * mq) / gl + ml * mq * t) * jB_c1.I[0]
        + (-((bl * mq) / gl) - (Cxl * mq) / gl + (ml * mq) / gl) * jB_c1l.I[0]
        + (bq * bx + Cxq + bx * Cxq + bq * Cxx + Cxq * Cxx + (bl * bq) / gl + Clq / gl + (bq * Cxl) / gl
           + (bl * Cxq) / gl + (Cxl * Cxq) / gl - (bq * ml) / gl - (Cxq * ml) / gl - bx * mq - Cxx * mq
           - (bl * mq) / gl - (Cxl * mq) / gl + (ml * mq) / gl + bq * ml * t + Cxq * ml * t - ml * mq * t) * jBm_c1.I[0]
        + (-((bl * bq) / gl) - Clq / gl - (bq * Cxl) / gl - (bl * Cxq) / gl - (Cxl * Cxq) / gl + (bq * ml) / gl
           + (Cxq * ml) / gl + (bl * mq) / gl + (Cxl * mq) / gl - (ml * mq) / gl) * jBm_c1l.I[0]
        - nm1 * nm0;

    // ---- cov_gl, mean_cov_model.h:166-176 ----
    double n_gl =
        (bg * bl) / ebgl + Cgl / ebgl + (bg * ml) / ebt - (bg * ml) / ebgl
        + Cll * mq * jB_c1l.I[1] + Clq * ml * jBm_c1.I[1]
        + (bq * Cll + bl * Clq + Clq * Cxl + Cll * Cxq - Clq * ml - Cll * mq) * jBm_c1l.I[1]
        + Cll * Clq * jBm_c1l.I[2] + ml * mq * jB_c1.I[0]
        + (bl * mq + Cxl * mq - ml * mq) * jB_c1l.I[0]
        + (bq * ml + Cxq * ml - ml * mq) * jBm_c1.I[0]
        + (bl * bq + Clq + bq * Cxl + bl * Cxq + Cxl * Cxq - bq * ml - Cxq * ml - bl * mq - Cxl * mq + ml * mq) * jBm_c1l.I[0]
        - nm1 * nm2;

    // ---- cov_gq, mean_cov_model.h:178-192 ----
    double n_gq;
    {
        const GgpInts<1> jB_c1q = ggp_group_0t<1>(k, B, nB, c1q, t, at2, B_0, B_t, M);
        const GgpInts<2> jBm_c1q = ggp_group_0t<2>(k, Bm, nBm, c1q, t, at2, Bm_0, Bm_t, M);
        const GgpInts<0> jBm_c1qw = ggp_group_0t<0>(k, Bm, nBm, c1qw, t, at2, Bm_0, Bm_t, M);
        const GgpInts<0> jBp_c1qw = ggp_group_0t<0>(k, Bp, 0.0, c1qw, t, at2, Bp_0, Bp_t, M);
        n_gq =
            (bg * bq) / ebgq + Cgq / ebgq + (bg * mq) / ebt - (bg * mq) / ebgq
            + Clq * mq * jB_c1q.I[1] + Clq * mq * jBm_c1.I[1]
            + (2 * bq * Clq + 2 * Clq * Cxq - 2 * Clq * mq) * jBm_c1q.I[1]
            + (Clq * Clq) * jBm_c1q.I[2] + (mq * mq) * jB_c1.I[0]
            + (bq * mq + Cxq * mq - (mq * mq)) * jB_c1q.I[0]
            + (bq * mq + Cxq * mq - (mq * mq)) * jBm_c1.I[0]
            - (sq2 * jBm_c1qw.I[0]) / (2. * gq)
            + ((bq * bq) + Cqq + 2 * bq * Cxq + (Cxq * Cxq) - 2 * bq * mq - 2 * Cxq * mq + (mq * mq)) * jBm_c1q.I[0]
            + (sq2 * jBp_c1qw.I[0]) / (2. * gq)
            - nm1 * nm3;
    }

    // ---- cov_gg, mean_cov_model.h:124-164 ----
    double n_gg;
    {
        const GgpInts<1> jB_c2 = ggp_group_0t<1>(k, B, nB, c2, t, at2, B_0, B_t, M);
        const GgpInts<2> jBm_c2 = ggp_group_0t<2>(k, Bm, nBm, c2, t, at2, Bm_0, Bm_t, M);
        GgpInts<1> jW_lo, jW_hi;
        ggp_group_both<1>(k, W, nW, d1, t, at2, a4t2, W_0, W_t, W_2t, M, jW_lo, jW_hi);
        GgpInts<3> jWm_lo, jWm_hi;
        ggp_group_both<3>(k, Wm, nWm, d1, t, at2, a4t2, Wm_0, Wm_t, Wm_2t, M, jWm_lo, jWm_hi);
        const GgpInts<1> jW_d2 = ggp_group_t2t<1>(k, W, nW, d2, t, at2, a4t2, W_t, W_2t, M);
        const GgpInts<1> jWm_d3 = ggp_group_t2t<1>(k, Wm, nWm, d3, t, at2, a4t2, Wm_t, Wm_2t, M);
        const GgpInts<0> jWp_d4 = ggp_group_t2t<0>(k, Wp, 0.0, d4, t, at2, a4t2, Wp_t, Wp_2t, M);
        const double mq2 = mq * mq, bq2 = bq * bq, Cxq2 = Cxq * Cxq, Clq2 = Clq * Clq, gq2 = gq * gq;
        n_gg =
            ((bg * bg) + Cgg) / e2bt
            + 2 * Cgl * mq * jB_c2.I[1]
            + (mq * (2 * Clq + gq * mq) * jW_lo.I[1]) / gq
            + 2 * (bq * Cgl + bg * Clq + Clq * Cxg + Cgl * Cxq - Cgl * mq) * jBm_c2.I[1]
            + ((bq2 * gq + Cqq * gq + 4 * bq * Cxq * gq + 4 * Cxq2 * gq - 2 * Clq * mq - 2 * bq * gq * mq
                - 4 * Cxq * gq * mq + gq * mq2) * jWm_lo.I[1]) / gq
            - mq2 * jW_hi.I[1]
            - (2 * Clq * mq * jW_d2.I[1]) / gq
            - (sq2 * jWm_lo.I[1]) / (2. * gq)
            + (sq2 * jWm_hi.I[1]) / (2. * gq)
            + (-bq2 - Cqq - 4 * bq * Cxq - 4 * Cxq2 + 2 * bq * mq + 4 * Cxq * mq - mq2 + 4 * bq * Clq * t
               + 8 * Clq * Cxq * t - 4 * Clq * mq * t) * jWm_hi.I[1]
            + (2 * Clq * mq * jWm_d3.I[1]) / gq
            + Clq2 * jWm_lo.I[3]
            - Clq2 * jWm_hi.I[3]
            + 2 * Cgl * Clq * jBm_c2.I[2]
            + (2 * bq * Clq + 4 * Clq * Cxq - 2 * Clq * mq) * jWm_lo.I[2]
            + (-2 * bq * Clq - 4 * Clq * Cxq + 2 * Clq * mq + 2 * Clq2 * t) * jWm_hi.I[2]
            + (2 * bg * mq + 2 * Cxg * mq) * jB_c2.I[0]
            + ((2 * bq * mq) / gq + (4 * Cxq * mq) / gq - (2 * mq2) / gq) * jW_lo.I[0]
            + (2 * bg * bq + 2 * Cgq + 2 * bq * Cxg + 2 * bg * Cxq + 2 * Cxg * Cxq - 2 * bg * mq - 2 * Cxg * mq) * jBm_c2.I[0]
            + ((-2 * bq * mq) / gq - (4 * Cxq * mq) / gq + (2 * mq2) / gq) * jWm_lo.I[0]
            + (sq2 * jW_lo.I[0]) / (2. * gq2)
            + (sq2 * jW_hi.I[0]) / (2. * gq2)
            + 2 * mq2 * t * jW_hi.I[0]
            + ((-2 * bq * mq) / gq - (4 * Cxq * mq) / gq + (2 * mq2) / gq) * jW_d2.I[0]
            - (sq2 * jWm_lo.I[0]) / (2. * gq2)
            - (sq2 * t * jWm_hi.I[0]) / gq
            + (2 * bq2 * t + 2 * Cqq * t + 8 * bq * Cxq * t + 8 * Cxq2 * t - 4 * bq * mq * t - 8 * Cxq * mq * t
               + 2 * mq2 * t) * jWm_hi.I[0]
            + ((2 * bq * mq) / gq + (4 * Cxq * mq) / gq - (2 * mq2) / gq) * jWm_d3.I[0]
            - (sq2 * jWp_d4.I[0]) / (2. * gq2)
            - (nm1 * nm1);
    }

    // ---- the elementary block, mean_cov_model.h:93-95, 117-122, 196-208 ----
    const double egl2 = egl * egl, egq2 = egq * egq;
    double n_xx = Cll * (omegl * omegl) / (gl * gl) + 2 * Cxl * omegl / gl + Cxx
                  + sl2 / (2 * ggp_pow(gl, 3.0, M)) * (2 * gl * t - 3 + 4 * egl - egl2);
    double n_xl = sl2 / (2 * (gl * gl)) * (omegl * omegl) + Cll * egl * omegl / gl + Cxl * egl;
    double n_xq = Clq * omegl * egq / gl + Cxq * egq;
    double n_ll = Cll * egl2 + sl2 / (2 * gl) * (1 - egl2);
    double n_lq = Clq * egl * egq;
    double n_qq = sq2 / (2 * gq) * (1 - egq2) + Cqq * egq2;

    s.m[0] = nm0; s.m[1] = nm1; s.m[2] = nm2; s.m[3] = nm3;
    s.c[0] = n_xx; s.c[1] = n_xg; s.c[2] = n_xl; s.c[3] = n_xq; s.c[4] = n_gg;
    s.c[5] = n_gl; s.c[6] = n_gq; s.c[7] = n_ll; s.c[8] = n_lq; s.c[9] = n_qq;
}

GGP_HD void ggp_propagate(GgpState& s, double t, const GgpOuParams& p, const GgpMathTables* __restrict__ M) {
    ggp_propagate_impl(s, t, p, M, nullptr);
}
