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
// EXECUTION PLAN: the 14 (B, t') pairs and the 17 groups are processed by LOOPS driven by small
// descriptor tables, with their inputs and results in a per-thread scratch (shared memory on the device),
// instead of 31 inlined copies: the instruction stream of a step drops from ~17k to ~5k SASS instructions
// (the straight-line version was instruction-cache bound, profiles/), and each loop body evaluates its 2-4
// exp / Dawson chains interleaved so a warp has independent FP64 work in flight.
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

// ---- per-thread scratch ---------------------------------------------------------------------------
// element i of the calling thread; on the device a column of a [GGP_SCRATCH][blockDim.x] shared array
// (conflict-free), on the host a plain array
struct GgpScratch {
    double* base;
    int stride;
    GGP_HDM double& operator[](int i) const { return base[i * stride]; }
};
enum {
    GGP_S_B = 0,     // 6 linear coefficients B, Bm, Bp, W, Wm, Wp
    GGP_S_NB = 6,    // 6: -(B^2)/(4a)
    GGP_S_C = 12,    // 9 constants c1, c1l, c1q, c1qw, c2, d1, d2, d3, d4
    GGP_S_U2 = 21,   // 14: u^2 of the (B, t') pairs
    GGP_S_D = 35,    // 14: u, then Dawson(u)
    GGP_S_G = 49,    // 14: t'(B + a t'), then its exp   (+6 elementary exp arguments/results at GGP_S_I)
    GGP_S_I = 63,    // 22: integrals of the current segment
    GGP_SCRATCH = 85
};

// (B, t') pairs: index of B and of t' in {0, t, 2t}
#define GGP_BT_B_INIT {0, 0, 1, 1, 2, 2, 3, 3, 3, 4, 4, 4, 5, 5}
#define GGP_BT_T_INIT {0, 1, 0, 1, 0, 1, 0, 1, 2, 0, 1, 2, 1, 2}
// groups: B index, c index, range (0: [0,t], 1: [t,2t]), highest order, pair at range start, pair at range end,
// chained (1: shares the t' = t exponentials of the preceding group), first output slot
struct GgpGroup { unsigned char b, c, hi, nk, i0, i1, chain, out; };
#define GGP_GROUP_INIT {                                                                         \
    {0, 0, 0, 1, 0, 1, 0, 0},  {1, 0, 0, 2, 2, 3, 0, 2},  {0, 1, 0, 1, 0, 1, 0, 5},  {1, 1, 0, 2, 2, 3, 0, 7},      \
    {0, 2, 0, 1, 0, 1, 0, 0},  {1, 2, 0, 2, 2, 3, 0, 2},  {1, 3, 0, 0, 2, 3, 0, 5},  {2, 3, 0, 0, 4, 5, 0, 6},      \
    {0, 4, 0, 1, 0, 1, 0, 0},  {1, 4, 0, 2, 2, 3, 0, 2},  {3, 5, 0, 1, 6, 7, 0, 5},  {3, 5, 1, 1, 7, 8, 1, 7},      \
    {4, 5, 0, 3, 9, 10, 0, 9}, {4, 5, 1, 3, 10, 11, 1, 13}, {3, 6, 1, 1, 7, 8, 0, 17}, {4, 7, 1, 1, 10, 11, 0, 19}, \
    {5, 8, 1, 0, 12, 13, 0, 21}}
#if defined(__CUDACC__)
__constant__ unsigned char ggp_bt_b_dev[14] = GGP_BT_B_INIT;
__constant__ unsigned char ggp_bt_t_dev[14] = GGP_BT_T_INIT;
__constant__ GgpGroup ggp_group_dev[17] = GGP_GROUP_INIT;
#endif
static const unsigned char ggp_bt_b_host[14] = GGP_BT_B_INIT;
static const unsigned char ggp_bt_t_host[14] = GGP_BT_T_INIT;
static const GgpGroup ggp_group_host[17] = GGP_GROUP_INIT;
#if defined(__CUDA_ARCH__)
#define GGP_BT_B ggp_bt_b_dev
#define GGP_BT_T ggp_bt_t_dev
#define GGP_GROUPS ggp_group_dev
#else
#define GGP_BT_B ggp_bt_b_host
#define GGP_BT_T ggp_bt_t_host
#define GGP_GROUPS ggp_group_host
#endif

struct GgpStepCommon {
    double a, sqa, twoa, m2sqa, p2sqa;
    GgpDivisor two_sqa, foura;
    GgpDivisor den0, den1, den2, den3;   // 2 sqrt a, 4 a^1.5, 8 a^2.5, 16 a^3.5
    double foura2;                   // 4 a^2
    double t, t2, at2, a4t2;
};

template <int NK>
struct GgpInts { double I[NK + 1]; };

template <int NK>
GGP_HD GgpInts<NK> ggp_load_ints(const GgpScratch& S, int slot) {
    GgpInts<NK> r;
#pragma unroll
    for (int k = 0; k <= NK; ++k) r.I[k] = S[GGP_S_I + slot + k];
    return r;
}

// groups [g0, g1): exponentials E(t') = exp(a t'^2 + B t' + c), H(t') = exp(-B^2/(4a) + c + u(t')^2), then the
// integrals of order 0..nk (mean_cov_model.h:9-67) into the scratch
GGP_HD void ggp_eval_groups(const GgpScratch& S, int g0, int g1, const GgpStepCommon& k, const GgpMathTables* __restrict__ M) {
    double pEc = 0, pE1 = 0, pH1 = 0;   // previous group's exponentials (chained groups)
#pragma unroll 1
    for (int g = g0; g < g1; ++g) {
        const GgpGroup d = GGP_GROUPS[g];
        const double B = S[GGP_S_B + d.b], c = S[GGP_S_C + d.c];
        const double u2_0 = S[GGP_S_U2 + d.i0], u2_1 = S[GGP_S_U2 + d.i1];
        const double D0 = S[GGP_S_D + d.i0], D1 = S[GGP_S_D + d.i1];
        const double t0 = d.hi ? k.t : 0.0, t1 = d.hi ? k.t2 : k.t;
        const double aE1 = (d.hi ? k.a4t2 : k.at2) + B * t1 + c;
        double Ec, E0, E1, H0 = 0, H1 = 0;
        if (d.nk >= 1) {
            const double nbc = S[GGP_S_NB + d.b] + c;
            if (d.chain) {
                double x[2] = {aE1, nbc + u2_1}, y[2];
                ggp_exp_n<2>(x, y, M);
                Ec = pEc; E0 = pE1; H0 = pH1; E1 = y[0]; H1 = y[1];
            } else if (d.hi) {
                double x[5] = {c, k.at2 + B * k.t + c, aE1, nbc + u2_0, nbc + u2_1}, y[5];
                ggp_exp_n<5>(x, y, M);
                Ec = y[0]; E0 = y[1]; E1 = y[2]; H0 = y[3]; H1 = y[4];
            } else {
                double x[4] = {c, aE1, nbc + u2_0, nbc + u2_1}, y[4];
                ggp_exp_n<4>(x, y, M);
                Ec = y[0]; E0 = y[0]; E1 = y[1]; H0 = y[2]; H1 = y[3];
            }
        } else {
            double x[2] = {d.hi ? k.at2 + B * k.t + c : c, aE1}, y[2];
            ggp_exp_n<2>(x, y, M);
            Ec = y[0]; E0 = y[0]; E1 = y[1];
        }
        pEc = Ec; pE1 = E1; pH1 = H1;
        {   // order 0, mean_cov_model.h:9-21
            const double x = 2. * (-E0 * D0 + E1 * D1);
            S[GGP_S_I + d.out] = x / k.den0;
        }
        if (d.nk >= 1) {
            const double G0 = S[GGP_S_G + d.i0], G1 = S[GGP_S_G + d.i1];
            {   // order 1, mean_cov_model.h:23-34
                const double x = (k.m2sqa * Ec * (G0 - G1) + B * 2. * (H0 * D0 - H1 * D1));
                S[GGP_S_I + d.out + 1] = x / k.den1;
            }
            if (d.nk >= 2) {   // order 2, mean_cov_model.h:36-49
                const double B2 = B * B;
                const double x = (k.p2sqa * Ec * (G0 * (B - k.twoa * t0) - G1 * (B - k.twoa * t1))
                                  + (H0 * (k.twoa - B2) * 2. * D0 + H1 * (-k.twoa + B2) * 2. * D1));
                S[GGP_S_I + d.out + 2] = x / k.den2;
                if (d.nk >= 3) {   // order 3, mean_cov_model.h:51-67
                    const double x3 = (k.m2sqa * Ec *
                                       (B2 * (G0 - G1) - k.twoa * G0 * (2. + B * t0) + k.twoa * G1 * (2 + B * t1)
                                        + k.foura2 * (G0 * (t0 * t0) - G1 * (t1 * t1))))
                                      + H0 * B * (-6. * k.a + B2) * 2. * D0
                                      - H1 * B * (-6 * k.a + B2) * 2. * D1;
                    S[GGP_S_I + d.out + 3] = x3 / k.den3;
                }
            }
        }
    }
}

// The step.  If `cross` is non-null it receives Cov(z_{n+1}, z_n) row-major 4x4 (mean_cov_model.h:380-432).
GGP_HD void ggp_propagate_impl(GgpState& s, double t, const GgpOuParams& p, const GgpMathTables* __restrict__ M,
                               const GgpScratch& S, double* __restrict__ cross) {
    const double bx = s.m[0], bg = s.m[1], bl = s.m[2], bq = s.m[3];
    const double Cxx = s.c[0], Cxg = s.c[1], Cxl = s.c[2], Cxq = s.c[3], Cgg = s.c[4], Cgl = s.c[5], Cgq = s.c[6],
                 Cll = s.c[7], Clq = s.c[8], Cqq = s.c[9];
    const double ml = p.ml, gl = p.gl, sl2 = p.sl2, mq = p.mq, gq = p.gq, sq2 = p.sq2, b = p.b;

    // ---- quantities common to all integrals ----
    GgpStepCommon k;
    k.a = Cll / 2.;
    k.sqa = GGP_SQRT(k.a);
    k.twoa = 2. * k.a;
    k.two_sqa = ggp_divisor(2. * k.sqa);
    k.m2sqa = -2. * k.sqa;
    k.p2sqa = 2. * k.sqa;
    k.foura = ggp_divisor(4. * k.a);
    k.foura2 = 4 * (k.a * k.a);
    k.den0 = k.two_sqa;
    k.den1 = ggp_divisor(4. * ggp_pow(k.a, 1.5, M));
    k.den2 = ggp_divisor(8. * ggp_pow(k.a, 2.5, M));
    k.den3 = ggp_divisor(16. * ggp_pow(k.a, 3.5, M));
    k.t = t;
    k.t2 = 2 * t;
    k.at2 = k.a * (t * t);
    k.a4t2 = k.a * (k.t2 * k.t2);

    // ---- the six linear coefficients and the nine constants ----
    {
        const double B = b + bl + Cxl, Bm = b + bl + Cxl - gq, Bp = b + bl + Cxl + gq;
        const double W = b + bl + 2 * Cxl, Wm = b + bl + 2 * Cxl - gq, Wp = b + bl + 2 * Cxl + gq;
        S[GGP_S_B + 0] = B; S[GGP_S_B + 1] = Bm; S[GGP_S_B + 2] = Bp;
        S[GGP_S_B + 3] = W; S[GGP_S_B + 4] = Wm; S[GGP_S_B + 5] = Wp;
        S[GGP_S_NB + 0] = -(B * B) / k.foura; S[GGP_S_NB + 1] = -(Bm * Bm) / k.foura;
        S[GGP_S_NB + 3] = -(W * W) / k.foura; S[GGP_S_NB + 4] = -(Wm * Wm) / k.foura;
        S[GGP_S_C + 0] = bx + Cxx / 2. - b * t;
        S[GGP_S_C + 1] = bx + Cxx / 2. - b * t - gl * t;
        S[GGP_S_C + 2] = bx + Cxx / 2. - b * t - gq * t;
        S[GGP_S_C + 3] = -b * t + bx + Cxx / 2. - gq * t;   // the reference's second spelling (mean_cov_model.h:184,186)
        S[GGP_S_C + 4] = bx + Cxx / 2. - 2 * b * t;
        S[GGP_S_C + 5] = 2 * (bx + Cxx - b * t);            // == 2*bx + 2*Cxx - 2*b*t bit for bit (scaling by 2 is exact)
        S[GGP_S_C + 6] = 2 * bx + 2 * Cxx - (2 * b + gq) * t;
        S[GGP_S_C + 7] = 2 * bx + 2 * Cxx - 2 * b * t + gq * t;
        S[GGP_S_C + 8] = 2 * bx + 2 * Cxx - 2 * b * t - 2 * gq * t;
    }
    // ---- the 14 (B, t') pairs: u, u^2, argument of G ----
#pragma unroll 1
    for (int i = 0; i < 14; i += 2) {
        double u[2], D[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const double B = S[GGP_S_B + GGP_BT_B[i + j]];
            const int ts = GGP_BT_T[i + j];
            const double tp = ts == 0 ? 0.0 : (ts == 1 ? t : k.t2);
            u[j] = (B + k.twoa * tp) / k.two_sqa;
            S[GGP_S_U2 + i + j] = u[j] * u[j];
            S[GGP_S_G + i + j] = tp * (B + k.a * tp);
        }
        ggp_dawson_n<2>(u, D, M);
        S[GGP_S_D + i] = D[0];
        S[GGP_S_D + i + 1] = D[1];
    }
    // ---- exponentials: G of the 14 pairs and the six elementary ones ----
    S[GGP_S_G + 14] = -gl * t;          // slots 63.. belong to the integrals later
    S[GGP_S_G + 15] = -gq * t;
    S[GGP_S_G + 16] = b * t;
    S[GGP_S_G + 17] = (b + gl) * t;
    S[GGP_S_G + 18] = (b + gq) * t;
    S[GGP_S_G + 19] = 2 * b * t;
#pragma unroll 1
    for (int i = 0; i < 20; i += 4) {
        double x[4] = {S[GGP_S_G + i], S[GGP_S_G + i + 1], S[GGP_S_G + i + 2], S[GGP_S_G + i + 3]}, y[4];
        ggp_exp_n<4>(x, y, M);
#pragma unroll
        for (int j = 0; j < 4; ++j) S[GGP_S_G + i + j] = y[j];
    }
    const double egl = S[GGP_S_G + 14];    // exp(-gl t)
    const double egq = S[GGP_S_G + 15];    // exp(-gq t)
    const double ebt = S[GGP_S_G + 16];    // exp(b t)
    const double ebgl = S[GGP_S_G + 17];   // exp((b+gl) t)
    const double ebgq = S[GGP_S_G + 18];   // exp((b+gq) t)
    const double e2bt = S[GGP_S_G + 19];   // exp(2 b t)
    const double omegl = 1 - egl;
    // divisors met more than once (see GgpDivisor): the reciprocal refinement is shared, the quotients are IEEE
    const GgpDivisor gl_ = ggp_divisor(gl), gq_ = ggp_divisor(gq), ebt_ = ggp_divisor(ebt), ebgl_ = ggp_divisor(ebgl),
                     ebgq_ = ggp_divisor(ebgq), ebt_gl_ = ggp_divisor(ebt * gl), ebgl_gl_ = ggp_divisor(ebgl * gl),
                     two_gq_ = ggp_divisor(2. * gq), two_gq2_ = ggp_divisor(2. * (gq * gq));

    // ---- segment A: the (., c1) and (., c1 - gl t) groups over [0, t] ----
    ggp_eval_groups(S, 0, 4, k, M);
    const GgpInts<1> jB_c1 = ggp_load_ints<1>(S, 0);
    const GgpInts<2> jBm_c1 = ggp_load_ints<2>(S, 2);
    const GgpInts<1> jB_c1l = ggp_load_ints<1>(S, 5);
    const GgpInts<2> jBm_c1l = ggp_load_ints<2>(S, 7);

    // new mean (mean_cov_model.h:73-87); needed by the covariance terms below
    double nm0 = bx + ml * t + (bl - ml) * omegl / gl_;
    double nm1 = bg / ebt_ + Clq * jBm_c1.I[1] + mq * jB_c1.I[0] + (bq + Cxq - mq) * jBm_c1.I[0];
    double nm2 = ml + (bl - ml) * egl;
    double nm3 = mq + (bq - mq) * egq;

    if (cross) {   // mean_cov_model.h:282-377: uses only the two (., c1) groups over [0, t]
        cross[0] = Cxx + Cxl * omegl / gl_;
        cross[1] = Cxg + Cgl * omegl / gl_;
        cross[2] = Cxl + Cll * omegl / gl_;
        cross[3] = Cxq + Clq * omegl / gl_;
        cross[4] = (bg * bx) / ebt_ + Cxg / ebt_ + Cxl * mq * jB_c1.I[1]
                   + (bx * Clq + bq * Cxl + Cxl * Cxq + Clq * Cxx - Cxl * mq) * jBm_c1.I[1]
                   + Clq * Cxl * jBm_c1.I[2] + (bx * mq + Cxx * mq) * jB_c1.I[0]
                   + (bq * bx + Cxq + bx * Cxq + bq * Cxx + Cxq * Cxx - bx * mq - Cxx * mq) * jBm_c1.I[0] - nm1 * bx;
        cross[5] = (bg * bg) / ebt_ + Cgg / ebt_ + Cgl * mq * jB_c1.I[1]
                   + (bq * Cgl + bg * Clq + Clq * Cxg + Cgl * Cxq - Cgl * mq) * jBm_c1.I[1]
                   + Cgl * Clq * jBm_c1.I[2] + (bg * mq + Cxg * mq) * jB_c1.I[0]
                   + (bg * bq + Cgq + bq * Cxg + bg * Cxq + Cxg * Cxq - bg * mq - Cxg * mq) * jBm_c1.I[0] - nm1 * bg;
        cross[6] = (bg * bl) / ebt_ + Cgl / ebt_ + Cll * mq * jB_c1.I[1]
                   + (bq * Cll + bl * Clq + Clq * Cxl + Cll * Cxq - Cll * mq) * jBm_c1.I[1]
                   + Cll * Clq * jBm_c1.I[2] + (bl * mq + Cxl * mq) * jB_c1.I[0]
                   + (bl * bq + Clq + bq * Cxl + bl * Cxq + Cxl * Cxq - bl * mq - Cxl * mq) * jBm_c1.I[0] - nm1 * bl;
        cross[7] = (bg * bq) / ebt_ + Cgq / ebt_ + Clq * mq * jB_c1.I[1]
                   + (2 * bq * Clq + 2 * Clq * Cxq - Clq * mq) * jBm_c1.I[1]
                   + (Clq * Clq) * jBm_c1.I[2] + (bq * mq + Cxq * mq) * jB_c1.I[0]
                   + ((bq * bq) + Cqq + 2 * bq * Cxq + (Cxq * Cxq) - bq * mq - Cxq * mq) * jBm_c1.I[0] - nm1 * bq;
        cross[8] = Cxl * egl; cross[9] = Cgl * egl; cross[10] = Cll * egl; cross[11] = Clq * egl;
        cross[12] = Cxq * egq; cross[13] = Cgq * egq; cross[14] = Clq * egq; cross[15] = Cqq * egq;
    }

    // ---- cov_xg, mean_cov_model.h:97-115 ----
    double n_xg =
        (bg * bx) / ebt_ + Cxg / ebt_ + (bg * bl) / ebt_gl_ + Cgl / ebt_gl_ - (bg * bl) / ebgl_gl_
        - Cgl / ebgl_gl_ - (bg * ml) / ebt_gl_ + (bg * ml) / ebgl_gl_ + (bg * ml * t) / ebt_
        + (Cxl * mq + (Cll * mq) / gl_) * jB_c1.I[1]
        - (Cll * mq * jB_c1l.I[1]) / gl_
        + (bx * Clq + bq * Cxl + Cxl * Cxq + Clq * Cxx + (bq * Cll) / gl_ + (bl * Clq) / gl_ + (Clq * Cxl) / gl_
           + (Cll * Cxq) / gl_ - (Clq * ml) / gl_ - Cxl * mq - (Cll * mq) / gl_ + Clq * ml * t) * jBm_c1.I[1]
        + (-((bq * Cll) / gl_) - (bl * Clq) / gl_ - (Clq * Cxl) / gl_ - (Cll * Cxq) / gl_ + (Clq * ml) / gl_
           + (Cll * mq) / gl_) * jBm_c1l.I[1]
        + (Clq * Cxl + (Cll * Clq) / gl_) * jBm_c1.I[2]
        - (Cll * Clq * jBm_c1l.I[2]) / gl_
        + (bx * mq + Cxx * mq + (bl * mq) / gl_ + (Cxl * mq) / gl_ - (ml * mq) / gl_ + ml * mq * t) * jB_c1.I[0]
        + (-((bl * mq) / gl_) - (Cxl * mq) / gl_ + (ml * mq) / gl_) * jB_c1l.I[0]
        + (bq * bx + Cxq + bx * Cxq + bq * Cxx + Cxq * Cxx + (bl * bq) / gl_ + Clq / gl_ + (bq * Cxl) / gl_
           + (bl * Cxq) / gl_ + (Cxl * Cxq) / gl_ - (bq * ml) / gl_ - (Cxq * ml) / gl_ - bx * mq - Cxx * mq
           - (bl * mq) / gl_ - (Cxl * mq) / gl_ + (ml * mq) / gl_ + bq * ml * t + Cxq * ml * t - ml * mq * t) * jBm_c1.I[0]
        + (-((bl * bq) / gl_) - Clq / gl_ - (bq * Cxl) / gl_ - (bl * Cxq) / gl_ - (Cxl * Cxq) / gl_ + (bq * ml) / gl_
           + (Cxq * ml) / gl_ + (bl * mq) / gl_ + (Cxl * mq) / gl_ - (ml * mq) / gl_) * jBm_c1l.I[0]
        - nm1 * nm0;

    // ---- cov_gl, mean_cov_model.h:166-176 ----
    double n_gl =
        (bg * bl) / ebgl_ + Cgl / ebgl_ + (bg * ml) / ebt_ - (bg * ml) / ebgl_
        + Cll * mq * jB_c1l.I[1] + Clq * ml * jBm_c1.I[1]
        + (bq * Cll + bl * Clq + Clq * Cxl + Cll * Cxq - Clq * ml - Cll * mq) * jBm_c1l.I[1]
        + Cll * Clq * jBm_c1l.I[2] + ml * mq * jB_c1.I[0]
        + (bl * mq + Cxl * mq - ml * mq) * jB_c1l.I[0]
        + (bq * ml + Cxq * ml - ml * mq) * jBm_c1.I[0]
        + (bl * bq + Clq + bq * Cxl + bl * Cxq + Cxl * Cxq - bq * ml - Cxq * ml - bl * mq - Cxl * mq + ml * mq) * jBm_c1l.I[0]
        - nm1 * nm2;

    // ---- segment B: cov_gq, mean_cov_model.h:178-192 ----
    ggp_eval_groups(S, 4, 8, k, M);
    double n_gq;
    {
        const GgpInts<1> jB_c1q = ggp_load_ints<1>(S, 0);
        const GgpInts<2> jBm_c1q = ggp_load_ints<2>(S, 2);
        const GgpInts<0> jBm_c1qw = ggp_load_ints<0>(S, 5);
        const GgpInts<0> jBp_c1qw = ggp_load_ints<0>(S, 6);
        n_gq =
        n_gq =
            (bg * bq) / ebgq_ + Cgq / ebgq_ + (bg * mq) / ebt_ - (bg * mq) / ebgq_
            + Clq * mq * jB_c1q.I[1] + Clq * mq * jBm_c1.I[1]
            + (2 * bq * Clq + 2 * Clq * Cxq - 2 * Clq * mq) * jBm_c1q.I[1]
            + (Clq * Clq) * jBm_c1q.I[2] + (mq * mq) * jB_c1.I[0]
            + (bq * mq + Cxq * mq - (mq * mq)) * jB_c1q.I[0]
            + (bq * mq + Cxq * mq - (mq * mq)) * jBm_c1.I[0]
            - (sq2 * jBm_c1qw.I[0]) / two_gq_
            + ((bq * bq) + Cqq + 2 * bq * Cxq + (Cxq * Cxq) - 2 * bq * mq - 2 * Cxq * mq + (mq * mq)) * jBm_c1q.I[0]
            + (sq2 * jBp_c1qw.I[0]) / two_gq_
            - nm1 * nm3;
    }

    // ---- segment C: cov_gg, mean_cov_model.h:124-164 ----
    ggp_eval_groups(S, 8, 17, k, M);
    double n_gg;
    {
        const GgpInts<1> jB_c2 = ggp_load_ints<1>(S, 0);
        const GgpInts<2> jBm_c2 = ggp_load_ints<2>(S, 2);
        const GgpInts<1> jW_lo = ggp_load_ints<1>(S, 5), jW_hi = ggp_load_ints<1>(S, 7);
        const GgpInts<3> jWm_lo = ggp_load_ints<3>(S, 9), jWm_hi = ggp_load_ints<3>(S, 13);
        const GgpInts<1> jW_d2 = ggp_load_ints<1>(S, 17);
        const GgpInts<1> jWm_d3 = ggp_load_ints<1>(S, 19);
        const GgpInts<0> jWp_d4 = ggp_load_ints<0>(S, 21);
        const double mq2 = mq * mq, bq2 = bq * bq, Cxq2 = Cxq * Cxq, Clq2 = Clq * Clq;
        n_gg =
            ((bg * bg) + Cgg) / e2bt
            + 2 * Cgl * mq * jB_c2.I[1]
            + (mq * (2 * Clq + gq * mq) * jW_lo.I[1]) / gq_
            + 2 * (bq * Cgl + bg * Clq + Clq * Cxg + Cgl * Cxq - Cgl * mq) * jBm_c2.I[1]
            + ((bq2 * gq + Cqq * gq + 4 * bq * Cxq * gq + 4 * Cxq2 * gq - 2 * Clq * mq - 2 * bq * gq * mq
                - 4 * Cxq * gq * mq + gq * mq2) * jWm_lo.I[1]) / gq_
            - mq2 * jW_hi.I[1]
            - (2 * Clq * mq * jW_d2.I[1]) / gq_
            - (sq2 * jWm_lo.I[1]) / two_gq_
            + (sq2 * jWm_hi.I[1]) / two_gq_
            + (-bq2 - Cqq - 4 * bq * Cxq - 4 * Cxq2 + 2 * bq * mq + 4 * Cxq * mq - mq2 + 4 * bq * Clq * t
               + 8 * Clq * Cxq * t - 4 * Clq * mq * t) * jWm_hi.I[1]
            + (2 * Clq * mq * jWm_d3.I[1]) / gq_
            + Clq2 * jWm_lo.I[3]
            - Clq2 * jWm_hi.I[3]
            + 2 * Cgl * Clq * jBm_c2.I[2]
            + (2 * bq * Clq + 4 * Clq * Cxq - 2 * Clq * mq) * jWm_lo.I[2]
            + (-2 * bq * Clq - 4 * Clq * Cxq + 2 * Clq * mq + 2 * Clq2 * t) * jWm_hi.I[2]
            + (2 * bg * mq + 2 * Cxg * mq) * jB_c2.I[0]
            + ((2 * bq * mq) / gq_ + (4 * Cxq * mq) / gq_ - (2 * mq2) / gq_) * jW_lo.I[0]
            + (2 * bg * bq + 2 * Cgq + 2 * bq * Cxg + 2 * bg * Cxq + 2 * Cxg * Cxq - 2 * bg * mq - 2 * Cxg * mq) * jBm_c2.I[0]
            + ((-2 * bq * mq) / gq_ - (4 * Cxq * mq) / gq_ + (2 * mq2) / gq_) * jWm_lo.I[0]
            + (sq2 * jW_lo.I[0]) / two_gq2_
            + (sq2 * jW_hi.I[0]) / two_gq2_
            + 2 * mq2 * t * jW_hi.I[0]
            + ((-2 * bq * mq) / gq_ - (4 * Cxq * mq) / gq_ + (2 * mq2) / gq_) * jW_d2.I[0]
            - (sq2 * jWm_lo.I[0]) / two_gq2_
            - (sq2 * t * jWm_hi.I[0]) / gq_
            + (2 * bq2 * t + 2 * Cqq * t + 8 * bq * Cxq * t + 8 * Cxq2 * t - 4 * bq * mq * t - 8 * Cxq * mq * t
               + 2 * mq2 * t) * jWm_hi.I[0]
            + ((2 * bq * mq) / gq_ + (4 * Cxq * mq) / gq_ - (2 * mq2) / gq_) * jWm_d3.I[0]
            - (sq2 * jWp_d4.I[0]) / two_gq2_
            - (nm1 * nm1);
    }

    // ---- the elementary block, mean_cov_model.h:93-95, 117-122, 196-208 ----
    const double egl2 = egl * egl, egq2 = egq * egq;
    double n_xx = Cll * (omegl * omegl) / (gl * gl) + 2 * Cxl * omegl / gl_ + Cxx
                  + sl2 / (2 * ggp_pow(gl, 3.0, M)) * (2 * gl * t - 3 + 4 * egl - egl2);
    double n_xl = sl2 / (2 * (gl * gl)) * (omegl * omegl) + Cll * egl * omegl / gl_ + Cxl * egl;
    double n_xq = Clq * omegl * egq / gl_ + Cxq * egq;
    double n_ll = Cll * egl2 + sl2 / (2 * gl) * (1 - egl2);
    double n_lq = Clq * egl * egq;
    double n_qq = sq2 / two_gq_ * (1 - egq2) + Cqq * egq2;

    s.m[0] = nm0; s.m[1] = nm1; s.m[2] = nm2; s.m[3] = nm3;
    s.c[0] = n_xx; s.c[1] = n_xg; s.c[2] = n_xl; s.c[3] = n_xq; s.c[4] = n_gg;
    s.c[5] = n_gl; s.c[6] = n_gq; s.c[7] = n_ll; s.c[8] = n_lq; s.c[9] = n_qq;
}

GGP_HD void ggp_propagate(GgpState& s, double t, const GgpOuParams& p, const GgpMathTables* __restrict__ M,
                          const GgpScratch& S) {
    ggp_propagate_impl(s, t, p, M, S, nullptr);
}
