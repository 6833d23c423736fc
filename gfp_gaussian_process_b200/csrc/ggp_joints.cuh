// ggp_joints.cuh — pairwise joint posteriors P(z_{n+m}, z_n | D) along the lineage (the reference's -j mode).
//
// Replaces, from the reference (paths under src/):
//   Gaussian / Affine_gaussian / Seperated_gaussian, seperate_gaussian, flip_xy      Gaussians.h:24-158
//   include_measurement                                                              correlation_tree.h:132-154
//   consecutive_joint[_cell_division], consecutive_conditional[_cell_division]       correlation_tree.h:160-396
//   calc_x / calc_X / calc_Y / propagation / next_joint                              correlation_tree.h:403-454
//   incorporate_backward_prob, crosscovariance_is_small                              correlation_tree.h:457-493
//   calc_joint_distributions, joint_distributions_recr, sc_joint_distributions       correlation_tree.h:499-626
//
// The reference recomputes, for every start point n, the conditional P(z_{k+1} | z_k, D_k) of every later
// point k it passes (one mean_cov_model + one cross_cov_model each time) and keeps 8x8 Eigen matrices on
// the heap.  Here the two quantities that depend on a single point only — the first joint
// J0(k) = P(z_{k+1}, z_k | D_k) and the transformed conditional — are computed ONCE per cell-timepoint by a
// first kernel (ggp_ctp_joint_prep: one propagation step with cross covariance per point) and cached in HBM,
// together with the two other per-point quantities every passing walk would recompute: the inverse of the
// conditional's F and the backward prediction divided by the prior.  A second, persistent kernel walks the
// start points: a thread takes the next start point from a device counter, walks forward in time and
// depth-first into both daughters with the 8-dim Gaussian in registers/local memory and an explicit stack of
// pending daughter branches, appends the joints it emits to a sparse list, and takes the next start point when
// its walk ends - the walk is one flat loop, one point per iteration, so the lanes of a warp stay on the
// step body together whatever the cell boundaries and walk lengths are.
// Eigen semantics kept: dynamic inverses = partial-pivot LU (also the 2x2 of
// include_measurement, which inverts a MatrixXd), dynamic products coefficient-wise left to right, matrix *
// vector = column-major GEMV (four columns at a time, remaining columns one by one).
// Host+device; compile with FMA contraction off.
#pragma once
#include "ggp_cell.cuh"

struct GgpGauss4 { double m[4]; double C[16]; };
struct GgpAffine { double a[4]; double F[16]; double A[16]; };   // N(y | a + F x, A)
struct GgpGauss8 { double m[8]; double C[64]; };
struct GgpSep { GgpGauss4 marg; GgpAffine cond; };

#define GGP_JOINT_PREP 144   // doubles cached per ctp: J0 (8 + 64), the conditional (4 + 16 + 16), F^-1 (16), backward / prior (4 + 16)
#define GGP_JP_COND 72
#define GGP_JP_FINV 108
#define GGP_JP_BW 124

// C = op(A) * op(B), 4x4, coefficient-wise, inner sum left to right
template <bool TA, bool TB>
GGP_HD void ggp_mm4(const double* __restrict__ A, const double* __restrict__ B, double* __restrict__ C) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            double s = (TA ? A[i] : A[4 * i]) * (TB ? B[4 * j] : B[j]);
#pragma unroll
            for (int k = 1; k < 4; ++k) s = s + (TA ? A[4 * k + i] : A[4 * i + k]) * (TB ? B[4 * j + k] : B[4 * k + j]);
            C[4 * i + j] = s;
        }
}

GGP_HD void ggp_gemv4t(const double* __restrict__ A, const double* __restrict__ x, double* __restrict__ y) {   // y = A x
    ggp_gemv4(A, x, y);
}

// seperate_gaussian (Gaussians.h:127-145): N([x y] | m, C) -> N(x | a, A) N(y | b' + F x, B')
GGP_HD void ggp_separate(const GgpGauss8& J, GgpSep& S) {
    double A[16], K[16], B[16];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            A[4 * i + j] = J.C[8 * i + j];
            K[4 * i + j] = J.C[8 * i + 4 + j];
            B[4 * i + j] = J.C[8 * (4 + i) + 4 + j];
        }
    double Ai[16], KtAi[16], t[4], KAK[16];
    ggp_inv_lu<4>(A, Ai);
    ggp_mm4<true, false>(K, Ai, KtAi);
    ggp_gemv4(KtAi, J.m, t);
    ggp_mm4<false, false>(KtAi, K, KAK);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        S.marg.m[i] = J.m[i];
        S.cond.a[i] = J.m[4 + i] - t[i];
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        S.marg.C[i] = A[i];
        S.cond.F[i] = KtAi[i];
        S.cond.A[i] = B[i] - KAK[i];
    }
}

// Seperated_gaussian::to_joint (Gaussians.h:109-124): N(x | m, C) N(y | a + F x, A) -> N([x y] | ., .)
GGP_HD void ggp_to_joint(const GgpGauss4& g, const GgpAffine& c, GgpGauss8& J) {
    double Fm[4], CtFt[16], FC[16], FCt[16], FCtFt[16];
    ggp_gemv4(c.F, g.m, Fm);
    ggp_mm4<true, true>(g.C, c.F, CtFt);
    ggp_mm4<false, false>(c.F, g.C, FC);
    ggp_mm4<false, true>(c.F, g.C, FCt);
    ggp_mm4<false, true>(FCt, c.F, FCtFt);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        J.m[i] = g.m[i];
        J.m[4 + i] = c.a[i] + Fm[i];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            J.C[8 * i + j] = g.C[4 * i + j];
            J.C[8 * i + 4 + j] = CtFt[4 * i + j];
            J.C[8 * (4 + i) + j] = FC[4 * i + j];
            J.C[8 * (4 + i) + 4 + j] = c.A[4 * i + j] + FCtFt[4 * i + j];
        }
}

// Gaussian::flip_xy (Gaussians.h:147-158)
GGP_HD void ggp_flip_xy(const GgpGauss8& J, GgpGauss8& R) {
#pragma unroll
    for (int i = 0; i < 4; ++i) { R.m[i] = J.m[4 + i]; R.m[4 + i] = J.m[i]; }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            R.C[8 * i + j] = J.C[8 * (4 + i) + 4 + j];
            R.C[8 * i + 4 + j] = J.C[8 * j + 4 + i];          // C.block(0,n,n,n)^T
            R.C[8 * (4 + i) + j] = J.C[8 * (4 + j) + i];      // C.block(n,0,n,n)^T
            R.C[8 * (4 + i) + 4 + j] = J.C[8 * i + j];
        }
}

// Gaussian::multiply (Gaussians.h:42-49)
GGP_HD void ggp_gauss_multiply(const GgpGauss4& n1, const double* __restrict__ m2, const double* __restrict__ C2, GgpGauss4& out) {
    double S[16], Si[16], C2Si[16], C1Si[16], a[4], b[4];
#pragma unroll
    for (int i = 0; i < 16; ++i) S[i] = n1.C[i] + C2[i];
    ggp_inv_lu<4>(S, Si);
    ggp_mm4<false, false>(C2, Si, C2Si);
    ggp_mm4<false, false>(n1.C, Si, C1Si);
    ggp_gemv4(C2Si, n1.m, a);
    ggp_gemv4(C1Si, m2, b);
#pragma unroll
    for (int i = 0; i < 4; ++i) out.m[i] = a[i] + b[i];
    ggp_mm4<false, false>(C1Si, C2, out.C);
}

// Affine_gaussian::transform() (Gaussians.h:71-81): N(y | a + F x, A) -> N(x | a' + F' y, A')
GGP_HD void ggp_affine_transform(const GgpAffine& c, GgpAffine& out) {
    double Fi[16], t[4], FiA[16];
    ggp_inv_lu<4>(c.F, Fi);
    ggp_gemv4(Fi, c.a, t);
    ggp_mm4<false, false>(Fi, c.A, FiA);
#pragma unroll
    for (int i = 0; i < 4; ++i) out.a[i] = -t[i];
#pragma unroll
    for (int i = 0; i < 16; ++i) out.F[i] = Fi[i];
    ggp_mm4<false, true>(FiA, Fi, out.A);
}

// include_measurement (correlation_tree.h:132-154) on the 8-dim joint; D = diag(D00, D11)
GGP_HD void ggp_include_measurement(GgpGauss8& J, double D00, double D11, double x, double g) {
    double S[4] = {J.C[0] + D00, J.C[1] + 0.0, J.C[8] + 0.0, J.C[9] + D11}, Si[4];
    const double xg0 = x - J.m[0], xg1 = g - J.m[1];
    ggp_inv_lu<2>(S, Si);
    double T0[8], T1[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        T0[i] = J.C[i] * Si[0] + J.C[8 + i] * Si[2];
        T1[i] = J.C[i] * Si[1] + J.C[8 + i] * Si[3];
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) J.m[i] = J.m[i] + ((0.0 + T0[i] * xg0) + T1[i] * xg1);
    double K0[8], K1[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { K0[j] = J.C[j]; K1[j] = J.C[8 + j]; }
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) J.C[8 * i + j] = J.C[8 * i + j] - (T0[i] * K0[j] + T1[i] * K1[j]);
}

// next_joint (correlation_tree.h:426-454): P(z_{k+1}, z_n | D_{k+1}) x P(z_{k+2} | z_{k+1}, D_{k+1}) -> P(z_{k+2}, z_n | D_{k+1}).
// sep = seperate_gaussian of the joint after include_measurement (shared with incorporate_backward_prob, which
// separates the same joint); c = the cached block of the point: cond.a [4], cond.F [16], cond.A [16], cond.F^-1 [16]
GGP_HD void ggp_next_joint(const GgpSep& sep, const double* __restrict__ c, GgpGauss8& out) {
    const double* __restrict__ ca = c;
    const double* __restrict__ cF = c + 4;
    const double* __restrict__ cA = c + 20;
    const double* __restrict__ Fi = c + 36;
    // calc_x / calc_X / calc_Y (correlation_tree.h:403-415)
    double S[16], Si[16], ASi[16], CSi[16], u[4], v[4];
    GgpAffine NX;
#pragma unroll
    for (int i = 0; i < 16; ++i) S[i] = sep.marg.C[i] + cA[i];
    ggp_inv_lu<4>(S, Si);
    ggp_mm4<false, false>(cA, Si, ASi);
    ggp_mm4<false, false>(sep.marg.C, Si, CSi);
    ggp_gemv4(ASi, sep.marg.m, u);
    ggp_gemv4(CSi, ca, v);
#pragma unroll
    for (int i = 0; i < 4; ++i) NX.a[i] = u[i] + v[i];
    ggp_mm4<false, false>(CSi, cF, NX.F);
    ggp_mm4<false, false>(CSi, cA, NX.A);
    // G.transform(marginal.m) (Gaussians.h:83-87) with G = (cond.a, cond.F, marginal.C + cond.A)
    GgpGauss4 nm;
    {
        double d[4], FiA[16];
#pragma unroll
        for (int i = 0; i < 4; ++i) d[i] = sep.marg.m[i] - ca[i];
        ggp_gemv4(Fi, d, nm.m);
        ggp_mm4<false, false>(Fi, S, FiA);
        ggp_mm4<false, true>(FiA, Fi, nm.C);
    }
    // propagation(sep.conditional, NX) (correlation_tree.h:418-423)
    GgpAffine nc;
    {
        double Fx[4], FA[16], FAFt[16];
        ggp_gemv4(sep.cond.F, NX.a, Fx);
#pragma unroll
        for (int i = 0; i < 4; ++i) nc.a[i] = sep.cond.a[i] + Fx[i];
        ggp_mm4<false, false>(sep.cond.F, NX.F, nc.F);
        ggp_mm4<false, false>(sep.cond.F, NX.A, FA);
        ggp_mm4<false, true>(FA, sep.cond.F, FAFt);
#pragma unroll
        for (int i = 0; i < 16; ++i) nc.A[i] = sep.cond.A[i] + FAFt[i];
    }
    ggp_to_joint(nm, nc, out);
}

// incorporate_backward_prob (correlation_tree.h:457-482); bw = the backward prediction at the point already
// divided by the prior (divide_by_prior depends on the point only: cached by the preparation), mean [4] then cov [16]
GGP_HD void ggp_incorporate_backward(const GgpSep& sep, const double* __restrict__ bw, GgpGauss8& out) {
    GgpGauss4 marg;
    ggp_gauss_multiply(sep.marg, bw, bw + 4, marg);
    ggp_to_joint(marg, sep.cond, out);
}

// crosscovariance_is_small (correlation_tree.h:484-493); a NaN ratio counts as small, as there
GGP_HD bool ggp_crosscov_small(const GgpGauss8& J, double tol) {
    bool small = true;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 4; j < 8; ++j)
            if (fabs(J.C[8 * i + j] / (J.m[i] * J.m[j])) > tol) small = false;
    return small;
}

// ------------------------------------------------------------------------------------------------
// per-ctp preparation: J0(k) = P(z_{k+1}, z_k | D_k) (consecutive_joint, correlation_tree.h:325-357, or
// consecutive_joint_cell_division :160-238 at a cell's last point) and the transformed conditional
// (consecutive_conditional :360-396 / _cell_division :241-319), from the forward posterior at k.
// prep = [J0.m 8][J0.C 64][cond.a 4][cond.F 16][cond.A 16][cond.F^-1 16][bw.m 4][bw.C 16]; bw = the backward
// prediction at k divided by the prior (incorporate_backward_prob's first step, correlation_tree.h:466-468);
// only bw is written for the last point of a leaf.
// ------------------------------------------------------------------------------------------------
struct GgpJointArgs {
    const double* params;    // [n_seg][11]
    const double* fwd;       // [n_ctp][20]
    const double* bwd;       // [n_ctp][20]
    const double* bstate;    // [n_cells][20] by slot
    double* prep;            // [n_ctp][GGP_JOINT_PREP]
    const int32_t* ctp_slot; // [n_ctp] slot of the cell a ctp belongs to
    double tol;
    // output list
    int64_t row_begin, row_end;   // start points handled by this launch (ctp range)
    long long cap;
    unsigned long long* count;
    long long* row_ctp;
    long long* col_ctp;
    double* rec44;
    unsigned long long* next_row;   // start points handed out so far (relative to row_begin)
    double* stack;           // [n_threads][stack_depth][72] pending daughter branches
    int32_t* stack_slot;     // [n_threads][stack_depth]
    int stack_depth;
};

GGP_HD void ggp_ctp_joint_prep(const GgpDevForest& F, const GgpJointArgs& A, int64_t k, const GgpMathTables* __restrict__ T,
                               const GgpScratch& S) {
    const int slot = A.ctp_slot[k];
    const int64_t off = F.s_off[slot];
    const int n = F.s_n[slot];
    const int t = (int)(k - off);
    const bool last = (t == n - 1);
    const int d1 = F.s_d1[slot];
    const double* p = A.params + GGP_NP * F.seg[k];
    double* out = A.prep + (int64_t)GGP_JOINT_PREP * k;
    {
        GgpGauss4 bw;
        const double* b = A.bwd + 20 * k;
#pragma unroll
        for (int i = 0; i < 4; ++i) bw.m[i] = b[i];
#pragma unroll
        for (int i = 0; i < 16; ++i) bw.C[i] = b[4 + i];
        ggp_divide_by_prior(bw.m, bw.C, p);
#pragma unroll
        for (int i = 0; i < 4; ++i) out[GGP_JP_BW + i] = bw.m[i];
#pragma unroll
        for (int i = 0; i < 16; ++i) out[GGP_JP_BW + 4 + i] = bw.C[i];
    }
    if (last && d1 < 0) return;   // joints starting at / passing a leaf's last point go nowhere
    const double* f = A.fwd + 20 * k;
    double mean1[4], cov1[16];
#pragma unroll
    for (int i = 0; i < 4; ++i) mean1[i] = f[i];
#pragma unroll
    for (int i = 0; i < 16; ++i) cov1[i] = f[4 + i];
    GgpState s;
    ggp_state_from16(s, mean1, cov1);
    GgpGauss8 J0;
    GgpAffine cond;
    if (!last) {
        double cross[16], cov2[16];
        const double dt = F.time[k + 1] - F.time[k];
        ggp_propagate_impl(s, dt, ggp_ou(p, false), T, S, cross);
        ggp_state_to16(s, cov2);
        GgpGauss8 J2;   // [z_n, z_n+1] for the conditional
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            J0.m[i] = s.m[i]; J0.m[4 + i] = mean1[i];
            J2.m[i] = mean1[i]; J2.m[4 + i] = s.m[i];
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                J0.C[8 * i + j] = cov2[4 * i + j];
                J0.C[8 * i + 4 + j] = cross[4 * i + j];
                J0.C[8 * (4 + i) + j] = cross[4 * j + i];
                J0.C[8 * (4 + i) + 4 + j] = cov1[4 * i + j];
                J2.C[8 * i + j] = cov1[4 * i + j];
                J2.C[8 * i + 4 + j] = cross[4 * j + i];
                J2.C[8 * (4 + i) + j] = cross[4 * i + j];
                J2.C[8 * (4 + i) + 4 + j] = cov2[4 * i + j];
            }
        GgpSep sep;
        ggp_separate(J2, sep);
        ggp_affine_transform(sep.cond, cond);
    } else {
        const double var_dx = p[9], var_dg = p[10];
        const double log2v = GGP_LOG2;
        if (F.model.division_binomial) {
            const double dt = F.time[F.s_off[d1]] - F.time[k];
            ggp_propagate(s, dt, ggp_ou(p, false), T, S);
            ggp_divide(s, var_dx, var_dg, F.model);
            double cov2[16], cross[16];
            ggp_state_to16(s, cov2);
#pragma unroll
            for (int i = 0; i < 16; ++i) cross[i] = cov1[i];
#pragma unroll
            for (int j = 0; j < 4; ++j) cross[4 + j] = cross[4 + j] / 2.;
            GgpGauss8 J2;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                J0.m[i] = s.m[i]; J0.m[4 + i] = mean1[i];
                J2.m[i] = mean1[i]; J2.m[4 + i] = s.m[i];
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    J0.C[8 * i + j] = cov2[4 * i + j];
                    J0.C[8 * i + 4 + j] = cross[4 * i + j];
                    J0.C[8 * (4 + i) + j] = cross[4 * j + i];
                    J0.C[8 * (4 + i) + 4 + j] = cov1[4 * i + j];
                    J2.C[8 * i + j] = cov1[4 * i + j];
                    J2.C[8 * i + 4 + j] = cross[4 * j + i];
                    J2.C[8 * (4 + i) + j] = cross[4 * i + j];
                    J2.C[8 * (4 + i) + 4 + j] = cov2[4 * i + j];
                }
            GgpSep sep;
            ggp_separate(J2, sep);
            ggp_affine_transform(sep.cond, cond);
        } else {
            // gauss: the model's own conditional N(z_n+1 | f + F z_n, D) (correlation_tree.h:224-236, :302-318)
            GgpAffine c;
#pragma unroll
            for (int i = 0; i < 16; ++i) { c.F[i] = 0.0; c.A[i] = 0.0; }
            c.F[0] = 1.0; c.F[5] = 0.5; c.F[10] = 1.0; c.F[15] = 1.0;
            c.A[0] = var_dx; c.A[5] = var_dg;
            c.a[0] = -log2v; c.a[1] = 0.0; c.a[2] = 0.0; c.a[3] = 0.0;
            GgpGauss4 marg;
#pragma unroll
            for (int i = 0; i < 4; ++i) marg.m[i] = mean1[i];
#pragma unroll
            for (int i = 0; i < 16; ++i) marg.C[i] = cov1[i];
            GgpGauss8 Jyx;
            ggp_to_joint(marg, c, Jyx);
            ggp_flip_xy(Jyx, J0);
            ggp_affine_transform(c, cond);
        }
    }
    double Fi[16];
    ggp_inv_lu<4>(cond.F, Fi);   // next_joint's G.transform inverts the conditional's F at every pass (Gaussians.h:83-87)
#pragma unroll
    for (int i = 0; i < 8; ++i) out[i] = J0.m[i];
#pragma unroll
    for (int i = 0; i < 64; ++i) out[8 + i] = J0.C[i];
#pragma unroll
    for (int i = 0; i < 4; ++i) out[GGP_JP_COND + i] = cond.a[i];
#pragma unroll
    for (int i = 0; i < 16; ++i) { out[GGP_JP_COND + 4 + i] = cond.F[i]; out[GGP_JP_COND + 20 + i] = cond.A[i]; out[GGP_JP_FINV + i] = Fi[i]; }
}

// ------------------------------------------------------------------------------------------------
// one start point (row): sc_joint_distributions' body for one n (correlation_tree.h:598-616) with
// joint_distributions_recr / calc_joint_distributions unrolled into a loop with an explicit stack
// ------------------------------------------------------------------------------------------------
GGP_HD void ggp_emit_joint(const GgpJointArgs& A, int64_t row, int64_t col, const GgpGauss8& J) {
#if defined(__CUDA_ARCH__)
    const unsigned long long idx = atomicAdd(A.count, 1ull);
#else
    const unsigned long long idx = (*A.count)++;
#endif
    if ((long long)idx >= A.cap) return;
#if defined(__CUDA_ARCH__)
    // streaming stores: a walker keeps ~600 bytes of spilled state in L1, records are written once and read by another kernel
    __stcs(A.row_ctp + idx, (long long)row);
    __stcs(A.col_ctp + idx, (long long)col);
#else
    A.row_ctp[idx] = row;
    A.col_ctp[idx] = col;
#endif
    double* r = A.rec44 + 44 * idx;
    double v[44];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = J.m[i];
    int q = 8;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = i; j < 8; ++j) v[q++] = J.C[8 * i + j];
#if defined(__CUDA_ARCH__)
    // a record is 352 bytes at a multiple of 352 from a 256-byte aligned base: 22 sixteen-byte stores instead of 44 eight-byte ones
    double2* r2 = reinterpret_cast<double2*>(r);
#pragma unroll
    for (int i = 0; i < 22; ++i) __stcs(r2 + i, make_double2(v[2 * i], v[2 * i + 1]));
#else
    for (int i = 0; i < 44; ++i) r[i] = v[i];
#endif
}

GGP_HD void ggp_load_joint(const double* __restrict__ src, GgpGauss8& J) {
    for (int i = 0; i < 8; ++i) J.m[i] = src[i];
    for (int i = 0; i < 64; ++i) J.C[i] = src[8 + i];
}
GGP_HD void ggp_store_joint(double* __restrict__ dst, const GgpGauss8& J) {
    for (int i = 0; i < 8; ++i) dst[i] = J.m[i];
    for (int i = 0; i < 64; ++i) dst[8 + i] = J.C[i];
}

// All walkers of a block start each point together (device): the step body is ~7 000 instructions of straight-line
// code, and warps spread over it each fetch it on their own.  Returns whether any walker of the block is still walking.
GGP_HD bool ggp_walkers_align(bool walking) {
#if defined(__CUDA_ARCH__)
    return __syncthreads_or(walking) != 0;
#else
    return walking;
#endif
}

GGP_HD int64_t ggp_next_start_point(const GgpJointArgs& A) {
#if defined(__CUDA_ARCH__)
    return A.row_begin + (int64_t)atomicAdd(A.next_row, 1ull);
#else
    return A.row_begin + (int64_t)(*A.next_row)++;
#endif
}

// Walks start points [row_begin, row_end) handed out by A.next_row until none is left.  `lane` selects this
// thread's stack of pending daughter branches.
GGP_HD void ggp_walk_start_points(const GgpDevForest& F, const GgpJointArgs& A, int64_t lane) {
    double* stack = A.stack + lane * A.stack_depth * 72;
    int32_t* stack_slot = A.stack_slot + lane * A.stack_depth;
    GgpGauss8 J, comb;
    int64_t row = 0, off = 0;
    int slot = 0, t = 0, n = 0, sp = 0, d1 = -1;
    double stale_g = 0.0;
    bool walking = false, drained = false;
    for (;;) {
        // ---- settle on the next point to process: next start point when the walk has ended, daughters or a
        //      pending branch at the end of a cell (joint_distributions_recr, correlation_tree.h:566-585) ----
        while (!drained && (!walking || t >= n)) {
            bool enter = false;
            if (!walking) {
                row = ggp_next_start_point(A);
                if (row >= A.row_end) { drained = true; break; }
                const int slot0 = A.ctp_slot[row];
                const int n0 = (int)(row - F.s_off[slot0]);
                const bool at_division = (n0 == F.s_n[slot0] - 1);
                // the last point of a leaf has an empty row; consecutive_joint_cell_division needs daughter1
                // (correlation_tree.h:606-612), both daughters start from it
                if (at_division && F.s_d1[slot0] < 0) continue;
                ggp_load_joint(A.prep + (int64_t)GGP_JOINT_PREP * row, J);
                sp = 0;
                walking = true;
                if (!at_division) {
                    slot = slot0;
                    t = n0 + 1;
                } else {
                    if (F.s_d2[slot0] >= 0) {
                        ggp_store_joint(stack, J);
                        stack_slot[0] = F.s_d2[slot0];
                        sp = 1;
                    }
                    slot = F.s_d1[slot0];
                    t = 0;
                }
                enter = true;
            } else {
                const int d2 = F.s_d2[slot];
                if (d1 >= 0) {   // daughter1 first, daughter2 continues from the same joint later
                    if (d2 >= 0 && sp < A.stack_depth) {
                        ggp_store_joint(stack + (int64_t)sp * 72, J);
                        stack_slot[sp] = d2;
                        ++sp;
                    }
                    slot = d1;
                    t = 0;
                    enter = true;
                } else if (d2 >= 0) {   // daughter2 without daughter1: unreachable with build_cell_genealogy
                    slot = d2;
                    t = 0;
                    enter = true;
                } else if (sp > 0) {
                    --sp;
                    ggp_load_joint(stack + (int64_t)sp * 72, J);
                    slot = stack_slot[sp];
                    t = 0;
                    enter = true;
                } else {
                    walking = false;
                }
            }
            if (enter) {
                off = F.s_off[slot];
                n = F.s_n[slot];
                d1 = F.s_d1[slot];
                stale_g = A.bstate[20 * (int64_t)slot + 1];   // MOMAdata::mean(1) left by the backward pass
            }
        }
        if (!ggp_walkers_align(walking)) return;
        if (!walking) continue;   // out of start points; the block's other walkers are still busy
        // ---- one point of calc_joint_distributions (correlation_tree.h:499-558) ----
        const int64_t k = off + t;
        const double* p = A.params + GGP_NP * F.seg[k];
        const double* c = A.prep + (int64_t)GGP_JOINT_PREP * k;
        const double D11 = F.model.noise_scaled ? p[8] * (stale_g + F.model.fp_auto) : p[8];
        ggp_include_measurement(J, p[7], D11, F.x[k], F.g[k]);
        GgpSep sep;
        ggp_separate(J, sep);
        ggp_incorporate_backward(sep, c + GGP_JP_BW, comb);
        if (ggp_crosscov_small(comb, A.tol)) {
            // this branch ends here; a pending daughter2 branch continues
            if (sp > 0) {
                --sp;
                ggp_load_joint(stack + (int64_t)sp * 72, J);
                slot = stack_slot[sp];
                t = 0;
                off = F.s_off[slot];
                n = F.s_n[slot];
                d1 = F.s_d1[slot];
                stale_g = A.bstate[20 * (int64_t)slot + 1];
            } else {
                walking = false;
            }
            continue;
        }
        ggp_emit_joint(A, row, k, comb);
        if (t < n - 1 || d1 >= 0) ggp_next_joint(sep, c + GGP_JP_COND, J);
        ++t;
    }
}

#if defined(__CUDACC__)
// ---- kernels: one thread per ctp (preparation), persistent threads taking start points from a counter (walk) ----
// resident blocks loop over the points and start every round together: the body is ~16 000 instructions of
// straight-line code, warps that drift apart each fetch it on their own
__global__ void __launch_bounds__(GGP_BLOCK) ggp_joint_prep_kernel(const GgpDevForest F, const GgpJointArgs A) {
    GgpMathTables& T = *reinterpret_cast<GgpMathTables*>(ggp_smem);
    ggp_stage_tables(&T);
    for (int64_t base = (int64_t)blockIdx.x * GGP_BLOCK; base < F.n_ctp; base += (int64_t)gridDim.x * GGP_BLOCK) {
        __syncthreads();
        const int64_t k = base + threadIdx.x;
        if (k < F.n_ctp) ggp_ctp_joint_prep(F, A, k, &T, ggp_thread_scratch());
    }
}

#ifndef GGP_WALK_BLOCK
#define GGP_WALK_BLOCK 256
#endif
__global__ void __launch_bounds__(GGP_WALK_BLOCK, 1) ggp_joint_walk_kernel(const GgpDevForest F, const GgpJointArgs A) {
    ggp_walk_start_points(F, A, (int64_t)blockIdx.x * GGP_WALK_BLOCK + threadIdx.x);
}
#endif
