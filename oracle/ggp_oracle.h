/* oracle/ggp_oracle.h — C interface of the CPU oracle.
 *
 * TEST INFRASTRUCTURE.  The oracle is a plain, single-threaded C++ restatement of the reference's
 * algorithm (bjks/gfp_gaussian_process, src/{mean_cov_model,predictions,likelihood,correlation_tree,
 * Gaussians,moma_input}.h) used ONLY by tests/, __graft_entry__.smoke() and bench.py's CPU baseline
 * to check the CUDA path.  Nothing in gfp_gaussian_process_b200/ may include, link or call it.
 *
 * Parity status: the arithmetic core (integrals, 14 moments, 16 cross-covariances, Dawson) is pinned
 * bit-for-bit against the reference's own sources compiled in oracle/_ref (tests/test_oracle_vs_ref.py)
 * and Dawson additionally against the 48 Maple values of Faddeeva.cc:2380-2512.  The Eigen-dependent
 * wrappers (filter step, division, backward pass, combine, joints) are restated from the reference
 * source with Eigen 3.3's algorithm choices (fixed 2x2 inverse = cofactor; dynamic determinant/inverse
 * = partial-pivot LU; products left to right); Eigen itself is not installable here, so that part is
 * "parity unpinned by the reference" (DESIGN.md).
 */
#ifndef GGP_ORACLE_H
#define GGP_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    long n_cells;
    long n_ctp;
    const long* cell_offset;   /* [n_cells+1] first ctp of each cell (cells in input-file order) */
    const int* parent;         /* [n_cells] index or -1   (moma_input.h:125-151) */
    const int* daughter1;      /* [n_cells] first matching cell in file order, or -1 */
    const int* daughter2;      /* [n_cells] second, or -1 */
    const double* time;        /* [n_ctp] */
    const double* log_length;  /* [n_ctp] */
    const double* fp;          /* [n_ctp] */
    const int* segment;        /* [n_ctp] */
    int noise_model;           /* 0 = const, 1 = scaled   (likelihood.h:59-64) */
    int division_model;        /* 0 = gauss, 1 = binomial (predictions.h:40-60) */
    double fp_auto;
    double init_f[4];          /* mean_x0, mean_g0, var_x0, var_g0 of first points (moma_input.h:675-704) */
    double init_r[4];          /* same of last points (moma_input.h:706-735) */
} ggp_oracle_forest;

/* init_cells_f / init_cells_r statistics (moma_input.h:663-735) into f->init_f / init_r */
void ggp_oracle_init_stats(ggp_oracle_forest* f);

/* one propagation step (mean_cov_model.h:211-274).  cov is row-major 4x4, only the upper triangle is read. */
void ggp_oracle_mean_cov_model(const double* mean, const double* cov, double t, const double* p7,
                               double* mean_out, double* cov_out);
/* cross covariance (mean_cov_model.h:380-432), row-major 4x4 */
void ggp_oracle_cross_cov_model(const double* mean, const double* cov, double t, const double* p7, double* out);
double ggp_oracle_dawson(double x);
double ggp_oracle_tauint(int k, double a, double b, double c, double t1, double t0);

/* total_likelihood (likelihood.h:125-174), returns +log-likelihood (the :170 overload).
 * cell_mean [n_cells*4], cell_cov [n_cells*16] are the persistent MOMAdata::mean/cov (in/out):
 * zero-filled = first evaluation ("fresh"); reuse across calls reproduces the reference's
 * history dependence.  per_cell_ll (optional) receives each cell's own sum.
 * nan_cell/nan_t (optional) = first (cell, time index) in depth-first order after which the running
 * sum is NaN (likelihood.h:71), or -1. */
double ggp_oracle_total_loglik(const ggp_oracle_forest* f, const double* params11,
                               double* cell_mean, double* cell_cov,
                               double* per_cell_ll, long* nan_cell, long* nan_t);

/* prediction_forward / prediction_backward / combine_predictions (predictions.h:166, 438, 466).
 * params: [n_seg*11]. Outputs per ctp: mean [n_ctp*4], cov [n_ctp*16] row-major. */
void ggp_oracle_prediction_forward(const ggp_oracle_forest* f, const double* params, int n_seg,
                                   double* cell_mean, double* cell_cov, double* mean_f, double* cov_f);
void ggp_oracle_prediction_backward(const ggp_oracle_forest* f, const double* params, int n_seg,
                                    double* cell_mean, double* cell_cov, double* mean_b, double* cov_b);
void ggp_oracle_combine_predictions(const ggp_oracle_forest* f, const double* params, int n_seg,
                                    const double* mean_f, const double* cov_f,
                                    const double* mean_b, const double* cov_b,
                                    double* mean_p, double* cov_p);

/* collect_joint_distributions (correlation_tree.h:629-648) as a sparse list.
 * Needs the outputs of the three passes above and the cell_mean left behind by the backward pass
 * (the reference reads a stale cell.mean(1) there).  Each record: row ctp (start point),
 * col ctp (later point), 8 means + 36 upper-triangular covariances (row-major order of the 8x8).
 * Returns the number of records; writes at most `cap` of them. */
long ggp_oracle_joints(const ggp_oracle_forest* f, const double* params, int n_seg, double tol,
                       const double* cell_mean_after_backward,
                       const double* mean_f, const double* cov_f,
                       const double* mean_b, const double* cov_b,
                       long cap, long* row_ctp, long* col_ctp, double* rec44);

#ifdef __cplusplus
}
#endif
#endif
