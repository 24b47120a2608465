/*
 * oracle/ba_oracle.h -- CPU restatement (plain C, FP64) of ScaViSLAM's
 * double-window bundle-adjustment iteration.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under scavislam_b200/ may include, link
 * or call this.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, as the checker / CPU baseline.
 *
 * PARITY UNPINNED: the reference tree holds no golden vectors for this path
 * and its arithmetic lives in un-vendored third-party libraries (g2o, Sophus,
 * CSparse, Eigen) that are absent here, so this file restates their published
 * algorithms anchored on the reference's
 * own call sites:
 *   scavislam/g2o_types/anchored_points.cpp:33-58,78-83,148-189,207-235
 *   scavislam/transformations.h:52-95
 *   scavislam/maths_utils.h:66-69
 *   scavislam/slam_graph.cpp:319-355,907-1080  scavislam/slam_graph-impl.cpp:29-126
 * Secondary pins used by tests/: finite-difference Jacobians, a dense numpy
 * re-derivation of the normal equations, scipy Cholesky.
 */
#ifndef SVS_BA_ORACLE_H
#define SVS_BA_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  int P, L, E, C;
  const double *pose_qt;       /* [P][7]  qx qy qz qw tx ty tz  (T_me_from_world) */
  const unsigned char *fixed;  /* [P] or NULL */
  const double *psi;           /* [L][3]  inverse-depth point in anchor frame */
  const int *e_point;          /* [E] vertex 0 */
  const int *e_pose;           /* [E] vertex 1 */
  const int *e_anchor;         /* [E] vertex 2 */
  const double *e_obs;         /* [E][3]  (u, v, u_right) */
  const double *e_info;        /* [E][3]  diagonal of Lambda */
  const int *c_i;              /* [C] vertex 0 of G2oEdgeSE3 (pose 1) */
  const int *c_j;              /* [C] vertex 1 (pose 2) */
  const double *c_T;           /* [C][7]  T_2_from_1 */
  const double *c_Lambda;      /* [C][36] row-major 6x6 information */
  double f, px, py, b;         /* G2oCameraParameters */
} oba_problem;

#define OBA_MAX_ITERS 64
typedef struct {
  int iterations;              /* g2o optimize() return value */
  int trials_total;
  double chi2_init;
  double chi2_final;
  double lambda_final;
  double chi2_iter[OBA_MAX_ITERS];   /* robust chi2 after each outer iteration */
  double lambda_iter[OBA_MAX_ITERS];
  int trials_iter[OBA_MAX_ITERS];
  int nnzb_S;                  /* upper blocks of the reduced system incl. diagonal */
  int nnzb_L;                  /* blocks of its Cholesky factor */
} oba_stats;

/* SE3 helpers (Sophus a621ff semantics), T = qx qy qz qw tx ty tz, tangent = (upsilon, omega). */
void oba_se3_exp(const double d[6], double T[7]);
void oba_se3_log(const double T[7], double d[6]);
void oba_se3_mul(const double A[7], const double B[7], double AB[7]);
void oba_se3_inv(const double A[7], double Ainv[7]);
void oba_se3_act(const double A[7], const double x[3], double y[3]);
void oba_se3_adj(const double A[7], double Adj[36]);
void oba_invert_depth(const double x[3], double y[3]);

/* cam = f px py b.  anchored_points.cpp:148-166 / :168-189 */
void oba_edge_error(const double cam[4], const double Tp[7], const double Ta[7],
                    const double psi[3], const double obs[3], double err[3]);
void oba_edge_jacobians(const double cam[4], const double Tp[7], const double Ta[7],
                        const double psi[3], double Jpsi[9], double Jp[18], double Ja[18]);
/* anchored_points.cpp:207-235 */
void oba_posepose_error(const double T21[7], const double T1[7], const double T2[7], double err[6]);
void oba_posepose_jacobians(const double T21[7], const double err[6], double Ji[36], double Jj[36]);

/* robust chi2 of the whole problem at the given state (g2o activeRobustChi2) */
double oba_chi2(const oba_problem *p, int robust, double huber_delta);

/* Dense reduced camera system at the problem's state: S (6P x 6P row-major, full
 * symmetric, lambda added to all pose and landmark diagonals before elimination)
 * and bs (6P).  Returns robust chi2. */
double oba_reduced_system(const oba_problem *p, int robust, double huber_delta,
                          double lambda, double *S, double *bs);

/* Dense full normal equations H ((6P+3L)^2 row-major), b: poses first then points. */
double oba_full_system(const oba_problem *p, int robust, double huber_delta,
                       double *H, double *b);

/* g2o SparseOptimizer::optimize(num_iters) with OptimizationAlgorithmLevenberg,
 * BlockSolver_6_3 (Schur) and a sparse block Cholesky.  pose_out [P][7], psi_out [L][3].
 * Returns iterations performed (0 on solver failure in the last iteration, -1 if empty). */
/* OpenMP threads of the build and Schur loops: 1 (default) = the sequential restatement every parity test uses;
 * > 1 = a timing variant whose sums are accumulated in per-thread copies (last bits differ) */
void oba_set_threads(int n);
int oba_optimize(const oba_problem *p, int num_iters, int robust, double huber_delta,
                 double lambda_init, int max_trials,
                 double *pose_out, double *psi_out, oba_stats *stats);

#ifdef __cplusplus
}
#endif
#endif
