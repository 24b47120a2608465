/*
 * oracle/ba_oracle.c -- CPU restatement of the ScaViSLAM double-window BA
 * iteration.  TEST INFRASTRUCTURE ONLY (see ba_oracle.h).  PARITY UNPINNED
 * (no reference golden vectors exist; g2o/Sophus/CSparse/Eigen are absent).
 *
 * What follows what:
 *   se3_*            Sophus a621ff SE3/SO3 (README:145-147); call sites
 *                    anchored_points.cpp:57,164-165,178-188,209-223
 *   edge error/jac   anchored_points.cpp:148-189, transformations.h:62-95,
 *                    maths_utils.h:66-69, G2oCameraParameters :33-50
 *   pose-pose edge   anchored_points.cpp:207-235
 *   problem layout   slam_graph.cpp:907-1032, slam_graph-impl.cpp:29-126
 *   LM / Schur       g2o OptimizationAlgorithmLevenberg::solve, BlockSolver<6,3>,
 *                    BaseMultiEdge::constructQuadraticForm, RobustKernelHuber as
 *                    configured at slam_graph.cpp:336-346,1063-1080
 */
#include "ba_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define SMALL_EPS 1e-10

/* ------------------------------------------------------------------ SE3 */

static void quat_to_R(const double q[4], double R[9]) {
  /* Eigen::Quaterniond::toRotationMatrix, q = x y z w */
  const double x = q[0], y = q[1], z = q[2], w = q[3];
  const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w;
  const double txx = tx * x, txy = ty * x, txz = tz * x;
  const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
  R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}

static void quat_mul(const double a[4], const double b[4], double c[4]) {
  const double ax = a[0], ay = a[1], az = a[2], aw = a[3];
  const double bx = b[0], by = b[1], bz = b[2], bw = b[3];
  c[3] = aw * bw - ax * bx - ay * by - az * bz;
  c[0] = aw * bx + ax * bw + ay * bz - az * by;
  c[1] = aw * by + ay * bw + az * bx - ax * bz;
  c[2] = aw * bz + az * bw + ax * by - ay * bx;
}

static void quat_normalize(double q[4]) {
  const double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  q[0] /= n; q[1] /= n; q[2] /= n; q[3] /= n;
}

static void mat3_vec(const double R[9], const double x[3], double y[3]) {
  y[0] = R[0] * x[0] + R[1] * x[1] + R[2] * x[2];
  y[1] = R[3] * x[0] + R[4] * x[1] + R[5] * x[2];
  y[2] = R[6] * x[0] + R[7] * x[1] + R[8] * x[2];
}

static void hat(const double v[3], double M[9]) {
  M[0] = 0;     M[1] = -v[2]; M[2] = v[1];
  M[3] = v[2];  M[4] = 0;     M[5] = -v[0];
  M[6] = -v[1]; M[7] = v[0];  M[8] = 0;
}

static void mat3_mul(const double A[9], const double B[9], double C[9]) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      C[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
}

void oba_se3_act(const double A[7], const double x[3], double y[3]) {
  double R[9];
  quat_to_R(A, R);
  mat3_vec(R, x, y);
  y[0] += A[4]; y[1] += A[5]; y[2] += A[6];
}

void oba_se3_mul(const double A[7], const double B[7], double AB[7]) {
  /* Sophus: t = tA + RA*tB ; q = normalize(qA*qB) */
  double R[9], t[3], q[4];
  quat_to_R(A, R);
  mat3_vec(R, B + 4, t);
  quat_mul(A, B, q);
  quat_normalize(q);
  AB[0] = q[0]; AB[1] = q[1]; AB[2] = q[2]; AB[3] = q[3];
  AB[4] = A[4] + t[0]; AB[5] = A[5] + t[1]; AB[6] = A[6] + t[2];
}

void oba_se3_inv(const double A[7], double Ai[7]) {
  double q[4] = {-A[0], -A[1], -A[2], A[3]}, R[9], mt[3] = {-A[4], -A[5], -A[6]}, t[3];
  quat_to_R(q, R);
  mat3_vec(R, mt, t);
  Ai[0] = q[0]; Ai[1] = q[1]; Ai[2] = q[2]; Ai[3] = q[3];
  Ai[4] = t[0]; Ai[5] = t[1]; Ai[6] = t[2];
}

void oba_se3_exp(const double d[6], double T[7]) {
  const double *ups = d, *om = d + 3;
  const double theta = sqrt(om[0] * om[0] + om[1] * om[1] + om[2] * om[2]);
  const double half = 0.5 * theta;
  double imag, real = cos(half);
  if (theta < SMALL_EPS) {
    const double t2 = theta * theta, t4 = t2 * t2;
    imag = 0.5 - 0.0208333 * t2 + 0.000260417 * t4;
  } else {
    imag = sin(half) / theta;
  }
  double q[4] = {imag * om[0], imag * om[1], imag * om[2], real};
  double Om[9], Om2[9], V[9];
  hat(om, Om);
  mat3_mul(Om, Om, Om2);
  if (theta < SMALL_EPS) {
    quat_to_R(q, V);
  } else {
    const double t2 = theta * theta;
    const double a = (1 - cos(theta)) / t2, b = (theta - sin(theta)) / (t2 * theta);
    for (int i = 0; i < 9; ++i) V[i] = a * Om[i] + b * Om2[i];
    V[0] += 1; V[4] += 1; V[8] += 1;
  }
  double t[3];
  mat3_vec(V, ups, t);
  T[0] = q[0]; T[1] = q[1]; T[2] = q[2]; T[3] = q[3];
  T[4] = t[0]; T[5] = t[1]; T[6] = t[2];
}

void oba_se3_log(const double T[7], double d[6]) {
  const double n = sqrt(T[0] * T[0] + T[1] * T[1] + T[2] * T[2]);
  const double w = T[3];
  double k;
  if (n < SMALL_EPS) {
    k = 2. / w - 2. * (n * n) / (w * w * w);
  } else if (fabs(w) < SMALL_EPS) {
    k = (w > 0 ? M_PI : -M_PI) / n;
  } else {
    k = 2 * atan(n / w) / n;
  }
  const double theta = k * n;
  double om[3] = {k * T[0], k * T[1], k * T[2]};
  double Om[9], Om2[9], Vi[9];
  hat(om, Om);
  mat3_mul(Om, Om, Om2);
  double c;
  if (theta < SMALL_EPS) c = 1. / 12.;
  else c = (1 - theta / (2 * tan(theta / 2))) / (theta * theta);
  for (int i = 0; i < 9; ++i) Vi[i] = -0.5 * Om[i] + c * Om2[i];
  Vi[0] += 1; Vi[4] += 1; Vi[8] += 1;
  mat3_vec(Vi, T + 4, d);
  d[3] = om[0]; d[4] = om[1]; d[5] = om[2];
}

void oba_se3_adj(const double A[7], double Adj[36]) {
  double R[9], tx[9], tR[9];
  quat_to_R(A, R);
  hat(A + 4, tx);
  mat3_mul(tx, R, tR);
  memset(Adj, 0, 36 * sizeof(double));
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      Adj[i * 6 + j] = R[i * 3 + j];
      Adj[(i + 3) * 6 + (j + 3)] = R[i * 3 + j];
      Adj[i * 6 + (j + 3)] = tR[i * 3 + j];
    }
}

void oba_invert_depth(const double x[3], double y[3]) {
  /* maths_utils.h:66-69: unproject2d(x.head<2>())/x[2] */
  y[0] = x[0] / x[2]; y[1] = x[1] / x[2]; y[2] = 1. / x[2];
}

/* ------------------------------------------------- reprojection edge */

static void stereo_map(const double cam[4], const double y[3], double uvu[3]) {
  /* anchored_points.cpp:33-50 */
  const double f = cam[0], px = cam[1], py = cam[2], b = cam[3];
  uvu[0] = (y[0] / y[2]) * f + px;
  uvu[1] = (y[1] / y[2]) * f + py;
  uvu[2] = ((y[0] - b) / y[2]) * f + px;
}

/* T_ca = Tp * Ta^-1 as R (row-major) and t */
static void rel_pose(const double Tp[7], const double Ta[7], double R[9], double t[3]) {
  double Ai[7], T[7];
  oba_se3_inv(Ta, Ai);
  oba_se3_mul(Tp, Ai, T);
  quat_to_R(T, R);
  t[0] = T[4]; t[1] = T[5]; t[2] = T[6];
}

void oba_edge_error(const double cam[4], const double Tp[7], const double Ta[7],
                    const double psi[3], const double obs[3], double err[3]) {
  double R[9], t[3], xa[3], y[3], z[3];
  rel_pose(Tp, Ta, R, t);
  oba_invert_depth(psi, xa);
  mat3_vec(R, xa, y);
  y[0] += t[0]; y[1] += t[1]; y[2] += t[2];
  stereo_map(cam, y, z);
  err[0] = obs[0] - z[0]; err[1] = obs[1] - z[1]; err[2] = obs[2] - z[2];
}

void oba_edge_jacobians(const double cam[4], const double Tp[7], const double Ta[7],
                        const double psi[3], double Jpsi[9], double Jp[18], double Ja[18]) {
  const double f = cam[0], b = cam[3];
  double R[9], t[3], xa[3], y[3];
  rel_pose(Tp, Ta, R, t);
  oba_invert_depth(psi, xa);
  mat3_vec(R, xa, y);
  y[0] += t[0]; y[1] += t[1]; y[2] += t[2];
  /* transformations.h:62-71 d_stereoproj_d_y */
  const double zsq = y[2] * y[2];
  const double Jc[9] = {f / y[2], 0, -(f * y[0]) / zsq,
                        0, f / y[2], -(f * y[1]) / zsq,
                        f / y[2], 0, -(f * (y[0] - b)) / zsq};
  /* transformations.h:82-95 d_Tinvpsi_d_psi */
  double Rx[3], M[9];
  mat3_vec(R, xa, Rx);
  const double ipz = 1. / psi[2];
  for (int i = 0; i < 3; ++i) {
    M[i * 3 + 0] = R[i * 3 + 0] * ipz;
    M[i * 3 + 1] = R[i * 3 + 1] * ipz;
    M[i * 3 + 2] = -Rx[i] * ipz;
  }
  double JM[9];
  mat3_mul(Jc, M, JM);
  for (int i = 0; i < 9; ++i) Jpsi[i] = -JM[i];
  /* transformations.h:73-80 d_expy_d_y = [I | -hat(y)] */
  double hy[9], hx[9], A[18], B[18];
  hat(y, hy);
  hat(xa, hx);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      A[i * 6 + j] = (i == j);
      A[i * 6 + 3 + j] = -hy[i * 3 + j];
      B[i * 6 + j] = (i == j);
      B[i * 6 + 3 + j] = -hx[i * 3 + j];
    }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 6; ++j) {
      double s = 0;
      for (int k = 0; k < 3; ++k) s += Jc[i * 3 + k] * A[k * 6 + j];
      Jp[i * 6 + j] = -s;
    }
  double JR[9];
  mat3_mul(Jc, R, JR);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 6; ++j) {
      double s = 0;
      for (int k = 0; k < 3; ++k) s += JR[i * 3 + k] * B[k * 6 + j];
      Ja[i * 6 + j] = s;
    }
}

/* ------------------------------------------------- pose-pose edge */

void oba_posepose_error(const double T21[7], const double T1[7], const double T2[7], double err[6]) {
  double T2i[7], A[7], B[7];
  oba_se3_inv(T2, T2i);
  oba_se3_mul(T21, T1, A);
  oba_se3_mul(A, T2i, B);
  oba_se3_log(B, err);
}

static void mat6_mul(const double A[36], const double B[36], double C[36]) {
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) {
      double s = 0;
      for (int k = 0; k < 6; ++k) s += A[i * 6 + k] * B[k * 6 + j];
      C[i * 6 + j] = s;
    }
}

static void third(const double A[7], const double d[6], double out[36]) {
  /* anchored_points.cpp:207-215 */
  double Adj[36], dl[36], t1[36], t2[36];
  oba_se3_adj(A, Adj);
  memset(dl, 0, sizeof dl);
  double hu[9], ho[9];
  hat(d, hu);
  hat(d + 3, ho);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      dl[i * 6 + j] = -ho[i * 3 + j];
      dl[i * 6 + 3 + j] = -hu[i * 3 + j];
      dl[(i + 3) * 6 + 3 + j] = -ho[i * 3 + j];
    }
  mat6_mul(dl, Adj, t1);
  mat6_mul(dl, t1, t2);
  for (int i = 0; i < 36; ++i) out[i] = Adj[i] + 0.5 * t1[i] + (1. / 12.) * t2[i];
}

void oba_posepose_jacobians(const double T21[7], const double err[6], double Ji[36], double Jj[36]) {
  const double I[7] = {0, 0, 0, 1, 0, 0, 0};
  double md[6], t[36];
  third(T21, err, Ji);
  for (int i = 0; i < 6; ++i) md[i] = -err[i];
  third(I, md, t);
  for (int i = 0; i < 36; ++i) Jj[i] = -t[i];
}

/* ------------------------------------------------- robust kernel */

static void huber(double e2, double delta, double rho[2]) {
  /* g2o RobustKernelHuber::robustify */
  const double dsqr = delta * delta;
  if (e2 <= dsqr) {
    rho[0] = e2; rho[1] = 1.;
  } else {
    const double sq = sqrt(e2);
    rho[0] = 2 * sq * delta - dsqr;
    rho[1] = delta / sq;
  }
}

static double edge_chi2_terms(const oba_problem *p, const double *poses, const double *psi,
                              int e, int robust, double delta, double err[3], double *w) {
  const double cam[4] = {p->f, p->px, p->py, p->b};
  oba_edge_error(cam, poses + 7 * p->e_pose[e], poses + 7 * p->e_anchor[e],
                 psi + 3 * p->e_point[e], p->e_obs + 3 * e, err);
  const double *om = p->e_info + 3 * e;
  const double e2 = err[0] * err[0] * om[0] + err[1] * err[1] * om[1] + err[2] * err[2] * om[2];
  if (robust) {
    double rho[2];
    huber(e2, delta, rho);
    *w = rho[1];
    return rho[0];
  }
  *w = 1.;
  return e2;
}

static double state_chi2(const oba_problem *p, const double *poses, const double *psi,
                         int robust, double delta) {
  double chi = 0;
  for (int e = 0; e < p->E; ++e) {
    double err[3], w;
    chi += edge_chi2_terms(p, poses, psi, e, robust, delta, err, &w);
  }
  for (int c = 0; c < p->C; ++c) {
    double err[6];
    oba_posepose_error(p->c_T + 7 * c, poses + 7 * p->c_i[c], poses + 7 * p->c_j[c], err);
    const double *Lm = p->c_Lambda + 36 * c;
    for (int i = 0; i < 6; ++i)
      for (int j = 0; j < 6; ++j) chi += err[i] * Lm[i * 6 + j] * err[j];
  }
  return chi;
}

double oba_chi2(const oba_problem *p, int robust, double huber_delta) {
  return state_chi2(p, p->pose_qt, p->psi, robust, huber_delta);
}

/* ------------------------------------------------- block-sparse system */

typedef struct {
  int P, L;
  /* landmark CSR */
  int *lm_ptr, *lm_edge;
  int *slot_ptr, *slot_pose, *edge_slot;  /* slot 0 of each landmark = anchor */
  /* pose-pose blocks, upper (i<=j) */
  int *tbl;            /* P*P -> block id or -1 */
  int nblk, cap;
  int *blk_i, *blk_j;
  double *Hpp;         /* nblk*36 */
  double *S;           /* nblk*36 */
  int **row; int *row_n, *row_cap;  /* row[i] = list of j>i with block (i,j) */
  double *bp;          /* 6P */
  double *W;           /* nslots*18  Hpl blocks (6x3 row-major) */
  double *D;           /* L*9 */
  double *bl;          /* L*3 */
  double *x;           /* 6P + 3L */
  double *bs;          /* 6P */
  int nnzb_S;
} sys_t;

static int sys_block(sys_t *s, int i, int j) {
  if (i > j) { int t = i; i = j; j = t; }
  int id = s->tbl[(size_t)i * s->P + j];
  if (id >= 0) return id;
  if (s->nblk == s->cap) {
    s->cap = s->cap ? 2 * s->cap : 1024;
    s->blk_i = realloc(s->blk_i, s->cap * sizeof(int));
    s->blk_j = realloc(s->blk_j, s->cap * sizeof(int));
    s->Hpp = realloc(s->Hpp, (size_t)s->cap * 36 * sizeof(double));
    s->S = realloc(s->S, (size_t)s->cap * 36 * sizeof(double));
  }
  id = s->nblk++;
  s->blk_i[id] = i; s->blk_j[id] = j;
  memset(s->Hpp + (size_t)id * 36, 0, 36 * sizeof(double));
  memset(s->S + (size_t)id * 36, 0, 36 * sizeof(double));
  s->tbl[(size_t)i * s->P + j] = id;
  if (i != j) {
    if (s->row_n[i] == s->row_cap[i]) {
      s->row_cap[i] = s->row_cap[i] ? 2 * s->row_cap[i] : 16;
      s->row[i] = realloc(s->row[i], s->row_cap[i] * sizeof(int));
    }
    s->row[i][s->row_n[i]++] = j;
  }
  return id;
}

static void sys_free(sys_t *s) {
  free(s->lm_ptr); free(s->lm_edge); free(s->slot_ptr); free(s->slot_pose); free(s->edge_slot);
  free(s->tbl); free(s->blk_i); free(s->blk_j); free(s->Hpp); free(s->S);
  if (s->row) for (int i = 0; i < s->P; ++i) free(s->row[i]);
  free(s->row); free(s->row_n); free(s->row_cap);
  free(s->bp); free(s->W); free(s->D); free(s->bl); free(s->x); free(s->bs);
}

/* returns 0 ok, -2 if edges of one landmark disagree on the anchor */
static int sys_init(sys_t *s, const oba_problem *p) {
  memset(s, 0, sizeof *s);
  const int P = p->P, L = p->L, E = p->E;
  s->P = P; s->L = L;
  s->lm_ptr = calloc(L + 1, sizeof(int));
  s->lm_edge = malloc((E > 0 ? E : 1) * sizeof(int));
  for (int e = 0; e < E; ++e) s->lm_ptr[p->e_point[e] + 1]++;
  for (int l = 0; l < L; ++l) s->lm_ptr[l + 1] += s->lm_ptr[l];
  int *fill = malloc((L + 1) * sizeof(int));
  memcpy(fill, s->lm_ptr, (L + 1) * sizeof(int));
  for (int e = 0; e < E; ++e) s->lm_edge[fill[p->e_point[e]]++] = e;
  free(fill);
  s->slot_ptr = calloc(L + 1, sizeof(int));
  s->edge_slot = malloc((E > 0 ? E : 1) * sizeof(int));
  s->slot_pose = malloc((E + L + 1) * sizeof(int));
  int ns = 0;
  for (int l = 0; l < L; ++l) {
    s->slot_ptr[l] = ns;
    const int b = s->lm_ptr[l], en = s->lm_ptr[l + 1];
    if (b == en) continue;
    const int a = p->e_anchor[s->lm_edge[b]];
    s->slot_pose[ns++] = a;
    for (int k = b; k < en; ++k) {
      const int e = s->lm_edge[k];
      if (p->e_anchor[e] != a) return -2;
      if (p->e_pose[e] == a) s->edge_slot[e] = 0;
      else { s->edge_slot[e] = ns - s->slot_ptr[l]; s->slot_pose[ns++] = p->e_pose[e]; }
    }
  }
  s->slot_ptr[L] = ns;
  s->tbl = malloc((size_t)P * P * sizeof(int));
  for (size_t i = 0; i < (size_t)P * P; ++i) s->tbl[i] = -1;
  s->row = calloc(P, sizeof(int *));
  s->row_n = calloc(P, sizeof(int));
  s->row_cap = calloc(P, sizeof(int));
  for (int i = 0; i < P; ++i) sys_block(s, i, i);
  for (int l = 0; l < L; ++l)
    for (int a = s->slot_ptr[l]; a < s->slot_ptr[l + 1]; ++a)
      for (int b = a + 1; b < s->slot_ptr[l + 1]; ++b)
        if (s->slot_pose[a] != s->slot_pose[b]) sys_block(s, s->slot_pose[a], s->slot_pose[b]);
  for (int c = 0; c < p->C; ++c) sys_block(s, p->c_i[c], p->c_j[c]);
  s->nnzb_S = s->nblk;
  s->bp = malloc(6 * P * sizeof(double));
  s->bs = malloc(6 * P * sizeof(double));
  s->W = malloc((size_t)(ns + 1) * 18 * sizeof(double));
  s->D = malloc((size_t)(L + 1) * 9 * sizeof(double));
  s->bl = malloc((size_t)(L + 1) * 3 * sizeof(double));
  s->x = malloc((size_t)(6 * P + 3 * L + 1) * sizeof(double));
  return 0;
}

/* C(6x6, ld 6) += sgn * A^T(3x6)^T ... helpers on small row-major blocks */
static void add_AtWB(double *C, const double *A, int na, const double *B, int nb,
                     const double w[3], int transpose_out) {
  /* C (na x nb) += A^T diag(w) B, with A 3 x na, B 3 x nb (row-major) */
  for (int i = 0; i < na; ++i)
    for (int j = 0; j < nb; ++j) {
      double s = 0;
      for (int k = 0; k < 3; ++k) s += A[k * na + i] * w[k] * B[k * nb + j];
      if (transpose_out) C[j * na + i] += s; else C[i * nb + j] += s;
    }
}

/* one landmark's edges of g2o BlockSolver::buildSystem; Hpp / bp are the shared arrays (single thread) or a
 * thread's private copies (oba_set_threads > 1); D, bl, W rows belong to the landmark alone */
static void build_landmark(sys_t *s, const oba_problem *p, const double *poses, const double *psi, int robust,
                           double delta, int l, double *Hpp, double *bp, double *chi) {
  const int P = s->P;
  const double cam[4] = {p->f, p->px, p->py, p->b};
    const int sb = s->slot_ptr[l];
    for (int k = s->lm_ptr[l]; k < s->lm_ptr[l + 1]; ++k) {
      const int e = s->lm_edge[k];
      const int ip = p->e_pose[e], ia = p->e_anchor[e];
      double err[3], w;
      *chi += edge_chi2_terms(p, poses, psi, e, robust, delta, err, &w);
      double Jpsi[9], Jp[18], Ja[18];
      oba_edge_jacobians(cam, poses + 7 * ip, poses + 7 * ia, psi + 3 * l, Jpsi, Jp, Ja);
      const int fp = p->fixed && p->fixed[ip], fa = p->fixed && p->fixed[ia];
      if (fp) memset(Jp, 0, sizeof Jp);
      if (fa) memset(Ja, 0, sizeof Ja);
      const double *om = p->e_info + 3 * e;
      const double wo[3] = {w * om[0], w * om[1], w * om[2]};        /* rho' * Omega */
      const double we[3] = {-wo[0] * err[0], -wo[1] * err[1], -wo[2] * err[2]};  /* -rho' Omega e */
      const double one[3] = {1, 1, 1};
      /* vertex 0: point */
      add_AtWB(s->D + 9 * l, Jpsi, 3, Jpsi, 3, wo, 0);
      for (int i = 0; i < 3; ++i)
        for (int k2 = 0; k2 < 3; ++k2) s->bl[3 * l + i] += Jpsi[k2 * 3 + i] * we[k2];
      (void)one;
      /* vertex 1: pose ; vertex 2: anchor */
      double *App = Hpp + 36 * (size_t)s->tbl[(size_t)ip * P + ip];
      double *Aaa = Hpp + 36 * (size_t)s->tbl[(size_t)ia * P + ia];
      add_AtWB(App, Jp, 6, Jp, 6, wo, 0);
      add_AtWB(Aaa, Ja, 6, Ja, 6, wo, 0);
      for (int i = 0; i < 6; ++i)
        for (int k2 = 0; k2 < 3; ++k2) {
          bp[6 * ip + i] += Jp[k2 * 6 + i] * we[k2];
          bp[6 * ia + i] += Ja[k2 * 6 + i] * we[k2];
        }
      /* off-diagonal (1,2): upper block of Hpp; when ip==ia it aliases the diagonal
       * block and receives Jp^T W Ja once (g2o mapHessianMemory quirk) */
      {
        const int lo = ip <= ia ? ip : ia, hi = ip <= ia ? ia : ip;
        double *Apa = Hpp + 36 * (size_t)s->tbl[(size_t)lo * P + hi];
        add_AtWB(Apa, Jp, 6, Ja, 6, wo, ip > ia);
      }
      /* Hpl blocks (pose x point) */
      add_AtWB(s->W + 18 * (size_t)(sb + s->edge_slot[e]), Jp, 6, Jpsi, 3, wo, 0);
      add_AtWB(s->W + 18 * (size_t)sb, Ja, 6, Jpsi, 3, wo, 0);
    }
}

static int g_threads = 1;
/* number of OpenMP threads of the build and Schur loops (1 = the sequential restatement, the default) */
void oba_set_threads(int n) { g_threads = n > 1 ? n : 1; }

/* g2o BlockSolver::buildSystem: Hpp, Hpl(W), Hll(D), b at the given state.  Returns robust chi2. */
static double sys_build(sys_t *s, const oba_problem *p, const double *poses, const double *psi,
                        int robust, double delta) {
  const int P = s->P, L = s->L;
  memset(s->Hpp, 0, (size_t)s->nblk * 36 * sizeof(double));
  memset(s->bp, 0, 6 * P * sizeof(double));
  memset(s->W, 0, (size_t)s->slot_ptr[L] * 18 * sizeof(double));
  memset(s->D, 0, (size_t)L * 9 * sizeof(double));
  memset(s->bl, 0, (size_t)L * 3 * sizeof(double));
  double chi = 0;
  if (g_threads <= 1) {
    for (int l = 0; l < L; ++l) build_landmark(s, p, poses, psi, robust, delta, l, s->Hpp, s->bp, &chi);
  } else {
    /* timing variant: landmarks split over threads, private pose blocks summed afterwards (the sum order, and
     * with it the last bits, differ from the sequential restatement) */
#pragma omp parallel num_threads(g_threads)
    {
      double *H = calloc((size_t)s->nblk * 36, sizeof(double)), *b = calloc(6 * (size_t)P, sizeof(double));
      double c = 0;
#pragma omp for schedule(static)
      for (int l = 0; l < L; ++l) build_landmark(s, p, poses, psi, robust, delta, l, H, b, &c);
#pragma omp critical
      {
        for (size_t i = 0; i < (size_t)s->nblk * 36; ++i) s->Hpp[i] += H[i];
        for (int i = 0; i < 6 * P; ++i) s->bp[i] += b[i];
        chi += c;
      }
      free(H); free(b);
    }
  }
  for (int c = 0; c < p->C; ++c) {
    const int i = p->c_i[c], j = p->c_j[c];
    double err[6], Ji[36], Jj[36];
    oba_posepose_error(p->c_T + 7 * c, poses + 7 * i, poses + 7 * j, err);
    oba_posepose_jacobians(p->c_T + 7 * c, err, Ji, Jj);
    if (p->fixed && p->fixed[i]) memset(Ji, 0, sizeof Ji);
    if (p->fixed && p->fixed[j]) memset(Jj, 0, sizeof Jj);
    const double *Lm = p->c_Lambda + 36 * c;
    double Oe[6], AtO[36], BtO[36];
    for (int a = 0; a < 6; ++a) {
      double sv = 0;
      for (int b2 = 0; b2 < 6; ++b2) { sv += Lm[a * 6 + b2] * err[b2]; chi += err[a] * Lm[a * 6 + b2] * err[b2]; }
      Oe[a] = -sv;
    }
    for (int a = 0; a < 6; ++a)
      for (int b2 = 0; b2 < 6; ++b2) {
        double sa = 0, sb2 = 0;
        for (int k = 0; k < 6; ++k) { sa += Ji[k * 6 + a] * Lm[k * 6 + b2]; sb2 += Jj[k * 6 + a] * Lm[k * 6 + b2]; }
        AtO[a * 6 + b2] = sa; BtO[a * 6 + b2] = sb2;
      }
    double *Aii = s->Hpp + 36 * (size_t)s->tbl[(size_t)i * P + i];
    double *Ajj = s->Hpp + 36 * (size_t)s->tbl[(size_t)j * P + j];
    const int lo = i <= j ? i : j, hi = i <= j ? j : i;
    double *Aij = s->Hpp + 36 * (size_t)s->tbl[(size_t)lo * P + hi];
    for (int a = 0; a < 6; ++a) {
      for (int b2 = 0; b2 < 6; ++b2) {
        double sii = 0, sjj = 0, sij = 0;
        for (int k = 0; k < 6; ++k) {
          sii += AtO[a * 6 + k] * Ji[k * 6 + b2];
          sjj += BtO[a * 6 + k] * Jj[k * 6 + b2];
          sij += AtO[a * 6 + k] * Jj[k * 6 + b2];
        }
        Aii[a * 6 + b2] += sii;
        Ajj[a * 6 + b2] += sjj;
        if (i <= j) Aij[a * 6 + b2] += sij; else Aij[b2 * 6 + a] += sij;
      }
      double bi = 0, bj = 0;
      for (int k = 0; k < 6; ++k) { bi += Ji[k * 6 + a] * Oe[k]; bj += Jj[k * 6 + a] * Oe[k]; }
      s->bp[6 * i + a] += bi;
      s->bp[6 * j + a] += bj;
    }
  }
  return chi;
}

static void inv3(const double A[9], double Ai[9]) {
  /* Eigen 3x3 inverse: cofactors / determinant */
  const double c00 = A[4] * A[8] - A[5] * A[7], c01 = A[5] * A[6] - A[3] * A[8], c02 = A[3] * A[7] - A[4] * A[6];
  const double det = A[0] * c00 + A[1] * c01 + A[2] * c02;
  const double id = 1. / det;
  Ai[0] = c00 * id; Ai[1] = (A[2] * A[7] - A[1] * A[8]) * id; Ai[2] = (A[1] * A[5] - A[2] * A[4]) * id;
  Ai[3] = c01 * id; Ai[4] = (A[0] * A[8] - A[2] * A[6]) * id; Ai[5] = (A[2] * A[3] - A[0] * A[5]) * id;
  Ai[6] = c02 * id; Ai[7] = (A[1] * A[6] - A[0] * A[7]) * id; Ai[8] = (A[0] * A[4] - A[1] * A[3]) * id;
}

/* one landmark of the Schur complement; S / bs are the shared arrays or a thread's private copies */
static void schur_landmark(sys_t *s, double lambda, int l, double *S, double *bs) {
  const int P = s->P;
    const int sb = s->slot_ptr[l], se = s->slot_ptr[l + 1];
    if (sb == se) return;
    double Dl[9], Di[9], db[3];
    memcpy(Dl, s->D + 9 * l, sizeof Dl);
    Dl[0] += lambda; Dl[4] += lambda; Dl[8] += lambda;
    inv3(Dl, Di);
    mat3_vec(Di, s->bl + 3 * l, db);
    for (int a = sb; a < se; ++a) {
      const int ia = s->slot_pose[a];
      const double *Ba = s->W + 18 * (size_t)a;
      double Y[18];
      for (int r = 0; r < 6; ++r)
        for (int c = 0; c < 3; ++c)
          Y[r * 3 + c] = Ba[r * 3] * Di[c] + Ba[r * 3 + 1] * Di[3 + c] + Ba[r * 3 + 2] * Di[6 + c];
      for (int r = 0; r < 6; ++r)
        bs[6 * ia + r] -= Ba[r * 3] * db[0] + Ba[r * 3 + 1] * db[1] + Ba[r * 3 + 2] * db[2];
      for (int b = sb; b < se; ++b) {
        const int ib = s->slot_pose[b];
        if (ib < ia) continue;
        if (ib == ia && b != a) continue; /* cannot happen: slots hold distinct poses */
        const double *Bb = s->W + 18 * (size_t)b;
        double *Sab = S + 36 * (size_t)s->tbl[(size_t)ia * P + ib];
        for (int r = 0; r < 6; ++r)
          for (int c = 0; c < 6; ++c)
            Sab[r * 6 + c] -= Y[r * 3] * Bb[c * 3] + Y[r * 3 + 1] * Bb[c * 3 + 1] + Y[r * 3 + 2] * Bb[c * 3 + 2];
      }
    }
}

/* g2o BlockSolver::solve, Schur part: S = Hpp + lambda I - Hpl (Hll + lambda I)^-1 Hpl^T */
static void sys_schur(sys_t *s, const oba_problem *p, double lambda) {
  const int P = s->P, L = s->L;
  memcpy(s->S, s->Hpp, (size_t)s->nblk * 36 * sizeof(double));
  for (int i = 0; i < P; ++i) {
    double *d = s->S + 36 * (size_t)s->tbl[(size_t)i * P + i];
    for (int a = 0; a < 6; ++a) d[a * 7] += lambda;
    if (p->fixed && p->fixed[i]) for (int a = 0; a < 6; ++a) d[a * 7] += 1.;
  }
  memcpy(s->bs, s->bp, 6 * P * sizeof(double));
  if (g_threads <= 1) {
    for (int l = 0; l < L; ++l) schur_landmark(s, lambda, l, s->S, s->bs);
  } else {
#pragma omp parallel num_threads(g_threads)
    {
      double *St = calloc((size_t)s->nblk * 36, sizeof(double)), *bt = calloc(6 * (size_t)P, sizeof(double));
#pragma omp for schedule(static)
      for (int l = 0; l < L; ++l) schur_landmark(s, lambda, l, St, bt);
#pragma omp critical
      {
        for (size_t i = 0; i < (size_t)s->nblk * 36; ++i) s->S[i] += St[i];
        for (int i = 0; i < 6 * P; ++i) s->bs[i] += bt[i];
      }
      free(St); free(bt);
    }
  }
}

/* Block right-looking Cholesky S = U^T U on the upper block pattern (with fill),
 * then U^T U x = bs.  Returns 0 on success, 1 if a pivot is not positive
 * (CSparse cs_chol failure => g2o treats the trial as failed). */
static int sys_solve_reduced(sys_t *s) {
  const int P = s->P;
  double *y = s->x; /* first 6P entries */
  memcpy(y, s->bs, 6 * P * sizeof(double));
  for (int k = 0; k < P; ++k) {
    double *Ukk = s->S + 36 * (size_t)s->tbl[(size_t)k * P + k];
    /* in-place upper Cholesky of the 6x6 diagonal block (upper triangle is authoritative) */
    for (int c = 0; c < 6; ++c) {
      double d = Ukk[c * 6 + c];
      for (int r = 0; r < c; ++r) d -= Ukk[r * 6 + c] * Ukk[r * 6 + c];
      if (!(d > 0)) return 1;
      d = sqrt(d);
      Ukk[c * 6 + c] = d;
      for (int c2 = c + 1; c2 < 6; ++c2) {
        double v = Ukk[c * 6 + c2];
        for (int r = 0; r < c; ++r) v -= Ukk[r * 6 + c] * Ukk[r * 6 + c2];
        Ukk[c * 6 + c2] = v / d;
      }
      for (int r = c + 1; r < 6; ++r) Ukk[r * 6 + c] = 0;
    }
    /* U_kj = U_kk^-T S_kj */
    for (int n = 0; n < s->row_n[k]; ++n) {
      double *B = s->S + 36 * (size_t)s->tbl[(size_t)k * P + s->row[k][n]];
      for (int c = 0; c < 6; ++c)
        for (int r = 0; r < 6; ++r) {
          double v = B[r * 6 + c];
          for (int q = 0; q < r; ++q) v -= Ukk[q * 6 + r] * B[q * 6 + c];
          B[r * 6 + c] = v / Ukk[r * 6 + r];
        }
    }
    /* y_k = U_kk^-T b_k */
    for (int r = 0; r < 6; ++r) {
      double v = y[6 * k + r];
      for (int q = 0; q < r; ++q) v -= Ukk[q * 6 + r] * y[6 * k + q];
      y[6 * k + r] = v / Ukk[r * 6 + r];
    }
    /* trailing update */
    const int nk = s->row_n[k];
    for (int a = 0; a < nk; ++a) {
      const int ia = s->row[k][a];
      const double *Ua = s->S + 36 * (size_t)s->tbl[(size_t)k * P + ia];
      for (int r = 0; r < 6; ++r) {
        double v = 0;
        for (int q = 0; q < 6; ++q) v += Ua[q * 6 + r] * y[6 * k + q];
        y[6 * ia + r] -= v;
      }
      for (int b = 0; b < nk; ++b) {
        const int ib = s->row[k][b];
        if (ib < ia) continue;
        const int id = sys_block(s, ia, ib);   /* may create fill (realloc!) */
        const double *Ua2 = s->S + 36 * (size_t)s->tbl[(size_t)k * P + ia];
        const double *Ub = s->S + 36 * (size_t)s->tbl[(size_t)k * P + ib];
        double *T = s->S + 36 * (size_t)id;
        for (int r = 0; r < 6; ++r)
          for (int c = 0; c < 6; ++c) {
            double v = 0;
            for (int q = 0; q < 6; ++q) v += Ua2[q * 6 + r] * Ub[q * 6 + c];
            T[r * 6 + c] -= v;
          }
      }
    }
  }
  for (int k = P - 1; k >= 0; --k) {
    const double *Ukk = s->S + 36 * (size_t)s->tbl[(size_t)k * P + k];
    double v[6];
    for (int r = 0; r < 6; ++r) v[r] = y[6 * k + r];
    for (int n = 0; n < s->row_n[k]; ++n) {
      const int j = s->row[k][n];
      const double *U = s->S + 36 * (size_t)s->tbl[(size_t)k * P + j];
      for (int r = 0; r < 6; ++r)
        for (int c = 0; c < 6; ++c) v[r] -= U[r * 6 + c] * y[6 * j + c];
    }
    for (int r = 5; r >= 0; --r) {
      double t = v[r];
      for (int c = r + 1; c < 6; ++c) t -= Ukk[r * 6 + c] * y[6 * k + c];
      y[6 * k + r] = t / Ukk[r * 6 + r];
    }
  }
  return 0;
}

/* landmark back-substitution: xl = (Hll+lambda)^-1 (bl - Hpl^T xp) */
static void sys_backsub(sys_t *s, double lambda) {
  const int P = s->P, L = s->L;
  double *xl = s->x + 6 * P;
  for (int l = 0; l < L; ++l) {
    double c[3] = {s->bl[3 * l], s->bl[3 * l + 1], s->bl[3 * l + 2]};
    for (int a = s->slot_ptr[l]; a < s->slot_ptr[l + 1]; ++a) {
      const double *B = s->W + 18 * (size_t)a;
      const double *xp = s->x + 6 * s->slot_pose[a];
      for (int r = 0; r < 6; ++r)
        for (int q = 0; q < 3; ++q) c[q] -= B[r * 3 + q] * xp[r];
    }
    if (s->slot_ptr[l] == s->slot_ptr[l + 1]) { xl[3 * l] = xl[3 * l + 1] = xl[3 * l + 2] = 0; continue; }
    double Dl[9], Di[9];
    memcpy(Dl, s->D + 9 * l, sizeof Dl);
    Dl[0] += lambda; Dl[4] += lambda; Dl[8] += lambda;
    inv3(Dl, Di);
    mat3_vec(Di, c, xl + 3 * l);
  }
}

double oba_reduced_system(const oba_problem *p, int robust, double huber_delta,
                          double lambda, double *Sd, double *bs) {
  sys_t s;
  if (sys_init(&s, p)) { sys_free(&s); return NAN; }
  const double chi = sys_build(&s, p, p->pose_qt, p->psi, robust, huber_delta);
  sys_schur(&s, p, lambda);
  const int P = p->P, n = 6 * P;
  memset(Sd, 0, (size_t)n * n * sizeof(double));
  for (int id = 0; id < s.nblk; ++id) {
    const int i = s.blk_i[id], j = s.blk_j[id];
    const double *B = s.S + 36 * (size_t)id;
    for (int r = 0; r < 6; ++r)
      for (int c = 0; c < 6; ++c) {
        if (i == j) {
          /* diagonal blocks: the upper triangle is authoritative (g2o works on it) */
          const double v = r <= c ? B[r * 6 + c] : B[c * 6 + r];
          Sd[(size_t)(6 * i + r) * n + 6 * j + c] = v;
        } else {
          Sd[(size_t)(6 * i + r) * n + 6 * j + c] = B[r * 6 + c];
          Sd[(size_t)(6 * j + c) * n + 6 * i + r] = B[r * 6 + c];
        }
      }
  }
  memcpy(bs, s.bs, n * sizeof(double));
  sys_free(&s);
  return chi;
}

double oba_full_system(const oba_problem *p, int robust, double huber_delta, double *H, double *b) {
  sys_t s;
  if (sys_init(&s, p)) { sys_free(&s); return NAN; }
  const double chi = sys_build(&s, p, p->pose_qt, p->psi, robust, huber_delta);
  const int P = p->P, L = p->L;
  const size_t n = 6 * (size_t)P + 3 * (size_t)L;
  memset(H, 0, n * n * sizeof(double));
  for (int id = 0; id < s.nblk; ++id) {
    const int i = s.blk_i[id], j = s.blk_j[id];
    const double *B = s.Hpp + 36 * (size_t)id;
    for (int r = 0; r < 6; ++r)
      for (int c = 0; c < 6; ++c) {
        if (i == j) H[(6 * i + r) * n + 6 * j + c] = r <= c ? B[r * 6 + c] : B[c * 6 + r];
        else { H[(6 * i + r) * n + 6 * j + c] = B[r * 6 + c]; H[(6 * j + c) * n + 6 * i + r] = B[r * 6 + c]; }
      }
  }
  for (int l = 0; l < L; ++l) {
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) H[(6 * P + 3 * l + r) * n + 6 * P + 3 * l + c] = s.D[9 * l + r * 3 + c];
    for (int a = s.slot_ptr[l]; a < s.slot_ptr[l + 1]; ++a) {
      const int ip = s.slot_pose[a];
      for (int r = 0; r < 6; ++r)
        for (int c = 0; c < 3; ++c) {
          H[(6 * ip + r) * n + 6 * P + 3 * l + c] += s.W[18 * (size_t)a + r * 3 + c];
          H[(6 * P + 3 * l + c) * n + 6 * ip + r] += s.W[18 * (size_t)a + r * 3 + c];
        }
    }
  }
  memcpy(b, s.bp, 6 * P * sizeof(double));
  memcpy(b + 6 * P, s.bl, 3 * (size_t)L * sizeof(double));
  sys_free(&s);
  return chi;
}

int oba_optimize(const oba_problem *p, int num_iters, int robust, double huber_delta,
                 double lambda_init, int max_trials,
                 double *pose_out, double *psi_out, oba_stats *st) {
  const int P = p->P, L = p->L;
  if (st) memset(st, 0, sizeof *st);
  memcpy(pose_out, p->pose_qt, 7 * (size_t)P * sizeof(double));
  memcpy(psi_out, p->psi, 3 * (size_t)L * sizeof(double));
  if (P == 0) return -1;
  sys_t s;
  int rc = sys_init(&s, p);
  if (rc) { sys_free(&s); return rc; }
  double *pose_bak = malloc(7 * (size_t)P * sizeof(double));
  double *psi_bak = malloc(3 * (size_t)(L + 1) * sizeof(double));
  double lambda = lambda_init, ni = 2;
  int iters = 0, ok = 1, failed = 0;
  if (st) { st->nnzb_S = s.nnzb_S; st->chi2_init = state_chi2(p, pose_out, psi_out, robust, huber_delta); }
  for (int it = 0; it < num_iters && ok; ++it) {
    double currentChi = sys_build(&s, p, pose_out, psi_out, robust, huber_delta);
    double tempChi = currentChi, rho = 0;
    if (it == 0) { lambda = lambda_init; ni = 2; }
    int qmax = 0;
    do {
      memcpy(pose_bak, pose_out, 7 * (size_t)P * sizeof(double));   /* push */
      memcpy(psi_bak, psi_out, 3 * (size_t)L * sizeof(double));
      sys_schur(&s, p, lambda);
      const int fail = sys_solve_reduced(&s);
      failed = fail;
      if (!fail) {
        sys_backsub(&s, lambda);
        for (int i = 0; i < P; ++i) {                               /* oplus */
          if (p->fixed && p->fixed[i]) continue;
          double dT[7], Tn[7];
          oba_se3_exp(s.x + 6 * i, dT);
          oba_se3_mul(dT, pose_out + 7 * i, Tn);
          memcpy(pose_out + 7 * i, Tn, sizeof Tn);
        }
        for (int i = 0; i < 3 * L; ++i) psi_out[i] += s.x[6 * P + i];
        tempChi = state_chi2(p, pose_out, psi_out, robust, huber_delta);
      } else {
        tempChi = DBL_MAX;
      }
      rho = currentChi - tempChi;
      double scale = 0;
      if (!fail) {
        for (int j = 0; j < 6 * P; ++j) scale += s.x[j] * (lambda * s.x[j] + s.bp[j]);
        for (int j = 0; j < 3 * L; ++j) scale += s.x[6 * P + j] * (lambda * s.x[6 * P + j] + s.bl[j]);
      }
      scale += 1e-3;
      rho /= scale;
      if (rho > 0 && isfinite(tempChi)) {
        double alpha = 1. - pow(2 * rho - 1, 3);
        alpha = alpha < 2. / 3. ? alpha : 2. / 3.;
        const double sf = alpha > 1. / 3. ? alpha : 1. / 3.;
        lambda *= sf;
        ni = 2;
        currentChi = tempChi;
      } else {
        lambda *= ni;
        ni *= 2;
        memcpy(pose_out, pose_bak, 7 * (size_t)P * sizeof(double)); /* pop */
        memcpy(psi_out, psi_bak, 3 * (size_t)L * sizeof(double));
      }
      qmax++;
    } while (rho < 0 && qmax < max_trials);
    if (st && it < OBA_MAX_ITERS) {
      st->chi2_iter[it] = currentChi; st->lambda_iter[it] = lambda; st->trials_iter[it] = qmax;
    }
    if (st) st->trials_total += qmax;
    ++iters;
    if (qmax == max_trials || rho == 0) ok = 0;  /* Terminate */
  }
  if (st) {
    st->iterations = iters;
    st->lambda_final = lambda;
    st->chi2_final = state_chi2(p, pose_out, psi_out, robust, huber_delta);
    st->nnzb_L = s.nblk;
  }
  (void)failed;
  free(pose_bak); free(psi_bak);
  sys_free(&s);
  return iters;
}
