/* pose_oracle.c -- see pose_oracle.h.  TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED.
 * Every function cites the reference lines it restates; the code is written from the algorithm,
 * sequentially and in the reference's summation order (obs_list order). */
#include "pose_oracle.h"

#include <math.h>
#include <string.h>

#include "ba_oracle.h"

static const double kEps = 0.0000000001; /* global.h:106 */

/* stereo_camera.cpp:36-44 (map_uvu) after SE3 action */
void opo_map(const double cam[4], const double T[7], const double xyz[3], double uvu[3]) {
  double p[3];
  oba_se3_act(T, xyz, p);
  uvu[0] = cam[0] * (p[0] / p[2]) + cam[1];
  uvu[1] = cam[0] * (p[1] / p[2]) + cam[2];
  uvu[2] = (p[0] - cam[3]) / p[2] * cam[0] + cam[1];
}

/* transformations.h:417-443 */
void opo_frame_jac(const double cam[4], const double T[7], const double xyz[3], double J[18]) {
  double p[3];
  oba_se3_act(T, xyz, p);
  const double x = p[0], y = p[1], z = p[2], f = cam[0];
  const double one_b_z = 1. / z, one_b_z_sq = 1. / (z * z);
  const double A = -f * one_b_z, B = -f * one_b_z;
  const double Cc = f * x * one_b_z_sq, D = f * y * one_b_z_sq, E = f * (x - cam[3]) * one_b_z_sq;
  const double rows[18] = {A, 0, Cc, y * Cc,      z * A - x * Cc, -y * A,
                           0, B, D,  -z * B + y * D, -x * D,      x * B,
                           A, 0, E,  y * E,       z * A - x * E,  -y * A};
  memcpy(J, rows, sizeof rows);
}

/* pose_optimizer.h:441-449 */
static double kernel(double delta, double b) {
  const double a = fabs(delta);
  return a < b ? delta * delta : 2 * b * a - b * b;
}

/* residual with the robust reweighting of pose_optimizer.h:172-177; returns sqrW(f) */
static double weighted_residual(const double cam[4], const double T[7], const double xyz[3], const double obs[3],
                                int robust, double kparam, double f[3]) {
  double m[3];
  opo_map(cam, T, xyz, m);
  for (int k = 0; k < 3; ++k) f[k] = obs[k] - m[k];
  if (robust) {
    const double nrm = fmax(kEps, sqrt(f[0] * f[0] + f[1] * f[1] + f[2] * f[2]));
    const double w = sqrt(kernel(nrm, kparam)) / nrm;
    for (int k = 0; k < 3; ++k) f[k] *= w;
  }
  return f[0] * f[0] + f[1] * f[1] + f[2] * f[2];
}

/* A x = b for symmetric positive definite 6x6 (Eigen LDLT in the reference, pose_optimizer.h:237) */
static void ldlt_solve6(const double A[36], const double b[6], double x[6]) {
  double L[36] = {0}, D[6];
  for (int j = 0; j < 6; ++j) {
    double d = A[j * 6 + j];
    for (int k = 0; k < j; ++k) d -= L[j * 6 + k] * L[j * 6 + k] * D[k];
    D[j] = d;
    for (int i = j + 1; i < 6; ++i) {
      double s = A[i * 6 + j];
      for (int k = 0; k < j; ++k) s -= L[i * 6 + k] * L[j * 6 + k] * D[k];
      L[i * 6 + j] = s / d;
    }
  }
  double y[6];
  for (int i = 0; i < 6; ++i) {
    double s = b[i];
    for (int k = 0; k < i; ++k) s -= L[i * 6 + k] * y[k];
    y[i] = s;
  }
  for (int i = 5; i >= 0; --i) {
    double s = y[i] / D[i];
    for (int k = i + 1; k < 6; ++k) s -= L[k * 6 + i] * x[k];
    x[i] = s;
  }
}

void opo_calc_fast_motion_only(int n, const int *pid, const double *obs, const double *pts, const double cam[4],
                               int robust, double kparam, int num_iter, double initial_mu, double tau, double T[7],
                               opo_stats *st) {
  double nu = 2, chi2 = 0, max_err = 0, norm_max_A = 0;
  int stop = 0, trial = 0;
  memset(st, 0, sizeof *st);
  /* :153-186 first pass */
  for (int i = 0; i < n; ++i) {
    const double *x = pts + 3 * pid[i];
    double J[18], f[3];
    opo_frame_jac(cam, T, x, J);
    for (int c = 0; c < 6; ++c) {
      const double d = J[c] * J[c] + J[6 + c] * J[6 + c] + J[12 + c] * J[12 + c];
      norm_max_A = fmax(norm_max_A, fabs(d));
    }
    chi2 += weighted_residual(cam, T, x, obs + 3 * i, robust, kparam, f);
    for (int k = 0; k < 3; ++k) max_err = fmax(max_err, fabs(f[k]));
  }
  st->num_obs = n;
  st->initial_chi2 = chi2;
  double mu = initial_mu;
  if (initial_mu == -1) mu = tau * norm_max_A; /* :193-196 */
  for (int ig = 0; ig < num_iter; ++ig) {
    double rho = 0;
    do {
      double A[36] = {0}, B[6] = {0};
      for (int k = 0; k < 6; ++k) A[k * 6 + k] = mu;
      for (int i = 0; i < n; ++i) { /* :213-236 */
        const double *x = pts + 3 * pid[i];
        double J[18], f[3];
        opo_frame_jac(cam, T, x, J);
        weighted_residual(cam, T, x, obs + 3 * i, robust, kparam, f);
        for (int r = 0; r < 6; ++r) {
          for (int c = 0; c < 6; ++c) A[r * 6 + c] += J[r] * J[c] + J[6 + r] * J[6 + c] + J[12 + r] * J[12 + c];
          B[r] -= J[r] * f[0] + J[6 + r] * f[1] + J[12 + r] * f[2];
        }
      }
      double delta[6], dT[7], Tn[7];
      ldlt_solve6(A, B, delta);
      oba_se3_exp(delta, dT); /* SE3_AbstractPoint::add, transformations.h:408-411 */
      oba_se3_mul(dT, T, Tn);
      double new_chi2 = 0, new_max_err = 0;
      for (int i = 0; i < n; ++i) { /* :245-263 */
        double f[3];
        new_chi2 += weighted_residual(cam, Tn, pts + 3 * pid[i], obs + 3 * i, robust, kparam, f);
        for (int k = 0; k < 3; ++k) new_max_err = fmax(new_max_err, fabs(f[k]));
      }
      st->trials++;
      if (isnan(new_chi2)) { st->nan_error = 1; goto done; } /* :265-268 */
      rho = chi2 - new_chi2;
      if (rho > 0) { /* :270-278 */
        memcpy(T, Tn, sizeof Tn);
        chi2 = new_chi2;
        max_err = new_max_err;
        double nb = 0;
        for (int k = 0; k < 6; ++k) nb = fmax(nb, fabs(B[k]));
        stop = nb <= kEps;
        const double c = 2 * rho - 1;
        mu *= fmax(1. / 3., 1 - c * c * c);
        nu = 2.;
        trial = 0;
        st->iterations++;
      } else { /* :280-293 */
        mu *= nu;
        nu *= 2.;
        ++trial;
        if (trial == 5) stop = 1;
      }
    } while (!(rho > 0 || stop));
    if (stop) break;
  }
done:
  st->chi2 = chi2;
  st->max_err = max_err;
}
