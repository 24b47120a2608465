/* constraint_oracle.c -- see constraint_oracle.h.  TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED. */
#include "constraint_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "ba_oracle.h"

static int cmp_double(const void *a, const void *b) {
  const double x = *(const double *)a, y = *(const double *)b;
  return (x > y) - (x < y);
}

/* median of a multiset, maths_utils.h:113-136 */
static double median_sorted(const double *d, int n) {
  if (n % 2 == 1) return d[n / 2];
  return 0.5 * (d[n / 2 - 1] + d[n / 2]);
}

void occ_compute_constraints(int P, const double *poses, const int *feat_ptr, const int *feat_point, int L,
                             const int *point_anchor, const double *xyz_anchor, int npairs, const int *v1,
                             const int *v2, double *T12, double *Lambda, int *strength) {
  (void)P; (void)L;
  for (int k = 0; k < npairs; ++k) {
    const double *T1 = poses + 7 * v1[k], *T2 = poses + 7 * v2[k];
    double T2i[7];
    oba_se3_inv(T2, T2i);
    oba_se3_mul(T1, T2i, T12 + 7 * k); /* :793 */
    const int a0 = feat_ptr[v1[k]], a1 = feat_ptr[v1[k] + 1], b0 = feat_ptr[v2[k]], b1 = feat_ptr[v2[k] + 1];
    double *depth = (double *)malloc(sizeof(double) * (size_t)(a1 - a0 + 1));
    int n = 0;
    for (int i = a0, j = b0; i < a1; ++i) { /* :797-832: points seen by both frames */
      const int p = feat_point[i];
      while (j < b1 && feat_point[j] < p) ++j;
      if (j >= b1 || feat_point[j] != p) continue;
      double Tai[7], A[7], x[3];
      oba_se3_inv(poses + 7 * point_anchor[p], Tai);
      oba_se3_mul(T1, Tai, A); /* v1.T_me_from_world * T_anchor_from_w.inverse() * p.xyz_anchor */
      oba_se3_act(A, xyz_anchor + 3 * p, x);
      depth[n++] = sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
    }
    strength[k] = n;
    double *Lm = Lambda + 36 * k;
    memset(Lm, 0, sizeof(double) * 36);
    if (n > 0) {
      qsort(depth, (size_t)n, sizeof(double), cmp_double);
      const double med = median_sorted(depth, n);
      const double *t = T12 + 7 * k + 4;
      const double norm_dist = sqrt(t[0] * t[0] + t[1] * t[1] + t[2] * t[2]) / med; /* :840-841 */
      const double a = 350 * 1. * norm_dist, b = 100 * 1.;
      for (int q = 0; q < 3; ++q) { Lm[q * 7] = (double)n * (a * a); Lm[(q + 3) * 7] = (double)n * (b * b); } /* :843-846 */
    }
    free(depth);
  }
}
