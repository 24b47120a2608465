/*
 * oracle/dt_oracle.c -- CPU restatement of ScaViSLAM's dense photometric tracker with the
 * semantics of its GPU path.  TEST INFRASTRUCTURE ONLY (see ba_oracle.h).  PARITY UNPINNED.
 *
 * Follows
 *   scavislam/gpu/dense_tracking.cu:24-80   matTimesVec, cameraProject, frameJacobian
 *   scavislam/gpu/dense_tracking.cu:82-122  pointcloud_kernel
 *   scavislam/gpu/dense_tracking.cu:172-263 jacobianReduction_kernel
 *   scavislam/gpu/dense_tracking.cu:376-453 chi2_kernel
 *   scavislam/dense_tracking.cpp:62-193     DenseTracker::denseTrackingGpu (LM loop)
 *   scavislam/dense_tracking.cpp:195-216    computeDensePointCloudGpu
 * Per-pixel arithmetic is IEEE single precision in the order written in the reference (this file
 * is compiled with -ffp-contract=off); sums over pixels are accumulated in double in raster
 * order (the reference's float tree + host sum depends on its block size).
 * The texture unit's bilinear filter (dense_tracking.cu:150-152, 285-287) is restated from the
 * CUDA programming guide: weights quantised to 8 fractional bits unless exact_bilinear.
 */
#include "dt_oracle.h"

#include <math.h>
#include <string.h>

#include "ba_oracle.h"

static float bilinear(const float *img, int stride, float u, float v, int exact) {
  const float x0 = floorf(u), y0 = floorf(v);
  float a = u - x0, b = v - y0;
  if (!exact) {
    a = floorf(a * 256.f + 0.5f) / 256.f;
    b = floorf(b * 256.f + 0.5f) / 256.f;
  }
  const int xi = (int)x0, yi = (int)y0;
  const float t00 = img[yi * stride + xi], t10 = img[yi * stride + xi + 1];
  const float t01 = img[(yi + 1) * stride + xi], t11 = img[(yi + 1) * stride + xi + 1];
  return ((1.f - a) * (1.f - b)) * t00 + (a * (1.f - b)) * t10 + ((1.f - a) * b) * t01 + (a * b) * t11;
}

/* SE3 (double) -> column-major 3x4 float, GpuMatrix34::set of T.matrix().topLeftCorner<3,4>() */
void odt_pose_to_m34(const double T[7], float m[12]) {
  const double x = T[0], y = T[1], z = T[2], w = T[3];
  double R[9];
  R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - z * w); R[2] = 2 * (x * z + y * w);
  R[3] = 2 * (x * y + z * w); R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - x * w);
  R[6] = 2 * (x * z - y * w); R[7] = 2 * (y * z + x * w); R[8] = 1 - 2 * (x * x + y * y);
  for (int c = 0; c < 3; ++c)
    for (int r = 0; r < 3; ++r) m[c * 3 + r] = (float)R[r * 3 + c];
  m[9] = (float)T[4]; m[10] = (float)T[5]; m[11] = (float)T[6];
}

/* dense_tracking.cu:24-80 on one pixel; returns 1 if the pixel contributes */
static int pixel_terms(const odt_level *L, const float m[12], int u, int v, int exact, int want_jac,
                       float *res_out, float jac[6]) {
  const float *p = L->cloud + 4 * ((size_t)v * L->cloud_stride + u);
  if (!(p[3] > 0)) return 0;
  /* matTimesVec(GpuMatrix34): dotStride3 */
  const float cx = p[0] * m[0] + p[1] * m[3] + p[2] * m[6] + p[3] * m[9];
  const float cy = p[0] * m[1] + p[1] * m[4] + p[2] * m[7] + p[3] * m[10];
  const float cz = p[0] * m[2] + p[1] * m[5] + p[2] * m[8] + p[3] * m[11];
  /* cameraProject */
  const float uc = L->f * cx / cz + L->px;
  const float vc = L->f * cy / cz + L->py;
  if (!(uc >= 1.f && vc >= 1.f && uc <= (float)(L->w - 2) && vc <= (float)(L->h - 2))) return 0;
  const float ip = L->prev[(size_t)v * L->stride + u];
  const float ic = bilinear(L->cur, L->stride, uc, vc, exact);
  *res_out = ip - ic;
  if (want_jac) {
    float dx = 0.5f * bilinear(L->dx, L->stride, uc, vc, exact);
    float dy = 0.5f * bilinear(L->dy, L->stride, uc, vc, exact);
    /* frameJacobian, literally (including the double-typed literals) */
    const float z_sq = cz * cz;
    dx *= L->f;
    dy *= L->f;
    jac[0] = (float)(-dx * (1. / cz));
    jac[1] = (float)(-dy * 1. / cz);
    jac[2] = (dx * cx / z_sq + dy * cy / z_sq);
    jac[3] = (dx * (cx * cy) / z_sq + dy * (1.f + cy * cy / z_sq));
    jac[4] = (-dx * (1.f + (cx * cx / z_sq)) - dy * (cx * cy) / z_sq);
    jac[5] = (dx * cy / cz - dy * cx / cz);
  }
  return 1;
}

/* chi2 + (optionally) Hessian (21, GpuSymMatrix6 packing: for c: for r<=c) and J*res */
void odt_pass(const odt_level *L, const double T[7], int exact, double *chi2, double H21[21], double b6[6],
              int *n_valid) {
  float m[12];
  odt_pose_to_m34(T, m);
  double c2 = 0;
  int n = 0;
  if (H21) memset(H21, 0, 21 * sizeof(double));
  if (b6) memset(b6, 0, 6 * sizeof(double));
  for (int v = 0; v < L->h; ++v)
    for (int u = 0; u < L->w; ++u) {
      float res, jac[6];
      if (!pixel_terms(L, m, u, v, exact, H21 != 0, &res, jac)) continue;
      ++n;
      c2 += (double)(res * res);
      if (H21) {
        int i = 0;
        for (int r = 0; r < 6; ++r)       /* addOuter: data[i] += vec[r]*vec[c], c <= r */
          for (int c = 0; c <= r; ++c) H21[i++] += (double)(jac[r] * jac[c]);
        for (int r = 0; r < 6; ++r) b6[r] += (double)(jac[r] * res);
      }
    }
  *chi2 = c2;
  if (n_valid) *n_valid = n;
}

/* residualImage_kernel (gpu/dense_tracking.cu:494-541): out = w*h packed float4 */
void odt_residual_image(const odt_level *L, const double T[7], int exact, float *out) {
  float m[12];
  odt_pose_to_m34(T, m);
  for (int v = 0; v < L->h; ++v)
    for (int u = 0; u < L->w; ++u) {
      float *o = out + 4 * ((size_t)v * L->w + u);
      const float *p = L->cloud + 4 * ((size_t)v * L->cloud_stride + u);
      o[3] = 1.f;
      if (p[3] > 0) {
        float res, jac[6];
        if (pixel_terms(L, m, u, v, exact, 0, &res, jac)) {
          const float r2 = 1 - 50.f * res * res;
          const float g = r2 > 0.f ? r2 : 0.f;
          o[0] = o[1] = o[2] = g;
        } else {
          o[0] = 1.f; o[1] = 0.f; o[2] = 0.f;
        }
      } else {
        o[0] = 0.f; o[1] = 1.f; o[2] = 0.f;
      }
    }
}

/* solve (H + mu diag(H)) x = -b, H packed as above; plain Gaussian elimination with pivoting */
static void solve6(const double H21[21], const double b6[6], double mu, double x[6]) {
  double A[6][7];
  int i = 0;
  for (int r = 0; r < 6; ++r)
    for (int c = 0; c <= r; ++c) { A[r][c] = H21[i]; A[c][r] = H21[i]; ++i; }
  for (int r = 0; r < 6; ++r) { A[r][r] += mu * A[r][r]; A[r][6] = -b6[r]; }
  for (int c = 0; c < 6; ++c) {
    int p = c;
    for (int r = c + 1; r < 6; ++r) if (fabs(A[r][c]) > fabs(A[p][c])) p = r;
    if (p != c) for (int k = 0; k < 7; ++k) { double t = A[c][k]; A[c][k] = A[p][k]; A[p][k] = t; }
    if (A[c][c] == 0) continue;
    for (int r = c + 1; r < 6; ++r) {
      const double f = A[r][c] / A[c][c];
      for (int k = c; k < 7; ++k) A[r][k] -= f * A[c][k];
    }
  }
  for (int r = 5; r >= 0; --r) {
    double s = A[r][6];
    for (int k = r + 1; k < 6; ++k) s -= A[r][k] * x[k];
    x[r] = A[r][r] != 0 ? s / A[r][r] : 0.;
  }
}

/* DenseTracker::denseTrackingGpu (dense_tracking.cpp:62-193): levels n-1 .. 0 */
void odt_track(const odt_level *levels, int nlevels, double T[7], int exact, odt_stats *st) {
  if (st) memset(st, 0, sizeof *st);
  for (int l = nlevels - 1; l >= 0; --l) {
    const odt_level *L = &levels[l];
    double chi2, H[21], b[6];
    int passes = 1;
    odt_pass(L, T, exact, &chi2, H, b, 0);
    double nu = 2, mu = 0.01f;
    int stop = 0, trial = 0;
    for (int i = 0; i < 15; ++i) {
      double rho = 0;
      do {
        /* jacobianReduction at the accepted pose: (H, b) are those of the last accepted pass */
        double x[6], dT[7], Tn[7], chin, Hn[21], bn[6];
        solve6(H, b, mu, x);
        oba_se3_exp(x, dT);
        oba_se3_mul(dT, T, Tn);
        odt_pass(L, Tn, exact, &chin, Hn, bn, 0);
        ++passes;
        rho = chi2 - chin;
        if (rho > 0) {
          memcpy(T, Tn, sizeof Tn);
          chi2 = chin;
          double nm = 0;
          for (int k = 0; k < 6; ++k) nm = fmax(nm, fabs(b[k]));
          stop = nm <= 1e-10;          /* norm_max(b) <= EPS (global.h:107) */
          const double u = 2 * rho - 1;
          mu *= fmax(1. / 3., 1 - u * u * u);
          nu = 2.;
          trial = 0;
          memcpy(H, Hn, sizeof Hn);
          memcpy(b, bn, sizeof bn);
        } else {
          mu *= nu;
          nu *= 2.;
          ++trial;
          if (trial == 2) stop = 1;
        }
      } while (!(rho > 0 || stop));
      if (stop) break;
    }
    if (st && l < ODT_MAX_LEVELS) { st->chi2[l] = chi2; st->passes[l] = passes; }
  }
}

/* pointcloud_kernel (dense_tracking.cu:82-122) with TQ column-major float[16] */
void odt_point_cloud(const float TQ[16], const float *disp, int width, int height, int stride_in,
                     int stride_out, int factor, float *cloud) {
  for (int v = 0; v < height; ++v)
    for (int u = 0; u < width; ++u) {
      const int x = u * factor;
      const int idx_in = v * stride_in + x;   /* row is NOT scaled by factor */
      float *o = cloud + 4 * ((size_t)v * stride_out + u);
      const float d = disp[idx_in] * factor;
      if (d <= 0) {
        o[0] = 0.f; o[1] = 0.f; o[2] = 0.f; o[3] = -1.f;
      } else {
        const float uvd[4] = {(float)u, (float)v, d, 1.f};
        float p[4];
        for (int r = 0; r < 4; ++r)
          p[r] = uvd[0] * TQ[r] + uvd[1] * TQ[4 + r] + uvd[2] * TQ[8 + r] + uvd[3] * TQ[12 + r];
        o[0] = p[0] / p[3]; o[1] = p[1] / p[3]; o[2] = p[2] / p[3]; o[3] = 1.f;
      }
    }
}

/* TQ = T^-1 * Q (dense_tracking.cpp:204, stereo_camera.cpp:24-34) in double, cast to float col-major */
void odt_make_TQ(const double T_cur_from_actkey[7], double f, double px, double py, double b, float TQ[16]) {
  double Ti[7];
  oba_se3_inv(T_cur_from_actkey, Ti);
  const double x = Ti[0], y = Ti[1], z = Ti[2], w = Ti[3];
  double M[16] = {0};
  M[0] = 1 - 2 * (y * y + z * z); M[1] = 2 * (x * y - z * w); M[2] = 2 * (x * z + y * w); M[3] = Ti[4];
  M[4] = 2 * (x * y + z * w); M[5] = 1 - 2 * (x * x + z * z); M[6] = 2 * (y * z - x * w); M[7] = Ti[5];
  M[8] = 2 * (x * z - y * w); M[9] = 2 * (y * z + x * w); M[10] = 1 - 2 * (x * x + y * y); M[11] = Ti[6];
  M[15] = 1;
  const double Q[16] = {1, 0, 0, -px, 0, 1, 0, -py, 0, 0, 0, f, 0, 0, 1. / b, 0};
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) {
      double s = 0;
      for (int k = 0; k < 4; ++k) s += M[r * 4 + k] * Q[k * 4 + c];
      TQ[c * 4 + r] = (float)s;
    }
}

/* ================================================================ the reference's non-CUDA tracker (a18)
 * dense_tracking.cpp:222-423, maths_utils.cpp:33-65, transformations.h:116-140, stereo_camera.cpp:24-34.
 * TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED. */
#include "ba_oracle.h"

#define ODTC_NTH 4 /* DenseTracker::EVERY_NTH_PIXEL, dense_tracking.h:82 */

/* computeDensePointCloudCpu, dense_tracking.cpp:393-422 */
void odtc_point_cloud(const double T[7], double f, double px, double py, double b, const float *disp, int disp_stride,
                      int level, int w, int h, float *cloud) {
  double Ti[7], R[9];
  oba_se3_inv(T, Ti);
  {
    const double x = Ti[0], y = Ti[1], z = Ti[2], ww = Ti[3];
    R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - z * ww); R[2] = 2 * (x * z + y * ww);
    R[3] = 2 * (x * y + z * ww); R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - x * ww);
    R[6] = 2 * (x * z - y * ww); R[7] = 2 * (y * z + x * ww); R[8] = 1 - 2 * (x * x + y * y);
  }
  const double M[16] = {R[0], R[1], R[2], Ti[4], R[3], R[4], R[5], Ti[5], R[6], R[7], R[8], Ti[6], 0, 0, 0, 1};
  const double Q[16] = {1, 0, 0, -px, 0, 1, 0, -py, 0, 0, 0, f, 0, 0, 1. / b, 0};
  double TQ[16];
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) {
      double s = 0;
      for (int k = 0; k < 4; ++k) s += M[r * 4 + k] * Q[k * 4 + c];
      TQ[r * 4 + c] = s;
    }
  const int gw = w / ODTC_NTH, gh = h / ODTC_NTH;
  const double inv_factor = 1. / (double)(1 << level); /* pyrFromZero_d(1., level) */
  for (int v = 0; v < gh; ++v)
    for (int u = 0; u < gw; ++u) {
      /* interpolateDisparity(disp, (u*4, v*4), level), maths_utils.cpp:37-44 */
      const double d = (double)disp[(size_t)((v * 4) << level) * disp_stride + ((u * 4) << level)] * inv_factor;
      float *o = cloud + 4 * ((size_t)v * gw + u);
      if (d <= 0) {
        o[0] = 0.f; o[1] = 0.f; o[2] = 0.f; o[3] = -1.f;
      } else {
        const double uvd[4] = {(double)(u * ODTC_NTH), (double)(v * ODTC_NTH), d, 1.};
        double p[4];
        for (int r = 0; r < 4; ++r) p[r] = TQ[r * 4] * uvd[0] + TQ[r * 4 + 1] * uvd[1] + TQ[r * 4 + 2] * uvd[2] + TQ[r * 4 + 3] * uvd[3];
        o[0] = (float)(p[0] / p[3]); o[1] = (float)(p[1] / p[3]); o[2] = (float)(p[2] / p[3]); o[3] = 1.f;
      }
    }
}

/* interpolateMat_32f, maths_utils.cpp:46-65 */
static float interp32f(const float *img, int stride, float u, float v) {
  const float x = floorf(u), y = floorf(v);
  const float sx = u - x, sy = v - y;
  const float wx0 = 1 - sx, wx1 = sx, wy0 = 1 - sy, wy1 = sy;
  const int xi = (int)x, yi = (int)y;
  const float v00 = img[(size_t)yi * stride + xi], v01 = img[(size_t)(yi + 1) * stride + xi];
  const float v10 = img[(size_t)yi * stride + xi + 1], v11 = img[(size_t)(yi + 1) * stride + xi + 1];
  return (wx0 * wy0) * v00 + (wx0 * wy1) * v01 + (wx1 * wy0) * v10 + (wx1 * wy1) * v11;
}

void odtc_pass(const odtc_level *L, const double T[7], double *chi2, double H21[21], double Jres[6], int *n_valid) {
  const int gw = L->w / ODTC_NTH, gh = L->h / ODTC_NTH;
  double c2 = 0;
  int nv = 0;
  if (H21) memset(H21, 0, sizeof(double) * 21);
  if (Jres) memset(Jres, 0, sizeof(double) * 6);
  for (int v = 0; v < gh; ++v)
    for (int u = 0; u < gw; ++u) {
      const float *c4 = L->cloud + 4 * ((size_t)v * gw + u);
      if (!(c4[3] > 0)) continue;
      const double xp[3] = {c4[0], c4[1], c4[2]};
      double xc[3];
      oba_se3_act(T, xp, xc);
      /* cam.map(project2d(xyz_cur)).cast<float>() */
      const float uc = (float)(L->f * (xc[0] / xc[2]) + L->px), vc = (float)(L->f * (xc[1] / xc[2]) + L->py);
      const int ui = (int)uc, vi = (int)vc; /* cast<int>: truncation */
      if (!(ui >= 2 && ui < L->w - 2 && vi >= 2 && vi < L->h - 2)) continue; /* isInFrame(uv, 2) */
      const float ip = (float)((1. / 255.) * L->prev_u8[(size_t)(v * ODTC_NTH) * L->pitch_u8 + u * ODTC_NTH]);
      const float ic = interp32f(L->cur, L->stride, uc, vc);
      float res = ip - ic;
      if (res > 0.1) res = 0.1f;
      if (res < -0.1) res = -0.1f;
      c2 += (double)(res * res);
      ++nv;
      if (H21) {
        const float dx = (float)(0.5 * interp32f(L->dx, L->stride, uc, vc));
        const float dy = (float)(0.5 * interp32f(L->dy, L->stride, uc, vc));
        /* frame_jac_xyz2uv, transformations.h:116-140 */
        const double x = xc[0], y = xc[1], z = xc[2], z2 = z * z, f = L->f;
        const double r0[6] = {-1. / z * f, 0, x / z2 * f, x * y / z2 * f, -(1 + (x * x / z2)) * f, y / z * f};
        const double r1[6] = {0, -1. / z * f, y / z2 * f, (1 + y * y / z2) * f, -x * y / z2 * f, -x / z * f};
        double J[6];
        for (int k = 0; k < 6; ++k) J[k] = dx * r0[k] + dy * r1[k];
        int q = 0;
        for (int r = 0; r < 6; ++r) {
          for (int c = r; c < 6; ++c) H21[q++] += J[r] * J[c];
          Jres[r] += J[r] * res;
        }
      }
    }
  *chi2 = c2;
  if (n_valid) *n_valid = nv;
}

/* H x = -Jres, H from its upper triangle (Eigen ldlt() in the reference, dense_tracking.cpp:332) */
static void odtc_solve(const double H21[21], const double Jres[6], double x[6]) {
  double A[6][6], Lm[6][6], D[6], y[6];
  int q = 0;
  for (int r = 0; r < 6; ++r)
    for (int c = r; c < 6; ++c) { A[r][c] = A[c][r] = H21[q++]; }
  for (int j = 0; j < 6; ++j) {
    double d = A[j][j];
    for (int k = 0; k < j; ++k) d -= Lm[j][k] * Lm[j][k] * D[k];
    D[j] = d;
    for (int i = j + 1; i < 6; ++i) {
      double s = A[i][j];
      for (int k = 0; k < j; ++k) s -= Lm[i][k] * Lm[j][k] * D[k];
      Lm[i][j] = d != 0. ? s / d : 0.;
    }
  }
  for (int i = 0; i < 6; ++i) {
    double s = -Jres[i];
    for (int k = 0; k < i; ++k) s -= Lm[i][k] * y[k];
    y[i] = s;
  }
  for (int i = 5; i >= 0; --i) {
    double s = D[i] != 0. ? y[i] / D[i] : 0.;
    for (int k = i + 1; k < 6; ++k) s -= Lm[k][i] * x[k];
    x[i] = s;
  }
}

void odtc_track(const odtc_level *levels, int nlevels, double T[7], odt_stats *st) {
  if (st) memset(st, 0, sizeof *st);
  for (int level = nlevels - 1; level >= 0; --level) {
    const odtc_level *L = levels + level;
    double chi2, H[21], Jr[6];
    int passes = 1;
    odtc_pass(L, T, &chi2, H, Jr, NULL); /* :229-262 and the first :276-331 sweep see the same pose */
    for (int i = 0; i < 15; ++i) {
      double x[6], dT[7], Tn[7], chin, Hn[21], Jn[6];
      odtc_solve(H, Jr, x);
      oba_se3_exp(x, dT);
      oba_se3_mul(dT, T, Tn);
      odtc_pass(L, Tn, &chin, Hn, Jn, NULL);
      ++passes;
      const double rho = chi2 - chin;
      if (rho > 0) { /* :368-376 */
        memcpy(T, Tn, sizeof Tn);
        chi2 = chin;
        memcpy(H, Hn, sizeof H);
        memcpy(Jr, Jn, sizeof Jr);
        double nm = 0;
        for (int k = 0; k < 6; ++k) nm = fmax(nm, fabs(x[k]));
        if (nm <= 0.0000000001) break; /* stop = norm_max(x) <= EPS */
      } else {
        break; /* :378-385: the identical trial is rejected once more (trial == 2) and the level ends */
      }
    }
    if (st) { st->chi2[level] = chi2; st->passes[level] = passes; }
  }
}
