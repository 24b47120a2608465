// ba_dev.cuh -- device helpers shared by the BA kernels (edge residuals, small inverses,
// pose-pose constraint arithmetic).
#pragma once
#include "ba_types.cuh"
#include "se3_dev.cuh"

namespace svs {

// ------------------------------------------------------------------ helpers

__device__ __forceinline__ void load12(const double* __restrict__ Rt, int p, double R[9], double t[3]) {
  const double* q = Rt + 12 * (size_t)p;
#pragma unroll
  for (int i = 0; i < 9; ++i) R[i] = __ldg(q + i);
  t[0] = __ldg(q + 9); t[1] = __ldg(q + 10); t[2] = __ldg(q + 11);
}

// T_ca = T_c * T_a^-1  as rotation + translation
__device__ __forceinline__ void rel_pose(const double Rc[9], const double tc[3], const double Ra[9],
                                         const double ta[3], double R[9], double t[3]) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      R[i * 3 + j] = Rc[i * 3] * Ra[j * 3] + Rc[i * 3 + 1] * Ra[j * 3 + 1] + Rc[i * 3 + 2] * Ra[j * 3 + 2];
#pragma unroll
  for (int i = 0; i < 3; ++i) t[i] = tc[i] - (R[i * 3] * ta[0] + R[i * 3 + 1] * ta[1] + R[i * 3 + 2] * ta[2]);
}

// e = z - pi_stereo(y)   (G2oCameraParameters::stereocam_uvu_map, anchored_points.cpp:43-50)
__device__ __forceinline__ void stereo_residual(const BaDev& d, const double y[3], const double obs[3], double e[3]) {
  e[0] = obs[0] - ((y[0] / y[2]) * d.f + d.px);
  e[1] = obs[1] - ((y[1] / y[2]) * d.f + d.py);
  e[2] = obs[2] - (((y[0] - d.b) / y[2]) * d.f + d.px);
}

// robust cost of one observation at (pose, anchor, psi)
__device__ __forceinline__ double edge_cost(const BaDev& d, const double* __restrict__ Rt, int ip, const double Ra[9],
                                            const double ta[3], const double xa[3], const double obs[3],
                                            const double om[3], int robust, double delta) {
  double Rc[9], tc[3], R[9], t[3], y[3], e[3];
  load12(Rt, ip, Rc, tc);
  rel_pose(Rc, tc, Ra, ta, R, t);
  mat3_vec(R, xa, y);
  y[0] += t[0]; y[1] += t[1]; y[2] += t[2];
  stereo_residual(d, y, obs, e);
  const double e2 = e[0] * e[0] * om[0] + e[1] * e[1] * om[1] + e[2] * e[2] * om[2];
  if (!robust) return e2;
  double r0, r1;
  huber(e2, delta, r0, r1);
  return r0;
}

__device__ __forceinline__ void inv3_sym_lambda(const double* __restrict__ D6, double lambda, double Di[9]) {
  // (Hll + lambda I)^-1 by cofactors (Eigen's 3x3 inverse, as used by g2o's D->inverse())
  const double a00 = D6[0] + lambda, a01 = D6[1], a02 = D6[2], a11 = D6[3] + lambda, a12 = D6[4], a22 = D6[5] + lambda;
  const double c00 = a11 * a22 - a12 * a12, c01 = a12 * a02 - a01 * a22, c02 = a01 * a12 - a11 * a02;
  const double det = a00 * c00 + a01 * c01 + a02 * c02;
  const double id = 1. / det;
  Di[0] = c00 * id; Di[1] = c01 * id; Di[2] = c02 * id;
  Di[3] = Di[1];    Di[4] = (a00 * a22 - a02 * a02) * id; Di[5] = (a01 * a02 - a00 * a12) * id;
  Di[6] = Di[2];    Di[7] = Di[5]; Di[8] = (a00 * a11 - a01 * a01) * id;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------ pose-pose constraint (one thread)

__device__ inline void third(const double A[7], const double dd[6], double out[36]) {
  // anchored_points.cpp:207-215: Adj_A + 1/2 ad_d Adj_A + 1/12 ad_d^2 Adj_A,
  // ad_d = SE3::d_lieBracketab_by_d_a(d)
  double Adj[36], dl[36], t1[36];
  se3_adj(A, Adj);
  double hu[9], ho[9];
  hat3(dd, hu);
  hat3(dd + 3, ho);
#pragma unroll
  for (int i = 0; i < 36; ++i) dl[i] = 0;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      dl[i * 6 + j] = -ho[i * 3 + j];
      dl[i * 6 + 3 + j] = -hu[i * 3 + j];
      dl[(i + 3) * 6 + 3 + j] = -ho[i * 3 + j];
    }
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) {
      double s = 0;
      for (int k = 0; k < 6; ++k) s += dl[i * 6 + k] * Adj[k * 6 + j];
      t1[i * 6 + j] = s;
    }
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) {
      double s = 0;
      for (int k = 0; k < 6; ++k) s += dl[i * 6 + k] * t1[k * 6 + j];
      out[i * 6 + j] = Adj[i * 6 + j] + 0.5 * t1[i * 6 + j] + (1. / 12.) * s;
    }
}

__device__ inline void constraint_error(const BaDev& d, const double* __restrict__ pose, int c, double err[6]) {
  double T21[7], T1[7], T2[7], T2i[7], A[7], B[7];
#pragma unroll
  for (int i = 0; i < 7; ++i) {
    T21[i] = d.c_T[7 * (size_t)c + i];
    T1[i] = pose[7 * (size_t)d.c_i[c] + i];
    T2[i] = pose[7 * (size_t)d.c_j[c] + i];
  }
  se3_inv(T2, T2i);
  se3_mul(T21, T1, A);
  se3_mul(A, T2i, B);
  se3_log(B, err);
}

__device__ inline double constraint_chi2(const BaDev& d, const double* __restrict__ pose, int c) {
  double err[6];
  constraint_error(d, pose, c, err);
  const double* Lm = d.c_Lam + 36 * (size_t)c;
  double chi = 0;
  for (int a = 0; a < 6; ++a)
    for (int b = 0; b < 6; ++b) chi += err[a] * Lm[a * 6 + b] * err[b];
  return chi;
}

// g2o BaseBinaryEdge::constructQuadraticForm for one G2oEdgeSE3
__device__ inline void constraint_build(const BaDev& d, const double* __restrict__ pose, int c) {
  const int i = d.c_i[c], j = d.c_j[c];
  double err[6], Ji[36], Jj[36], T21[7];
  constraint_error(d, pose, c, err);
#pragma unroll
  for (int k = 0; k < 7; ++k) T21[k] = d.c_T[7 * (size_t)c + k];
  third(T21, err, Ji);
  {
    const double I7[7] = {0, 0, 0, 1, 0, 0, 0};
    double md[6];
    for (int k = 0; k < 6; ++k) md[k] = -err[k];
    third(I7, md, Jj);
    for (int k = 0; k < 36; ++k) Jj[k] = -Jj[k];
  }
  if (d.fixed[i]) for (int k = 0; k < 36; ++k) Ji[k] = 0;
  if (d.fixed[j]) for (int k = 0; k < 36; ++k) Jj[k] = 0;
  const double* Lm = d.c_Lam + 36 * (size_t)c;
  double chi = 0, Oe[6];
  for (int a = 0; a < 6; ++a) {
    double s = 0;
    for (int b = 0; b < 6; ++b) { s += Lm[a * 6 + b] * err[b]; chi += err[a] * Lm[a * 6 + b] * err[b]; }
    Oe[a] = -s;
  }
  d.chi_c[c] = chi;
  const int tii = d.tbl[(size_t)i * d.P + i], tjj = d.tbl[(size_t)j * d.P + j], tij = d.tbl[(size_t)i * d.P + j];
  double* Sii = d.S + 36 * (size_t)(tii >> 1);
  double* Sjj = d.S + 36 * (size_t)(tjj >> 1);
  double* Sij = d.S + 36 * (size_t)(tij >> 1);
  const int tr = tij & 1;
  for (int a = 0; a < 6; ++a) {
    double AtO[6], BtO[6];
    for (int b = 0; b < 6; ++b) {
      double sa = 0, sb = 0;
      for (int k = 0; k < 6; ++k) { sa += Ji[k * 6 + a] * Lm[k * 6 + b]; sb += Jj[k * 6 + a] * Lm[k * 6 + b]; }
      AtO[b] = sa; BtO[b] = sb;
    }
    double bi = 0, bj = 0;
    for (int k = 0; k < 6; ++k) { bi += Ji[k * 6 + a] * Oe[k]; bj += Jj[k * 6 + a] * Oe[k]; }
    atomicAdd(d.bp + 6 * i + a, bi);
    atomicAdd(d.bp + 6 * j + a, bj);
    for (int b = 0; b < 6; ++b) {
      double sii = 0, sjj = 0, sij = 0;
      for (int k = 0; k < 6; ++k) {
        sii += AtO[k] * Ji[k * 6 + b];
        sjj += BtO[k] * Jj[k * 6 + b];
        sij += AtO[k] * Jj[k * 6 + b];
      }
      atomicAdd(Sii + a * 6 + b, sii);
      atomicAdd(Sjj + a * 6 + b, sjj);
      atomicAdd(Sij + (tr ? b * 6 + a : a * 6 + b), sij);   // rows <-> pose i unless transposed
    }
  }
}

}  // namespace svs
