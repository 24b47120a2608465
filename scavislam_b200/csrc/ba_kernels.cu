// ba_kernels.cu -- sm_100a kernels of the double-window BA iteration.
//
//   k_build   fused linearise + J^T W J + 3x3 landmark elimination + Schur scatter
//             (G2oEdgeProjectPSI2UVU::computeError/linearizeOplus, anchored_points.cpp:148-189;
//              g2o BaseMultiEdge::constructQuadraticForm, BlockSolver<6,3>::buildSystem and the
//              Schur part of BlockSolver::solve) -- one warp per landmark; plus the pose-pose
//             constraints G2oEdgeSE3 (anchored_points.cpp:207-235) on trailing CTAs.
//   k_solve   block-sparse Cholesky of the reduced camera system, forward/backward solve,
//             pose update T <- exp(dx) T (LinearSolverCSparse::solve, slam_graph.cpp:55-60;
//             G2oVertexSE3::oplusImpl, anchored_points.cpp:53-58)
//   k_update  landmark back-substitution, psi += dpsi (G2oVertexPointXYZ::oplusImpl :78-83),
//             robust chi2 of the trial state
//   k_decide  Levenberg-Marquardt accept/reject logic of
//             g2o::OptimizationAlgorithmLevenberg::solve (lambda0/trials set at slam_graph.cpp:338-342,1073)
#include "ba_kernels.cuh"
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

__device__ void third(const double A[7], const double dd[6], double out[36]) {
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

__device__ void constraint_error(const BaDev& d, const double* __restrict__ pose, int c, double err[6]) {
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

__device__ double constraint_chi2(const BaDev& d, const double* __restrict__ pose, int c) {
  double err[6];
  constraint_error(d, pose, c, err);
  const double* Lm = d.c_Lam + 36 * (size_t)c;
  double chi = 0;
  for (int a = 0; a < 6; ++a)
    for (int b = 0; b < 6; ++b) chi += err[a] * Lm[a * 6 + b] * err[b];
  return chi;
}

// g2o BaseBinaryEdge::constructQuadraticForm for one G2oEdgeSE3
__device__ void constraint_build(const BaDev& d, const double* __restrict__ pose, int c) {
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

// ------------------------------------------------------------------ k_build

// Per-warp shared-memory plan (doubles): per edge lane J~p[18] J~a[18] (stride 19), J~psi[9], e~[3];
// per slot B[18] Y[18] (stride 19); 16 scratch.  Then ints: pair table, slot poses.
constexpr int kSJ = 19;
__host__ __device__ inline int build_warp_doubles(int Kmax) { return (2 * kSJ + 12) * Kmax + 2 * kSJ * Kmax + 16; }
__host__ __device__ inline int build_warp_ints(int Kmax) { return ((Kmax * (Kmax + 1) / 2 + Kmax + 1) / 2) * 2; }
size_t build_smem_bytes(int warps, int Kmax) {
  return (size_t)warps * ((size_t)build_warp_doubles(Kmax) * 8 + (size_t)build_warp_ints(Kmax) * 4);
}

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
k_build(BaDev d, int lm_begin, int lm_end, int Kmax, int robust, double delta, int n_lm_blocks) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const LmCtl* __restrict__ ctl = d.ctl;
  const int cur = ctl->cur;
  if ((int)blockIdx.x >= n_lm_blocks) {   // pose-pose constraints, one thread each
    const int c = ((int)blockIdx.x - n_lm_blocks) * (WARPS * 32) + (int)threadIdx.x;
    if (c < d.C) constraint_build(d, d.pose[cur], c);
    return;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int li = lm_begin + (int)blockIdx.x * WARPS + warp;
  if (li >= lm_end) return;
  const double lambda = ctl->lambda;

  const int wd = build_warp_doubles(Kmax);
  double* sm = reinterpret_cast<double*>(smem_raw) + (size_t)warp * wd;
  double* sJp = sm;                       // [k][19]
  double* sJa = sJp + kSJ * Kmax;         // [k][19]
  double* sJs = sJa + kSJ * Kmax;         // [k][9]
  double* sE = sJs + 9 * Kmax;            // [k][3]
  double* sB = sE + 3 * Kmax;             // [K][19]
  double* sY = sB + kSJ * Kmax;           // [K][19]
  double* sD = sY + kSJ * Kmax;           // 16: D(6) bl(3) Dinv? -> D6, bl3
  int* si = reinterpret_cast<int*>(reinterpret_cast<double*>(smem_raw) + (size_t)WARPS * wd) +
            (size_t)warp * build_warp_ints(Kmax);
  int* sPose = si;                        // [K]
  int* sPair = si + Kmax;                 // [npairs]

  const int e0 = d.lm_eptr[li], k = d.lm_eptr[li + 1] - e0;
  const int s0 = d.lm_sptr[li], K = d.lm_sptr[li + 1] - s0;
  double* Dbl = d.Dbl + 12 * (size_t)li;
  if (k == 0) {
    if (lane < 12) Dbl[lane] = 0;
    if (lane == 0) d.chi_l[li] = 0;
    return;
  }
  const int has_self = d.lm_self[li];
  const int off = has_self ? 0 : 1;       // slot of edge lane i is i + off
  const int ia = d.lm_anchor[li];
  const double* __restrict__ Rt = d.Rt[cur];
  const double* __restrict__ psi = d.psi[cur] + 3 * (size_t)li;

  double Ra[9], ta[3];
  load12(Rt, ia, Ra, ta);
  const double p0 = __ldg(psi), p1 = __ldg(psi + 1), p2 = __ldg(psi + 2);
  const double ipz = 1. / p2;
  const double xa[3] = {p0 * ipz, p1 * ipz, ipz};   // invert_depth (maths_utils.h:66-69)
  const int fa = d.fixed[ia];

  double chi = 0;
  if (lane == 0) sPose[0] = ia;
  if (lane < k) {
    const int e = e0 + lane;
    const int ip = d.e_pose[e];
    sPose[lane + off] = ip;   // for the self edge (lane 0, off 0) this rewrites the anchor
    const double obs[3] = {__ldg(d.e_obs + e), __ldg(d.e_obs + (size_t)d.E + e), __ldg(d.e_obs + 2 * (size_t)d.E + e)};
    const double om[3] = {__ldg(d.e_w + e), __ldg(d.e_w + (size_t)d.E + e), __ldg(d.e_w + 2 * (size_t)d.E + e)};
    double Rc[9], tc[3], R[9], t[3], y[3], er[3];
    load12(Rt, ip, Rc, tc);
    rel_pose(Rc, tc, Ra, ta, R, t);
    mat3_vec(R, xa, y);
    y[0] += t[0]; y[1] += t[1]; y[2] += t[2];
    stereo_residual(d, y, obs, er);
    const double e2 = er[0] * er[0] * om[0] + er[1] * er[1] * om[1] + er[2] * er[2] * om[2];
    double r0 = e2, r1 = 1.;
    if (robust) huber(e2, delta, r0, r1);
    chi = r0;
    const double sw[3] = {sqrt(r1 * om[0]), sqrt(r1 * om[1]), sqrt(r1 * om[2])};   // sqrt(rho' Omega)
    // d_stereoproj_d_y (transformations.h:62-71): rows (a 0 c0) (0 a c1) (a 0 c2)
    const double iz = 1. / y[2];
    const double a = d.f * iz;
    const double c0 = -(d.f * y[0]) * iz * iz, c1 = -(d.f * y[1]) * iz * iz, c2 = -(d.f * (y[0] - d.b)) * iz * iz;
    double* Jp = sJp + kSJ * lane;
    double* Ja = sJa + kSJ * lane;
    double* Js = sJs + 9 * lane;
    // J_pose = -Jcam [I | -hat(y)]  (anchored_points.cpp:187, transformations.h:73-80)
    const int fp = d.fixed[ip];
    const double zp = fp ? 0. : 1.;
    Jp[0] = zp * sw[0] * -a;  Jp[1] = 0;                 Jp[2] = zp * sw[0] * -c0;
    Jp[3] = zp * sw[0] * (-c0 * y[1]);  Jp[4] = zp * sw[0] * (-a * y[2] + c0 * y[0]);  Jp[5] = zp * sw[0] * (a * y[1]);
    Jp[6] = 0;                Jp[7] = zp * sw[1] * -a;   Jp[8] = zp * sw[1] * -c1;
    Jp[9] = zp * sw[1] * (a * y[2] - c1 * y[1]);  Jp[10] = zp * sw[1] * (c1 * y[0]);  Jp[11] = zp * sw[1] * (-a * y[0]);
    Jp[12] = zp * sw[2] * -a; Jp[13] = 0;                Jp[14] = zp * sw[2] * -c2;
    Jp[15] = zp * sw[2] * (-c2 * y[1]); Jp[16] = zp * sw[2] * (-a * y[2] + c2 * y[0]); Jp[17] = zp * sw[2] * (a * y[1]);
    // J_anchor = Jcam R [I | -hat(x_a)]  (anchored_points.cpp:188)
    const double za = fa ? 0. : 1.;
    double M[9];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      M[j] = a * R[j] + c0 * R[6 + j];
      M[3 + j] = a * R[3 + j] + c1 * R[6 + j];
      M[6 + j] = a * R[j] + c2 * R[6 + j];
    }
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const double s = za * sw[q];
      Ja[q * 6 + 0] = s * M[q * 3 + 0];
      Ja[q * 6 + 1] = s * M[q * 3 + 1];
      Ja[q * 6 + 2] = s * M[q * 3 + 2];
      Ja[q * 6 + 3] = s * -(M[q * 3 + 1] * xa[2] - M[q * 3 + 2] * xa[1]);
      Ja[q * 6 + 4] = s * -(-M[q * 3 + 0] * xa[2] + M[q * 3 + 2] * xa[0]);
      Ja[q * 6 + 5] = s * -(M[q * 3 + 0] * xa[1] - M[q * 3 + 1] * xa[0]);
    }
    // J_psi = -Jcam d_Tinvpsi_d_psi (anchored_points.cpp:186, transformations.h:82-95)
    double N[9];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      N[i * 3 + 0] = R[i * 3 + 0] * ipz;
      N[i * 3 + 1] = R[i * 3 + 1] * ipz;
      N[i * 3 + 2] = -(y[i] - t[i]) * ipz;
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      Js[j] = -sw[0] * (a * N[j] + c0 * N[6 + j]);
      Js[3 + j] = -sw[1] * (a * N[3 + j] + c1 * N[6 + j]);
      Js[6 + j] = -sw[2] * (a * N[j] + c2 * N[6 + j]);
    }
    sE[3 * lane + 0] = sw[0] * er[0];
    sE[3 * lane + 1] = sw[1] * er[1];
    sE[3 * lane + 2] = sw[2] * er[2];
    // own Hpl block B = J~p^T J~psi (6x3); the self edge's block is cancelled by its anchor part
    if (!(has_self && lane == 0)) {
      double* B = sB + kSJ * (lane + off);
#pragma unroll
      for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c)
          B[r * 3 + c] = Jp[r] * Js[c] + Jp[6 + r] * Js[3 + c] + Jp[12 + r] * Js[6 + c];
    }
  }
  chi = warp_sum(chi);
  if (lane == 0) d.chi_l[li] = chi;
  // pair table: (dst block, transpose, m, n) for all slot pairs m <= n
  __syncwarp();
  const int npairs = K * (K + 1) / 2;
  for (int pidx = lane; pidx < npairs; pidx += 32) {
    int m = 0, rem = pidx;
    while (rem >= K - m) { rem -= K - m; ++m; }
    const int n = m + rem;
    const int t = d.tbl[(size_t)sPose[m] * d.P + sPose[n]];
    sPair[pidx] = ((t >> 1) << 11) | ((t & 1) << 10) | (m << 5) | n;
  }
  // landmark sums: anchor Hpl block (18), Hll upper (6), b_l (3)
  const int i_first = has_self ? 1 : 0;   // first non-self edge lane
  if (lane < 27) {
    double s = 0;
    if (lane < 18) {
      const int r = lane / 3, c = lane % 3;
      for (int i = i_first; i < k; ++i)
        s += sJa[kSJ * i + r] * sJs[9 * i + c] + sJa[kSJ * i + 6 + r] * sJs[9 * i + 3 + c] +
             sJa[kSJ * i + 12 + r] * sJs[9 * i + 6 + c];
      sB[lane] = s;
    } else if (lane < 24) {
      const int t = lane - 18;
      const int r = t < 3 ? 0 : (t < 5 ? 1 : 2), c = t < 3 ? t : (t < 5 ? t - 2 : 2);
      for (int i = 0; i < k; ++i)
        s += sJs[9 * i + r] * sJs[9 * i + c] + sJs[9 * i + 3 + r] * sJs[9 * i + 3 + c] +
             sJs[9 * i + 6 + r] * sJs[9 * i + 6 + c];
      sD[t] = s;
    } else {
      const int c = lane - 24;
      for (int i = 0; i < k; ++i)
        s -= sJs[9 * i + c] * sE[3 * i] + sJs[9 * i + 3 + c] * sE[3 * i + 1] + sJs[9 * i + 6 + c] * sE[3 * i + 2];
      sD[6 + c] = s;
    }
  }
  __syncwarp();
  if (lane < 9) Dbl[lane] = sD[lane];
  double Di[9];
  inv3_sym_lambda(sD, lambda, Di);
  const double bl[3] = {sD[6], sD[7], sD[8]};
  // Y = B Dinv per slot; spill B to HBM (SoA) for the back-substitution
  for (int it = lane; it < K * 18; it += 32) {
    const int s = it / 18, rc = it % 18, r = rc / 3, c = rc % 3;
    const double* B = sB + kSJ * s + r * 3;
    sY[kSJ * s + rc] = B[0] * Di[c] + B[1] * Di[3 + c] + B[2] * Di[6 + c];
  }
  for (int c = 0; c < 18; ++c)
    if (lane < K) d.W[(size_t)c * d.nslots + s0 + lane] = sB[kSJ * lane + c];
  __syncwarp();
  // Schur scatter: for every slot pair, direct J^T W J part minus Y_m B_n^T
  const int skip_self = (d.flags & 1);
  const int total = npairs * 36;
  for (int fidx = lane; fidx < total; fidx += 32) {
    const int pidx = fidx / 36, el = fidx - pidx * 36;
    const int r = el / 6, c = el - r * 6;
    const int pk = sPair[pidx];
    const int m = (pk >> 5) & 31, n = pk & 31;
    const double* Ym = sY + kSJ * m + r * 3;
    const double* Bn = sB + kSJ * n + c * 3;
    double v = -(Ym[0] * Bn[0] + Ym[1] * Bn[1] + Ym[2] * Bn[2]);
    if (m == n) {
      if (m > 0) {
        const double* Jp = sJp + kSJ * (m - off);
        v += Jp[r] * Jp[c] + Jp[6 + r] * Jp[6 + c] + Jp[12 + r] * Jp[12 + c];
      } else {
        // anchor diagonal: all edges' J~a^T J~a; the self edge keeps g2o's J1^T W J1
        for (int i = (skip_self ? i_first : 0); i < k; ++i) {
          const double* Ja = sJa + kSJ * i;
          v += Ja[r] * Ja[c] + Ja[6 + r] * Ja[6 + c] + Ja[12 + r] * Ja[12 + c];
        }
      }
    } else if (m == 0) {
      const double* Ja = sJa + kSJ * (n - off);
      const double* Jp = sJp + kSJ * (n - off);
      v += Ja[r] * Jp[c] + Ja[6 + r] * Jp[6 + c] + Ja[12 + r] * Jp[12 + c];
    }
    double* dst = d.S + 36 * (size_t)(pk >> 11) + (((pk >> 10) & 1) ? c * 6 + r : el);
    atomicAdd(dst, v);
  }
  // gradients: bp = -J^T W e, bc = Y b_l
  for (int it = lane; it < K * 6; it += 32) {
    const int s = it / 6, r = it - s * 6;
    double g = 0;
    if (s > 0) {
      const int i = s - off;
      g = -(sJp[kSJ * i + r] * sE[3 * i] + sJp[kSJ * i + 6 + r] * sE[3 * i + 1] + sJp[kSJ * i + 12 + r] * sE[3 * i + 2]);
    } else {
      for (int i = i_first; i < k; ++i)
        g -= sJa[kSJ * i + r] * sE[3 * i] + sJa[kSJ * i + 6 + r] * sE[3 * i + 1] + sJa[kSJ * i + 12 + r] * sE[3 * i + 2];
    }
    const double* Y = sY + kSJ * s + r * 3;
    const double corr = Y[0] * bl[0] + Y[1] * bl[1] + Y[2] * bl[2];
    const int p = sPose[s];
    atomicAdd(d.bp + 6 * p + r, g);
    atomicAdd(d.bc + 6 * p + r, corr);
  }
}

void launch_build(const BaDev& d, int lm_begin, int lm_end, int Kmax, int robust, double delta, cudaStream_t st) {
  constexpr int WARPS = 8;
  const int nl = lm_end - lm_begin;
  const int n_lm_blocks = (nl + WARPS - 1) / WARPS;
  const int n_c_blocks = (d.C + WARPS * 32 - 1) / (WARPS * 32);
  const size_t smem = build_smem_bytes(WARPS, Kmax);
  static size_t configured = 0;
  if (smem > configured) {
    cudaFuncSetAttribute(k_build<WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    configured = smem;
  }
  if (n_lm_blocks + n_c_blocks == 0) return;
  k_build<WARPS><<<n_lm_blocks + n_c_blocks, WARPS * 32, smem, st>>>(d, lm_begin, lm_end, Kmax, robust, delta, n_lm_blocks);
}

// ------------------------------------------------------------------ k_solve

constexpr int kSolveThreads = 512;
constexpr int kPanelCap = 160;   // sub-diagonal blocks of one column staged in shared memory

// In-place lower Cholesky of a 6x6 (row-major, lower triangle authoritative) by one warp, then
// its inverse.  Lanes 0..5 own rows.  Returns false when a pivot is not positive.
__device__ bool chol6_inv(double* A, double* Ai, int lane) {
  bool ok = true;
  for (int c = 0; c < 6; ++c) {
    double dv = A[c * 6 + c];
    ok = ok && (dv > 0.);
    const double dd = sqrt(dv);
    __syncwarp();
    if (lane == c) A[c * 6 + c] = dd;
    if (lane > c && lane < 6) A[lane * 6 + c] /= dd;
    __syncwarp();
    // trailing update: row `lane`, columns c+1..lane
    if (lane > c && lane < 6)
      for (int c2 = c + 1; c2 <= lane; ++c2) A[lane * 6 + c2] -= A[lane * 6 + c] * A[c2 * 6 + c];
    __syncwarp();
  }
  // inverse of lower-triangular A: lane = column of the inverse
  if (lane < 6) {
    const int c = lane;
    for (int r = 0; r < 6; ++r) {
      double v = (r == c) ? 1. : 0.;
      for (int q = c; q < r; ++q) v -= A[r * 6 + q] * Ai[q * 6 + c];
      Ai[r * 6 + c] = (r < c) ? 0. : v / A[r * 6 + r];
    }
  }
  __syncwarp();
  return ok;
}

__global__ void __launch_bounds__(kSolveThreads)
k_solve(BaDev d) {
  __shared__ double sDg[36], sLi[36], sY[6];
  __shared__ double sPanel[kPanelCap * 36];
  __shared__ double sRed[kSolveThreads / 32];
  __shared__ int sFail;
  LmCtl* ctl = d.ctl;
  const int t = threadIdx.x, nt = blockDim.x, lane = t & 31, warp = t >> 5;
  const int P = d.P;
  const double lambda = ctl->lambda;
  const int cur = ctl->cur;
  if (t == 0) sFail = 0;
  // right-hand side in elimination order: bs = bp - bc
  for (int i = t; i < 6 * P; i += nt) {
    const int j = i / 6, r = i - 6 * j;
    const int p = d.perm[j];
    d.ywork[i] = d.bp[6 * p + r] - d.bc[6 * p + r];
  }
  __syncthreads();
  for (int j = 0; j < P; ++j) {
    const int base = d.col_ptr[j], nb = d.col_ptr[j + 1] - base - 1;
    const bool in_smem = nb <= kPanelCap;
    double* colS = d.S + 36 * (size_t)(base + 1);
    if (t < 36) {
      double v = d.S[36 * (size_t)base + t];
      if (t % 7 == 0) v += lambda + (d.fixed[d.perm[j]] ? 1. : 0.);
      sDg[t] = v;
    }
    if (t >= 64 && t < 70) sY[t - 64] = d.ywork[6 * j + (t - 64)];
    if (in_smem)
      for (int i = t; i < nb * 36; i += nt) sPanel[i] = colS[i];
    __syncthreads();
    if (warp == 0) {
      const bool ok = chol6_inv(sDg, sLi, lane);
      if (!ok && lane == 0) sFail = 1;
    }
    __syncthreads();
    if (sFail) break;
    // L_ij = S_ij L_jj^-T  (row r of block i times Linv^T); y_j = Linv b_j
    double* panel = in_smem ? sPanel : colS;
    for (int i = t; i < nb * 6; i += nt) {
      double* row = panel + 6 * (size_t)i;
      double v[6], o[6];
#pragma unroll
      for (int q = 0; q < 6; ++q) v[q] = row[q];
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        double s = 0;
#pragma unroll
        for (int q = 0; q < 6; ++q) s += (q <= c) ? v[q] * sLi[c * 6 + q] : 0.;
        o[c] = s;
      }
#pragma unroll
      for (int q = 0; q < 6; ++q) row[q] = o[q];
      if (in_smem) {
#pragma unroll
        for (int q = 0; q < 6; ++q) colS[6 * (size_t)i + q] = o[q];
      }
    }
    if (t < 36) d.Linv[36 * (size_t)j + t] = sLi[t];
    double yj = 0;
    if (t >= 64 && t < 70) {
      const int r = t - 64;
      for (int q = 0; q <= r; ++q) yj += sLi[r * 6 + q] * sY[q];
    }
    __syncthreads();
    if (t >= 64 && t < 70) { sY[t - 64] = yj; d.ywork[6 * j + (t - 64)] = yj; }
    __syncthreads();
    // trailing update S_ab -= L_aj L_bj^T for a >= b in column j; b_a -= L_aj y_j
    const int u0 = d.upd_ptr[j], nu = d.upd_ptr[j + 1] - u0;
    for (int w = t; w < nu * 36; w += nt) {
      const int pidx = w / 36, el = w - pidx * 36, r = el / 6, c = el - r * 6;
      const int ab = d.upd_ab[u0 + pidx];
      const double* La = panel + 36 * (size_t)(ab >> 16) + r * 6;
      const double* Lb = panel + 36 * (size_t)(ab & 0xffff) + c * 6;
      double s = 0;
#pragma unroll
      for (int q = 0; q < 6; ++q) s += La[q] * Lb[q];
      d.S[36 * (size_t)d.upd_dst[u0 + pidx] + el] -= s;
    }
    for (int w = t; w < nb * 6; w += nt) {
      const int a = w / 6, r = w - a * 6;
      const double* La = panel + 36 * (size_t)a + r * 6;
      double s = 0;
#pragma unroll
      for (int q = 0; q < 6; ++q) s += La[q] * sY[q];
      d.ywork[6 * d.row_idx[base + 1 + a] + r] -= s;
    }
    __syncthreads();
  }
  if (sFail) {
    if (t == 0) { ctl->chol_fail = 1; ctl->scale_pose = 0; }
    // trial state = accepted state (it will be rejected)
    for (int i = t; i < 7 * P; i += nt) d.pose[1 - cur][i] = d.pose[cur][i];
    for (int i = t; i < 12 * P; i += nt) d.Rt[1 - cur][i] = d.Rt[cur][i];
    for (int i = t; i < 6 * P; i += nt) d.x[i] = 0;
    return;
  }
  // backward solve L^T x = y, warp 0; x overwrites ywork
  if (warp == 0) {
    for (int j = P - 1; j >= 0; --j) {
      const int base = d.col_ptr[j], nb = d.col_ptr[j + 1] - base - 1;
      // v_r = y_r - sum_a sum_q L_aj[q][r] x_a[q]
      double acc = 0;
      const int r = lane % 6, g = lane / 6;   // 5 groups of 6 lanes, lanes 30,31 idle
      if (lane < 30)
        for (int a = g; a < nb; a += 5) {
          const double* La = d.S + 36 * (size_t)(base + 1 + a);
          const double* xa = d.ywork + 6 * d.row_idx[base + 1 + a];
#pragma unroll
          for (int q = 0; q < 6; ++q) acc += La[q * 6 + r] * xa[q];
        }
      // reduce over groups: lanes r, r+6, ..., r+24
      double tot = acc;
      tot += __shfl_down_sync(0xffffffffu, acc, 6);
      const double a12 = __shfl_down_sync(0xffffffffu, acc, 12);
      const double a18 = __shfl_down_sync(0xffffffffu, acc, 18);
      const double a24 = __shfl_down_sync(0xffffffffu, acc, 24);
      tot += a12 + a18 + a24;
      double v = 0;
      if (lane < 6) v = d.ywork[6 * j + lane] - tot;
      // x_r = sum_{q >= r} Linv[q][r] v_q
      double xr = 0;
#pragma unroll
      for (int q = 0; q < 6; ++q) {
        const double vq = __shfl_sync(0xffffffffu, v, q);
        if (lane < 6 && q >= lane) xr += d.Linv[36 * (size_t)j + q * 6 + lane] * vq;
      }
      if (lane < 6) d.ywork[6 * j + lane] = xr;
      __syncwarp();
    }
  }
  __syncthreads();
  // pose update (G2oVertexSE3::oplusImpl) into the trial buffer + scale = sum x (lambda x + b)
  double sc = 0;
  for (int p = t; p < P; p += nt) {
    const int j = d.pos[p];
    double dx[6], T[7], Tn[7];
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      dx[r] = d.fixed[p] ? 0. : d.ywork[6 * j + r];
      d.x[6 * p + r] = dx[r];
      sc += dx[r] * (lambda * dx[r] + d.bp[6 * p + r]);
    }
#pragma unroll
    for (int r = 0; r < 7; ++r) T[r] = d.pose[cur][7 * (size_t)p + r];
    if (d.fixed[p]) {
#pragma unroll
      for (int r = 0; r < 7; ++r) Tn[r] = T[r];
    } else {
      double dT[7];
      se3_exp(dx, dT);
      se3_mul(dT, T, Tn);
    }
    double R[9];
    quat_to_R(Tn, R);
#pragma unroll
    for (int r = 0; r < 7; ++r) d.pose[1 - cur][7 * (size_t)p + r] = Tn[r];
#pragma unroll
    for (int r = 0; r < 9; ++r) d.Rt[1 - cur][12 * (size_t)p + r] = R[r];
    d.Rt[1 - cur][12 * (size_t)p + 9] = Tn[4];
    d.Rt[1 - cur][12 * (size_t)p + 10] = Tn[5];
    d.Rt[1 - cur][12 * (size_t)p + 11] = Tn[6];
  }
  sc = warp_sum(sc);
  if (lane == 0) sRed[warp] = sc;
  __syncthreads();
  if (t == 0) {
    double s = 0;
    for (int w = 0; w < nt / 32; ++w) s += sRed[w];
    ctl->scale_pose = s;
    ctl->chol_fail = 0;
  }
}

void launch_solve(const BaDev& d, cudaStream_t st) { k_solve<<<1, kSolveThreads, 0, st>>>(d); }

// ------------------------------------------------------------------ k_update

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
k_update(BaDev d, int robust, double delta, int n_lm_blocks, int n_c_blocks) {
  LmCtl* ctl = d.ctl;
  const int cur = ctl->cur, trial = 1 - cur;
  // every CTA helps clearing the reduced system for the next build
  {
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, gn = (size_t)gridDim.x * blockDim.x;
    for (size_t i = gid; i < (size_t)d.nblk * 36; i += gn) d.S[i] = 0.;
    for (size_t i = gid; i < (size_t)6 * d.P; i += gn) { d.bp[i] = 0.; d.bc[i] = 0.; }
  }
  if ((int)blockIdx.x >= n_lm_blocks) {
    const int c = ((int)blockIdx.x - n_lm_blocks) * (WARPS * 32) + (int)threadIdx.x;
    if (c < d.C) d.chi_c_new[c] = constraint_chi2(d, d.pose[trial], c);
    return;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int li = (int)blockIdx.x * WARPS + warp;
  if (li >= d.L) return;
  const double lambda = ctl->lambda;
  const int e0 = d.lm_eptr[li], k = d.lm_eptr[li + 1] - e0;
  const int s0 = d.lm_sptr[li], K = d.lm_sptr[li + 1] - s0;
  const double* psi = d.psi[cur] + 3 * (size_t)li;
  double* psin = d.psi[trial] + 3 * (size_t)li;
  if (k == 0 || ctl->chol_fail) {
    if (lane < 3) psin[lane] = psi[lane];
    if (lane == 0) { d.chi_new_l[li] = 0; d.scale_l[li] = 0; }
    return;
  }
  const int off = d.lm_self[li] ? 0 : 1;
  const int ia = d.lm_anchor[li];
  // c = b_l - sum_slots B_s^T x_s  (slots walked in chunks of 32)
  double c3[3] = {0, 0, 0};
  for (int s = lane; s < K; s += 32) {
    const int p = (s == 0) ? ia : d.e_pose[e0 + s - off];
    const double* xs = d.x + 6 * p;
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      const double xr = xs[r];
#pragma unroll
      for (int q = 0; q < 3; ++q) c3[q] -= d.W[(size_t)(r * 3 + q) * d.nslots + s0 + s] * xr;
    }
  }
#pragma unroll
  for (int q = 0; q < 3; ++q) c3[q] = warp_sum(c3[q]);
  const double* Dbl = d.Dbl + 12 * (size_t)li;
  const double bl[3] = {Dbl[6], Dbl[7], Dbl[8]};
  double Di[9];
  inv3_sym_lambda(Dbl, lambda, Di);
  const double cc[3] = {bl[0] + c3[0], bl[1] + c3[1], bl[2] + c3[2]};
  double dpsi[3], pn[3];
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    dpsi[q] = Di[q * 3] * cc[0] + Di[q * 3 + 1] * cc[1] + Di[q * 3 + 2] * cc[2];
    pn[q] = psi[q] + dpsi[q];
  }
  if (lane < 3) psin[lane] = pn[lane];
  if (lane == 0)
    d.scale_l[li] = dpsi[0] * (lambda * dpsi[0] + bl[0]) + dpsi[1] * (lambda * dpsi[1] + bl[1]) +
                    dpsi[2] * (lambda * dpsi[2] + bl[2]);
  // robust chi2 of this landmark's observations at the trial state
  const double* __restrict__ Rt = d.Rt[trial];
  double Ra[9], ta[3];
  load12(Rt, ia, Ra, ta);
  const double ipz = 1. / pn[2];
  const double xa[3] = {pn[0] * ipz, pn[1] * ipz, ipz};
  double chi = 0;
  for (int i = lane; i < k; i += 32) {
    const int e = e0 + i;
    const double obs[3] = {__ldg(d.e_obs + e), __ldg(d.e_obs + (size_t)d.E + e), __ldg(d.e_obs + 2 * (size_t)d.E + e)};
    const double om[3] = {__ldg(d.e_w + e), __ldg(d.e_w + (size_t)d.E + e), __ldg(d.e_w + 2 * (size_t)d.E + e)};
    chi += edge_cost(d, Rt, d.e_pose[e], Ra, ta, xa, obs, om, robust, delta);
  }
  chi = warp_sum(chi);
  if (lane == 0) d.chi_new_l[li] = chi;
}

void launch_update(const BaDev& d, int robust, double delta, cudaStream_t st) {
  constexpr int WARPS = 8;
  const int n_lm_blocks = (d.L + WARPS - 1) / WARPS;
  const int n_c_blocks = (d.C + WARPS * 32 - 1) / (WARPS * 32);
  int nb = n_lm_blocks + n_c_blocks;
  if (nb == 0) nb = 1;   // still clears the reduced system
  k_update<WARPS><<<nb, WARPS * 32, 0, st>>>(d, robust, delta, n_lm_blocks, n_c_blocks);
}

// ------------------------------------------------------------------ k_chi2 (state cur)

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
k_chi2(BaDev d, int robust, double delta, int n_lm_blocks) {
  const int cur = d.ctl->cur;
  if ((int)blockIdx.x >= n_lm_blocks) {
    const int c = ((int)blockIdx.x - n_lm_blocks) * (WARPS * 32) + (int)threadIdx.x;
    if (c < d.C) d.chi_c[c] = constraint_chi2(d, d.pose[cur], c);
    return;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int li = (int)blockIdx.x * WARPS + warp;
  if (li >= d.L) return;
  const int e0 = d.lm_eptr[li], k = d.lm_eptr[li + 1] - e0;
  double chi = 0;
  if (k > 0) {
    const double* __restrict__ Rt = d.Rt[cur];
    const double* psi = d.psi[cur] + 3 * (size_t)li;
    double Ra[9], ta[3];
    load12(Rt, d.lm_anchor[li], Ra, ta);
    const double ipz = 1. / psi[2];
    const double xa[3] = {psi[0] * ipz, psi[1] * ipz, ipz};
    for (int i = lane; i < k; i += 32) {
      const int e = e0 + i;
      const double obs[3] = {__ldg(d.e_obs + e), __ldg(d.e_obs + (size_t)d.E + e), __ldg(d.e_obs + 2 * (size_t)d.E + e)};
      const double om[3] = {__ldg(d.e_w + e), __ldg(d.e_w + (size_t)d.E + e), __ldg(d.e_w + 2 * (size_t)d.E + e)};
      chi += edge_cost(d, Rt, d.e_pose[e], Ra, ta, xa, obs, om, robust, delta);
    }
    chi = warp_sum(chi);
  }
  if (lane == 0) d.chi_l[li] = chi;
}

void launch_chi2(const BaDev& d, int robust, double delta, cudaStream_t st) {
  constexpr int WARPS = 8;
  const int n_lm_blocks = (d.L + WARPS - 1) / WARPS;
  const int n_c_blocks = (d.C + WARPS * 32 - 1) / (WARPS * 32);
  if (n_lm_blocks + n_c_blocks == 0) return;
  k_chi2<WARPS><<<n_lm_blocks + n_c_blocks, WARPS * 32, 0, st>>>(d, robust, delta, n_lm_blocks);
}

// ------------------------------------------------------------------ k_prep: quaternion poses -> R,t

__global__ void k_prep(BaDev d, int buf) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= d.P) return;
  double T[7], R[9];
#pragma unroll
  for (int r = 0; r < 7; ++r) T[r] = d.pose[buf][7 * (size_t)p + r];
  quat_to_R(T, R);
#pragma unroll
  for (int r = 0; r < 9; ++r) d.Rt[buf][12 * (size_t)p + r] = R[r];
  d.Rt[buf][12 * (size_t)p + 9] = T[4];
  d.Rt[buf][12 * (size_t)p + 10] = T[5];
  d.Rt[buf][12 * (size_t)p + 11] = T[6];
}

void launch_prep(const BaDev& d, int buf, cudaStream_t st) {
  if (d.P == 0) return;
  k_prep<<<(d.P + 127) / 128, 128, 0, st>>>(d, buf);
}

// ------------------------------------------------------------------ k_decide

constexpr int kDecideThreads = 512;

__device__ double block_sum_det(const double* __restrict__ a, int n, double* red) {
  // fixed-order reduction: strided partial sums, then a serial pass over the partials
  double s = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += a[i];
  __syncthreads();
  red[threadIdx.x] = s;
  __syncthreads();
  for (int w = blockDim.x / 2; w > 0; w >>= 1) {
    if ((int)threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
    __syncthreads();
  }
  const double r = red[0];
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(kDecideThreads)
k_decide(BaDev d) {
  __shared__ double red[kDecideThreads];
  LmCtl* ctl = d.ctl;
  const double chi_cur = block_sum_det(d.chi_l, d.L, red) + block_sum_det(d.chi_c, d.C, red);
  const double chi_new = block_sum_det(d.chi_new_l, d.L, red) + block_sum_det(d.chi_c_new, d.C, red);
  const double scale_pts = block_sum_det(d.scale_l, d.L, red);
  if (threadIdx.x != 0) return;
  const int fail = ctl->chol_fail;
  double currentChi = chi_cur;
  const double tempChi = fail ? 1.7976931348623157e308 : chi_new;
  if (ctl->iter == 0 && ctl->qmax == 0) ctl->chi_init = currentChi;
  double lambda = ctl->lambda, ni = ctl->ni;
  double rho = currentChi - tempChi;
  double scale = fail ? 0. : ctl->scale_pose + scale_pts;   // computeScale(): sum x (lambda x + b)
  scale += 1e-3;
  rho /= scale;
  int cur = ctl->cur;
  if (rho > 0 && isfinite(tempChi)) {
    const double u = 2 * rho - 1;
    double alpha = 1. - u * u * u;
    alpha = fmin(alpha, 2. / 3.);
    const double sf = fmax(1. / 3., alpha);
    lambda *= sf;
    ni = 2;
    currentChi = tempChi;
    cur ^= 1;
  } else {
    lambda *= ni;
    ni *= 2;
  }
  int qmax = ctl->qmax + 1;
  ctl->trials_total += 1;
  const int again = (rho < 0 && qmax < ctl->max_trials) ? 1 : 0;
  ctl->lambda = lambda; ctl->ni = ni; ctl->cur = cur; ctl->rho = rho;
  ctl->chi_cur = currentChi; ctl->chi_new = tempChi;
  ctl->again = again;
  if (!again) {
    const int it = ctl->iter;
    if (it < kMaxIters) { ctl->chi_iter[it] = currentChi; ctl->lambda_iter[it] = lambda; ctl->trials_iter[it] = qmax; }
    ctl->stop = (qmax == ctl->max_trials || rho == 0) ? 1 : 0;
    ctl->iter = it + 1;
    qmax = 0;
  }
  ctl->qmax = qmax;
}

void launch_decide(const BaDev& d, cudaStream_t st) { k_decide<<<1, kDecideThreads, 0, st>>>(d); }

}  // namespace svs
