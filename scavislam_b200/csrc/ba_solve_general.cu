// ba_solve_general.cu -- k_solve_general: reference implementation of the reduced-system solve that
// works out of global memory for any column width (used when a factor column does not fit
// the shared-memory ring of k_solve, see ba_solve.cu).
#include "ba_dev.cuh"
#include "ba_kernels.cuh"

namespace svs {

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
k_solve_general(BaDev d) {
  __shared__ double sDg[36], sLi[36], sY[6];
  __shared__ double sPanel[kPanelCap * 36];
  __shared__ double sRed[kSolveThreads / 32];
  __shared__ int sFail;
  LmCtl* ctl = d.ctl;
  if (ctl->max_iters > 0 && (ctl->stop || ctl->iter >= ctl->max_iters)) return;   // speculatively enqueued trial: nothing left to do
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

void launch_solve_general(const BaDev& d, cudaStream_t st) { k_solve_general<<<1, kSolveThreads, 0, st>>>(d); }

}  // namespace svs
