// ba_solve.cu -- k_solve: block-sparse right-looking Cholesky of the reduced camera system
// S dx = bs, forward/backward solve and pose update, one CTA.
//   replaces g2o::LinearSolverCSparse<Matrix6d>::solve (instantiated at slam_graph.cpp:55-60) and
//   G2oVertexSE3::oplusImpl (anchored_points.cpp:53-58).
//
// The factor is a latency chain of P block columns.  What bounds it is the dependent chain per
// column (6 pivots: rsqrt -> mul -> fma), two CTA barriers and shared-memory round trips -- not
// HBM and not the tensor cores (6x6 blocks, FP64).  So the design removes everything else from
// that chain:
//   * blocks live in a shared-memory ring that covers the next `cap` blocks in column-major
//     order (the whole band of a SLAM window); trailing updates are shared-memory RMWs,
//     blocks outside the ring (far fill) fall back to global RMWs;
//   * ring refills are LDGSTS (cp.async) issued after a column's update and only waited for
//     before the next column's update, so they fly under the pivot chain;
//   * each thread that owns a panel row factors the 6x6 diagonal block redundantly in
//     registers (no warp-cooperative pivoting, no divisions: rsqrt + multiplies);
//   * the right-hand side rides along as one more row of the panel (forward solve for free).
#include "ba_dev.cuh"
#include "ba_kernels.cuh"

namespace svs {

constexpr int kSolveThreads = 256;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

// Lower Cholesky of a symmetric 6x6 (lower triangle read from `A`, row-major, lambda added to
// the diagonal) and the inverse of the factor, all in registers.  l[], li[] are packed lower
// triangles (index r*(r+1)/2 + c).  Returns false when a pivot is not positive.
__device__ __forceinline__ bool chol6_regs(const double* __restrict__ A, double dlam, double l[21], double li[21]) {
  double a[21];
#pragma unroll
  for (int r = 0; r < 6; ++r)
#pragma unroll
    for (int c = 0; c <= r; ++c) a[r * (r + 1) / 2 + c] = A[r * 6 + c] + (r == c ? dlam : 0.);
  bool ok = true;
  double rinv[6];
#pragma unroll
  for (int c = 0; c < 6; ++c) {
    const double dv = a[c * (c + 1) / 2 + c];
    ok = ok && (dv > 0.);
    rinv[c] = rsqrt(dv);
    l[c * (c + 1) / 2 + c] = dv * rinv[c];
#pragma unroll
    for (int r = c + 1; r < 6; ++r) l[r * (r + 1) / 2 + c] = a[r * (r + 1) / 2 + c] * rinv[c];
#pragma unroll
    for (int r = c + 1; r < 6; ++r)
#pragma unroll
      for (int c2 = c + 1; c2 <= r; ++c2)
        a[r * (r + 1) / 2 + c2] -= l[r * (r + 1) / 2 + c] * l[c2 * (c2 + 1) / 2 + c];
  }
  // inverse of the lower-triangular factor, column by column
#pragma unroll
  for (int c = 0; c < 6; ++c) {
    li[c * (c + 1) / 2 + c] = rinv[c];
#pragma unroll
    for (int r = c + 1; r < 6; ++r) {
      double v = 0.;
#pragma unroll
      for (int q = c; q < r; ++q) v -= l[r * (r + 1) / 2 + q] * li[q * (q + 1) / 2 + c];
      li[r * (r + 1) / 2 + c] = v * rinv[r];
    }
  }
  return ok;
}

constexpr int kUpdPf = 4;   // update-list entries per thread prefetched one column ahead

// smem layout: [ring: cap*36 doubles][y: 6P doubles if y_in_smem][meta ints if meta_in_smem:
//               col_ptr (P+1), upd_ptr (P+1), row_idx (nblk), fixed-by-position (P)]
__global__ void __launch_bounds__(kSolveThreads)
k_solve(BaDev d, int cap, int y_in_smem, int meta_in_smem) {
  extern __shared__ __align__(16) double sm_solve[];
  __shared__ int sFail;
  __shared__ double sRed[kSolveThreads / 32];
  double* ring = sm_solve;
  double* ysm = sm_solve + (size_t)cap * 36;
  double* yv = y_in_smem ? ysm : d.ywork;
  LmCtl* ctl = d.ctl;
  const int t = threadIdx.x, nt = kSolveThreads, lane = t & 31, warp = t >> 5;
  const int P = d.P, nblk = d.nblk;
  const double lambda = ctl->lambda;
  const int cur = ctl->cur;
  if (t == 0) sFail = 0;
  // index metadata: every use below sits on the per-column critical path, so keep it on chip
  int* meta = reinterpret_cast<int*>(ysm + (y_in_smem ? ((6 * (size_t)P + 1) / 2) * 2 : 0));
  const int* col_ptr = d.col_ptr;
  const int* upd_ptr = d.upd_ptr;
  const int* row_idx = d.row_idx;
  int* sfix = nullptr;
  if (meta_in_smem) {
    int* c = meta; int* u = meta + (P + 1); int* r = u + (P + 1); sfix = r + nblk;
    for (int i = t; i <= P; i += nt) { c[i] = d.col_ptr[i]; u[i] = d.upd_ptr[i]; }
    for (int i = t; i < nblk; i += nt) r[i] = d.row_idx[i];
    for (int i = t; i < P; i += nt) sfix[i] = d.fixed[d.perm[i]];
    col_ptr = c; upd_ptr = u; row_idx = r;
  }

  // initial ring fill: blocks [0, hi)
  int hi = min(nblk, cap);
  for (int c = t; c < hi * 18; c += nt) {
    const int id = c / 18, w = c - id * 18;
    cp_async16(ring + (size_t)(id % cap) * 36 + 2 * w, d.S + (size_t)id * 36 + 2 * w);
  }
  cp_async_commit();
  // right-hand side in elimination order: bs = bp - bc
  for (int i = t; i < 6 * P; i += nt) {
    const int j = i / 6, r = i - 6 * j;
    const int p = d.perm[j];
    yv[i] = d.bp[6 * p + r] - d.bc[6 * p + r];
  }
  cp_async_wait_all();
  __syncthreads();

  int pf_ab[kUpdPf], pf_dst[kUpdPf];   // update-list entries of the column about to be processed
  auto prefetch_upd = [&](int j) {
    const int u0 = upd_ptr[j], nu = upd_ptr[j + 1] - u0;
#pragma unroll
    for (int i = 0; i < kUpdPf; ++i) {
      const int w = t + i * nt;
      if (w < nu * 36) { pf_ab[i] = d.upd_ab[u0 + w / 36]; pf_dst[i] = d.upd_dst[u0 + w / 36]; }
    }
  };
  if (P > 0) prefetch_upd(0);
  for (int j = 0; j < P; ++j) {
    const int base = col_ptr[j], nb = col_ptr[j + 1] - base - 1;
    const double* diag = ring + (size_t)(base % cap) * 36;
    int cu_ab[kUpdPf], cu_dst[kUpdPf];
#pragma unroll
    for (int i = 0; i < kUpdPf; ++i) { cu_ab[i] = pf_ab[i]; cu_dst[i] = pf_dst[i]; }
    if (j + 1 < P) prefetch_upd(j + 1);   // in flight under this column's pivot chain
    const int nrows = nb * 6 + 1;   // panel rows + the right-hand side row
    // --- phase A+B: every thread that owns a row factors the diagonal block in registers, then
    //     row <- row * L_jj^-T   (block rows: L_ij ; rhs row: y_j = L_jj^-1 b_j)
    if (t < nrows || t == 0) {
      double l[21], li[21];
      const double dlam = lambda + ((sfix ? sfix[j] : (int)d.fixed[d.perm[j]]) ? 1. : 0.);
      const bool ok = chol6_regs(diag, dlam, l, li);
      if (!ok) sFail = 1;
      if (t == 0) {
        double* Lo = d.Linv + 36 * (size_t)j;
#pragma unroll
        for (int r = 0; r < 6; ++r)
#pragma unroll
          for (int c = 0; c < 6; ++c) Lo[r * 6 + c] = (c <= r) ? li[r * (r + 1) / 2 + c] : 0.;
      }
      for (int row = t; row < nrows; row += nt) {
        double* src;
        double* gdst = nullptr;
        if (row < nb * 6) {
          const int a = row / 6, r = row - a * 6;
          src = ring + (size_t)((base + 1 + a) % cap) * 36 + r * 6;
          gdst = d.S + (size_t)(base + 1 + a) * 36 + r * 6;
        } else {
          src = yv + 6 * j;
        }
        double v[6], o[6];
#pragma unroll
        for (int q = 0; q < 6; ++q) v[q] = src[q];
#pragma unroll
        for (int c = 0; c < 6; ++c) {
          double s = 0.;
#pragma unroll
          for (int q = 0; q <= c; ++q) s += v[q] * li[c * (c + 1) / 2 + q];
          o[c] = s;
        }
#pragma unroll
        for (int q = 0; q < 6; ++q) src[q] = o[q];
        if (gdst) {
#pragma unroll
          for (int q = 0; q < 6; ++q) gdst[q] = o[q];   // final factor column, read by the backward solve
        }
      }
    }
    cp_async_wait_all();   // refills issued after the previous column's update
    __syncthreads();
    if (sFail) break;
    // --- phase C: S_ab -= L_aj L_bj^T for a >= b in column j ; b_a -= L_aj y_j
    const int u0 = upd_ptr[j], nu = upd_ptr[j + 1] - u0;
    auto update_item = [&](int w, int ab, int dst) {
      const int pidx = w / 36, el = w - pidx * 36, r = el / 6, c = el - r * 6;
      const double* La = ring + (size_t)((base + 1 + (ab >> 16)) % cap) * 36 + r * 6;
      const double* Lb = ring + (size_t)((base + 1 + (ab & 0xffff)) % cap) * 36 + c * 6;
      double s = 0.;
#pragma unroll
      for (int q = 0; q < 6; ++q) s += La[q] * Lb[q];
      if (dst < hi) ring[(size_t)(dst % cap) * 36 + el] -= s;
      else d.S[(size_t)dst * 36 + el] -= s;
    };
#pragma unroll
    for (int i = 0; i < kUpdPf; ++i) {   // entries prefetched during the previous column
      const int w = t + i * nt;
      if (w < nu * 36) update_item(w, cu_ab[i], cu_dst[i]);
    }
    for (int w = t + kUpdPf * nt; w < nu * 36; w += nt)   // wide columns: the rest from L2
      update_item(w, d.upd_ab[u0 + w / 36], d.upd_dst[u0 + w / 36]);
    for (int w = t; w < nb * 6; w += nt) {
      const int a = w / 6, r = w - a * 6;
      const double* La = ring + (size_t)((base + 1 + a) % cap) * 36 + r * 6;
      const double* yj = yv + 6 * j;
      double s = 0.;
#pragma unroll
      for (int q = 0; q < 6; ++q) s += La[q] * yj[q];
      yv[6 * row_idx[base + 1 + a] + r] -= s;
    }
    __syncthreads();
    // --- refill the ring slots column j frees: blocks [hi, min(nblk, col_ptr[j+1] + cap))
    const int hi_new = min(nblk, col_ptr[j + 1] + cap);
    for (int c = t; c < (hi_new - hi) * 18; c += nt) {
      const int id = hi + c / 18, w = c % 18;
      cp_async16(ring + (size_t)(id % cap) * 36 + 2 * w, d.S + (size_t)id * 36 + 2 * w);
    }
    cp_async_commit();
    hi = hi_new;
  }
  cp_async_wait_all();
  __syncthreads();
  if (sFail) {
    if (t == 0) { ctl->chol_fail = 1; ctl->scale_pose = 0; }
    for (int i = t; i < 7 * P; i += nt) d.pose[1 - cur][i] = d.pose[cur][i];
    for (int i = t; i < 12 * P; i += nt) d.Rt[1 - cur][i] = d.Rt[cur][i];
    for (int i = t; i < 6 * P; i += nt) d.x[i] = 0;
    return;
  }
  // --- backward solve L^T x = y by warp 0; lanes (g, r): g = lane / 6 walks the column's blocks
  //     g, g+5, ...; the next column's blocks are prefetched while this one is reduced.
  if (warp == 0) {
    const int r = lane % 6, g = lane / 6;
    double pf[2][6];   // up to two prefetched blocks per lane group (column r of L_aj)
    int prow[2];
    double linv_c[6];  // column r of Linv_j (entries q >= r)
    auto prefetch = [&](int j) {
      const int base = col_ptr[j], nb = col_ptr[j + 1] - base - 1;
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int a = g + 5 * u;
        prow[u] = -1;
        if (lane < 30 && a < nb) {
          const double* La = d.S + 36 * (size_t)(base + 1 + a);
          prow[u] = row_idx[base + 1 + a];
#pragma unroll
          for (int q = 0; q < 6; ++q) pf[u][q] = La[q * 6 + r];
        }
      }
      if (lane < 6) {
#pragma unroll
        for (int q = 0; q < 6; ++q) linv_c[q] = d.Linv[36 * (size_t)j + q * 6 + lane];
      }
    };
    if (P > 0) prefetch(P - 1);
    for (int j = P - 1; j >= 0; --j) {
      const int base = col_ptr[j], nb = col_ptr[j + 1] - base - 1;
      double acc = 0.;
#pragma unroll
      for (int u = 0; u < 2; ++u)
        if (prow[u] >= 0) {
          const double* xa = yv + 6 * prow[u];
#pragma unroll
          for (int q = 0; q < 6; ++q) acc += pf[u][q] * xa[q];
        }
      if (lane < 30)
        for (int a = g + 10; a < nb; a += 5) {   // wide columns: the rest straight from L2
          const double* La = d.S + 36 * (size_t)(base + 1 + a);
          const double* xa = yv + 6 * row_idx[base + 1 + a];
#pragma unroll
          for (int q = 0; q < 6; ++q) acc += La[q * 6 + r] * xa[q];
        }
      double lc[6];
#pragma unroll
      for (int q = 0; q < 6; ++q) lc[q] = linv_c[q];
      if (j > 0) prefetch(j - 1);
      double tot = acc;
      tot += __shfl_down_sync(0xffffffffu, acc, 6);
      const double a12 = __shfl_down_sync(0xffffffffu, acc, 12);
      const double a18 = __shfl_down_sync(0xffffffffu, acc, 18);
      const double a24 = __shfl_down_sync(0xffffffffu, acc, 24);
      tot += a12 + a18 + a24;
      double v = 0.;
      if (lane < 6) v = yv[6 * j + lane] - tot;
      double xr = 0.;
#pragma unroll
      for (int q = 0; q < 6; ++q) {
        const double vq = __shfl_sync(0xffffffffu, v, q);
        if (q >= lane) xr += lc[q] * vq;
      }
      if (lane < 6) yv[6 * j + lane] = xr;
      __syncwarp();
    }
  }
  __syncthreads();
  // --- pose update (G2oVertexSE3::oplusImpl) into the trial buffer; scale = sum x (lambda x + b)
  double sc = 0;
  for (int p = t; p < P; p += nt) {
    const int j = d.pos[p];
    double dx[6], T[7], Tn[7];
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      dx[r] = d.fixed[p] ? 0. : yv[6 * j + r];
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

// Picks the ring capacity for a device and launches; falls back to the global-memory kernel
// when one factor column alone would not fit the ring.
void launch_solve(const BaDev& d, int max_col_blocks, cudaStream_t st) {
  static int smem_optin = -1;
  if (smem_optin < 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cudaFuncSetAttribute(k_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin - 1024);
  }
  const size_t budget = (size_t)smem_optin - 1024 - 256;
  const int y_in_smem = (size_t)6 * d.P * 8 <= budget / 4;
  const size_t ybytes = y_in_smem ? (((size_t)6 * d.P * 8 + 15) / 16) * 16 : 0;
  const size_t mbytes_want = ((size_t)(2 * (d.P + 1) + d.nblk + d.P) * 4 + 15) / 16 * 16;
  const int meta_in_smem = mbytes_want <= (budget - ybytes) / 3;
  const size_t mbytes = meta_in_smem ? mbytes_want : 0;
  int cap = (int)((budget - ybytes - mbytes) / 288);
  if (cap > d.nblk) cap = d.nblk > 0 ? d.nblk : 1;
  if (max_col_blocks + 1 > cap) {
    launch_solve_general(d, st);
    return;
  }
  const size_t smem = (size_t)cap * 288 + ybytes + mbytes;
  k_solve<<<1, kSolveThreads, smem, st>>>(d, cap, y_in_smem, meta_in_smem);
}

}  // namespace svs
