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

constexpr int kSolveThreads = 512;

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

constexpr int kUpdPf = 2;        // update-list entries per update thread prefetched one column ahead
constexpr int kPanelThreads = 128;  // warps 0-3: factor the next column (look-ahead); warp 0 owns the pivot chain
constexpr int kUpdThreads = kSolveThreads - kPanelThreads;

__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// 1/sqrt(a): hardware approximation + two Newton steps (the library rsqrt() carries special-case
// handling that would sit on the pivot chain)
__device__ __forceinline__ double fast_rsqrt(double a) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
  const double h = 0.5 * a;
  y = fma(y, fma(-h * y, y, 0.5), y);
  y = fma(y, fma(-h * y, y, 0.5), y);
  return y;
}

// Lower Cholesky of the 6x6 block at A (row-major, lower triangle read, dlam added to the
// diagonal) in registers: l = packed lower factor (r*(r+1)/2 + c), rinv = 1 / diagonal.
__device__ __forceinline__ bool chol6_lean(const double* __restrict__ A, double dlam, double l[21], double rinv[6]) {
  double a[21];
#pragma unroll
  for (int r = 0; r < 6; ++r)
#pragma unroll
    for (int c = 0; c <= r; ++c) a[r * (r + 1) / 2 + c] = A[r * 6 + c] + (r == c ? dlam : 0.);
  bool ok = true;
#pragma unroll
  for (int c = 0; c < 6; ++c) {
    const double dv = a[c * (c + 1) / 2 + c];
    ok = ok && (dv > 0.);
    rinv[c] = fast_rsqrt(dv);
    l[c * (c + 1) / 2 + c] = dv * rinv[c];
#pragma unroll
    for (int r = c + 1; r < 6; ++r) l[r * (r + 1) / 2 + c] = a[r * (r + 1) / 2 + c] * rinv[c];
#pragma unroll
    for (int r = c + 1; r < 6; ++r)
#pragma unroll
      for (int c2 = c + 1; c2 <= r; ++c2)
        a[r * (r + 1) / 2 + c2] -= l[r * (r + 1) / 2 + c] * l[c2 * (c2 + 1) / 2 + c];
  }
  return ok;
}

// smem layout: [ring: cap*36 doubles][y: 6P doubles if y_in_smem][meta ints: col_ptr (P+1),
//               upd_ptr (P+1), row_idx (nblk), urg_dst (nblk), fixed-by-position (P)]
// cap is a power of two >= 4 * (widest column + 1).
__global__ void __launch_bounds__(kSolveThreads)
k_solve(BaDev d, int cap, int y_in_smem, int refill_period) {
  extern __shared__ __align__(16) double sm_solve[];
  __shared__ int sFailBuf[2];   // indexed by column parity: written by the panel warps, read after the column barrier
  __shared__ double sRed[kSolveThreads / 32];
  __shared__ double sL[2][28];   // factor of the diagonal block of column j (l 21, rinv 6), double buffered
  double* ring = sm_solve;
  const int mask = cap - 1;
  double* ysm = sm_solve + (size_t)cap * 36;
  double* yv = y_in_smem ? ysm : d.ywork;
  LmCtl* ctl = d.ctl;
  if (ctl->max_iters > 0 && (ctl->stop || ctl->iter >= ctl->max_iters)) return;   // speculatively enqueued trial: nothing left to do
  const int t = threadIdx.x, nt = kSolveThreads, lane = t & 31, warp = t >> 5;
  const int P = d.P, nblk = d.nblk;
  const double lambda = ctl->lambda;
  const int cur = ctl->cur;
  if (t == 0) { sFailBuf[0] = 0; sFailBuf[1] = 0; }
  int* meta = reinterpret_cast<int*>(ysm + (y_in_smem ? ((6 * (size_t)P + 1) / 2) * 2 : 0));
  int* col_ptr = meta;
  int* upd_ptr = col_ptr + (P + 1);
  int* row_idx = upd_ptr + (P + 1);
  int* urg_dst = row_idx + nblk;
  int* sfix = urg_dst + nblk;
  for (int i = t; i <= P; i += nt) { col_ptr[i] = d.col_ptr[i]; upd_ptr[i] = d.upd_ptr[i]; }
  for (int i = t; i < nblk; i += nt) { row_idx[i] = d.row_idx[i]; urg_dst[i] = d.urg_dst[i]; }
  for (int i = t; i < P; i += nt) sfix[i] = d.fixed[d.perm[i]];
  // initial ring fill: blocks [0, hi)
  int hi = min(nblk, cap);
  for (int c = t; c < hi * 18; c += nt) {
    const int id = c / 18, w = c - id * 18;
    cp_async16(ring + (size_t)(id & mask) * 36 + 2 * w, d.S + (size_t)id * 36 + 2 * w);
  }
  cp_async_commit();
  for (int i = t; i < 6 * P; i += nt) {   // right-hand side in elimination order: bs = bp - bc
    const int j = i / 6, r = i - 6 * j;
    const int p = d.perm[j];
    yv[i] = d.bp[6 * p + r] - d.bc[6 * p + r];
  }
  cp_async_wait_all();
  __syncthreads();

  // factor + scale the panel of column jn (panel warps only): chol by warp 0, rows by both
  auto panel_column = [&](int jn) {
    const int base = col_ptr[jn], nb = col_ptr[jn + 1] - base - 1;
    double* sl = sL[jn & 1];
    if (warp == 0) {
      double l[21], rinv[6];
      const bool ok = chol6_lean(ring + (size_t)(base & mask) * 36, lambda + (sfix[jn] ? 1. : 0.), l, rinv);
      if (lane == 0) {
        if (!ok) sFailBuf[jn & 1] = 1;
#pragma unroll
        for (int i = 0; i < 21; ++i) sl[i] = l[i];
#pragma unroll
        for (int i = 0; i < 6; ++i) sl[21 + i] = rinv[i];
      }
    }
    bar_sync(1, kPanelThreads);
    // row <- row * L^-T by forward substitution (block rows: L_ij ; rhs row: y = L^-1 b)
    const int nrows = nb * 6 + 1;
    for (int row = t; row < nrows; row += kPanelThreads) {
      double* src;
      double* gdst = nullptr;
      if (row < nb * 6) {
        const int a = row / 6, r = row - a * 6;
        src = ring + (size_t)((base + 1 + a) & mask) * 36 + r * 6;
        gdst = d.S + (size_t)(base + 1 + a) * 36 + r * 6;
      } else {
        src = yv + 6 * jn;
      }
      double o[6];
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        double s = src[c];
#pragma unroll
        for (int q = 0; q < c; ++q) s -= o[q] * sl[c * (c + 1) / 2 + q];
        o[c] = s * sl[21 + c];
      }
#pragma unroll
      for (int q = 0; q < 6; ++q) src[q] = o[q];
      if (gdst) {
#pragma unroll
        for (int q = 0; q < 6; ++q) gdst[q] = o[q];   // final factor column, read by the backward solve
      }
    }
  };

  // one element of a pair update (used on the pivot chain, where width beats register tiling)
  auto update_item = [&](int base, int w, int ab, int dst, int hi_res) {
    const int pidx = w / 36, el = w - pidx * 36, r = el / 6, c = el - r * 6;
    (void)pidx;
    const double2* La = reinterpret_cast<const double2*>(ring + (size_t)((base + 1 + (ab >> 16)) & mask) * 36 + r * 6);
    const double2* Lb = reinterpret_cast<const double2*>(ring + (size_t)((base + 1 + (ab & 0xffff)) & mask) * 36 + c * 6);
    const double2 a0 = La[0], a1 = La[1], a2 = La[2], b0 = Lb[0], b1 = Lb[1], b2 = Lb[2];
    const double s = (a0.x * b0.x + a0.y * b0.y) + (a1.x * b1.x + a1.y * b1.y) + (a2.x * b2.x + a2.y * b2.y);
    if (dst < hi_res) ring[(size_t)(dst & mask) * 36 + el] -= s;
    else d.S[(size_t)dst * 36 + el] -= s;
  };
  // half of a pair update, register tiled: rows 3h..3h+2 of S_ab -= L_a L_b^T.  18 + 36 doubles loaded
  // for 108 FMAs -- off the pivot chain, what matters is not to flood the shared-memory pipe the
  // chain also lives on (the element-wise form loads 12 doubles per 6 FMAs).
  auto update_half = [&](int base, int u, int ab, int dst, int hi_res) {
    const int h = u & 1;
    const double2* La = reinterpret_cast<const double2*>(ring + (size_t)((base + 1 + (ab >> 16)) & mask) * 36 + h * 18);
    const double2* Lb = reinterpret_cast<const double2*>(ring + (size_t)((base + 1 + (ab & 0xffff)) & mask) * 36);
    double a[18];
#pragma unroll
    for (int q = 0; q < 9; ++q) { const double2 v = La[q]; a[2 * q] = v.x; a[2 * q + 1] = v.y; }
    double o[18];
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      const double2 b0 = Lb[3 * c], b1 = Lb[3 * c + 1], b2 = Lb[3 * c + 2];
#pragma unroll
      for (int r = 0; r < 3; ++r)
        o[r * 6 + c] = (a[r * 6] * b0.x + a[r * 6 + 1] * b0.y) + (a[r * 6 + 2] * b1.x + a[r * 6 + 3] * b1.y) +
                       (a[r * 6 + 4] * b2.x + a[r * 6 + 5] * b2.y);
    }
    double2* D = reinterpret_cast<double2*>((dst < hi_res ? ring + (size_t)(dst & mask) * 36 : d.S + (size_t)dst * 36) + h * 18);
#pragma unroll
    for (int q = 0; q < 9; ++q) { double2 v = D[q]; v.x -= o[2 * q]; v.y -= o[2 * q + 1]; D[q] = v; }
  };

  // number of "urgent" pairs of column j: those that land in column j+1 (pairs (a, 0) when the
  // first sub-diagonal row of column j is j+1; the update list is ordered b-major)
  auto urgent_of = [&](int j) {
    const int base = col_ptr[j], nb = col_ptr[j + 1] - base - 1;
    return (nb > 0 && row_idx[base + 1] == j + 1) ? nb : 0;
  };

  long long tm[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) tm[i] = 0;
  long long t_prev = clock64();
#define TICK(slot) { const long long now_ = clock64(); tm[slot] += now_ - t_prev; t_prev = now_; }
  if (P > 0 && t < kPanelThreads) panel_column(0);
  const int ut = t - kPanelThreads;   // index among the update threads
  int pf_ab[kUpdPf], pf_dst[kUpdPf];
  auto prefetch_upd = [&](int j) {
    const int u0 = upd_ptr[j] + urgent_of(j), nunits = (upd_ptr[j + 1] - u0) * 2;
#pragma unroll
    for (int i = 0; i < kUpdPf; ++i) {
      const int u = ut + i * kUpdThreads;
      if (u < nunits) { pf_ab[i] = d.upd_ab[u0 + (u >> 1)]; pf_dst[i] = d.upd_dst[u0 + (u >> 1)]; }
    }
  };
  if (P > 0 && t >= kPanelThreads) prefetch_upd(0);
  __syncthreads();
  int failed = P > 0 ? sFailBuf[0] : 0;
  int until_refill = refill_period;

  TICK(0);
  for (int j = 0; j < P && !failed; ++j) {
    const int base = col_ptr[j], nb = col_ptr[j + 1] - base - 1;
    const int urgent = urgent_of(j);
    const int u0 = upd_ptr[j];
    TICK(1);
    if (t < kPanelThreads) {
      // ---- panel warps: the part of column j's update that lands in column j+1, then factor it
      if (urgent) {
        if (warp == 0) {
          // S_{j+1,j+1} -= L L^T with L = L_{j+1,j}: symmetric, 21 lanes, one round; this is the only
          // part of column j's update the next pivot chain waits for
          if (lane < 21) {
            int r = 0, u = lane;
            while (u > r) { u -= r + 1; ++r; }
            const int c = u;
            const double2* La = reinterpret_cast<const double2*>(ring + (size_t)((base + 1) & mask) * 36 + r * 6);
            const double2* Lb = reinterpret_cast<const double2*>(ring + (size_t)((base + 1) & mask) * 36 + c * 6);
            const double2 a0 = La[0], a1 = La[1], a2 = La[2], b0 = Lb[0], b1 = Lb[1], b2 = Lb[2];
            const double sv = (a0.x * b0.x + a0.y * b0.y) + (a1.x * b1.x + a1.y * b1.y) + (a2.x * b2.x + a2.y * b2.y);
            double* D = ring + (size_t)(col_ptr[j + 1] & mask) * 36;
            const double nv = D[r * 6 + c] - sv;
            D[r * 6 + c] = nv;
            D[c * 6 + r] = nv;
          }
          __syncwarp();
        } else {
          if (warp == 1 && lane < 6) {   // b_{j+1} -= L_{j+1,j} y_j
            const double* La = ring + (size_t)((base + 1) & mask) * 36 + lane * 6;
            const double* yj = yv + 6 * j;
            double sv = 0.;
#pragma unroll
            for (int q = 0; q < 6; ++q) sv += La[q] * yj[q];
            yv[6 * (j + 1) + lane] -= sv;
          }
          for (int u = 2 + (t - 32); u < urgent * 2; u += kPanelThreads - 32) {
            const int a = u >> 1;
            update_half(base, u, a << 16, urg_dst[base + 1 + a], 0x7fffffff);
          }
        }
      }
      TICK(2);
      if (j + 1 < P) panel_column(j + 1);   // its barrier also orders warp 1's panel updates before the row scaling
      TICK(3);
    } else {
      // ---- update warps: the rest of column j's trailing update
      int cu_ab[kUpdPf], cu_dst[kUpdPf];
#pragma unroll
      for (int i = 0; i < kUpdPf; ++i) { cu_ab[i] = pf_ab[i]; cu_dst[i] = pf_dst[i]; }
      if (j + 1 < P) prefetch_upd(j + 1);
      const int uu = u0 + urgent, nunits = (upd_ptr[j + 1] - uu) * 2;
#pragma unroll
      for (int i = 0; i < kUpdPf; ++i) {
        const int u = ut + i * kUpdThreads;
        if (u < nunits) update_half(base, u, cu_ab[i], cu_dst[i], hi);
      }
      for (int u = ut + kUpdPf * kUpdThreads; u < nunits; u += kUpdThreads)
        update_half(base, u, d.upd_ab[uu + (u >> 1)], d.upd_dst[uu + (u >> 1)], hi);
      // b_a -= L_aj y_j for the rows the panel warps did not take
      for (int w = ut + (urgent ? 6 : 0); w < nb * 6; w += kUpdThreads) {
        const int a = w / 6, r = w - a * 6;
        const double* La = ring + (size_t)((base + 1 + a) & mask) * 36 + r * 6;
        const double* yj = yv + 6 * j;
        double s = 0.;
#pragma unroll
        for (int q = 0; q < 6; ++q) s += La[q] * yj[q];
        yv[6 * row_idx[base + 1 + a] + r] -= s;
      }
      // inverse of column j's diagonal factor for the backward solve (off the critical path)
      if (warp == kSolveThreads / 32 - 1 && lane < 6) {
        const double* sl = sL[j & 1];
        const int c = lane;
        double col[6];
#pragma unroll
        for (int r = 0; r < 6; ++r) {
          double v = (r == c) ? 1. : 0.;
#pragma unroll
          for (int q = 0; q < r; ++q) v -= (q >= c) ? sl[r * (r + 1) / 2 + q] * col[q] : 0.;
          col[r] = (r < c) ? 0. : v * sl[21 + r];
        }
#pragma unroll
        for (int r = 0; r < 6; ++r) d.Linv[36 * (size_t)j + r * 6 + c] = col[r];
      }
      TICK(4);
    }
    __syncthreads();
    TICK(5);
    failed = sFailBuf[(j + 1) & 1];
    // ---- every refill_period columns: reload the ring slots the finished columns freed.  Copies are
    //      never in flight while updates run, so a destination is either resident (< hi) or in HBM.
    if (--until_refill == 0) until_refill = refill_period;
    if (until_refill == refill_period && hi < nblk) {
      const int hi_new = min(nblk, col_ptr[j + 1] + cap);
      for (int c = t; c < (hi_new - hi) * 18; c += nt) {
        const int id = hi + c / 18, w = c % 18;
        cp_async16(ring + (size_t)(id & mask) * 36 + 2 * w, d.S + (size_t)id * 36 + 2 * w);
      }
      cp_async_commit();
      cp_async_wait_all();
      __syncthreads();
      hi = hi_new;
      TICK(6);
    }
  }
  cp_async_wait_all();
  __syncthreads();
  if (failed) {
    if (t == 0) { ctl->chol_fail = 1; ctl->scale_pose = 0; }
    for (int i = t; i < 7 * P; i += nt) d.pose[1 - cur][i] = d.pose[cur][i];
    for (int i = t; i < 12 * P; i += nt) d.Rt[1 - cur][i] = d.Rt[cur][i];
    for (int i = t; i < 6 * P; i += nt) d.x[i] = 0;
    return;
  }
  // --- backward solve L^T x = y.  The ring is free now: the factor is streamed back through it in
  //     chunks of columns (L blocks + the inverse diagonal factors), double buffered; warp 0 walks
  //     the dependency chain out of shared memory while the other warps fetch the next chunk.
  {
    const int half = cap / 2;
    double* bufs[2] = {ring, ring + (size_t)half * 36};
    auto chunk_lo = [&](int jhi) {   // largest [jlo, jhi) whose blocks + diagonal inverses fit one half
      int jlo = jhi;
      while (jlo > 0 && (col_ptr[jhi] - col_ptr[jlo - 1]) + (jhi - (jlo - 1)) <= half) --jlo;
      return jlo;
    };
    auto load_chunk = [&](double* buf, int jlo, int jhi, int tid, int nth) {
      const int nb16 = (col_ptr[jhi] - col_ptr[jlo]) * 18;
      const double* src = d.S + (size_t)col_ptr[jlo] * 36;
      for (int c = tid; c < nb16; c += nth) cp_async16(buf + 2 * (size_t)c, src + 2 * (size_t)c);
      double* lbuf = buf + (size_t)(col_ptr[jhi] - col_ptr[jlo]) * 36;
      const double* lsrc = d.Linv + (size_t)jlo * 36;
      for (int c = tid; c < (jhi - jlo) * 18; c += nth) cp_async16(lbuf + 2 * (size_t)c, lsrc + 2 * (size_t)c);
      cp_async_commit();
    };
    int jhi = P, which = 0;
    int jlo = chunk_lo(jhi);
    if (P > 0) load_chunk(bufs[0], jlo, jhi, t, nt);
    cp_async_wait_all();
    __syncthreads();
    while (jhi > 0) {
      const int njhi = jlo, njlo = njhi > 0 ? chunk_lo(njhi) : 0;
      if (njhi > 0 && warp > 0) load_chunk(bufs[which ^ 1], njlo, njhi, t - 32, nt - 32);
      if (warp == 0) {
        const double* buf = bufs[which];
        const double* lbuf = buf + (size_t)(col_ptr[jhi] - col_ptr[jlo]) * 36;
        const int r = lane % 6, g = lane / 6;   // 5 lane groups walk a column's blocks; lanes 30,31 idle
        for (int j = jhi - 1; j >= jlo; --j) {
          const int base = col_ptr[j], nb = col_ptr[j + 1] - base - 1;
          double acc = 0.;
          if (lane < 30)
            for (int a = g; a < nb; a += 5) {
              const double* La = buf + (size_t)(base + 1 + a - col_ptr[jlo]) * 36;
              const double* xa = yv + 6 * row_idx[base + 1 + a];
#pragma unroll
              for (int q = 0; q < 6; ++q) acc += La[q * 6 + r] * xa[q];
            }
          double tot = acc;
          tot += __shfl_down_sync(0xffffffffu, acc, 6);
          const double a12 = __shfl_down_sync(0xffffffffu, acc, 12);
          const double a18 = __shfl_down_sync(0xffffffffu, acc, 18);
          const double a24 = __shfl_down_sync(0xffffffffu, acc, 24);
          tot += a12 + a18 + a24;
          double v = 0.;
          if (lane < 6) v = yv[6 * j + lane] - tot;
          const double* Li = lbuf + (size_t)(j - jlo) * 36;
          double xr = 0.;   // x_r = sum_{q >= r} Linv[q][r] v_q
#pragma unroll
          for (int q = 0; q < 6; ++q) {
            const double vq = __shfl_sync(0xffffffffu, v, q);
            if (lane < 6 && q >= lane) xr += Li[q * 6 + lane] * vq;
          }
          if (lane < 6) yv[6 * j + lane] = xr;
          __syncwarp();
        }
      }
      cp_async_wait_all();
      __syncthreads();
      jhi = njhi; jlo = njlo; which ^= 1;
    }
  }
  __syncthreads();
  TICK(7);
  if (d.dbg && (t == 0 || t == kPanelThreads)) {
    for (int i = 0; i < 12; ++i) d.dbg[(t ? 12 : 0) + i] = tm[i];
  }
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
  const size_t mbytes = ((size_t)(2 * (d.P + 1) + 2 * d.nblk + d.P) * 4 + 15) / 16 * 16;
  int cap = 0;
  if (ybytes + mbytes < budget) {
    const int avail = (int)((budget - ybytes - mbytes) / 288);
    for (cap = 1; cap * 2 <= avail; cap *= 2) {}
    if (cap > avail) cap = 0;
  }
  // the look-ahead kernel keeps two columns live and never checks residency on them
  if (d.P == 0 || cap < 4 * (max_col_blocks + 1)) {
    launch_solve_general(d, st);
    return;
  }
  while (cap / 2 >= d.nblk && cap / 2 >= 4 * (max_col_blocks + 1)) cap /= 2;   // small problems: small ring
  // columns j+1 and j+2 must stay resident between refills: cap >= (period + 3) * widest column
  int period = cap / (max_col_blocks + 1) - 3;
  period = period < 1 ? 1 : (period > 32 ? 32 : period);
  const size_t smem = (size_t)cap * 288 + ybytes + mbytes;
  k_solve<<<1, kSolveThreads, smem, st>>>(d, cap, y_in_smem, period);
}

}  // namespace svs
