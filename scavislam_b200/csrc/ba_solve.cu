// ba_solve.cu -- k_solve: block-sparse right-looking Cholesky of the reduced camera system
// S dx = bs, forward/backward solve and pose update, one CTA.
//   replaces g2o::LinearSolverCSparse<Matrix6d>::solve (instantiated at slam_graph.cpp:55-60) and
//   G2oVertexSE3::oplusImpl (anchored_points.cpp:53-58).
//
// The factor is a latency chain of P block columns.  What bounds it is not HBM and not the tensor
// cores (6x6 blocks, FP64) but the instruction latency of the three things a column step waits for --
// the pivot chain (6 pivots: rsqrt -> mul -> fma), the part of the update that lands in the next
// column, and the rest of the trailing update -- plus two barriers.  The design:
//   * blocks live in a shared-memory ring that covers the next `cap` blocks in column-major order
//     (the whole band of a SLAM window); trailing updates are shared-memory RMWs, blocks outside the
//     ring (far fill) fall back to global RMWs; ring refills are LDGSTS (cp.async) every few columns;
//   * look-ahead: panel warps factor column j+1 while update warps apply column j's trailing update
//     in quarter-block units (at most one per thread), read from a list staged in shared memory;
//   * the diagonal block is updated by 21 lanes and factored redundantly in registers by every lane of
//     the chain warp (no shuffles or shared-memory round trips between pivots; rsqrt + multiplies);
//   * warp roles follow the schedulers (warp % 4): the chain warp shares its scheduler with the
//     least loaded update warp (scripts/ubench/chol.cu, fp64_lanes.cu);
//   * two teams factor the window from both ends concurrently and meet in a separator;
//   * the right-hand side rides along as one more row of the panel (forward solve for free); the
//     backward solve streams the transposed factor back through the ring.
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

constexpr int kUpdStage = 128;    // update-list entries of a column staged in shared memory one column ahead
constexpr int kMaxTeams = 4;      // independent branches of the elimination tree factored concurrently

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

// A team = a group of warps of the CTA that factors one contiguous range of columns.  With a
// nested-dissection ordering the ranges of different teams are branches of the elimination tree
// that only meet in the separator columns at the end: they run concurrently (the factorisation is
// a latency chain, so the win is the shorter chain, not the extra lanes), scatter into the separator
// blocks with FP64 atomics, and the separators are then factored by the whole CTA as one team.
struct Team {
  int tid, nth, npanel;       // thread index within the team, team size, threads on the look-ahead panel
  int bar_all, bar_panel;     // named barriers
  int slot;                   // index of the team's fail flags / diagonal-factor buffers
  int ring_off, cap;          // the team's share of the shared-memory ring (blocks), power of two
  int j0, j1;                 // column range [j0, j1)
  int sep_pos0, sep_blk0;     // first separator column / its first block (P / nblk when there is none)
};

struct SolveShared {
  double* ring; double* yv;
  int* col_ptr; int* upd_ptr; int* row_idx; int* urg_dst; int* sfix;
  int (*fail)[2];
  double (*sL)[2][28];
  int2 (*upd)[2][kUpdStage];   // [min(slot, 2)]
};

// Right-looking block Cholesky of columns [T.j0, T.j1) with look-ahead; forward solve rides along.
__device__ void factor_range(const BaDev& d, const Team& T, const SolveShared& S, double lambda, int refill_period) {
  const int t = T.tid, lane = threadIdx.x & 31, twarp = T.tid >> 5;
  // Panel threads are the team's first T.npanel threads (its warps 0..3); update units are dealt from the
  // team's LAST warp downwards, because warp w issues from scheduler w % 4: the pivot-chain warp (panel
  // warp 0) then shares its scheduler and FP64 pipe with the least loaded update warp (measured: -15 % per
  // column).  For the same reason the second team's pivot chain runs on its panel warp 1, not 0.
  const bool crit = t < T.npanel;
  const int npw = T.npanel >> 5;
  const int pw = crit ? (twarp + (T.slot == 1 ? npw - 1 : 0)) % npw : twarp;
  const int pt = pw * 32 + lane;
  const int kPanel = T.npanel;
  const int nupd = T.nth - T.npanel, ut = T.nth - 1 - t;
  const int mask = T.cap - 1;
  double* ring = S.ring + (size_t)T.ring_off * 36;
  double* yv = S.yv;
  const int* col_ptr = S.col_ptr; const int* upd_ptr = S.upd_ptr; const int* row_idx = S.row_idx;
  const int* urg_dst = S.urg_dst;
  const int blk_end = col_ptr[T.j1];
  if (T.j0 >= T.j1) return;

  // ring fill: blocks [col_ptr[j0], hi)
  int hi = min(blk_end, col_ptr[T.j0] + T.cap);
  for (int c = t; c < (hi - col_ptr[T.j0]) * 18; c += T.nth) {
    const int id = col_ptr[T.j0] + c / 18, w = c % 18;
    cp_async16(ring + (size_t)(id & mask) * 36 + 2 * w, d.S + (size_t)id * 36 + 2 * w);
  }
  cp_async_commit();
  cp_async_wait_all();
  bar_sync(T.bar_all, T.nth);

#ifdef SVS_SOLVE_PROFILE
  long long pacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long pc0 = 0, pc1 = clock64();
  long long* pprof = pacc;
  long long* pclk = &pc1;
#endif
  // factor + scale the panel of column jn (panel threads): pivot chain on warp 0, rows on all
  auto panel_column = [&](int jn, const double* Lsub) {
    const int base = col_ptr[jn], nb = col_ptr[jn + 1] - base - 1;
    double* sl = S.sL[T.slot][jn & 1];
    if (pw == 0) {
      double* D = ring + (size_t)(base & mask) * 36;
      if (Lsub) {
        // S_jj -= L L^T with L = L_{jn,jn-1} just scaled by the previous column: symmetric, 21 lanes, one
        // round -- the only part of that column's update the pivot chain waits for
        if (lane < 21) {
          const int r = (lane >= 1) + (lane >= 3) + (lane >= 6) + (lane >= 10) + (lane >= 15);
          const int c = lane - r * (r + 1) / 2;
          const double2* La = reinterpret_cast<const double2*>(Lsub + r * 6);
          const double2* Lb = reinterpret_cast<const double2*>(Lsub + c * 6);
          const double2 a0 = La[0], a1 = La[1], a2 = La[2], b0 = Lb[0], b1 = Lb[1], b2 = Lb[2];
          D[r * 6 + c] -= (a0.x * b0.x + a0.y * b0.y) + (a1.x * b1.x + a1.y * b1.y) + (a2.x * b2.x + a2.y * b2.y);
        }
        __syncwarp();
      }
      // every lane factors the block redundantly in registers: no shuffles or shared-memory round trips on
      // the pivot chain (scripts/ubench/chol.cu: both queue behind the update warps' traffic)
      double l[21], rinv[6];
      const bool ok = chol6_lean(D, lambda + (S.sfix[jn] ? 1. : 0.), l, rinv);
      if (lane == 0 && !ok) S.fail[T.slot][jn & 1] = 1;
      {   // sl[i] <- l[i] (i < 21), rinv[i - 21]: one store per lane instead of 27 by lane 0
        double v = 0.;
#pragma unroll
        for (int i = 0; i < 21; ++i) v = (lane == i) ? l[i] : v;
#pragma unroll
        for (int i = 0; i < 6; ++i) v = (lane == 21 + i) ? rinv[i] : v;
        if (lane < 27) sl[lane] = v;
      }
#ifdef SVS_SOLVE_PROFILE
      if (pprof) { const long long c_ = clock64(); pprof[1] += c_ - *pclk; *pclk = c_; }
#endif
    }
    bar_sync(T.bar_panel, kPanel);
#ifdef SVS_SOLVE_PROFILE
    if (pprof && sl[21] != 0.) { const long long c_ = clock64(); pprof[2] += c_ - *pclk; *pclk = c_; }
#endif
    // row <- row * L^-T by forward substitution (block rows: L_ij ; rhs row: y = L^-1 b)
    const int nrows = nb * 6 + 1;
    for (int row = pt; row < nrows; row += kPanel) {
      double* src;
      double* gdst = nullptr;
      if (row < nb * 6) {
        const int a = row / 6, r = row - a * 6;
        src = ring + (size_t)((base + 1 + a) & mask) * 36 + r * 6;
        gdst = d.S + (size_t)(base + 1 + a) * 36 + r;
      } else {
        src = yv + 6 * jn;
      }
      double o[6], v[6], L_[28];
      {
        const double2* s2 = reinterpret_cast<const double2*>(src);
        const double2 s0 = s2[0], s1 = s2[1], s3 = s2[2];
        v[0] = s0.x; v[1] = s0.y; v[2] = s1.x; v[3] = s1.y; v[4] = s3.x; v[5] = s3.y;
        const double2* l2 = reinterpret_cast<const double2*>(sl);
#pragma unroll
        for (int q = 0; q < 14; ++q) { const double2 t2 = l2[q]; L_[2 * q] = t2.x; L_[2 * q + 1] = t2.y; }
      }
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        double s = v[c];
#pragma unroll
        for (int q = 0; q < c; ++q) s -= o[q] * L_[c * (c + 1) / 2 + q];
        o[c] = s * L_[21 + c];
      }
      {
        double2* d2 = reinterpret_cast<double2*>(src);
        d2[0] = make_double2(o[0], o[1]); d2[1] = make_double2(o[2], o[3]); d2[2] = make_double2(o[4], o[5]);
        if (gdst) {   // final factor column for the backward solve, stored TRANSPOSED (L_ij^T row-major):
                      // the backward chain then reads rows of L^T with 16-byte loads
#pragma unroll
          for (int q = 0; q < 6; ++q) gdst[q * 6] = o[q];
        }
      }
    }
  };

  // a quarter of a pair update: rows 3h..3h+2, columns 3g..3g+2 of S_ab -= L_a L_b^T.  What bounds a column
  // step is the instruction count of its longest warp (a lone warp retires an instruction every 7-12
  // cycles here), so the unit is sized to give every update thread at most one: 18 16-byte loads, 54 FMAs
  // into accumulators preloaded with the destination, 12 loads/stores of the destination.
  auto update_quarter = [&](int base, int u, int ab, int dst, int hi_res) {
    const int h = (u >> 1) & 1, g = u & 1;
    const double2* La = reinterpret_cast<const double2*>(ring + (size_t)((base + 1 + (ab >> 16)) & mask) * 36 + h * 18);
    const double2* Lb = reinterpret_cast<const double2*>(ring + (size_t)((base + 1 + (ab & 0xffff)) & mask) * 36 + g * 18);
    double a[18], b[18], o[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) { const double2 v = La[q]; a[2 * q] = v.x; a[2 * q + 1] = v.y; }
#pragma unroll
    for (int q = 0; q < 9; ++q) { const double2 v = Lb[q]; b[2 * q] = v.x; b[2 * q + 1] = v.y; }
    const bool in_ring = dst < hi_res, shared_sep = !in_ring && dst >= T.sep_blk0;
    double* D = (in_ring ? ring + (size_t)(dst & mask) * 36 : d.S + (size_t)dst * 36) + h * 18 + g * 3;
    if (shared_sep) {
#pragma unroll
      for (int i = 0; i < 9; ++i) o[i] = 0.;
    } else {
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) o[r * 3 + c] = D[r * 6 + c];
    }
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int k = 0; k < 6; ++k) o[r * 3 + c] = fma(-a[r * 6 + k], b[c * 6 + k], o[r * 3 + c]);
    if (shared_sep) {   // separator block shared with the other teams
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) atomicAdd(D + r * 6 + c, o[r * 3 + c]);
    } else {
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) D[r * 6 + c] = o[r * 3 + c];
    }
  };

  // pairs of column j that land in column j+1 (pairs (a, 0) when the first sub-diagonal row is j+1)
  auto urgent_of = [&](int j) {
    const int base = col_ptr[j], nb = col_ptr[j + 1] - base - 1;
    return (j + 1 < T.j1 && nb > 0 && row_idx[base + 1] == j + 1) ? nb : 0;
  };

  if (crit) panel_column(T.j0, nullptr);
  // column j's list of non-urgent pairs is staged in shared memory during column j-1
  int2 (*stage)[kUpdStage] = S.upd[T.slot < 2 ? T.slot : 2];
  auto stage_load = [&](int j, int2& e) {
    const int v0 = upd_ptr[j] + urgent_of(j), np = upd_ptr[j + 1] - v0;
    const bool mine = ut < np && ut < kUpdStage;
    if (mine) e = make_int2(d.upd_ab[v0 + ut], d.upd_dst[v0 + ut]);
    return mine;
  };
  if (!crit) {
    int2 e;
    if (stage_load(T.j0, e)) stage[T.j0 & 1][ut] = e;
  }
  bar_sync(T.bar_all, T.nth);
  int failed = S.fail[T.slot][T.j0 & 1];
  int until_refill = refill_period;

#ifdef SVS_SOLVE_PROFILE
#define PSTAMP(i) do { const long long c_ = clock64(); pacc[i] += c_ - pc1; pc1 = c_; } while (0)
#else
#define PSTAMP(i) do {} while (0)
#endif
  for (int j = T.j0; j < T.j1 && !failed; ++j) {
    const int base = col_ptr[j], nb = col_ptr[j + 1] - base - 1;
    const int urgent = urgent_of(j);
    const int u0 = upd_ptr[j];
#ifdef SVS_SOLVE_PROFILE
    pc0 = pc1 = clock64();
#endif
    if (crit) {
      // ---- panel threads: the part of column j's update that lands in column j+1, then factor it
      if (urgent) {
        if (pw == 0) {
          // (the diagonal part of this update is done by panel_column below)
          PSTAMP(0);
        } else {
          if (pw == 1 && lane < 6) {   // b_{j+1} -= L_{j+1,j} y_j
            const double* La = ring + (size_t)((base + 1) & mask) * 36 + lane * 6;
            const double* yj = yv + 6 * j;
            double sv = 0.;
#pragma unroll
            for (int q = 0; q < 6; ++q) sv += La[q] * yj[q];
            yv[6 * (j + 1) + lane] -= sv;
          }
          for (int u = 4 + (pt - 32); u < urgent * 4; u += kPanel - 32) {
            const int a = u >> 2;
            update_quarter(base, u, a << 16, urg_dst[base + 1 + a], 0x7fffffff);
          }
        }
      }
      if (j + 1 < T.j1)   // its barrier also orders the other panel warps' updates before the row scaling
        panel_column(j + 1, urgent ? ring + (size_t)((base + 1) & mask) * 36 : nullptr);
      PSTAMP(3);
    } else {
      // ---- update threads: the rest of column j's trailing update
      int2 nxt;
      const bool have_nxt = j + 1 < T.j1 && stage_load(j + 1, nxt);
      const int uu = u0 + urgent, npairs = upd_ptr[j + 1] - uu;
      const int2* cur = stage[j & 1];
      for (int u = ut; u < npairs * 4; u += nupd) {
        const int pr = u >> 2;
        const int2 e = pr < kUpdStage ? cur[pr] : make_int2(d.upd_ab[uu + pr], d.upd_dst[uu + pr]);
        update_quarter(base, u, e.x, e.y, hi);
      }
      if (have_nxt) stage[(j + 1) & 1][ut] = nxt;
      // b_a -= L_aj y_j for the rows the panel threads did not take
      for (int w = ut + (urgent ? 6 : 0); w < nb * 6; w += nupd) {
        const int a = w / 6, r = w - a * 6;
        const double* La = ring + (size_t)((base + 1 + a) & mask) * 36 + r * 6;
        const double* yj = yv + 6 * j;
        double s = 0.;
#pragma unroll
        for (int q = 0; q < 6; ++q) s += La[q] * yj[q];
        const int row = row_idx[base + 1 + a];
        if (row >= T.sep_pos0) atomicAdd(yv + 6 * row + r, -s);   // separator row shared with the other teams
        else yv[6 * row + r] -= s;
      }
      // inverse of column j's diagonal factor for the backward solve (off the critical path)
      if (twarp == npw + 1 && lane < 6) {
        const double* sl = S.sL[T.slot][j & 1];
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
      PSTAMP(5);
    }
    bar_sync(T.bar_all, T.nth);
    failed = S.fail[T.slot][(j + 1) & 1];
#ifdef SVS_SOLVE_PROFILE
    if (failed >= 0) { if (crit) PSTAMP(4); else PSTAMP(6); }
#endif
    // ---- every refill_period columns: reload the ring slots the finished columns freed.  Copies are
    //      never in flight while updates run, so a destination is either resident (< hi) or in HBM.
    if (--until_refill == 0) until_refill = refill_period;
    if (until_refill == refill_period && hi < blk_end && j + 1 < T.j1) {
      const int hi_new = min(blk_end, col_ptr[j + 1] + T.cap);
      for (int c = t; c < (hi_new - hi) * 18; c += T.nth) {
        const int id = hi + c / 18, w = c % 18;
        cp_async16(ring + (size_t)(id & mask) * 36 + 2 * w, d.S + (size_t)id * 36 + 2 * w);
      }
      cp_async_commit();
      cp_async_wait_all();
      bar_sync(T.bar_all, T.nth);
      hi = hi_new;
    }
  }
  if (failed && t == 0) S.fail[T.slot][0] = S.fail[T.slot][1] = 1;
#ifdef SVS_SOLVE_PROFILE
  (void)pc0;
  if (d.dbg && T.slot == 0 && (t == 0 || t == T.npanel))
    for (int i = 0; i < 8; ++i) d.dbg[24 + (t ? 8 : 0) + i] = pacc[i];
#endif
}

// Backward solve L^T x = y for columns [T.j0, T.j1), descending: the factor is streamed back through
// the team's ring in chunks (L blocks + inverse diagonal factors), double buffered; the team's warp 0
// walks the dependency chain out of shared memory while its other warps fetch the next chunk.
__device__ void backsolve_range(const BaDev& d, const Team& T, const SolveShared& S) {
  if (T.j0 >= T.j1) return;
  const int t = T.tid, lane = threadIdx.x & 31, twarp = T.tid >> 5;
  const int* col_ptr = S.col_ptr; const int* row_idx = S.row_idx;
  double* yv = S.yv;
  const int half = T.cap / 2;
  double* bufs[2] = {S.ring + (size_t)T.ring_off * 36, S.ring + (size_t)(T.ring_off + half) * 36};
  auto chunk_lo = [&](int jhi) {   // largest [jlo, jhi) whose blocks + diagonal inverses fit one half
    int jlo = jhi - 1;
    while (jlo > T.j0 && (col_ptr[jhi] - col_ptr[jlo - 1]) + (jhi - (jlo - 1)) <= half) --jlo;
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
  int jhi = T.j1, which = 0;
  int jlo = chunk_lo(jhi);
  load_chunk(bufs[0], jlo, jhi, t, T.nth);
  cp_async_wait_all();
  bar_sync(T.bar_all, T.nth);
  while (jhi > T.j0) {
    const int njhi = jlo, njlo = njhi > T.j0 ? chunk_lo(njhi) : T.j0;
    if (njhi > T.j0 && twarp > 0) load_chunk(bufs[which ^ 1], njlo, njhi, t - 32, T.nth - 32);
    if (twarp == 0) {
      const double* buf = bufs[which];
      const double* lbuf = buf + (size_t)(col_ptr[jhi] - col_ptr[jlo]) * 36;
      const int r = lane % 6, g = lane / 6;   // 5 lane groups walk a column's blocks; lanes 30,31 idle
      for (int j = jhi - 1; j >= jlo; --j) {
        const int base = col_ptr[j], nb = col_ptr[j + 1] - base - 1;
        double acc = 0.;
        if (lane < 30)
          for (int a = g; a < nb; a += 5) {
            const double2* Lt = reinterpret_cast<const double2*>(buf + (size_t)(base + 1 + a - col_ptr[jlo]) * 36 + r * 6);
            const double2* xa = reinterpret_cast<const double2*>(yv + 6 * row_idx[base + 1 + a]);
            const double2 l0 = Lt[0], l1 = Lt[1], l2 = Lt[2], x0 = xa[0], x1 = xa[1], x2 = xa[2];
            acc += (l0.x * x0.x + l0.y * x0.y) + (l1.x * x1.x + l1.y * x1.y) + (l2.x * x2.x + l2.y * x2.y);
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
    bar_sync(T.bar_all, T.nth);
    jhi = njhi; jlo = njlo; which ^= 1;
  }
}

// smem layout: [ring: cap*36 doubles][y: 6P doubles][meta ints: col_ptr (P+1), upd_ptr (P+1),
//               row_idx (nblk), urg_dst (nblk), fixed-by-position (P)]
// cap is a power of two; every team's share is >= 4 * (widest column of its range + 1).
__global__ void __launch_bounds__(kSolveThreads)
k_solve(BaDev d, int cap, int refill_branch, int refill_sep) {
  extern __shared__ __align__(16) double sm_solve[];
  __shared__ int sFail[kMaxTeams + 1][2];
  __shared__ double sRed[kSolveThreads / 32];
  __shared__ __align__(16) double sLbuf[kMaxTeams + 1][2][28];
  __shared__ int2 sUpd[3][2][kUpdStage];
  LmCtl* ctl = d.ctl;
  if (ctl->max_iters > 0 && (ctl->stop || ctl->iter >= ctl->max_iters)) return;   // speculatively enqueued trial: nothing left to do
  const int t = threadIdx.x, nt = kSolveThreads, lane = t & 31, warp = t >> 5;
  const int P = d.P, nblk = d.nblk;
  const double lambda = ctl->lambda;
  const int cur = ctl->cur;
  SolveShared S;
  S.ring = sm_solve;
  S.yv = sm_solve + (size_t)cap * 36;
  int* meta = reinterpret_cast<int*>(S.yv + ((6 * (size_t)P + 1) / 2) * 2);
  S.col_ptr = meta; S.upd_ptr = S.col_ptr + (P + 1); S.row_idx = S.upd_ptr + (P + 1);
  S.urg_dst = S.row_idx + nblk; S.sfix = S.urg_dst + nblk;
  S.fail = sFail; S.sL = sLbuf; S.upd = sUpd;
  if (t < 2 * (kMaxTeams + 1)) sFail[t >> 1][t & 1] = 0;
  for (int i = t; i <= P; i += nt) { S.col_ptr[i] = d.col_ptr[i]; S.upd_ptr[i] = d.upd_ptr[i]; }
  for (int i = t; i < nblk; i += nt) { S.row_idx[i] = d.row_idx[i]; S.urg_dst[i] = d.urg_dst[i]; }
  for (int i = t; i < P; i += nt) S.sfix[i] = d.fixed[d.perm[i]];
  for (int i = t; i < 6 * P; i += nt) {   // right-hand side in elimination order: bs = bp - bc
    const int j = i / 6, r = i - 6 * j;
    const int p = d.perm[j];
    S.yv[i] = d.bp[6 * p + r] - d.bc[6 * p + r];
  }
  __syncthreads();
  long long tk[8];
  tk[0] = clock64();

  const int G = d.nbranch;                 // 1: a single chain (no nested dissection)
  const int sep0 = d.branch_ptr[G];        // first separator column (= P when G == 1)
  Team whole;
  whole.tid = t; whole.nth = nt; whole.npanel = 128; whole.bar_all = 0; whole.bar_panel = 9; whole.slot = kMaxTeams;
  whole.ring_off = 0; whole.cap = cap; whole.sep_pos0 = P; whole.sep_blk0 = nblk;
  Team mine = whole;
  if (G > 1) {
    const int tn = nt / G, g = t / tn;
    mine.tid = t - g * tn; mine.nth = tn; mine.npanel = tn / 2; mine.bar_all = 1 + 2 * g; mine.bar_panel = 2 + 2 * g;
    mine.slot = g; mine.cap = cap / G; mine.ring_off = g * (cap / G);
    mine.j0 = d.branch_ptr[g]; mine.j1 = d.branch_ptr[g + 1];
    mine.sep_pos0 = sep0; mine.sep_blk0 = S.col_ptr[sep0];
    factor_range(d, mine, S, lambda, refill_branch);
    tk[1] = clock64();
    __threadfence_block();
    __syncthreads();
    tk[2] = clock64();
    whole.j0 = sep0; whole.j1 = P;
    factor_range(d, whole, S, lambda, refill_sep);
  } else {
    tk[1] = tk[2] = tk[0];
    whole.j0 = 0; whole.j1 = P;
    factor_range(d, whole, S, lambda, refill_sep);
  }
  cp_async_wait_all();
  __syncthreads();
  tk[3] = clock64();
  int failed = 0;
  for (int g = 0; g <= kMaxTeams; ++g) failed |= sFail[g][0] | sFail[g][1];
  if (failed) {
    if (t == 0) { ctl->chol_fail = 1; ctl->scale_pose = 0; }
    for (int i = t; i < 7 * P; i += nt) d.pose[1 - cur][i] = d.pose[cur][i];
    for (int i = t; i < 12 * P; i += nt) d.Rt[1 - cur][i] = d.Rt[cur][i];
    for (int i = t; i < 6 * P; i += nt) d.x[i] = 0;
    return;
  }
  // backward: separators first, then the branches concurrently
  if (G > 1) {
    backsolve_range(d, whole, S);
    __syncthreads();
    tk[4] = clock64();
    backsolve_range(d, mine, S);
    tk[5] = clock64();
  } else {
    backsolve_range(d, whole, S);
    tk[4] = tk[5] = clock64();
  }
  __syncthreads();
  tk[6] = clock64();
  if (d.dbg && (t & 127) == 0) {   // one thread per team: phase boundaries in cycles since the setup
    for (int i = 1; i < 7; ++i) d.dbg[(t >> 7) * 6 + i - 1] = tk[i] - tk[0];
  }
  double* yv = S.yv;
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

// Shared-memory budget of k_solve: ring capacity (blocks, power of two) for a problem, 0 if the
// right-hand side and the index metadata alone do not fit.
int solve_ring_capacity(int P, int nblk) {
  static int smem_optin = -1;
  if (smem_optin < 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cudaFuncSetAttribute(k_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin - 10240);   // 9.6 KB static
  }
  const size_t budget = (size_t)smem_optin - 10240 - 256;
  const size_t ybytes = (((size_t)6 * P * 8 + 15) / 16) * 16;
  const size_t mbytes = ((size_t)(2 * (P + 1) + 2 * nblk + P) * 4 + 15) / 16 * 16;
  if (ybytes + mbytes >= budget || ybytes > budget / 4) return 0;
  const int avail = (int)((budget - ybytes - mbytes) / 288);
  int cap = 1;
  while (cap * 2 <= avail) cap *= 2;
  return cap <= avail ? cap : 0;
}

// Launches the look-ahead kernel when every team's ring share holds 4 of its widest columns
// (it keeps two columns live and never checks residency on them); otherwise the global-memory kernel.
void launch_solve(const BaDev& d, int max_col_branch, int max_col_sep, cudaStream_t st) {
  int cap = d.P > 0 ? solve_ring_capacity(d.P, d.nblk) : 0;
  const int G = d.nbranch;
  if (cap == 0 || cap / G < 4 * (max_col_branch + 1) || cap < 4 * (max_col_sep + 1)) {
    launch_solve_general(d, st);
    return;
  }
  while (G == 1 && cap / 2 >= d.nblk && cap / 2 >= 4 * (max_col_sep + 1)) cap /= 2;   // small problems: small ring
  auto period = [](int c, int widest) { int p = c / (widest + 1) - 3; return p < 1 ? 1 : (p > 32 ? 32 : p); };
  const size_t ybytes = (((size_t)6 * d.P * 8 + 15) / 16) * 16;
  const size_t mbytes = ((size_t)(2 * (d.P + 1) + 2 * d.nblk + d.P) * 4 + 15) / 16 * 16;
  const size_t smem = (size_t)cap * 288 + ybytes + mbytes;
  k_solve<<<1, kSolveThreads, smem, st>>>(d, cap, period(cap / G, max_col_branch), period(cap, max_col_sep));
}

}  // namespace svs
