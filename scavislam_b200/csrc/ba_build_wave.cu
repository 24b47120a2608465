// ba_build_wave.cu -- k_build_wave: the fused linearise + J^T W J + landmark elimination + Schur
// scatter for landmark groups that share one (anchor, observer set) with at most 8 frames --
// the bulk of a SLAM window.  Same mathematics as k_build (ba_build.cu), different mapping:
//
//   * a warp owns a TASK: up to `chunk` consecutive landmarks with identical slot lists, so all
//     of them scatter into the same K(K+1)/2 blocks of the reduced camera system;
//   * edges are linearised 32 at a time (one lane per edge, several landmarks per wave), which
//     keeps every lane busy where one-warp-per-landmark leaves 3/4 of them idle;
//   * the Schur products Y_m B_n^T (and the direct J^T W J terms) are accumulated over all
//     landmarks of the task in registers -- each lane owns up to 7 (pair, row) units of 6 doubles --
//     and flushed with ONE set of RED.F64 per task instead of one per landmark
//     (measured on B200: 450 G coalesced FP64 reductions/s, scripts/ubench/lat.cu -- the
//     per-landmark scatter alone would cost 39 us on the 200-keyframe window).
//
// Reference semantics: G2oEdgeProjectPSI2UVU::linearizeOplus (anchored_points.cpp:168-189), g2o
// BaseMultiEdge::constructQuadraticForm, BlockSolver<6,3>::buildSystem / solve (Schur part).
#include <algorithm>
#include <cstdlib>

#include "ba_dev.cuh"
#include "ba_kernels.cuh"

namespace svs {

constexpr int kWvWarps = 4;
constexpr int kWvLm = 8;       // landmarks per wave
constexpr int kWvSlots = 40;   // slots per wave
constexpr int kWvJ = 19;       // row stride of the per-edge Jacobian rows (odd: conflict-free 64-bit stores)
constexpr int kWvLmD = 56;     // per-landmark scratch: D(6) bl(3) Dinv(9) A_aa(36) + pad
constexpr int kWvDoubles = 2 * 32 * kWvJ + 32 * 9 + 32 * 3 + 2 * kWvSlots * 18 + kWvLm * kWvLmD + 32;
constexpr int kWvInts = 8 + 40;   // slot poses, pair table (36) padded

size_t build_wave_smem_bytes() { return (size_t)kWvWarps * (kWvDoubles * 8 + kWvInts * 4); }

__global__ void __launch_bounds__(kWvWarps * 32)
k_build_wave(BaDev d, int robust, double delta, int n_task_blocks, int prof, int persist) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const LmCtl* __restrict__ ctl = d.ctl;
  if (ctl->max_iters > 0 && (ctl->stop || ctl->iter >= ctl->max_iters)) return;   // speculatively enqueued trial: nothing left to do
  const int cur = ctl->cur;
  // pose-pose constraints (G2oEdgeSE3), one thread each, riding on CTAs of this launch so that their long
  // serial 6x6 arithmetic overlaps the landmark work instead of following it: the trailing CTAs of a
  // one-task-per-warp grid, the LEADING ones of a persistent grid (its task CTAs stay until the list is empty)
  {
    const int c_blocks = (int)gridDim.x - n_task_blocks;
    const int cb = persist ? (int)blockIdx.x : (int)blockIdx.x - n_task_blocks;
    if (cb >= 0 && cb < c_blocks) {
      const int c = cb * (kWvWarps * 32) + (int)threadIdx.x;
      if (c < d.C) constraint_build(d, d.pose[cur], c);
      return;
    }
  }
  // persistent grid: the warps draw tasks from one counter (longest tasks first, set_problem sorts them), so a
  // warp slot is never idle while tasks remain; the last warp to leave resets the counter pair for the next launch
  unsigned* const tctr = d.ticket + 1;
  auto next_task = [&]() {
    int t = 0;
    if (lane == 0) t = (int)atomicAdd(tctr, 1u);
    return __shfl_sync(0xffffffffu, t, 0);
  };
  const double lambda = ctl->lambda;
  // (Drawing the next ticket early was measured, profiles/r02_build_spill_prefetch_ab.txt: when a task starts, the
  //  atomic's round trip is hidden but a reserved task waits for its owner while other warps run dry at the end of
  //  the list -- 0.105 instead of 0.082 ms on the 200-keyframe window; before the flush: no difference.)
  for (int task = persist ? next_task() : (int)blockIdx.x * kWvWarps + warp; task < d.ntasks;
       task = persist ? next_task() : d.ntasks) {
  double* sm = reinterpret_cast<double*>(smem_raw) + (size_t)warp * kWvDoubles;
  double* sJp = sm;
  double* sJa = sJp + 32 * kWvJ;
  double* sJs = sJa + 32 * kWvJ;
  double* sE = sJs + 32 * 9;
  double* sB = sE + 32 * 3;
  double* sY = sB + kWvSlots * 18;
  double* sLm = sY + kWvSlots * 18;
  double* sChi = sLm + kWvLm * kWvLmD;
  int* si = reinterpret_cast<int*>(reinterpret_cast<double*>(smem_raw) + (size_t)kWvWarps * kWvDoubles) + warp * kWvInts;
  int* sPose = si;
  int* sPair = si + 8;

  const int lm0 = d.task_lm[task], nlm = d.task_cnt[task];
  const int e_base = d.lm_eptr[lm0], k = d.lm_eptr[lm0 + 1] - e_base;
  const int s_base = d.lm_sptr[lm0], K = d.lm_sptr[lm0 + 1] - s_base;
  const int has_self = d.lm_self[lm0];
  const int off = has_self ? 0 : 1;
  const int i_first = has_self ? 1 : 0;
  const int ia = d.lm_anchor[lm0];
  const int fa = d.fixed[ia];
  const int skip_self = d.flags & 1;
  const double* __restrict__ Rt = d.Rt[cur];
  double Ra[9], ta[3];
  load12(Rt, ia, Ra, ta);

  // slot poses and the pair table, ordered by kind so that the 32 units of a round mostly share a
  // code path: anchor-row pairs (0,n), diagonal pairs (m,m), (0,0), then the plain pairs
  if (lane == 0) sPose[0] = ia;
  if (lane < k && !(has_self && lane == 0)) sPose[lane + off] = d.e_pose[e_base + lane];
  __syncwarp();
  const int npairs = K * (K + 1) / 2;
  for (int p = lane; p < npairs; p += 32) {
    int m, n;
    if (p < K - 1) { m = 0; n = 1 + p; }
    else if (p < 2 * K - 2) { m = n = 1 + p - (K - 1); }
    else if (p == 2 * K - 2) { m = n = 0; }
    else {
      int rem = p - (2 * K - 1);
      m = 1;
      while (rem >= K - 1 - m) { rem -= K - 1 - m; ++m; }
      n = m + 1 + rem;
    }
    const int t = d.tbl[(size_t)sPose[m] * d.P + sPose[n]];
    sPair[p] = ((t >> 1) << 11) | ((t & 1) << 10) | (m << 5) | n;
  }
  __syncwarp();

  const int nunits = npairs * 6;
  double acc[7][6];
#pragma unroll
  for (int q = 0; q < 7; ++q)
#pragma unroll
    for (int c = 0; c < 6; ++c) acc[q][c] = 0.;
  double accg[2] = {0., 0.}, accc[2] = {0., 0.};

  int nw_max = 32 / k;
  if (nw_max > kWvSlots / K) nw_max = kWvSlots / K;
  if (nw_max > kWvLm) nw_max = kWvLm;

  long long pacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, pclk = prof ? clock64() : 0;
#define PBW(i) do { if (prof) { __syncwarp(); const long long c_ = clock64(); pacc[i] += c_ - pclk; pclk = c_; } } while (0)
  PBW(0);
  for (int w0 = 0; w0 < nlm; w0 += nw_max) {
    const int nw = min(nw_max, nlm - w0);
    // ---- phase 1: one lane per edge of the wave
    double chi = 0.;
    if (lane < nw * k) {
      const int j = lane / k, i = lane - j * k;
      const int li = lm0 + w0 + j;
      const int e = e_base + (w0 + j) * k + i;
      const double* __restrict__ psi = d.psi[cur] + 3 * (size_t)li;
      const double p0 = __ldg(psi), p1 = __ldg(psi + 1), p2 = __ldg(psi + 2);
      const double ipz = fast_inv(p2);
      const double xa[3] = {p0 * ipz, p1 * ipz, ipz};
      const int ip = (has_self && i == 0) ? ia : sPose[i + off];
      double* Jp = sJp + kWvJ * lane;
      double* Js = sJs + 9 * lane;
      chi = linearize_edge(d, Rt, e, ip, Ra, ta, xa, ipz, fa, robust, delta, Jp, sJa + kWvJ * lane, Js, sE + 3 * lane);
      if (!(has_self && i == 0)) {   // own Hpl block B = J~p^T J~psi
        double* B = sB + 18 * (j * K + i + off);
#pragma unroll
        for (int r = 0; r < 6; ++r)
#pragma unroll
          for (int c = 0; c < 3; ++c) B[r * 3 + c] = Jp[r] * Js[c] + Jp[6 + r] * Js[3 + c] + Jp[12 + r] * Js[6 + c];
      }
    }
    sChi[lane] = chi;
    __syncwarp();
    PBW(1);
    if (lane < nw) {
      double s = 0.;
      for (int i = 0; i < k; ++i) s += sChi[lane * k + i];
      d.chi_l[lm0 + w0 + lane] = s;
    }
    // ---- phase 2: per-landmark sums over its edges: anchor Hpl block (18), Hll (6), b_l (3), A_aa (21);
    //      one loop per kind so that the lanes of a round share a code path
    for (int it = lane; it < nw * 18; it += 32) {
      const int j = it / 18, t = it - j * 18, r = t / 3, c = t - r * 3;
      const double* Ja = sJa + kWvJ * (j * k + i_first);
      const double* Js = sJs + 9 * (j * k + i_first);
      double s = 0.;
      for (int i = i_first; i < k; ++i, Ja += kWvJ, Js += 9) s += Ja[r] * Js[c] + Ja[6 + r] * Js[3 + c] + Ja[12 + r] * Js[6 + c];
      sB[18 * (j * K) + t] = s;
    }
    for (int it = lane; it < nw * 6; it += 32) {
      const int j = it / 6, u = it - j * 6;
      const int r = u < 3 ? 0 : (u < 5 ? 1 : 2), c = u < 3 ? u : (u < 5 ? u - 2 : 2);
      const double* Js = sJs + 9 * (j * k);
      double s = 0.;
      for (int i = 0; i < k; ++i, Js += 9) s += Js[r] * Js[c] + Js[3 + r] * Js[3 + c] + Js[6 + r] * Js[6 + c];
      sLm[j * kWvLmD + u] = s;
    }
    for (int it = lane; it < nw * 3; it += 32) {
      const int j = it / 3, c = it - j * 3;
      const double* Js = sJs + 9 * (j * k);
      const double* Ee = sE + 3 * (j * k);
      double s = 0.;
      for (int i = 0; i < k; ++i, Js += 9, Ee += 3) s -= Js[c] * Ee[0] + Js[3 + c] * Ee[1] + Js[6 + c] * Ee[2];
      sLm[j * kWvLmD + 6 + c] = s;
    }
    for (int it = lane; it < nw * 21; it += 32) {
      // anchor diagonal: all edges' J~a^T J~a; the self edge keeps g2o's J1^T W J1
      const int j = it / 21, u = it - j * 21;
      const int r = (u >= 1) + (u >= 3) + (u >= 6) + (u >= 10) + (u >= 15), c = u - r * (r + 1) / 2;
      const int i0 = skip_self ? i_first : 0;
      const double* Ja = sJa + kWvJ * (j * k + i0);
      double s = 0.;
      for (int i = i0; i < k; ++i, Ja += kWvJ) s += Ja[r] * Ja[c] + Ja[6 + r] * Ja[6 + c] + Ja[12 + r] * Ja[12 + c];
      sLm[j * kWvLmD + 18 + r * 6 + c] = s;
      sLm[j * kWvLmD + 18 + c * 6 + r] = s;
    }
    __syncwarp();
    PBW(2);
    // ---- phase 3: (Hll + lambda I)^-1 per landmark; Hll / b_l to HBM for the back-substitution
    if (lane < nw) {
      double Di[9];
      inv3_sym_lambda(sLm + lane * kWvLmD, lambda, Di);
#pragma unroll
      for (int q = 0; q < 9; ++q) sLm[lane * kWvLmD + 9 + q] = Di[q];
    }
    for (int it = lane; it < nw * 9; it += 32) {
      const int j = it / 9, t = it - j * 9;
      d.Dbl[12 * (size_t)(lm0 + w0 + j) + t] = sLm[j * kWvLmD + t];
    }
    __syncwarp();
    // ---- phase 4: Y = B Dinv per slot; spill B (Hpl) to HBM, SoA over slots
    const int nslots_w = nw * K;
    for (int it = lane; it < nslots_w * 6; it += 32) {   // one row of a slot's block per lane
      const int sg = it / 6, r = it - sg * 6;
      const double* B = sB + 18 * sg + r * 3;
      const double* Di = sLm + (sg / K) * kWvLmD + 9;
      const double b0 = B[0], b1 = B[1], b2 = B[2];
      double* Y = sY + 18 * sg + r * 3;
#pragma unroll
      for (int c = 0; c < 3; ++c) Y[c] = b0 * Di[c] + b1 * Di[3 + c] + b2 * Di[6 + c];
    }
    {
      // (ncu source view, round 2: written as a double loop over (c, sg) with the address formed per element this
      //  spill was the hottest line of the kernel -- 16 % of the stall samples, 9 % of the instructions.  A wave has at
      //  most 40 slots: two per lane, one pointer bumped by nslots per component, the 18 values of a slot read with
      //  16-byte shared-memory loads)
      const bool v0 = lane < nslots_w, v1 = lane + 32 < nslots_w;
      const double2* b0 = reinterpret_cast<const double2*>(sB + 18 * lane);
      const double2* b1 = reinterpret_cast<const double2*>(sB + 18 * (lane + 32));
      double* w = d.W + (size_t)s_base + (size_t)w0 * K + lane;
      const size_t ns = (size_t)d.nslots;
      if (v0) {
#pragma unroll
        for (int c = 0; c < 9; ++c) {
          const double2 t2 = b0[c];
          w[(size_t)(2 * c) * ns] = t2.x;
          w[(size_t)(2 * c + 1) * ns] = t2.y;
        }
      }
      if (v1) {
#pragma unroll
        for (int c = 0; c < 9; ++c) {
          const double2 t2 = b1[c];
          w[(size_t)(2 * c) * ns + 32] = t2.x;
          w[(size_t)(2 * c + 1) * ns + 32] = t2.y;
        }
      }
    }
    __syncwarp();
    PBW(3);
    // ---- phase 5: accumulate the task's contribution to the reduced system in registers.
    //      The Schur product common to every unit runs branch-free (9 16-byte loads of B_n, FMAs straight
    //      into the accumulators); the direct J^T W J terms of the few special pairs follow in their own loops.
#pragma unroll
    for (int q = 0; q < 7; ++q) {
      const int u = lane + 32 * q;
      if (u < nunits) {
        const int p = u / 6, r = u - p * 6;
        const int pk = sPair[p];
        const int m = (pk >> 5) & 31, n = pk & 31;
        {
          const double* Ym = sY + 18 * m + r * 3;
          const double2* Bn = reinterpret_cast<const double2*>(sB + 18 * n);
#pragma unroll 1
          for (int j = 0; j < nw; ++j, Ym += 18 * K, Bn += 9 * K) {
            const double y0 = -Ym[0], y1 = -Ym[1], y2 = -Ym[2];
            double bb[18];
#pragma unroll
            for (int i = 0; i < 9; ++i) { const double2 t2 = Bn[i]; bb[2 * i] = t2.x; bb[2 * i + 1] = t2.y; }
#pragma unroll
            for (int c = 0; c < 6; ++c)
              acc[q][c] = fma(y2, bb[c * 3 + 2], fma(y1, bb[c * 3 + 1], fma(y0, bb[c * 3], acc[q][c])));
          }
        }
        if (m == n) {
          if (m > 0) {   // own J~p^T J~p of the observer
            const double* Jp = sJp + kWvJ * (m - off);
#pragma unroll 1
            for (int j = 0; j < nw; ++j, Jp += kWvJ * k) {
              const double a0 = Jp[r], a1 = Jp[6 + r], a2 = Jp[12 + r];
#pragma unroll
              for (int c = 0; c < 6; ++c)
                acc[q][c] = fma(a2, Jp[12 + c], fma(a1, Jp[6 + c], fma(a0, Jp[c], acc[q][c])));
            }
          } else {       // anchor diagonal, summed over the edges in phase 2
            const double* A = sLm + 18 + r * 6;
#pragma unroll 1
            for (int j = 0; j < nw; ++j, A += kWvLmD) {
#pragma unroll
              for (int c = 0; c < 6; ++c) acc[q][c] += A[c];
            }
          }
        } else if (m == 0) {   // anchor x observer: J~a^T J~p
          const double* Ja = sJa + kWvJ * (n - off);
          const double* Jp = sJp + kWvJ * (n - off);
#pragma unroll 1
          for (int j = 0; j < nw; ++j, Ja += kWvJ * k, Jp += kWvJ * k) {
            const double a0 = Ja[r], a1 = Ja[6 + r], a2 = Ja[12 + r];
#pragma unroll
            for (int c = 0; c < 6; ++c)
              acc[q][c] = fma(a2, Jp[12 + c], fma(a1, Jp[6 + c], fma(a0, Jp[c], acc[q][c])));
          }
        }
      }
    }
    PBW(4);
    // ---- phase 6: gradients bp = -J^T W e, bc = Y b_l
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int it = lane + 32 * q;
      if (it < K * 6) {
        const int s = it / 6, r = it - s * 6;
        for (int j = 0; j < nw; ++j) {
          double g = 0.;
          if (s > 0) {
            const int l = j * k + s - off;
            g = -(sJp[kWvJ * l + r] * sE[3 * l] + sJp[kWvJ * l + 6 + r] * sE[3 * l + 1] + sJp[kWvJ * l + 12 + r] * sE[3 * l + 2]);
          } else {
            for (int i = i_first; i < k; ++i) {
              const int l = j * k + i;
              g -= sJa[kWvJ * l + r] * sE[3 * l] + sJa[kWvJ * l + 6 + r] * sE[3 * l + 1] + sJa[kWvJ * l + 12 + r] * sE[3 * l + 2];
            }
          }
          const double* Y = sY + 18 * (j * K + s) + r * 3;
          const double* bl = sLm + j * kWvLmD + 6;
          accg[q] += g;
          accc[q] += Y[0] * bl[0] + Y[1] * bl[1] + Y[2] * bl[2];
        }
      }
    }
    __syncwarp();
    PBW(5);
  }
  // ---- flush: one RED.F64 per accumulated element for the whole task
#pragma unroll
  for (int q = 0; q < 7; ++q) {
    const int u = lane + 32 * q;
    if (u < nunits) {
      const int p = u / 6, r = u - p * 6;
      const int pk = sPair[p];
      double* blk = d.S + 36 * (size_t)(pk >> 11);
      const int tr = (pk >> 10) & 1;
#pragma unroll
      for (int c = 0; c < 6; ++c) atomicAdd(blk + (tr ? c * 6 + r : r * 6 + c), acc[q][c]);
    }
  }
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const int it = lane + 32 * q;
    if (it < K * 6) {
      const int s = it / 6, r = it - s * 6;
      const int p = sPose[s];
      atomicAdd(d.bp + 6 * p + r, accg[q]);
      atomicAdd(d.bc + 6 * p + r, accc[q]);
    }
  }
  PBW(6);
  if (prof && lane == 0)
    for (int i = 0; i < 7; ++i) atomicAdd(reinterpret_cast<unsigned long long*>(d.dbg) + 48 + i, (unsigned long long)pacc[i]);
#undef PBW
    __syncwarp();   // the next task reuses this warp's shared-memory tables
  }
  if (persist && lane == 0) {
    const unsigned nwarps_total = (unsigned)n_task_blocks * kWvWarps;
    if (atomicAdd(tctr + 1, 1u) == nwarps_total - 1) { tctr[0] = 0u; tctr[1] = 0u; }   // every other warp has drawn its last ticket
  }
}

// The same kernel with a TEAM of two warps per task: the 32 edge lanes of a wave are warp 0's, every later phase is
// spread over the 64 lanes, and each lane accumulates at most 4 (pair, row) units instead of 7 -- 24 accumulator
// registers instead of 42, so that the kernel fits 128 registers and sixteen warps per SM are resident instead of
// eight (round 1's ncu capture: 255 registers, 10 % of the warp slots active, issue slots 24 % busy).
constexpr int kWvTeams = 4;                 // teams per CTA, two warps each; a team owns what a warp owned before
__device__ __forceinline__ void team_sync(int team) { asm volatile("bar.sync %0, 64;" ::"r"(team + 1) : "memory"); }

__global__ void __launch_bounds__(kWvTeams * 64, 2)
k_build_wave2(BaDev d, int robust, double delta, int n_task_blocks, int prof) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 6, lane = threadIdx.x & 63;   // warp = team, lane = lane of the team (0..63)
  const LmCtl* __restrict__ ctl = d.ctl;
  if (ctl->max_iters > 0 && (ctl->stop || ctl->iter >= ctl->max_iters)) return;   // speculatively enqueued trial: nothing left to do
  const int cur = ctl->cur;
  if ((int)blockIdx.x >= n_task_blocks) {
    // pose-pose constraints (G2oEdgeSE3), one thread each, riding on trailing CTAs of this launch so
    // that their long serial 6x6 arithmetic overlaps the landmark work instead of following it
    const int c = ((int)blockIdx.x - n_task_blocks) * (kWvTeams * 64) + (int)threadIdx.x;
    if (c < d.C) constraint_build(d, d.pose[cur], c);
    return;
  }
  const int task = (int)blockIdx.x * kWvTeams + warp;
  if (task >= d.ntasks) return;
  const double lambda = ctl->lambda;
  double* sm = reinterpret_cast<double*>(smem_raw) + (size_t)warp * kWvDoubles;
  double* sJp = sm;
  double* sJa = sJp + 32 * kWvJ;
  double* sJs = sJa + 32 * kWvJ;
  double* sE = sJs + 32 * 9;
  double* sB = sE + 32 * 3;
  double* sY = sB + kWvSlots * 18;
  double* sLm = sY + kWvSlots * 18;
  double* sChi = sLm + kWvLm * kWvLmD;
  int* si = reinterpret_cast<int*>(reinterpret_cast<double*>(smem_raw) + (size_t)kWvTeams * kWvDoubles) + warp * kWvInts;
  int* sPose = si;
  int* sPair = si + 8;

  const int lm0 = d.task_lm[task], nlm = d.task_cnt[task];
  const int e_base = d.lm_eptr[lm0], k = d.lm_eptr[lm0 + 1] - e_base;
  const int s_base = d.lm_sptr[lm0], K = d.lm_sptr[lm0 + 1] - s_base;
  const int has_self = d.lm_self[lm0];
  const int off = has_self ? 0 : 1;
  const int i_first = has_self ? 1 : 0;
  const int ia = d.lm_anchor[lm0];
  const int fa = d.fixed[ia];
  const int skip_self = d.flags & 1;
  const double* __restrict__ Rt = d.Rt[cur];
  double Ra[9], ta[3];
  load12(Rt, ia, Ra, ta);

  // slot poses and the pair table, ordered by kind so that the 32 units of a round mostly share a
  // code path: anchor-row pairs (0,n), diagonal pairs (m,m), (0,0), then the plain pairs
  if (lane == 0) sPose[0] = ia;
  if (lane < k && !(has_self && lane == 0)) sPose[lane + off] = d.e_pose[e_base + lane];
  team_sync(warp);
  const int npairs = K * (K + 1) / 2;
  for (int p = lane; p < npairs; p += 64) {
    int m, n;
    if (p < K - 1) { m = 0; n = 1 + p; }
    else if (p < 2 * K - 2) { m = n = 1 + p - (K - 1); }
    else if (p == 2 * K - 2) { m = n = 0; }
    else {
      int rem = p - (2 * K - 1);
      m = 1;
      while (rem >= K - 1 - m) { rem -= K - 1 - m; ++m; }
      n = m + 1 + rem;
    }
    const int t = d.tbl[(size_t)sPose[m] * d.P + sPose[n]];
    sPair[p] = ((t >> 1) << 11) | ((t & 1) << 10) | (m << 5) | n;
  }
  team_sync(warp);

  const int nunits = npairs * 6;
  double acc[4][6];
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int c = 0; c < 6; ++c) acc[q][c] = 0.;
  double accg[2] = {0., 0.}, accc[2] = {0., 0.};

  int nw_max = 32 / k;
  if (nw_max > kWvSlots / K) nw_max = kWvSlots / K;
  if (nw_max > kWvLm) nw_max = kWvLm;

  long long pacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, pclk = prof ? clock64() : 0;
#define PBW(i) do { if (prof) { team_sync(warp); const long long c_ = clock64(); pacc[i] += c_ - pclk; pclk = c_; } } while (0)
  PBW(0);
  for (int w0 = 0; w0 < nlm; w0 += nw_max) {
    const int nw = min(nw_max, nlm - w0);
    // ---- phase 1: one lane per edge of the wave
    double chi = 0.;
    if (lane < nw * k) {
      const int j = lane / k, i = lane - j * k;
      const int li = lm0 + w0 + j;
      const int e = e_base + (w0 + j) * k + i;
      const double* __restrict__ psi = d.psi[cur] + 3 * (size_t)li;
      const double p0 = __ldg(psi), p1 = __ldg(psi + 1), p2 = __ldg(psi + 2);
      const double ipz = fast_inv(p2);
      const double xa[3] = {p0 * ipz, p1 * ipz, ipz};
      const int ip = (has_self && i == 0) ? ia : sPose[i + off];
      double* Jp = sJp + kWvJ * lane;
      double* Js = sJs + 9 * lane;
      chi = linearize_edge(d, Rt, e, ip, Ra, ta, xa, ipz, fa, robust, delta, Jp, sJa + kWvJ * lane, Js, sE + 3 * lane);
      if (!(has_self && i == 0)) {   // own Hpl block B = J~p^T J~psi
        double* B = sB + 18 * (j * K + i + off);
#pragma unroll
        for (int r = 0; r < 6; ++r)
#pragma unroll
          for (int c = 0; c < 3; ++c) B[r * 3 + c] = Jp[r] * Js[c] + Jp[6 + r] * Js[3 + c] + Jp[12 + r] * Js[6 + c];
      }
    }
    if (lane < 32) sChi[lane] = chi;
    team_sync(warp);
    PBW(1);
    if (lane < nw) {
      double s = 0.;
      for (int i = 0; i < k; ++i) s += sChi[lane * k + i];
      d.chi_l[lm0 + w0 + lane] = s;
    }
    // ---- phase 2: per-landmark sums over its edges: anchor Hpl block (18), Hll (6), b_l (3), A_aa (21);
    //      one loop per kind so that the lanes of a round share a code path
    for (int it = lane; it < nw * 18; it += 64) {
      const int j = it / 18, t = it - j * 18, r = t / 3, c = t - r * 3;
      const double* Ja = sJa + kWvJ * (j * k + i_first);
      const double* Js = sJs + 9 * (j * k + i_first);
      double s = 0.;
      for (int i = i_first; i < k; ++i, Ja += kWvJ, Js += 9) s += Ja[r] * Js[c] + Ja[6 + r] * Js[3 + c] + Ja[12 + r] * Js[6 + c];
      sB[18 * (j * K) + t] = s;
    }
    for (int it = lane; it < nw * 6; it += 64) {
      const int j = it / 6, u = it - j * 6;
      const int r = u < 3 ? 0 : (u < 5 ? 1 : 2), c = u < 3 ? u : (u < 5 ? u - 2 : 2);
      const double* Js = sJs + 9 * (j * k);
      double s = 0.;
      for (int i = 0; i < k; ++i, Js += 9) s += Js[r] * Js[c] + Js[3 + r] * Js[3 + c] + Js[6 + r] * Js[6 + c];
      sLm[j * kWvLmD + u] = s;
    }
    for (int it = lane; it < nw * 3; it += 64) {
      const int j = it / 3, c = it - j * 3;
      const double* Js = sJs + 9 * (j * k);
      const double* Ee = sE + 3 * (j * k);
      double s = 0.;
      for (int i = 0; i < k; ++i, Js += 9, Ee += 3) s -= Js[c] * Ee[0] + Js[3 + c] * Ee[1] + Js[6 + c] * Ee[2];
      sLm[j * kWvLmD + 6 + c] = s;
    }
    for (int it = lane; it < nw * 21; it += 64) {
      // anchor diagonal: all edges' J~a^T J~a; the self edge keeps g2o's J1^T W J1
      const int j = it / 21, u = it - j * 21;
      const int r = (u >= 1) + (u >= 3) + (u >= 6) + (u >= 10) + (u >= 15), c = u - r * (r + 1) / 2;
      const int i0 = skip_self ? i_first : 0;
      const double* Ja = sJa + kWvJ * (j * k + i0);
      double s = 0.;
      for (int i = i0; i < k; ++i, Ja += kWvJ) s += Ja[r] * Ja[c] + Ja[6 + r] * Ja[6 + c] + Ja[12 + r] * Ja[12 + c];
      sLm[j * kWvLmD + 18 + r * 6 + c] = s;
      sLm[j * kWvLmD + 18 + c * 6 + r] = s;
    }
    team_sync(warp);
    PBW(2);
    // ---- phase 3: (Hll + lambda I)^-1 per landmark; Hll / b_l to HBM for the back-substitution
    if (lane < nw) {
      double Di[9];
      inv3_sym_lambda(sLm + lane * kWvLmD, lambda, Di);
#pragma unroll
      for (int q = 0; q < 9; ++q) sLm[lane * kWvLmD + 9 + q] = Di[q];
    }
    for (int it = lane; it < nw * 9; it += 64) {
      const int j = it / 9, t = it - j * 9;
      d.Dbl[12 * (size_t)(lm0 + w0 + j) + t] = sLm[j * kWvLmD + t];
    }
    team_sync(warp);
    // ---- phase 4: Y = B Dinv per slot; spill B (Hpl) to HBM, SoA over slots
    const int nslots_w = nw * K;
    for (int it = lane; it < nslots_w * 6; it += 64) {   // one row of a slot's block per lane
      const int sg = it / 6, r = it - sg * 6;
      const double* B = sB + 18 * sg + r * 3;
      const double* Di = sLm + (sg / K) * kWvLmD + 9;
      const double b0 = B[0], b1 = B[1], b2 = B[2];
      double* Y = sY + 18 * sg + r * 3;
#pragma unroll
      for (int c = 0; c < 3; ++c) Y[c] = b0 * Di[c] + b1 * Di[3 + c] + b2 * Di[6 + c];
    }
    {
      const size_t s0 = (size_t)s_base + (size_t)w0 * K;
      for (int c = 0; c < 18; ++c)
        for (int sg = lane; sg < nslots_w; sg += 64) d.W[(size_t)c * d.nslots + s0 + sg] = sB[18 * sg + c];
    }
    team_sync(warp);
    PBW(3);
    // ---- phase 5: accumulate the task's contribution to the reduced system in registers.
    //      The Schur product common to every unit runs branch-free (9 16-byte loads of B_n, FMAs straight
    //      into the accumulators); the direct J^T W J terms of the few special pairs follow in their own loops.
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int u = lane + 64 * q;
      if (u < nunits) {
        const int p = u / 6, r = u - p * 6;
        const int pk = sPair[p];
        const int m = (pk >> 5) & 31, n = pk & 31;
        {
          const double* Ym = sY + 18 * m + r * 3;
          const double2* Bn = reinterpret_cast<const double2*>(sB + 18 * n);
#pragma unroll 1
          for (int j = 0; j < nw; ++j, Ym += 18 * K, Bn += 9 * K) {
            const double y0 = -Ym[0], y1 = -Ym[1], y2 = -Ym[2];
            double bb[18];
#pragma unroll
            for (int i = 0; i < 9; ++i) { const double2 t2 = Bn[i]; bb[2 * i] = t2.x; bb[2 * i + 1] = t2.y; }
#pragma unroll
            for (int c = 0; c < 6; ++c)
              acc[q][c] = fma(y2, bb[c * 3 + 2], fma(y1, bb[c * 3 + 1], fma(y0, bb[c * 3], acc[q][c])));
          }
        }
        if (m == n) {
          if (m > 0) {   // own J~p^T J~p of the observer
            const double* Jp = sJp + kWvJ * (m - off);
#pragma unroll 1
            for (int j = 0; j < nw; ++j, Jp += kWvJ * k) {
              const double a0 = Jp[r], a1 = Jp[6 + r], a2 = Jp[12 + r];
#pragma unroll
              for (int c = 0; c < 6; ++c)
                acc[q][c] = fma(a2, Jp[12 + c], fma(a1, Jp[6 + c], fma(a0, Jp[c], acc[q][c])));
            }
          } else {       // anchor diagonal, summed over the edges in phase 2
            const double* A = sLm + 18 + r * 6;
#pragma unroll 1
            for (int j = 0; j < nw; ++j, A += kWvLmD) {
#pragma unroll
              for (int c = 0; c < 6; ++c) acc[q][c] += A[c];
            }
          }
        } else if (m == 0) {   // anchor x observer: J~a^T J~p
          const double* Ja = sJa + kWvJ * (n - off);
          const double* Jp = sJp + kWvJ * (n - off);
#pragma unroll 1
          for (int j = 0; j < nw; ++j, Ja += kWvJ * k, Jp += kWvJ * k) {
            const double a0 = Ja[r], a1 = Ja[6 + r], a2 = Ja[12 + r];
#pragma unroll
            for (int c = 0; c < 6; ++c)
              acc[q][c] = fma(a2, Jp[12 + c], fma(a1, Jp[6 + c], fma(a0, Jp[c], acc[q][c])));
          }
        }
      }
    }
    PBW(4);
    // ---- phase 6: gradients bp = -J^T W e, bc = Y b_l
#pragma unroll
    for (int q = 0; q < 1; ++q) {
      const int it = lane;
      if (it < K * 6) {
        const int s = it / 6, r = it - s * 6;
        for (int j = 0; j < nw; ++j) {
          double g = 0.;
          if (s > 0) {
            const int l = j * k + s - off;
            g = -(sJp[kWvJ * l + r] * sE[3 * l] + sJp[kWvJ * l + 6 + r] * sE[3 * l + 1] + sJp[kWvJ * l + 12 + r] * sE[3 * l + 2]);
          } else {
            for (int i = i_first; i < k; ++i) {
              const int l = j * k + i;
              g -= sJa[kWvJ * l + r] * sE[3 * l] + sJa[kWvJ * l + 6 + r] * sE[3 * l + 1] + sJa[kWvJ * l + 12 + r] * sE[3 * l + 2];
            }
          }
          const double* Y = sY + 18 * (j * K + s) + r * 3;
          const double* bl = sLm + j * kWvLmD + 6;
          accg[q] += g;
          accc[q] += Y[0] * bl[0] + Y[1] * bl[1] + Y[2] * bl[2];
        }
      }
    }
    team_sync(warp);
    PBW(5);
  }
  // ---- flush: one RED.F64 per accumulated element for the whole task
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int u = lane + 64 * q;
    if (u < nunits) {
      const int p = u / 6, r = u - p * 6;
      const int pk = sPair[p];
      double* blk = d.S + 36 * (size_t)(pk >> 11);
      const int tr = (pk >> 10) & 1;
#pragma unroll
      for (int c = 0; c < 6; ++c) atomicAdd(blk + (tr ? c * 6 + r : r * 6 + c), acc[q][c]);
    }
  }
#pragma unroll
  for (int q = 0; q < 1; ++q) {
    const int it = lane;
    if (it < K * 6) {
      const int s = it / 6, r = it - s * 6;
      const int p = sPose[s];
      atomicAdd(d.bp + 6 * p + r, accg[q]);
      atomicAdd(d.bc + 6 * p + r, accc[q]);
    }
  }
  PBW(6);
  if (prof && lane == 0)
    for (int i = 0; i < 7; ++i) atomicAdd(reinterpret_cast<unsigned long long*>(d.dbg) + 48 + i, (unsigned long long)pacc[i]);
#undef PBW
}


// resident CTAs of k_build_wave on the current device (occupancy x SM count), cached per device
static int build_wave_resident_ctas() {
  static int cache[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return 1 << 30;
  if (cache[dev] == 0) {
    int per_sm = 0, sms = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_build_wave, kWvWarps * 32, build_wave_smem_bytes());
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cache[dev] = per_sm > 0 && sms > 0 ? per_sm * sms : 1 << 30;
  }
  return cache[dev];
}

void launch_build_wave(const BaDev& d, int robust, double delta, cudaStream_t st) {
  if (d.ntasks == 0 && d.C == 0) return;
  static const int prof = getenv("SVS_BUILD_TIMING") ? 1 : 0;
  // A/B switch.  Measured on B200 (C2 / C5, per 10 launches): one warp per task 1.20 / 7.08 ms, a team of two warps
  // per task 1.66 / 8.98 ms -- the team halves the accumulators (128 registers, sixteen warps per SM instead of
  // eight) but pays for it with 1.1 KB of spills, five named barriers per wave and an idle second warp while the
  // 32 edge lanes linearise; the one-warp kernel stays the default
  static const int one_warp = getenv("SVS_BUILD_WAVE2") ? 0 : 1;
  // the opt-in above 48 KB of dynamic shared memory is per device: handles may live on several GPUs of one process
  if (one_warp) {
    if (device_needs_smem_optin(0, build_wave_smem_bytes()))
      cudaFuncSetAttribute(k_build_wave, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)build_wave_smem_bytes());
    int task_blocks = (d.ntasks + kWvWarps - 1) / kWvWarps;
    const int c_blocks = (d.C + kWvWarps * 32 - 1) / (kWvWarps * 32);
    // persistent grid (default): as many task CTAs as are resident at once (255 registers and 111 KB of shared memory
    // per CTA: two per SM), tasks drawn from a counter; SVS_BUILD_STATIC=1 keeps one task per warp for the A/B
    static const int persist_on = getenv("SVS_BUILD_STATIC") ? 0 : 1;
    const int resident = build_wave_resident_ctas();
    const int persist = (persist_on && task_blocks > resident) ? 1 : 0;
    if (persist) task_blocks = resident;
    k_build_wave<<<task_blocks + c_blocks, kWvWarps * 32, build_wave_smem_bytes(), st>>>(d, robust, delta, task_blocks, prof, persist);
    return;
  }
  if (device_needs_smem_optin(2, build_wave_smem_bytes()))
    cudaFuncSetAttribute(k_build_wave2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)build_wave_smem_bytes());
  const int task_blocks = (d.ntasks + kWvTeams - 1) / kWvTeams;
  const int c_blocks = (d.C + kWvTeams * 64 - 1) / (kWvTeams * 64);
  k_build_wave2<<<task_blocks + c_blocks, kWvTeams * 64, build_wave_smem_bytes(), st>>>(d, robust, delta, task_blocks, prof);
}

}  // namespace svs
