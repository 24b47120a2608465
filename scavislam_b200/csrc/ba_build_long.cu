// ba_build_long.cu -- k_build_long: the fused linearise + J^T W J + landmark elimination + Schur scatter
// for landmarks with more than 32 slots (anchor + observers).  The reference puts no bound on a track:
// copyDataToG2o adds one edge per frame of Point::vis_set that lies in the double window
// (slam_graph.cpp:1001-1027), and with an outer window of 200 keyframes a slowly moving camera produces
// such tracks.  Same mathematics as k_build (ba_build.cu); nothing is staged in shared memory, so the
// track length is unbounded:
//   pass A  one lane per edge, 32 edges at a time: linearise, J~p^T J~psi -> the landmark's Hpl column in HBM
//           (d.W, needed there for the back-substitution anyway), direct J^T W J terms and gradients
//           straight into the reduced system; Hll, b_l and the anchor's sums are reduced over the warp;
//   pass B  for every slot m: Y_m = B_m (Hll + lambda I)^-1 (all lanes), then one lane per slot n >= m:
//           S_mn -= Y_m B_n^T with B read back from d.W (L1/L2 resident: 144 B per slot).
#include "ba_dev.cuh"
#include "ba_kernels.cuh"

namespace svs {

constexpr int kLongWarps = 4;

__device__ __forceinline__ void add_block(double* __restrict__ blk, int transpose, const double v[36]) {
#pragma unroll
  for (int r = 0; r < 6; ++r)
#pragma unroll
    for (int c = 0; c < 6; ++c) atomicAdd(blk + (transpose ? c * 6 + r : r * 6 + c), v[r * 6 + c]);
}

__global__ void __launch_bounds__(kLongWarps * 32)
k_build_long(BaDev d, int robust, double delta) {
  const LmCtl* __restrict__ ctl = d.ctl;
  if (ctl->max_iters > 0 && (ctl->stop || ctl->iter >= ctl->max_iters)) return;   // speculatively enqueued trial: nothing left to do
  const int cur = ctl->cur;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int idx = (int)blockIdx.x * kLongWarps + warp;
  if (idx >= d.nlong) return;
  const int li = d.long_lm[idx];
  const double lambda = ctl->lambda;
  const int e0 = d.lm_eptr[li], k = d.lm_eptr[li + 1] - e0;
  const int s0 = d.lm_sptr[li], K = d.lm_sptr[li + 1] - s0;
  const int has_self = d.lm_self[li];
  const int off = has_self ? 0 : 1;       // slot of edge i is i + off
  const int ia = d.lm_anchor[li];
  const int fa = d.fixed[ia];
  const int skip_self = d.flags & 1;
  const double* __restrict__ Rt = d.Rt[cur];
  const double* __restrict__ psi = d.psi[cur] + 3 * (size_t)li;
  double Ra[9], ta[3];
  load12(Rt, ia, Ra, ta);
  const double p0 = __ldg(psi), p1 = __ldg(psi + 1), p2 = __ldg(psi + 2);
  const double ipz = 1. / p2;
  const double xa[3] = {p0 * ipz, p1 * ipz, ipz};   // invert_depth (maths_utils.h:66-69)
  const int taa = d.tbl[(size_t)ia * d.P + ia];

  // ---- pass A
  double hll[6] = {0, 0, 0, 0, 0, 0}, bl[3] = {0, 0, 0}, Ba[18], Aaa[21], ga[6] = {0, 0, 0, 0, 0, 0}, chi = 0;
#pragma unroll
  for (int q = 0; q < 18; ++q) Ba[q] = 0;
#pragma unroll
  for (int q = 0; q < 21; ++q) Aaa[q] = 0;
  for (int i0 = 0; i0 < k; i0 += 32) {
    const int i = i0 + lane;
    if (i >= k) continue;
    const int e = e0 + i;
    const int ip = d.e_pose[e];
    const bool self = has_self && i == 0;
    double Jp[18], Ja[18], Js[9], Ee[3];
    chi += linearize_edge(d, Rt, e, ip, Ra, ta, xa, ipz, fa, robust, delta, Jp, Ja, Js, Ee);
    {
      int u = 0;
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = r; c < 3; ++c) hll[u++] += Js[r] * Js[c] + Js[3 + r] * Js[3 + c] + Js[6 + r] * Js[6 + c];
#pragma unroll
      for (int c = 0; c < 3; ++c) bl[c] -= Js[c] * Ee[0] + Js[3 + c] * Ee[1] + Js[6 + c] * Ee[2];
    }
    if (!(self && skip_self)) {   // anchor diagonal: every edge's J~a^T J~a; the self edge keeps g2o's J1^T W J1
      int u = 0;
#pragma unroll
      for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int c = 0; c <= r; ++c) Aaa[u++] += Ja[r] * Ja[c] + Ja[6 + r] * Ja[6 + c] + Ja[12 + r] * Ja[12 + c];
    }
    if (self) continue;   // its Hpl block and gradient are cancelled by its anchor part
    {   // own Hpl block B = J~p^T J~psi -> HBM; anchor's share J~a^T J~psi
      const size_t slot = (size_t)s0 + i + off;
#pragma unroll
      for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          d.W[(size_t)(r * 3 + c) * d.nslots + slot] = Jp[r] * Js[c] + Jp[6 + r] * Js[3 + c] + Jp[12 + r] * Js[6 + c];
          Ba[r * 3 + c] += Ja[r] * Js[c] + Ja[6 + r] * Js[3 + c] + Ja[12 + r] * Js[6 + c];
        }
    }
    double v[36];
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
      for (int c = 0; c < 6; ++c) v[r * 6 + c] = Jp[r] * Jp[c] + Jp[6 + r] * Jp[6 + c] + Jp[12 + r] * Jp[12 + c];
    add_block(d.S + 36 * (size_t)(d.tbl[(size_t)ip * d.P + ip] >> 1), 0, v);
    {
      const int t = d.tbl[(size_t)ia * d.P + ip];   // rows <-> anchor unless transposed
#pragma unroll
      for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int c = 0; c < 6; ++c) v[r * 6 + c] = Ja[r] * Jp[c] + Ja[6 + r] * Jp[6 + c] + Ja[12 + r] * Jp[12 + c];
      add_block(d.S + 36 * (size_t)(t >> 1), t & 1, v);
    }
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      atomicAdd(d.bp + 6 * ip + r, -(Jp[r] * Ee[0] + Jp[6 + r] * Ee[1] + Jp[12 + r] * Ee[2]));
      ga[r] -= Ja[r] * Ee[0] + Ja[6 + r] * Ee[1] + Ja[12 + r] * Ee[2];
    }
  }
  __syncwarp();
  chi = warp_sum(chi);
#pragma unroll
  for (int q = 0; q < 6; ++q) { hll[q] = warp_sum(hll[q]); ga[q] = warp_sum(ga[q]); }
#pragma unroll
  for (int q = 0; q < 3; ++q) bl[q] = warp_sum(bl[q]);
#pragma unroll
  for (int q = 0; q < 18; ++q) Ba[q] = warp_sum(Ba[q]);
#pragma unroll
  for (int q = 0; q < 21; ++q) Aaa[q] = warp_sum(Aaa[q]);
  double* Dbl = d.Dbl + 12 * (size_t)li;
  if (lane == 0) {
    d.chi_l[li] = chi;
#pragma unroll
    for (int q = 0; q < 6; ++q) Dbl[q] = hll[q];
#pragma unroll
    for (int q = 0; q < 3; ++q) Dbl[6 + q] = bl[q];
#pragma unroll
    for (int q = 0; q < 18; ++q) d.W[(size_t)q * d.nslots + s0] = Ba[q];   // slot 0 = anchor
#pragma unroll
    for (int r = 0; r < 6; ++r) atomicAdd(d.bp + 6 * ia + r, ga[r]);
  }
  {   // anchor diagonal block (full 6x6), one element per lane and round
    double* Saa = d.S + 36 * (size_t)(taa >> 1);
    for (int el = lane; el < 36; el += 32) {
      const int r = el / 6, c = el - r * 6;
      const int rr = r >= c ? r : c, cc = r >= c ? c : r;
      double v = 0;
#pragma unroll
      for (int q = 0; q < 21; ++q) v = (q == rr * (rr + 1) / 2 + cc) ? Aaa[q] : v;
      atomicAdd(Saa + el, v);
    }
  }
  if (k == 0) return;
  double hd[6];
#pragma unroll
  for (int q = 0; q < 6; ++q) hd[q] = hll[q];
  double Di[9];
  inv3_sym_lambda(hd, lambda, Di);
  __syncwarp();   // this warp's Hpl blocks in HBM are read back below

  // ---- pass B: Schur complement of the landmark
  for (int m = 0; m < K; ++m) {
    const int pm = m == 0 ? ia : d.e_pose[e0 + m - off];
    double Bm[18], Ym[18];
#pragma unroll
    for (int q = 0; q < 18; ++q) Bm[q] = d.W[(size_t)q * d.nslots + s0 + m];
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) Ym[r * 3 + c] = Bm[r * 3] * Di[c] + Bm[r * 3 + 1] * Di[3 + c] + Bm[r * 3 + 2] * Di[6 + c];
    if (lane < 6)
      atomicAdd(d.bc + 6 * pm + lane, Ym[lane * 3] * bl[0] + Ym[lane * 3 + 1] * bl[1] + Ym[lane * 3 + 2] * bl[2]);
    for (int n = m + lane; n < K; n += 32) {
      const int pn = n == 0 ? ia : d.e_pose[e0 + n - off];
      double Bn[18], v[36];
#pragma unroll
      for (int q = 0; q < 18; ++q) Bn[q] = d.W[(size_t)q * d.nslots + s0 + n];
#pragma unroll
      for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int c = 0; c < 6; ++c)
          v[r * 6 + c] = -(Ym[r * 3] * Bn[c * 3] + Ym[r * 3 + 1] * Bn[c * 3 + 1] + Ym[r * 3 + 2] * Bn[c * 3 + 2]);
      const int t = d.tbl[(size_t)pm * d.P + pn];
      add_block(d.S + 36 * (size_t)(t >> 1), t & 1, v);
    }
  }
}

void launch_build_long(const BaDev& d, int robust, double delta, cudaStream_t st) {
  if (d.nlong == 0) return;
  k_build_long<<<(d.nlong + kLongWarps - 1) / kLongWarps, kLongWarps * 32, 0, st>>>(d, robust, delta);
}

}  // namespace svs
