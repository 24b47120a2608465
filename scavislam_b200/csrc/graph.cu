// graph.cu -- the SLAM map kept on the device and the assembly of a double-window problem from it
// (a "next" row): SlamGraph::copyDataToG2o / copyPosesToG2o / addPointToG2o /
// addObsToG2o (scavislam/slam_graph.cpp:907-1032, slam_graph-impl.cpp:29-126) without the O(E) walk over
// hash maps on the host: vertices (T_me_from_world), points (anchorframe_id, xyz_anchor) and the
// observations (vis_set + feature_table: centre and pyramid level) live in device memory; a window is
// assembled by three kernels (count the visible poses of every active point that lie in the window,
// exclusive scan, emit the edges in the reference's order: active points in list order, vis_set order inside)
// and handed to the bundle adjuster where it lies -- only the edge index triples come back to the host,
// for the structure analysis of svs_ba_set_problem.
//
// Round 2 adds what surrounds the assembly: the choice of the double window from the pose graph
// (computeInitialDoubleWin + computeActivePointsAndExtendOuterWindow, slam_graph.cpp:556-663, and the pair selection
// of copyContraintsToG2o, :938-981) and the growth of the map by one keyframe (addKeyframe, :144-186, 359-421), both
// on the tables where they lie.
#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/svs_b200.h"
#include "internal.cuh"
#include "se3_dev.cuh"
#include "svs_nvtx.hpp"

namespace {

struct MapDev {
  int V, Np;
  const double* pose;       // [V][7]
  const int* anchor;        // [Np] vertex index
  const double* xyz;        // [Np][3]
  const int* vis_ptr;       // [Np+1]
  const int* vis_pose;      // [nnz] vertex index
  const double* center;     // [nnz][3]  (u, v, u_right) at level 0
  const int* level;         // [nnz]
};

// edges of active point l = observations whose pose is in the window (slam_graph.cpp:1001-1027)
__global__ void k_count(MapDev m, const int* __restrict__ win_pos, const int* __restrict__ active, int L, int* __restrict__ cnt,
                        int* __restrict__ bad) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= L) return;
  const int p = active[l];
  if (win_pos[m.anchor[p]] < 0) atomicAdd(bad, 1);   // the anchor frame must be a vertex of the problem
  int c = 0;
  for (int i = m.vis_ptr[p]; i < m.vis_ptr[p + 1]; ++i) c += win_pos[m.vis_pose[i]] >= 0;
  cnt[l] = c;
}

// single-CTA exclusive scan (L is a few 10^4..10^5)
__global__ void k_scan(const int* __restrict__ cnt, int L, int* __restrict__ ptr) {
  __shared__ int sw[32];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < L; base += blockDim.x) {
    const int i = base + threadIdx.x;
    const int v = i < L ? cnt[i] : 0;
    int s = v;
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, s, o); if ((threadIdx.x & 31) >= o) s += t; }
    if ((threadIdx.x & 31) == 31) sw[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
      int w = threadIdx.x < (blockDim.x >> 5) ? sw[threadIdx.x] : 0;
      for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, w, o); if (threadIdx.x >= o) w += t; }
      sw[threadIdx.x] = w;
    }
    __syncthreads();
    const int before = (threadIdx.x >> 5) ? sw[(threadIdx.x >> 5) - 1] : 0;
    if (i < L) ptr[i] = carry + before + s - v;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry += before + s;
    __syncthreads();
  }
  if (threadIdx.x == 0) ptr[L] = carry;
}

__global__ void k_emit(MapDev m, const int* __restrict__ win_pos, const int* __restrict__ active, int L,
                       const int* __restrict__ ptr, int E, int* __restrict__ e_point, int* __restrict__ e_pose,
                       int* __restrict__ e_anchor, double* __restrict__ obs_info, double* __restrict__ psi) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= L) return;
  const int p = active[l];
  {   // addPointToG2o: psi = invert_depth(xyz_anchor) (slam_graph.cpp:907-920, maths_utils.h:66-69)
    const double x = m.xyz[3 * (size_t)p], y = m.xyz[3 * (size_t)p + 1], z = m.xyz[3 * (size_t)p + 2];
    psi[3 * (size_t)l] = x / z; psi[3 * (size_t)l + 1] = y / z; psi[3 * (size_t)l + 2] = 1. / z;
  }
  const int a = win_pos[m.anchor[p]];
  int at = ptr[l];
  for (int i = m.vis_ptr[p]; i < m.vis_ptr[p + 1]; ++i) {
    const int w = win_pos[m.vis_pose[i]];
    if (w < 0) continue;
    e_point[at] = l; e_pose[at] = w; e_anchor[at] = a;
    double* o = obs_info + 3 * (size_t)at;
    o[0] = m.center[3 * (size_t)i]; o[1] = m.center[3 * (size_t)i + 1]; o[2] = m.center[3 * (size_t)i + 2];
    // Lambda = diag(s, s, 0.333^2), s = (2^-level)^2 (slam_graph.cpp:1010-1015)
    const double f = 1. / (double)(1 << m.level[i]), s = f * f;
    double* wq = obs_info + 3 * (size_t)E + 3 * (size_t)at;
    wq[0] = s; wq[1] = s; wq[2] = 0.333 * 0.333;
    ++at;
  }
}

__global__ void k_gather_poses(MapDev m, const int* __restrict__ window, int P, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 7 * P) return;
  out[i] = m.pose[7 * (size_t)window[i / 7] + i % 7];
}

// restoreDataFromG2o (slam_graph.cpp:1037-1058) without leaving the device: vertex and point estimates of the
// optimised window go back into the map (xyz_anchor = invert_depth(psi), maths_utils.h:66-69)
__global__ void k_absorb(double* __restrict__ map_pose, double* __restrict__ map_xyz, const double* __restrict__ pose0,
                         const double* __restrict__ pose1, const double* __restrict__ psi0, const double* __restrict__ psi1,
                         const int* __restrict__ lm_user, const int* __restrict__ cur, const int* __restrict__ window,
                         const int* __restrict__ active, int P, int L) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int c = *cur;
  const double* pose = c ? pose1 : pose0;
  const double* psi = c ? psi1 : psi0;
  if (i < 7 * P) map_pose[7 * (size_t)window[i / 7] + i % 7] = pose[i];
  if (i < L) {
    const double a = psi[3 * (size_t)i], b = psi[3 * (size_t)i + 1], w = psi[3 * (size_t)i + 2];
    double* x = map_xyz + 3 * (size_t)active[lm_user[i]];
    x[0] = a / w; x[1] = b / w; x[2] = 1. / w;
  }
}

// scatter of n records of `width` doubles into rows `index[i]` of a table
__global__ void k_scatter_rows(double* __restrict__ table, const int* __restrict__ index, const double* __restrict__ rows, int n,
                               int width) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * width) return;
  table[(size_t)width * index[i / width] + i % width] = rows[i];
}


// ------------------------------------------------------------------ window selection
struct GraphDev {
  const int* nbr_ptr;       // [V+1]
  const int* nbr_id;        // [nnzN] neighbours of a vertex, strongest first (the order computeInitialDoubleWin pushes them)
  const double* nbr_T;      // [nnzN][7]  T_nbr_from_me of the directed entry (or nullptr)
  const double* nbr_Lam;    // [nnzN][36]
};

// computeInitialDoubleWin (slam_graph.cpp:556-598).  The queue discipline IS the algorithm (a vertex joins when it
// is popped, not when it is pushed), so one thread walks it; a window is a few hundred vertices.
__global__ void k_bfs(int V, GraphDev g, int root, int inner, int dbl, int* __restrict__ type, int* __restrict__ queue, int qcap) {
  if (blockIdx.x || threadIdx.x) return;
  int head = 0, tail = 0, count = 0;
  queue[tail++] = root;
  while (count < dbl && head < tail) {
    const int v = queue[head++];
    if (type[v]) continue;                       // "Avoid cycles!"
    type[v] = count < inner ? 1 : 2;
    ++count;
    for (int i = g.nbr_ptr[v]; i < g.nbr_ptr[v + 1] && tail < qcap; ++i) queue[tail++] = g.nbr_id[i];
  }
}

__device__ __forceinline__ bool has_edge(const GraphDev& g, int a, int b) {
  for (int i = g.nbr_ptr[a]; i < g.nbr_ptr[a + 1]; ++i)
    if (g.nbr_id[i] == b) return true;
  for (int i = g.nbr_ptr[b]; i < g.nbr_ptr[b + 1]; ++i)
    if (g.nbr_id[i] == a) return true;
  return false;
}

// computeActivePointsAndExtendOuterWindow (slam_graph.cpp:600-663), one thread per map point: active when an INNER
// frame sees it and its anchor frame is in the window, or that inner frame has an edge to the anchor frame (which then
// joins the outer window: ext[anchor] = 1)
__global__ void k_active(MapDev m, GraphDev g, const int* __restrict__ type, int* __restrict__ ext, int* __restrict__ active) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= m.Np) return;
  const int a = m.anchor[p];
  const bool inwin = type[a] != 0;
  int act = 0;
  for (int i = m.vis_ptr[p]; i < m.vis_ptr[p + 1] && !act; ++i) {
    const int f = m.vis_pose[i];
    if (type[f] != 1) continue;
    if (inwin) act = 1;
    else if (has_edge(g, f, a)) { act = 1; ext[a] = 1; }
  }
  active[p] = act;
}

__global__ void k_window_flags(int V, const int* __restrict__ type, const int* __restrict__ ext, int* __restrict__ wtype,
                               int* __restrict__ flag) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  const int t = type[v] ? type[v] : (ext[v] ? 2 : 0);
  wtype[v] = t;
  flag[v] = t != 0;
}

// out[ptr[i]] = i for flagged i (ascending: the order of a std::map / of the sorted point ids); pos[i] = ptr[i] or -1
__global__ void k_compact(int n, const int* __restrict__ flag, const int* __restrict__ ptr, int* __restrict__ out, int* __restrict__ pos) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (flag[i]) out[ptr[i]] = i;
  if (pos) pos[i] = flag[i] ? ptr[i] : -1;
}

// the pair loop of copyContraintsToG2o (slam_graph.cpp:938-981) over the directed neighbour entries
__device__ __forceinline__ bool pair_selected(const int* wtype, int a, int b) {
  return b != a && wtype[a] && wtype[b] && (wtype[a] == 2 || wtype[b] == 2);
}
__global__ void k_pair_count(int V, GraphDev g, const int* __restrict__ wtype, int* __restrict__ cnt) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= V) return;
  int c = 0;
  for (int i = g.nbr_ptr[a]; i < g.nbr_ptr[a + 1]; ++i) c += pair_selected(wtype, a, g.nbr_id[i]);
  cnt[a] = c;
}
__global__ void k_pair_emit(int V, GraphDev g, const int* __restrict__ wtype, const int* __restrict__ win_pos,
                            const int* __restrict__ ptr, int* __restrict__ c_i, int* __restrict__ c_j, double* __restrict__ c_T,
                            double* __restrict__ c_Lam) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= V) return;
  for (int i = g.nbr_ptr[a]; i < g.nbr_ptr[a + 1]; ++i) {
    const int b = g.nbr_id[i];
    if (!pair_selected(wtype, a, b)) continue;
    int rank = 0;   // ascending id2 inside id1 (the inner loop of the reference runs over a std::map)
    for (int k = g.nbr_ptr[a]; k < g.nbr_ptr[a + 1]; ++k) rank += pair_selected(wtype, a, g.nbr_id[k]) && g.nbr_id[k] < b;
    const int at = ptr[a] + rank;
    c_i[at] = win_pos[a]; c_j[at] = win_pos[b];
    for (int q = 0; q < 7; ++q) c_T[7 * (size_t)at + q] = g.nbr_T ? g.nbr_T[7 * (size_t)i + q] : (q == 3 ? 1. : 0.);
    for (int q = 0; q < 36; ++q) c_Lam[36 * (size_t)at + q] = g.nbr_Lam ? g.nbr_Lam[36 * (size_t)i + q] : 0.;
  }
}

// ------------------------------------------------------------------ growth by one keyframe
// T_new = T_newkey_from_oldkey * T_oldkey_from_world (slam_graph.cpp:153-156)
__global__ void k_new_pose(const double* __restrict__ old_pose, int oldkey, const double* __restrict__ T_rel, double* __restrict__ out) {
  if (blockIdx.x || threadIdx.x) return;
  double A[7], B[7], AB[7];
  for (int q = 0; q < 7; ++q) { A[q] = T_rel[q]; B[q] = old_pose[7 * (size_t)oldkey + q]; }
  svs::se3_mul(A, B, AB);
  for (int q = 0; q < 7; ++q) out[q] = AB[q];
}
__global__ void k_grow_count(int Np_old, int Np_new, const int* __restrict__ old_ptr, const int* __restrict__ add_idx, int* __restrict__ cnt) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= Np_new) return;
  cnt[p] = p < Np_old ? old_ptr[p + 1] - old_ptr[p] + (add_idx[p] >= 0) : 2;
}
__global__ void k_mark_tracks(int n, const int* __restrict__ track_point, int* __restrict__ add_idx) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) add_idx[track_point[t]] = t;
}
// one thread per point moves its observations to their new place and appends the new keyframe's (its id is the
// largest, so it is the last of the point's ascending list); a new point is seen by its anchor frame and the keyframe
__global__ void k_grow_move(MapDev m, int Np_new, int newkey, const int* __restrict__ add_idx, const int* __restrict__ new_ptr,
                            const double* __restrict__ track_center, const int* __restrict__ track_level,
                            const int* __restrict__ new_anchor, const double* __restrict__ new_anchor_center,
                            const int* __restrict__ new_anchor_level, const double* __restrict__ new_center,
                            const int* __restrict__ new_level, int* __restrict__ vis_pose, double* __restrict__ center,
                            int* __restrict__ level) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= Np_new) return;
  int at = new_ptr[p];
  auto put = [&](int v, const double* c, int l) {
    vis_pose[at] = v; level[at] = l;
    center[3 * (size_t)at] = c[0]; center[3 * (size_t)at + 1] = c[1]; center[3 * (size_t)at + 2] = c[2];
    ++at;
  };
  if (p < m.Np) {
    for (int i = m.vis_ptr[p]; i < m.vis_ptr[p + 1]; ++i) put(m.vis_pose[i], m.center + 3 * (size_t)i, m.level[i]);
    const int t = add_idx[p];
    if (t >= 0) put(newkey, track_center + 3 * (size_t)t, track_level[t]);
  } else {
    const int q = p - m.Np;
    put(new_anchor[q], new_anchor_center + 3 * (size_t)q, new_anchor_level[q]);
    put(newkey, new_center + 3 * (size_t)q, new_level[q]);
  }
}

}  // namespace

struct svs_map {
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  int V = 0, Np = 0, nnz = 0;
  char* d_map = nullptr; size_t map_cap = 0;
  MapDev m{};
  char* d_work = nullptr; size_t work_cap = 0;
  std::vector<int> h_ep, h_es, h_ea, h_winpos;
  std::vector<double> h_pose, h_psi;
  const double* d_oi_last = nullptr;   // [E][3] observations, [E][3] weights of the last assembly
  int last_E = 0;
  const int* d_win_last = nullptr; const int* d_act_last = nullptr; int last_P = 0, last_L = 0;   // the last assembled window
  char* d_upd = nullptr; size_t upd_cap = 0;   // staging of svs_map_update_*
  char* d_graph = nullptr; size_t graph_cap = 0; GraphDev g{}; int nnzN = 0;   // svs_map_set_graph
  char* d_sel = nullptr; size_t sel_cap = 0;   // work buffers of svs_map_select_window
};

#define GCK(call)                                                       \
  do {                                                                  \
    cudaError_t e_ = (call);                                            \
    if (e_ != cudaSuccess) {                                            \
      h->err = std::string(#call) + ": " + cudaGetErrorString(e_);      \
      return SVS_ERR_CUDA;                                              \
    }                                                                   \
  } while (0)

static size_t al256(size_t x) { return (x + 255) / 256 * 256; }

struct MapLayout { size_t o_pose, o_anch, o_xyz, o_vptr, o_vpose, o_cen, o_lvl, total; };
static MapLayout map_layout(int V, int Np, int nnz) {
  MapLayout lo{};
  size_t off = 0;
  lo.o_pose = off; off += al256(sizeof(double) * 7 * (size_t)V);
  lo.o_anch = off; off += al256(sizeof(int) * (size_t)std::max(Np, 1));
  lo.o_xyz = off; off += al256(sizeof(double) * 3 * (size_t)std::max(Np, 1));
  lo.o_vptr = off; off += al256(sizeof(int) * ((size_t)Np + 1));
  lo.o_vpose = off; off += al256(sizeof(int) * (size_t)std::max(nnz, 1));
  lo.o_cen = off; off += al256(sizeof(double) * 3 * (size_t)std::max(nnz, 1));
  lo.o_lvl = off; off += al256(sizeof(int) * (size_t)std::max(nnz, 1));
  lo.total = off;
  return lo;
}
static void map_bind(svs_map* h, char* B, const MapLayout& lo, int V, int Np, int nnz) {
  h->V = V; h->Np = Np; h->nnz = nnz;
  h->m.V = V; h->m.Np = Np;
  h->m.pose = reinterpret_cast<const double*>(B + lo.o_pose); h->m.anchor = reinterpret_cast<const int*>(B + lo.o_anch);
  h->m.xyz = reinterpret_cast<const double*>(B + lo.o_xyz); h->m.vis_ptr = reinterpret_cast<const int*>(B + lo.o_vptr);
  h->m.vis_pose = reinterpret_cast<const int*>(B + lo.o_vpose); h->m.center = reinterpret_cast<const double*>(B + lo.o_cen);
  h->m.level = reinterpret_cast<const int*>(B + lo.o_lvl);
}

extern "C" {

int svs_map_create(int device, svs_map** out) {
  if (!out) return SVS_ERR_INVALID;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) return SVS_ERR_NOGPU;
  svs_map* h = new svs_map();
  if (device < 0) cudaGetDevice(&device);
  h->device = device;
  if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete h;
    return SVS_ERR_CUDA;
  }
  *out = h;
  return SVS_OK;
}

void svs_map_destroy(svs_map* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  cudaFree(h->d_map); cudaFree(h->d_work); cudaFree(h->d_upd); cudaFree(h->d_graph); cudaFree(h->d_sel);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

const char* svs_map_last_error(const svs_map* h) { return h ? h->err.c_str() : "null handle"; }

int svs_map_set(svs_map* h, int V, const double* T_me_from_world, int Np, const int* point_anchor, const double* xyz_anchor,
                const int* vis_ptr, const int* vis_pose, const double* feat_center, const int* feat_level) {
  if (!h || V <= 0 || Np < 0 || !T_me_from_world || (Np && (!point_anchor || !xyz_anchor || !vis_ptr))) return SVS_ERR_INVALID;
  const int nnz = Np ? vis_ptr[Np] : 0;
  if (nnz < 0 || (nnz && (!vis_pose || !feat_center || !feat_level))) return SVS_ERR_INVALID;
  for (int p = 0; p < Np; ++p) {
    if (point_anchor[p] < 0 || point_anchor[p] >= V) { h->err = "point anchored in a vertex outside [0, V)"; return SVS_ERR_INVALID; }
    if (vis_ptr[p + 1] < vis_ptr[p]) { h->err = "vis_ptr not ascending"; return SVS_ERR_INVALID; }
  }
  for (int i = 0; i < nnz; ++i)
    if (vis_pose[i] < 0 || vis_pose[i] >= V || feat_level[i] < 0 || feat_level[i] > 30) {
      h->err = "observation names a vertex outside [0, V) or a bad pyramid level";
      return SVS_ERR_INVALID;
    }
  cudaSetDevice(h->device);
  const MapLayout lo = map_layout(V, Np, nnz);
  const size_t off = lo.total;
  const size_t o_pose = lo.o_pose, o_anch = lo.o_anch, o_xyz = lo.o_xyz, o_vptr = lo.o_vptr, o_vpose = lo.o_vpose, o_cen = lo.o_cen,
               o_lvl = lo.o_lvl;
  GCK(cudaStreamSynchronize(h->stream));
  if (off > h->map_cap) {
    cudaFree(h->d_map); h->d_map = nullptr; h->map_cap = 0;
    GCK(cudaMalloc(&h->d_map, off + off / 4));
    h->map_cap = off + off / 4;
  }
  char* B = h->d_map;
  GCK(cudaMemcpyAsync(B + o_pose, T_me_from_world, sizeof(double) * 7 * (size_t)V, cudaMemcpyHostToDevice, h->stream));
  if (Np) {
    GCK(cudaMemcpyAsync(B + o_anch, point_anchor, sizeof(int) * (size_t)Np, cudaMemcpyHostToDevice, h->stream));
    GCK(cudaMemcpyAsync(B + o_xyz, xyz_anchor, sizeof(double) * 3 * (size_t)Np, cudaMemcpyHostToDevice, h->stream));
    GCK(cudaMemcpyAsync(B + o_vptr, vis_ptr, sizeof(int) * ((size_t)Np + 1), cudaMemcpyHostToDevice, h->stream));
  }
  if (nnz) {
    GCK(cudaMemcpyAsync(B + o_vpose, vis_pose, sizeof(int) * (size_t)nnz, cudaMemcpyHostToDevice, h->stream));
    GCK(cudaMemcpyAsync(B + o_cen, feat_center, sizeof(double) * 3 * (size_t)nnz, cudaMemcpyHostToDevice, h->stream));
    GCK(cudaMemcpyAsync(B + o_lvl, feat_level, sizeof(int) * (size_t)nnz, cudaMemcpyHostToDevice, h->stream));
  }
  GCK(cudaStreamSynchronize(h->stream));
  map_bind(h, B, lo, V, Np, nnz);
  h->g = GraphDev{}; h->nnzN = 0;   // a new map: its pose graph comes with svs_map_set_graph
  return SVS_OK;
}

// n records of `width` doubles go to the device in ONE copy and are scattered there
static int map_scatter(svs_map* h, double* table, int rows_in_table, int n, const int* index, const double* rows, int width,
                       const char* what) {
  for (int i = 0; i < n; ++i)
    if (index[i] < 0 || index[i] >= rows_in_table) { h->err = std::string(what) + " index out of range"; return SVS_ERR_INVALID; }
  if (n == 0) return SVS_OK;
  cudaSetDevice(h->device);
  const size_t ib = al256(sizeof(int) * (size_t)n), rb = sizeof(double) * (size_t)n * width;
  if (ib + rb > h->upd_cap) {
    GCK(cudaStreamSynchronize(h->stream));
    cudaFree(h->d_upd); h->d_upd = nullptr; h->upd_cap = 0;
    GCK(cudaMalloc(&h->d_upd, 2 * (ib + rb)));
    h->upd_cap = 2 * (ib + rb);
  }
  GCK(cudaMemcpyAsync(h->d_upd, index, sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, h->stream));
  GCK(cudaMemcpyAsync(h->d_upd + ib, rows, rb, cudaMemcpyHostToDevice, h->stream));
  k_scatter_rows<<<(n * width + 255) / 256, 256, 0, h->stream>>>(table, reinterpret_cast<const int*>(h->d_upd),
                                                                 reinterpret_cast<const double*>(h->d_upd + ib), n, width);
  GCK(cudaGetLastError());
  GCK(cudaStreamSynchronize(h->stream));   // the caller's arrays may go away
  return SVS_OK;
}

int svs_map_update_poses(svs_map* h, int n, const int* vertex, const double* T_me_from_world) {
  if (!h || n < 0 || (n && (!vertex || !T_me_from_world)) || !h->d_map) return SVS_ERR_INVALID;
  return map_scatter(h, const_cast<double*>(h->m.pose), h->V, n, vertex, T_me_from_world, 7, "vertex");
}

int svs_map_update_points(svs_map* h, int n, const int* point, const double* xyz_anchor) {
  if (!h || n < 0 || (n && (!point || !xyz_anchor)) || !h->d_map) return SVS_ERR_INVALID;
  return map_scatter(h, const_cast<double*>(h->m.xyz), h->Np, n, point, xyz_anchor, 3, "point");
}

int svs_map_get(svs_map* h, double* T_me_from_world, double* xyz_anchor) {
  if (!h || !h->d_map) return SVS_ERR_INVALID;
  cudaSetDevice(h->device);
  if (T_me_from_world) GCK(cudaMemcpyAsync(T_me_from_world, h->m.pose, sizeof(double) * 7 * (size_t)h->V, cudaMemcpyDeviceToHost, h->stream));
  if (xyz_anchor && h->Np) GCK(cudaMemcpyAsync(xyz_anchor, h->m.xyz, sizeof(double) * 3 * (size_t)h->Np, cudaMemcpyDeviceToHost, h->stream));
  GCK(cudaStreamSynchronize(h->stream));
  return SVS_OK;
}

// SlamGraph::restoreDataFromG2o (slam_graph.cpp:1037-1058): the optimised window of `ba` (loaded with
// svs_ba_set_problem_from_map from THIS map) goes back into the map, device to device
int svs_map_absorb(svs_map* h, svs_ba* ba) {
  svs::NvtxRange nvtx_("restoreDataFromG2o");
  if (!h || !ba || !h->d_map || !h->d_win_last) return SVS_ERR_INVALID;
  if (svs::ba_device(ba) != h->device) { h->err = "map and bundle adjuster live on different devices"; return SVS_ERR_INVALID; }
  const double* const* pose; const double* const* psi; const int* lm_user; const int* cur; cudaStream_t st; int P, L;
  if (svs::ba_state_on_device(ba, &pose, &psi, &lm_user, &cur, &st, &P, &L) != SVS_OK || P != h->last_P || L != h->last_L) {
    h->err = "the bundle adjuster does not hold the window this map assembled last";
    return SVS_ERR_STATE;
  }
  cudaSetDevice(h->device);
  GCK(cudaStreamSynchronize(h->stream));
  const int n = std::max(7 * P, L);
  k_absorb<<<(n + 255) / 256, 256, 0, st>>>(const_cast<double*>(h->m.pose), const_cast<double*>(h->m.xyz), pose[0], pose[1], psi[0], psi[1],
                                            lm_user, cur, h->d_win_last, h->d_act_last, P, L);
  GCK(cudaGetLastError());
  GCK(cudaStreamSynchronize(st));
  return SVS_OK;
}

int svs_ba_set_problem_from_map(svs_ba* ba, svs_map* h, int P, const int* window_vertex, const unsigned char* fixed, int L,
                                const int* active_point, int C, const int* c_i, const int* c_j, const double* c_T,
                                const double* c_Lambda, const svs_cam* cam, int* num_edges) {
  svs::NvtxRange nvtx_("copyDataToG2o");
  if (!ba || !h || P <= 0 || L < 0 || C < 0 || !window_vertex || (L && !active_point) || !cam || !h->d_map) return SVS_ERR_INVALID;
  if (svs::ba_device(ba) != h->device) { h->err = "map and bundle adjuster live on different devices"; return SVS_ERR_INVALID; }
  // window position of every vertex (-1 = outside the double window)
  h->h_winpos.assign(h->V, -1);
  for (int i = 0; i < P; ++i) {
    const int v = window_vertex[i];
    if (v < 0 || v >= h->V || h->h_winpos[v] >= 0) { h->err = "window names a vertex twice or outside [0, V)"; return SVS_ERR_INVALID; }
    h->h_winpos[v] = i;
  }
  for (int l = 0; l < L; ++l)
    if (active_point[l] < 0 || active_point[l] >= h->Np) { h->err = "active point outside [0, Np)"; return SVS_ERR_INVALID; }
  cudaSetDevice(h->device);
  // work arena: win_pos, window, active, cnt, ptr, bad | poses, psi | edges (sized for every observation of the map)
  size_t off = 0;
  const size_t o_wp = off; off += al256(sizeof(int) * (size_t)h->V);
  const size_t o_win = off; off += al256(sizeof(int) * (size_t)P);
  const size_t o_act = off; off += al256(sizeof(int) * (size_t)std::max(L, 1));
  const size_t o_cnt = off; off += al256(sizeof(int) * (size_t)std::max(L, 1));
  const size_t o_ptr = off; off += al256(sizeof(int) * ((size_t)L + 1));
  const size_t o_bad = off; off += 256;
  const size_t o_pose = off; off += al256(sizeof(double) * 7 * (size_t)P);
  const size_t o_psi = off; off += al256(sizeof(double) * 3 * (size_t)std::max(L, 1));
  const size_t emax = (size_t)std::max(h->nnz, 1);
  const size_t o_ep = off; off += al256(sizeof(int) * emax);
  const size_t o_es = off; off += al256(sizeof(int) * emax);
  const size_t o_ea = off; off += al256(sizeof(int) * emax);
  const size_t o_oi = off; off += al256(sizeof(double) * 6 * emax);
  GCK(cudaStreamSynchronize(h->stream));
  if (off > h->work_cap) {
    cudaFree(h->d_work); h->d_work = nullptr; h->work_cap = 0;
    GCK(cudaMalloc(&h->d_work, off + off / 4));
    h->work_cap = off + off / 4;
  }
  char* W = h->d_work;
  int* d_wp = reinterpret_cast<int*>(W + o_wp); int* d_win = reinterpret_cast<int*>(W + o_win);
  int* d_act = reinterpret_cast<int*>(W + o_act); int* d_cnt = reinterpret_cast<int*>(W + o_cnt);
  int* d_ptr = reinterpret_cast<int*>(W + o_ptr); int* d_bad = reinterpret_cast<int*>(W + o_bad);
  double* d_pose = reinterpret_cast<double*>(W + o_pose); double* d_psi = reinterpret_cast<double*>(W + o_psi);
  int* d_ep = reinterpret_cast<int*>(W + o_ep); int* d_es = reinterpret_cast<int*>(W + o_es); int* d_ea = reinterpret_cast<int*>(W + o_ea);
  double* d_oi = reinterpret_cast<double*>(W + o_oi);
  GCK(cudaMemcpyAsync(d_wp, h->h_winpos.data(), sizeof(int) * (size_t)h->V, cudaMemcpyHostToDevice, h->stream));
  GCK(cudaMemcpyAsync(d_win, window_vertex, sizeof(int) * (size_t)P, cudaMemcpyHostToDevice, h->stream));
  if (L) GCK(cudaMemcpyAsync(d_act, active_point, sizeof(int) * (size_t)L, cudaMemcpyHostToDevice, h->stream));
  GCK(cudaMemsetAsync(d_bad, 0, sizeof(int), h->stream));
  int E = 0, bad = 0;
  if (L) {
    k_count<<<(L + 255) / 256, 256, 0, h->stream>>>(h->m, d_wp, d_act, L, d_cnt, d_bad);
    k_scan<<<1, 1024, 0, h->stream>>>(d_cnt, L, d_ptr);
    GCK(cudaMemcpyAsync(&E, d_ptr + L, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    GCK(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    GCK(cudaStreamSynchronize(h->stream));
    if (bad) { h->err = "an active point is anchored in a frame outside the window"; return SVS_ERR_INVALID; }
    // obs_info = [E][3] observations followed by [E][3] weights: the emit kernel needs E for the second half
    k_emit<<<(L + 255) / 256, 256, 0, h->stream>>>(h->m, d_wp, d_act, L, d_ptr, E, d_ep, d_es, d_ea, d_oi, d_psi);
  }
  k_gather_poses<<<(7 * P + 255) / 256, 256, 0, h->stream>>>(h->m, d_win, P, d_pose);
  GCK(cudaGetLastError());
  // only the index triples, psi and the window's poses go back: structure analysis and initial state of the BA handle
  h->h_ep.resize(std::max(E, 1)); h->h_es.resize(std::max(E, 1)); h->h_ea.resize(std::max(E, 1));
  h->h_pose.resize(7 * (size_t)P); h->h_psi.resize(3 * (size_t)std::max(L, 1));
  if (E) {
    GCK(cudaMemcpyAsync(h->h_ep.data(), d_ep, sizeof(int) * (size_t)E, cudaMemcpyDeviceToHost, h->stream));
    GCK(cudaMemcpyAsync(h->h_es.data(), d_es, sizeof(int) * (size_t)E, cudaMemcpyDeviceToHost, h->stream));
    GCK(cudaMemcpyAsync(h->h_ea.data(), d_ea, sizeof(int) * (size_t)E, cudaMemcpyDeviceToHost, h->stream));
  }
  if (L) GCK(cudaMemcpyAsync(h->h_psi.data(), d_psi, sizeof(double) * 3 * (size_t)L, cudaMemcpyDeviceToHost, h->stream));
  GCK(cudaMemcpyAsync(h->h_pose.data(), d_pose, sizeof(double) * 7 * (size_t)P, cudaMemcpyDeviceToHost, h->stream));
  GCK(cudaStreamSynchronize(h->stream));
  if (num_edges) *num_edges = E;
  h->d_oi_last = d_oi; h->last_E = E;
  h->d_win_last = d_win; h->d_act_last = d_act; h->last_P = P; h->last_L = L;
  const int rc = svs::ba_set_problem_device_obs(ba, P, h->h_pose.data(), fixed, L, h->h_psi.data(), E, h->h_ep.data(), h->h_es.data(),
                                                h->h_ea.data(), d_oi, C, c_i, c_j, c_T, c_Lambda, cam);
  if (rc != SVS_OK) h->err = std::string("svs_ba_set_problem: ") + svs_last_error(ba);
  return rc;
}


// ------------------------------------------------------------------ pose graph, window selection, growth

int svs_map_set_graph(svs_map* h, const int* nbr_ptr, const int* nbr_id, const double* nbr_T, const double* nbr_Lambda) {
  if (!h || !h->d_map || !nbr_ptr) return SVS_ERR_INVALID;
  const int V = h->V, nn = nbr_ptr[V];
  if (nbr_ptr[0] != 0 || nn < 0 || (nn && !nbr_id) || ((nbr_T == nullptr) != (nbr_Lambda == nullptr))) return SVS_ERR_INVALID;
  for (int v = 0; v < V; ++v)
    if (nbr_ptr[v + 1] < nbr_ptr[v]) { h->err = "nbr_ptr not ascending"; return SVS_ERR_INVALID; }
  for (int i = 0; i < nn; ++i)
    if (nbr_id[i] < 0 || nbr_id[i] >= V) { h->err = "neighbour outside [0, V)"; return SVS_ERR_INVALID; }
  cudaSetDevice(h->device);
  size_t off = 0;
  const size_t o_ptr = off; off += al256(sizeof(int) * ((size_t)V + 1));
  const size_t o_id = off; off += al256(sizeof(int) * (size_t)std::max(nn, 1));
  const size_t o_T = off; off += al256(sizeof(double) * 7 * (size_t)std::max(nn, 1));
  const size_t o_L = off; off += al256(sizeof(double) * 36 * (size_t)std::max(nn, 1));
  GCK(cudaStreamSynchronize(h->stream));
  if (off > h->graph_cap) {
    cudaFree(h->d_graph); h->d_graph = nullptr; h->graph_cap = 0;
    GCK(cudaMalloc(&h->d_graph, off + off / 4));
    h->graph_cap = off + off / 4;
  }
  char* B = h->d_graph;
  GCK(cudaMemcpyAsync(B + o_ptr, nbr_ptr, sizeof(int) * ((size_t)V + 1), cudaMemcpyHostToDevice, h->stream));
  if (nn) {
    GCK(cudaMemcpyAsync(B + o_id, nbr_id, sizeof(int) * (size_t)nn, cudaMemcpyHostToDevice, h->stream));
    if (nbr_T) {
      GCK(cudaMemcpyAsync(B + o_T, nbr_T, sizeof(double) * 7 * (size_t)nn, cudaMemcpyHostToDevice, h->stream));
      GCK(cudaMemcpyAsync(B + o_L, nbr_Lambda, sizeof(double) * 36 * (size_t)nn, cudaMemcpyHostToDevice, h->stream));
    }
  }
  GCK(cudaStreamSynchronize(h->stream));
  h->g.nbr_ptr = reinterpret_cast<const int*>(B + o_ptr); h->g.nbr_id = reinterpret_cast<const int*>(B + o_id);
  h->g.nbr_T = nbr_T ? reinterpret_cast<const double*>(B + o_T) : nullptr;
  h->g.nbr_Lam = nbr_T ? reinterpret_cast<const double*>(B + o_L) : nullptr;
  h->nnzN = nn;
  return SVS_OK;
}

int svs_map_select_window(svs_map* h, int root, int inner_window_size, int double_window_size, int cap_P, int* P_out,
                          int* window_vertex, unsigned char* inner, int cap_L, int* L_out, int* active_point, int cap_C,
                          int* C_out, int* c_i, int* c_j, double* c_T, double* c_Lambda) {
  if (!h || !h->d_map || !P_out || !window_vertex || !L_out || (cap_L && !active_point) || cap_P <= 0 || cap_L < 0 || cap_C < 0)
    return SVS_ERR_INVALID;
  if (!h->g.nbr_ptr) { h->err = "svs_map_set_graph has not been called for this map"; return SVS_ERR_STATE; }
  const int V = h->V, Np = h->Np, nn = h->nnzN;
  if (root < 0 || root >= V || inner_window_size < 0 || inner_window_size >= double_window_size) {   // assert at slam_graph.cpp:563
    h->err = "root outside [0, V) or inner_window_size >= double_window_size";
    return SVS_ERR_INVALID;
  }
  cudaSetDevice(h->device);
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t o = off; off += al256(bytes); return o; };
  const size_t o_type = take(sizeof(int) * V), o_ext = take(sizeof(int) * V), o_wtype = take(sizeof(int) * V);
  const size_t o_flag = take(sizeof(int) * V), o_ptrV = take(sizeof(int) * ((size_t)V + 1)), o_win = take(sizeof(int) * V);
  const size_t o_pos = take(sizeof(int) * V), o_act = take(sizeof(int) * (size_t)std::max(Np, 1));
  const size_t o_ptrP = take(sizeof(int) * ((size_t)Np + 1)), o_actl = take(sizeof(int) * (size_t)std::max(Np, 1));
  const size_t o_q = take(sizeof(int) * ((size_t)nn + 1)), o_cc = take(sizeof(int) * V), o_cp = take(sizeof(int) * ((size_t)V + 1));
  const size_t o_ci = take(sizeof(int) * (size_t)std::max(nn, 1)), o_cj = take(sizeof(int) * (size_t)std::max(nn, 1));
  const size_t o_cT = take(sizeof(double) * 7 * (size_t)std::max(nn, 1)), o_cL = take(sizeof(double) * 36 * (size_t)std::max(nn, 1));
  GCK(cudaStreamSynchronize(h->stream));
  if (off > h->sel_cap) {
    cudaFree(h->d_sel); h->d_sel = nullptr; h->sel_cap = 0;
    GCK(cudaMalloc(&h->d_sel, off + off / 4));
    h->sel_cap = off + off / 4;
  }
  char* W = h->d_sel;
  auto I = [&](size_t o) { return reinterpret_cast<int*>(W + o); };
  GCK(cudaMemsetAsync(I(o_type), 0, sizeof(int) * V, h->stream));
  GCK(cudaMemsetAsync(I(o_ext), 0, sizeof(int) * V, h->stream));
  const int bV = (V + 255) / 256, bP = (std::max(Np, 1) + 255) / 256;
  k_bfs<<<1, 32, 0, h->stream>>>(V, h->g, root, inner_window_size, double_window_size, I(o_type), I(o_q), nn + 1);
  if (Np) k_active<<<bP, 256, 0, h->stream>>>(h->m, h->g, I(o_type), I(o_ext), I(o_act));
  k_window_flags<<<bV, 256, 0, h->stream>>>(V, I(o_type), I(o_ext), I(o_wtype), I(o_flag));
  k_scan<<<1, 1024, 0, h->stream>>>(I(o_flag), V, I(o_ptrV));
  k_compact<<<bV, 256, 0, h->stream>>>(V, I(o_flag), I(o_ptrV), I(o_win), I(o_pos));
  if (Np) {
    k_scan<<<1, 1024, 0, h->stream>>>(I(o_act), Np, I(o_ptrP));
    k_compact<<<bP, 256, 0, h->stream>>>(Np, I(o_act), I(o_ptrP), I(o_actl), nullptr);
  }
  k_pair_count<<<bV, 256, 0, h->stream>>>(V, h->g, I(o_wtype), I(o_cc));
  k_scan<<<1, 1024, 0, h->stream>>>(I(o_cc), V, I(o_cp));
  k_pair_emit<<<bV, 256, 0, h->stream>>>(V, h->g, I(o_wtype), I(o_pos), I(o_cp), I(o_ci), I(o_cj),
                                        reinterpret_cast<double*>(W + o_cT), reinterpret_cast<double*>(W + o_cL));
  GCK(cudaGetLastError());
  int P = 0, L = 0, C = 0;
  GCK(cudaMemcpyAsync(&P, I(o_ptrV) + V, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  if (Np) GCK(cudaMemcpyAsync(&L, I(o_ptrP) + Np, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  GCK(cudaMemcpyAsync(&C, I(o_cp) + V, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  GCK(cudaStreamSynchronize(h->stream));
  *P_out = P; *L_out = L;
  if (C_out) *C_out = C;
  if (P > cap_P || L > cap_L || (c_i && C > cap_C)) { h->err = "window, active points or constraints exceed the caller's capacity"; return SVS_ERR_INVALID; }
  GCK(cudaMemcpyAsync(window_vertex, I(o_win), sizeof(int) * (size_t)P, cudaMemcpyDeviceToHost, h->stream));
  if (L) GCK(cudaMemcpyAsync(active_point, I(o_actl), sizeof(int) * (size_t)L, cudaMemcpyDeviceToHost, h->stream));
  h->h_winpos.resize(V);
  GCK(cudaMemcpyAsync(h->h_winpos.data(), I(o_wtype), sizeof(int) * (size_t)V, cudaMemcpyDeviceToHost, h->stream));
  if (c_i && C) {
    if (!c_j || !c_T || !c_Lambda) return SVS_ERR_INVALID;
    GCK(cudaMemcpyAsync(c_i, I(o_ci), sizeof(int) * (size_t)C, cudaMemcpyDeviceToHost, h->stream));
    GCK(cudaMemcpyAsync(c_j, I(o_cj), sizeof(int) * (size_t)C, cudaMemcpyDeviceToHost, h->stream));
    GCK(cudaMemcpyAsync(c_T, W + o_cT, sizeof(double) * 7 * (size_t)C, cudaMemcpyDeviceToHost, h->stream));
    GCK(cudaMemcpyAsync(c_Lambda, W + o_cL, sizeof(double) * 36 * (size_t)C, cudaMemcpyDeviceToHost, h->stream));
  }
  GCK(cudaStreamSynchronize(h->stream));
  if (inner)
    for (int i = 0; i < P; ++i) inner[i] = h->h_winpos[window_vertex[i]] == 1;
  return SVS_OK;
}

int svs_map_add_keyframe(svs_map* h, int oldkey, const double* T_newkey_from_oldkey, int n_new, const int* new_anchor,
                         const double* new_xyz_anchor, const double* new_anchor_center, const int* new_anchor_level,
                         const double* new_center, const int* new_level, int n_track, const int* track_point,
                         const double* track_center, const int* track_level, int* vertex_index, int* first_new_point) {
  if (!h || !h->d_map || !T_newkey_from_oldkey || n_new < 0 || n_track < 0 ||
      (n_new && (!new_anchor || !new_xyz_anchor || !new_anchor_center || !new_anchor_level || !new_center || !new_level)) ||
      (n_track && (!track_point || !track_center || !track_level)))
    return SVS_ERR_INVALID;
  const int V = h->V, Np = h->Np, nnz = h->nnz;
  if (oldkey < 0 || oldkey >= V) { h->err = "oldkey outside [0, V)"; return SVS_ERR_INVALID; }
  for (int q = 0; q < n_new; ++q)
    if (new_anchor[q] < 0 || new_anchor[q] >= V || new_anchor_level[q] < 0 || new_anchor_level[q] > 30 || new_level[q] < 0 || new_level[q] > 30) {
      h->err = "new point anchored outside [0, V) or bad pyramid level";
      return SVS_ERR_INVALID;
    }
  {
    std::vector<int> tp(track_point, track_point + n_track);
    std::sort(tp.begin(), tp.end());
    for (int t = 0; t < n_track; ++t)
      if (tp[t] < 0 || tp[t] >= Np || (t && tp[t] == tp[t - 1]) || track_level[t] < 0 || track_level[t] > 30) {
        h->err = "tracked point outside [0, Np), listed twice, or bad pyramid level";
        return SVS_ERR_INVALID;
      }
  }
  cudaSetDevice(h->device);
  GCK(cudaStreamSynchronize(h->stream));
  const int V2 = V + 1, Np2 = Np + n_new, nnz2 = nnz + n_track + 2 * n_new;
  const MapLayout lo = map_layout(V2, Np2, nnz2);
  char* B2 = nullptr;
  GCK(cudaMalloc(&B2, lo.total + lo.total / 4));
  // staging: everything the kernels read from the caller, in one pinned-less copy (keyframe rate, a few 10 KB)
  std::vector<char> st;
  auto push = [&](const void* src, size_t bytes) { const size_t o = st.size(); st.resize(o + al256(bytes)); if (bytes) memcpy(st.data() + o, src, bytes); return o; };
  const size_t s_T = push(T_newkey_from_oldkey, sizeof(double) * 7);
  const size_t s_na = push(new_anchor, sizeof(int) * (size_t)n_new), s_nx = push(new_xyz_anchor, sizeof(double) * 3 * (size_t)n_new);
  const size_t s_nac = push(new_anchor_center, sizeof(double) * 3 * (size_t)n_new), s_nal = push(new_anchor_level, sizeof(int) * (size_t)n_new);
  const size_t s_nc = push(new_center, sizeof(double) * 3 * (size_t)n_new), s_nl = push(new_level, sizeof(int) * (size_t)n_new);
  const size_t s_tp = push(track_point, sizeof(int) * (size_t)n_track), s_tc = push(track_center, sizeof(double) * 3 * (size_t)n_track);
  const size_t s_tl = push(track_level, sizeof(int) * (size_t)n_track);
  const size_t s_add = st.size(); st.resize(s_add + al256(sizeof(int) * (size_t)Np2));
  const size_t s_cnt = st.size(); st.resize(s_cnt + al256(sizeof(int) * (size_t)Np2));
  if (st.size() > h->upd_cap) {
    cudaFree(h->d_upd); h->d_upd = nullptr; h->upd_cap = 0;
    if (cudaMalloc(&h->d_upd, 2 * st.size()) != cudaSuccess) { cudaFree(B2); h->err = "cudaMalloc"; return SVS_ERR_CUDA; }
    h->upd_cap = 2 * st.size();
  }
  char* U = h->d_upd;
  cudaError_t e = cudaMemcpyAsync(U, st.data(), s_add, cudaMemcpyHostToDevice, h->stream);
  auto D = [&](size_t o) { return reinterpret_cast<double*>(U + o); };
  auto Ii = [&](size_t o) { return reinterpret_cast<int*>(U + o); };
  if (e == cudaSuccess) e = cudaMemsetAsync(U + s_add, 0xff, sizeof(int) * (size_t)Np2, h->stream);
  // vertices
  if (e == cudaSuccess) e = cudaMemcpyAsync(B2 + lo.o_pose, h->m.pose, sizeof(double) * 7 * (size_t)V, cudaMemcpyDeviceToDevice, h->stream);
  k_new_pose<<<1, 32, 0, h->stream>>>(h->m.pose, oldkey, D(s_T), reinterpret_cast<double*>(B2 + lo.o_pose) + 7 * (size_t)V);
  // points
  if (Np && e == cudaSuccess) e = cudaMemcpyAsync(B2 + lo.o_anch, h->m.anchor, sizeof(int) * (size_t)Np, cudaMemcpyDeviceToDevice, h->stream);
  if (Np && e == cudaSuccess) e = cudaMemcpyAsync(B2 + lo.o_xyz, h->m.xyz, sizeof(double) * 3 * (size_t)Np, cudaMemcpyDeviceToDevice, h->stream);
  if (n_new && e == cudaSuccess) e = cudaMemcpyAsync(B2 + lo.o_anch + sizeof(int) * (size_t)Np, U + s_na, sizeof(int) * (size_t)n_new, cudaMemcpyDeviceToDevice, h->stream);
  if (n_new && e == cudaSuccess) e = cudaMemcpyAsync(B2 + lo.o_xyz + sizeof(double) * 3 * (size_t)Np, U + s_nx, sizeof(double) * 3 * (size_t)n_new, cudaMemcpyDeviceToDevice, h->stream);
  // observations
  if (Np2) {
    if (n_track) k_mark_tracks<<<(n_track + 255) / 256, 256, 0, h->stream>>>(n_track, Ii(s_tp), Ii(s_add));
    k_grow_count<<<(Np2 + 255) / 256, 256, 0, h->stream>>>(Np, Np2, h->m.vis_ptr, Ii(s_add), Ii(s_cnt));
    k_scan<<<1, 1024, 0, h->stream>>>(Ii(s_cnt), Np2, reinterpret_cast<int*>(B2 + lo.o_vptr));
    k_grow_move<<<(Np2 + 255) / 256, 256, 0, h->stream>>>(h->m, Np2, V, Ii(s_add), reinterpret_cast<const int*>(B2 + lo.o_vptr), D(s_tc),
                                                        Ii(s_tl), Ii(s_na), D(s_nac), Ii(s_nal), D(s_nc), Ii(s_nl),
                                                        reinterpret_cast<int*>(B2 + lo.o_vpose), reinterpret_cast<double*>(B2 + lo.o_cen),
                                                        reinterpret_cast<int*>(B2 + lo.o_lvl));
  }
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  if (e != cudaSuccess) { cudaFree(B2); h->err = std::string("svs_map_add_keyframe: ") + cudaGetErrorString(e); return SVS_ERR_CUDA; }
  cudaFree(h->d_map);
  h->d_map = B2; h->map_cap = lo.total + lo.total / 4;
  map_bind(h, B2, lo, V2, Np2, nnz2);
  h->g = GraphDev{}; h->nnzN = 0;          // the pose graph changed with the new vertex: svs_map_set_graph again
  h->d_win_last = nullptr;                 // (a window assembled before the growth can no longer be absorbed)
  if (vertex_index) *vertex_index = V;
  if (first_new_point) *first_new_point = Np;
  return SVS_OK;
}

// the assembled edge list of the last svs_ba_set_problem_from_map, for inspection
int svs_map_last_edges(svs_map* h, int E, int* e_point, int* e_pose, int* e_anchor, double* e_obs, double* e_info) {
  if (!h || E != h->last_E || !h->d_work) return SVS_ERR_INVALID;
  if (E == 0) return SVS_OK;
  if (e_point) memcpy(e_point, h->h_ep.data(), sizeof(int) * (size_t)E);
  if (e_pose) memcpy(e_pose, h->h_es.data(), sizeof(int) * (size_t)E);
  if (e_anchor) memcpy(e_anchor, h->h_ea.data(), sizeof(int) * (size_t)E);
  cudaSetDevice(h->device);
  if (e_obs) GCK(cudaMemcpy(e_obs, h->d_oi_last, sizeof(double) * 3 * (size_t)E, cudaMemcpyDeviceToHost));
  if (e_info) GCK(cudaMemcpy(e_info, h->d_oi_last + 3 * (size_t)E, sizeof(double) * 3 * (size_t)E, cudaMemcpyDeviceToHost));
  return SVS_OK;
}

}  // extern "C"
