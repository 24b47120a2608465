// Exercises the C++ host layer (include/svs_b200.hpp) the way a ScaViSLAM maintainer would call it
// from SlamGraph::optimize: read a dumped double window, add vertices/edges by id, optimise, write
// the poses back.  Usage: shim_main <in.bin> <out.bin> <num_iters>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "svs_b200.hpp"

template <typename T>
static bool rd(FILE* f, std::vector<T>& v, size_t n) { v.resize(n); return n == 0 || fread(v.data(), sizeof(T), n, f) == n; }

int main(int argc, char** argv) {
  if (argc < 4) { fprintf(stderr, "usage\n"); return 2; }
  svs::StereoGraph g;
  if (!g.valid()) { printf("NO_GPU %s\n", g.last_error()); return 3; }
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 2;
  int hdr[4]; double cam[4];
  if (fread(hdr, sizeof(int), 4, f) != 4 || fread(cam, sizeof(double), 4, f) != 4) return 2;
  const int P = hdr[0], L = hdr[1], E = hdr[2], C = hdr[3];
  std::vector<double> T, xyz, obs, info, cT, cL;
  std::vector<int> ep, ef, ea, ci, cj;
  if (!rd(f, T, 7 * (size_t)P) || !rd(f, xyz, 3 * (size_t)L) || !rd(f, ep, E) || !rd(f, ef, E) || !rd(f, ea, E) ||
      !rd(f, obs, 3 * (size_t)E) || !rd(f, info, 3 * (size_t)E) || !rd(f, ci, C) || !rd(f, cj, C) ||
      !rd(f, cT, 7 * (size_t)C) || !rd(f, cL, 36 * (size_t)C)) return 2;
  fclose(f);
  g.setCamera(cam[0], cam[1], cam[2], cam[3]);
  // frame ids / point ids are arbitrary integers in the reference: offset them to prove the id mapping
  for (int i = 0; i < P; ++i) { svs::SE3d S; memcpy(S.q, &T[7 * i], 32); memcpy(S.t, &T[7 * i + 4], 24); g.addPose(1000 + 3 * i, S); }
  for (int l = 0; l < L; ++l) g.addPoint(50000 + l, &xyz[3 * l]);
  for (int e = 0; e < E; ++e) g.addObs(&obs[3 * e], &info[3 * e], 50000 + ep[e], 1000 + 3 * ef[e], 1000 + 3 * ea[e]);
  for (int c = 0; c < C; ++c) { svs::SE3d S; memcpy(S.q, &cT[7 * c], 32); memcpy(S.t, &cT[7 * c + 4], 24);
                                g.addConstraint(S, &cL[36 * c], 1000 + 3 * ci[c], 1000 + 3 * cj[c]); }
  svs::Statistics st;
  const int it = g.optimize(svs::OptParams(atoi(argv[3]), true, 3), &st);   // OptParams(2,true,3): backend.cpp:187
  if (it <= -100) { printf("ERROR %d %s\n", it, g.last_error()); return 4; }
  FILE* o = fopen(argv[2], "wb");
  for (size_t i = 0; i < g.num_poses(); ++i) { svs::SE3d S = g.pose(i); fwrite(S.q, 8, 4, o); fwrite(S.t, 8, 3, o); }
  for (size_t l = 0; l < g.num_points(); ++l) { double x[3]; g.point_xyz_anchor(l, x); fwrite(x, 8, 3, o); }
  fclose(o);
  printf("OK iterations=%d frames=%d points=%d point_edges=%d frame_edges=%d calc_time=%.6f\n", it, st.num_frames,
         st.num_points, st.num_point_edges, st.num_frame_edges, st.calc_time);
  return 0;
}
