// Drives the front-end and the device-resident map through the C++ layer (include/svs_b200.hpp) with the
// reference's class and method names: FramePreprocessor (FrameGrabber::preprocessing) -> FastGrid::detectAdaptively
// -> DenseTracker::{computeDensePointCloudGpu, denseTrackingGpu} -> GuidedMatcher::match (corners handed over on the
// device) -> BA_SE3_XYZ_STEREO::calcFastMotionOnly -> DeviceMap::{copyDataToG2o, restoreDataFromG2o}.
// usage: frontend_main in.bin out.bin      (tests/test_cpp_shim.py writes in.bin and checks out.bin against the
// same calls made through the C ABI from Python)
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "svs_b200.hpp"

template <typename T>
static std::vector<T> rd(FILE* f, size_t n) {
  std::vector<T> v(n);
  if (n && fread(v.data(), sizeof(T), n, f) != n) { fprintf(stderr, "short read\n"); exit(2); }
  return v;
}
template <typename T>
static void wr(FILE* f, const std::vector<T>& v) { fwrite(v.data(), sizeof(T), v.size(), f); }

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 2;
  const auto hd = rd<int>(f, 8);   // W H V Np nnz P L nlevel_cams(=3)
  const int W = hd[0], H = hd[1], V = hd[2], Np = hd[3], nnz = hd[4], P = hd[5], L = hd[6];
  const auto cams_flat = rd<double>(f, 12);   // 3 levels x (f px py b)
  const auto img0 = rd<unsigned char>(f, (size_t)W * H), img1 = rd<unsigned char>(f, (size_t)W * H);
  const auto disp0 = rd<float>(f, (size_t)W * H), disp1 = rd<float>(f, (size_t)W * H);
  const auto m_pose = rd<double>(f, 7 * (size_t)V);
  const auto m_anchor = rd<int>(f, Np);
  const auto m_xyz = rd<double>(f, 3 * (size_t)Np);
  const auto m_vptr = rd<int>(f, (size_t)Np + 1), m_vpose = rd<int>(f, nnz);
  const auto m_center = rd<double>(f, 3 * (size_t)nnz);
  const auto m_level = rd<int>(f, nnz);
  const auto win = rd<int>(f, P), act = rd<int>(f, L);
  fclose(f);
  svs_cam cams[3];
  for (int l = 0; l < 3; ++l) cams[l] = svs_cam{cams_flat[4 * l], cams_flat[4 * l + 1], cams_flat[4 * l + 2], cams_flat[4 * l + 3]};

  svs::FramePreprocessor pp0(W, H, 3), pp1(W, H, 3);
  if (!pp0.valid()) { printf("NO_GPU\n"); return 3; }
  if (!pp0.preprocessing(img0.data(), W) || !pp1.preprocessing(img1.data(), W)) return 4;
  svs::FramePreprocessor::Level l0[3], l1[3];
  for (int l = 0; l < 3; ++l)
    if (!pp0.level(l, &l0[l]) || !pp1.level(l, &l1[l])) return 4;

  // FAST on both frames (level 0 and 1), images taken where the preprocessing left them
  svs::FastGrid fg0(W, H, 222, 74, 25, 3, 3), fg1(W / 2, H / 2, 55, 18, 25, 3, 3);
  std::vector<int> kxy0, koff0, xy[2], off[2];
  if (fg0.detectAdaptivelyDevice(l0[0].u8, l0[0].pitch_u8, W, H, 6, &kxy0, &koff0) < 0) return 5;   // keyframe corners
  svs::FastGrid fc0(W, H, 222, 74, 25, 3, 3);
  if (fc0.detectAdaptivelyDevice(l1[0].u8, l1[0].pitch_u8, W, H, 6, &xy[0], &off[0]) < 0) return 5;
  if (fg1.detectAdaptivelyDevice(l1[1].u8, l1[1].pitch_u8, W / 2, H / 2, 6, &xy[1], &off[1]) < 0) return 5;

  // dense tracking frame 0 -> frame 1
  svs::DenseTracker dt(W, H, 3);
  for (int l = 0; l < 3; ++l) {
    svs_dt_set_intrinsics(dt.handle(), l, (float)cams[l].f, (float)cams[l].px, (float)cams[l].py);
    if (svs_dt_set_images_device(dt.handle(), l, l0[l].f32, l1[l].f32, l1[l].dx, l1[l].dy, l1[l].stride_f32) != SVS_OK) return 6;
  }
  if (svs_dt_set_disparity(dt.handle(), disp0.data(), W, W, H) != SVS_OK) return 6;
  svs::SE3d T;   // identity
  if (!dt.computeDensePointCloudGpu(T, cams) || !dt.denseTrackingGpu(&T)) return 6;
  std::vector<double> T_track = {T.q[0], T.q[1], T.q[2], T.q[3], T.t[0], T.t[1], T.t[2]};
  // residual image of the coarsest level at the tracked pose: every pixel is one of the three classes
  std::vector<float> res_img;
  if (!dt.residualImage(2, T, &res_img, W / 4, H / 4)) return 6;
  for (size_t i = 0; i < res_img.size(); i += 4)
    if (res_img[i + 3] != 1.f || res_img[i] < 0.f || res_img[i] > 1.f) return 6;

  // guided matching of the keyframe's corners (level 0) into frame 1
  std::vector<svs_match_level> lv = {{W, H, cams[0].f, cams[0].px, cams[0].py}, {W / 2, H / 2, cams[1].f, cams[1].px, cams[1].py}};
  svs::GuidedMatcher gm(lv);
  const double I7[7] = {0, 0, 0, 1, 0, 0, 0};
  const unsigned char* kp[2] = {l0[0].u8, l0[1].u8};
  const unsigned char* cp[2] = {l1[0].u8, l1[1].u8};
  const int kpitch[2] = {l0[0].pitch_u8, l0[1].pitch_u8}, cpitch[2] = {l1[0].pitch_u8, l1[1].pitch_u8};
  if (svs_matcher_set_pyramid_device(gm.handle(), 0, I7, kp, kpitch) != SVS_OK ||
      svs_matcher_set_pyramid_device(gm.handle(), -1, nullptr, cp, cpitch) != SVS_OK ||
      svs_matcher_set_current(gm.handle(), nullptr, nullptr, disp1.data(), W) != SVS_OK)
    return 7;
  if (!gm.setFeatureTree(0, fc0) || !gm.setFeatureTree(1, fg1)) return 7;
  std::vector<svs_match_point> ap;
  for (size_t i = 0; 2 * i < kxy0.size(); ++i) {
    const int u = kxy0[2 * i], v = kxy0[2 * i + 1];
    const float d = disp0[(size_t)v * W + u];
    if (!(d > 0)) continue;
    const double z = cams[0].f * cams[0].b / d;
    svs_match_point q{};
    q.keyframe = 0; q.anchor_level = 0;
    q.xyz_anchor[0] = (u - cams[0].px) / cams[0].f * z; q.xyz_anchor[1] = (v - cams[0].py) / cams[0].f * z; q.xyz_anchor[2] = z;
    q.anchor_obs_pyr[0] = u; q.anchor_obs_pyr[1] = v;
    ap.push_back(q);
  }
  double T7[7] = {T.q[0], T.q[1], T.q[2], T.q[3], T.t[0], T.t[1], T.t[2]};
  std::vector<svs_match_result> track;
  const int nm = gm.match(T7, I7, ap, 4, 22, 10, &track);
  if (nm < 0) return 7;

  // motion-only refinement on the matcher's device results
  svs::BA_SE3_XYZ_STEREO pose_opt;
  svs::PoseOptimizerParams prm(true, 2.0, 15);
  svs::OptimizerStatistics ost = pose_opt.calcFastMotionOnly(gm, cams[0], prm, &T);
  std::vector<double> T_pose = {T.q[0], T.q[1], T.q[2], T.q[3], T.t[0], T.t[1], T.t[2]};

  // device-resident map: assemble the window, two LM iterations, write back on the device
  svs::DeviceMap dm;
  svs::StereoGraph graph;
  if (!dm.set(m_pose, m_anchor, m_xyz, m_vptr, m_vpose, m_center, m_level)) return 8;
  const int E = dm.copyDataToG2o(graph.handle(), win, act, cams[0]);
  if (E < 0) { fprintf(stderr, "%s\n", dm.last_error()); return 8; }
  svs_ba_stats st{};
  if (svs_ba_optimize(graph.handle(), 2, 1, 1.0, 50., 5, &st) != 2) return 8;
  if (!dm.restoreDataFromG2o(graph.handle())) return 8;
  std::vector<double> o_pose, o_xyz;
  if (!dm.get(&o_pose, &o_xyz)) return 8;

  // window selection from the pose graph (a ring over the vertices) and growth by one keyframe, on the device tables
  const int Vm = (int)(m_pose.size() / 7);
  std::vector<int> nbr_ptr(Vm + 1), nbr_id;
  for (int v = 0; v < Vm; ++v) {
    nbr_ptr[v] = (int)nbr_id.size();
    nbr_id.push_back((v + 1) % Vm); nbr_id.push_back((v + Vm - 1) % Vm);
  }
  nbr_ptr[Vm] = (int)nbr_id.size();
  if (!dm.setGraph(nbr_ptr, nbr_id)) return 9;
  svs::DeviceMap::DoubleWindow dw;
  if (!dm.computeDoubleWindow(win[0], 3, 6, &dw)) { fprintf(stderr, "%s\n", dm.last_error()); return 9; }
  const double T_rel[7] = {0, 0, 0, 1, 0.25, -0.5, 0.125};
  std::vector<int> tracked = {0, 1, 2, 3, 4}, tlevel = {0, 1, 0, 1, 0};
  std::vector<double> tcenter(15, 100.);
  const int newv = dm.addKeyframe(win[1], T_rel, {}, {}, {}, {}, {}, {}, tracked, tcenter, tlevel);
  if (newv != Vm) { fprintf(stderr, "%s\n", dm.last_error()); return 9; }
  std::vector<double> g_pose, g_xyz;
  if (!dm.get(&g_pose, &g_xyz)) return 9;
  std::vector<double> new_pose(g_pose.end() - 7, g_pose.end());

  FILE* o = fopen(argv[2], "wb");
  std::vector<int> counts = {(int)(xy[0].size() / 2), (int)(xy[1].size() / 2), (int)ap.size(), nm, E, ost.num_obs};
  wr(o, counts); wr(o, xy[0]); wr(o, xy[1]);
  std::vector<int> midx(track.size());
  for (size_t i = 0; i < track.size(); ++i) midx[i] = track[i].matched ? track[i].index : -1;
  wr(o, midx); wr(o, T_track); wr(o, T_pose); wr(o, o_pose); wr(o, o_xyz);
  std::vector<int> counts2 = {(int)dw.window_vertex.size(), (int)dw.active_point.size(), (int)dw.c_i.size(), newv};
  std::vector<int> inner_i(dw.inner.begin(), dw.inner.end());
  wr(o, counts2); wr(o, dw.window_vertex); wr(o, inner_i); wr(o, dw.active_point); wr(o, dw.c_i); wr(o, dw.c_j); wr(o, new_pose);
  fclose(o);
  printf("OK corners=%d/%d candidates=%d matched=%d edges=%d\n", counts[0], counts[1], counts[2], nm, E);
  return 0;
}
