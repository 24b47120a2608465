// svs_b200.hpp -- header-only C++ host layer above the C ABI (svs_b200.h), dependency-free.
//
// Mirrors the reference's interfaces for the hot path so that a maintainer can swap the bodies
// of the corresponding ScaViSLAM methods for calls into this header:
//
//   svs::OptParams / svs::Statistics   <->  ScaViSLAM::OptParams (slam_graph.hpp:36-50),
//                                           SlamGraph::Statistics (slam_graph.hpp:366-386)
//   svs::StereoGraph::optimize         <->  SlamGraph<SE3,StereoCamera,SE3XYZ_STEREO,3>::optimize
//                                           (slam_graph.cpp:319-355) = copyDataToG2o + g2o + restore
//   svs::FastGrid                      <->  ScaViSLAM::FastGrid (fast_grid.h:30-64)
//   svs::DenseTracker                  <->  ScaViSLAM::DenseTracker / GpuTracker (dense_tracking.h:40-96)
//   svs::GuidedMatcher                 <->  ScaViSLAM::GuidedMatcher<StereoCamera> (matcher.hpp:62-186)
//
// "We do not use C++ exceptions" (reference README:295): errors come back as bool / int; the
// text is available from last_error().  There is no CPU fallback anywhere in this layer.
#ifndef SVS_B200_HPP
#define SVS_B200_HPP

#include <cmath>
#include <cstring>
#include <stdexcept>
#include <string>
#include <algorithm>
#include <utility>
#include <vector>

#include "svs_b200.h"

namespace svs {

struct SE3d {           // Sophus::SE3 as the C ABI carries it: unit quaternion (x y z w) + translation
  double q[4] = {0, 0, 0, 1};
  double t[3] = {0, 0, 0};
};

struct OptParams {      // slam_graph.hpp:36-50
  OptParams(int num_iters, bool use_robust_kernel = false, double huber_kernel_width = 1)
      : num_iters(num_iters), use_robust_kernel(use_robust_kernel), huber_kernel_width(huber_kernel_width) {}
  int num_iters;
  bool use_robust_kernel;
  double huber_kernel_width;
};

struct Statistics {     // slam_graph.hpp:366-386
  int num_frame_edges = 0, num_point_edges = 0, num_frames = 0, num_points = 0;
  double calc_time = 0;
};

// The double window handed to the optimiser: what SlamGraph::copyDataToG2o walks
// (slam_graph.cpp:985-1032).  Ids are the caller's (frame ids / point ids, any integers).
class StereoGraph {
 public:
  StereoGraph() { svs_ba_opts o{-1, 0, {0, 0, 0, 0, 0, 0}}; ok_ = svs_ba_create(&o, &h_) == SVS_OK; }
  ~StereoGraph() { if (h_) svs_ba_destroy(h_); }
  StereoGraph(const StereoGraph&) = delete;
  StereoGraph& operator=(const StereoGraph&) = delete;
  bool valid() const { return ok_; }
  svs_ba* handle() { return h_; }
  const char* last_error() const { return h_ ? svs_last_error(h_) : "svs_ba_create failed (no CUDA device)"; }

  void clear() { pose_id_.clear(); T_.clear(); fixed_.clear(); point_id_.clear(); psi_.clear(); ep_.clear(); ef_.clear();
                 ea_.clear(); obs_.clear(); info_.clear(); ci_.clear(); cj_.clear(); cT_.clear(); cL_.clear(); }
  void setCamera(double f, double px, double py, double b) { cam_ = svs_cam{f, px, py, b}; }
  // addPoseToG2o (slam_graph-impl.cpp:29-42)
  void addPose(int frame_id, const SE3d& T_me_from_w, bool fixed = false) {
    pose_id_.push_back(frame_id); push7(T_, T_me_from_w); fixed_.push_back(fixed ? 1 : 0);
  }
  // addPointToG2o (slam_graph.cpp:907-920): xyz in the anchor frame; stored as psi = invert_depth(xyz)
  void addPoint(int point_id, const double xyz_anchor[3]) {
    point_id_.push_back(point_id);
    psi_.push_back(xyz_anchor[0] / xyz_anchor[2]); psi_.push_back(xyz_anchor[1] / xyz_anchor[2]); psi_.push_back(1. / xyz_anchor[2]);
  }
  // addObsToG2o (slam_graph-impl.cpp:44-97): obs = (u, v, u_right), Lambda = diag(lambda)
  void addObs(const double obs[3], const double lambda_diag[3], int point_id, int frame_id, int anchor_id) {
    ep_.push_back(point_id); ef_.push_back(frame_id); ea_.push_back(anchor_id);
    for (int k = 0; k < 3; ++k) { obs_.push_back(obs[k]); info_.push_back(lambda_diag[k]); }
  }
  // addConstraintToG2o (slam_graph-impl.cpp:99-126): Lambda row-major 6x6
  void addConstraint(const SE3d& T_2_from_1, const double Lambda[36], int frame_id_1, int frame_id_2) {
    ci_.push_back(frame_id_1); cj_.push_back(frame_id_2); push7(cT_, T_2_from_1);
    cL_.insert(cL_.end(), Lambda, Lambda + 36);
  }

  // SlamGraph::optimize(const OptParams&, Statistics*) (slam_graph.cpp:319-355).  lambda0 = 50 and
  // 5 trials as the reference configures g2o (:338, :1073).  Note the reference never applies
  // huber_kernel_width (slam_graph-impl.cpp:86-90, delta stays 1): pass apply_huber_width = true
  // for the corrected behaviour.  Returns g2o's optimize() value.
  int optimize(const OptParams& p, Statistics* stats = nullptr, bool apply_huber_width = false) {
    if (!ok_) return -100 + SVS_ERR_NOGPU;
    std::vector<int> ep(ep_.size()), ef(ef_.size()), ea(ea_.size()), ci(ci_.size()), cj(cj_.size());
    if (!remap(ep_, point_id_, ep) || !remap(ef_, pose_id_, ef) || !remap(ea_, pose_id_, ea) ||
        !remap(ci_, pose_id_, ci) || !remap(cj_, pose_id_, cj))
      return -100 + SVS_ERR_INVALID;
    svs_ba_stats st{};
    const int it = svs_optimiseInnerAndOuterWindow(
        h_, (int)pose_id_.size(), T_.data(), fixed_.data(), (int)point_id_.size(), psi_.data(), (int)ep.size(), ep.data(),
        ef.data(), ea.data(), obs_.data(), info_.data(), (int)ci.size(), ci.data(), cj.data(), cT_.data(), cL_.data(), &cam_,
        p.num_iters, p.use_robust_kernel ? 1 : 0, apply_huber_width ? p.huber_kernel_width : 1.0, &st);
    if (stats && it > -100) {
      stats->num_frames = st.num_frames; stats->num_points = st.num_points;
      stats->num_point_edges = st.num_point_edges; stats->num_frame_edges = st.num_frame_edges;
      stats->calc_time = st.ms_total * 1e-3;
    }
    last_ = st;
    return it;
  }
  // restoreDataFromG2o (slam_graph.cpp:1037-1058)
  SE3d pose(size_t i) const { SE3d T; memcpy(T.q, &T_[7 * i], 32); memcpy(T.t, &T_[7 * i + 4], 24); return T; }
  void point_xyz_anchor(size_t i, double xyz[3]) const {   // invert_depth(psi)
    xyz[0] = psi_[3 * i] / psi_[3 * i + 2]; xyz[1] = psi_[3 * i + 1] / psi_[3 * i + 2]; xyz[2] = 1. / psi_[3 * i + 2];
  }
  size_t num_poses() const { return pose_id_.size(); }
  size_t num_points() const { return point_id_.size(); }
  const svs_ba_stats& last_stats() const { return last_; }

 private:
  static void push7(std::vector<double>& v, const SE3d& T) { v.insert(v.end(), T.q, T.q + 4); v.insert(v.end(), T.t, T.t + 3); }
  // id -> index through a sorted copy of the id table (ids may be sparse or hashed); a duplicate id is an error
  static bool remap(const std::vector<int>& ids, const std::vector<int>& table, std::vector<int>& out) {
    std::vector<std::pair<int, int>> sorted(table.size());
    for (size_t k = 0; k < table.size(); ++k) sorted[k] = std::make_pair(table[k], (int)k);
    std::sort(sorted.begin(), sorted.end());
    for (size_t k = 1; k < sorted.size(); ++k)
      if (sorted[k].first == sorted[k - 1].first) return false;
    for (size_t k = 0; k < ids.size(); ++k) {
      auto it = std::lower_bound(sorted.begin(), sorted.end(), std::make_pair(ids[k], 0),
                                 [](const std::pair<int, int>& a, const std::pair<int, int>& b) { return a.first < b.first; });
      if (it == sorted.end() || it->first != ids[k]) return false;
      out[k] = it->second;
    }
    return true;
  }
  svs_ba* h_ = nullptr;
  bool ok_ = false;
  svs_cam cam_{1, 0, 0, 0.5};
  std::vector<int> pose_id_, point_id_, ep_, ef_, ea_, ci_, cj_;
  std::vector<double> T_, psi_, obs_, info_, cT_, cL_;
  std::vector<unsigned char> fixed_;
  svs_ba_stats last_{};
};

// ScaViSLAM::FastGrid (fast_grid.h:30-64).  Keypoints come back as flat (x, y) pairs grouped by
// cell; the quadtree content of keypoint i of cell c is i - cell_off[c].
class FastGrid {
 public:
  FastGrid(int img_w, int img_h, int num_features_per_cell, int boundary_per_cell, int fast_thr, int grid_w, int grid_h,
           int fast_min = 10, int fast_max = 40, int max_keypoints = 200000)
      : cells_((size_t)grid_w * grid_h), max_kp_(max_keypoints) {
    ok_ = svs_fast_create(-1, img_w, img_h, max_keypoints, &h_) == SVS_OK &&
          svs_fast_grid_init(img_w, img_h, num_features_per_cell, boundary_per_cell, fast_thr, grid_w, grid_h, fast_min,
                             fast_max, &grid_, cells_.data()) == SVS_OK;
  }
  ~FastGrid() { if (h_) svs_fast_destroy(h_); }
  FastGrid(const FastGrid&) = delete;
  FastGrid& operator=(const FastGrid&) = delete;
  bool valid() const { return ok_; }
  svs_fast* handle() { return h_; }
  const std::vector<svs_fast_cell>& cell_grid2d() const { return cells_; }
  // the same on an image that already lies on the device (a FramePreprocessor level)
  int detectAdaptivelyDevice(const unsigned char* d_img, int pitch, int w, int h, int trials, std::vector<int>* xy,
                             std::vector<int>* cell_off) {
    if (!ok_ || svs_fast_set_image_device(h_, d_img, pitch, w, h) != SVS_OK) return -1;
    xy->resize(2 * (size_t)max_kp_); cell_off->resize(cells_.size() + 1);
    const int n = svs_fast_detect_adaptively(h_, &grid_, cells_.data(), trials, xy->data(), max_kp_, cell_off->data());
    if (n >= 0) xy->resize(2 * (size_t)(n < max_kp_ ? n : max_kp_));
    return n;
  }
  // detectAdaptively(img, trials, qt)
  int detectAdaptively(const unsigned char* img, int pitch, int w, int h, int trials, std::vector<int>* xy,
                       std::vector<int>* cell_off) {
    if (!ok_ || svs_fast_set_image(h_, img, pitch, w, h) != SVS_OK) return -1;
    xy->resize(2 * (size_t)max_kp_); cell_off->resize(cells_.size() + 1);
    const int n = svs_fast_detect_adaptively(h_, &grid_, cells_.data(), trials, xy->data(), max_kp_, cell_off->data());
    if (n >= 0) xy->resize(2 * (size_t)(n < max_kp_ ? n : max_kp_));
    return n;
  }
  // static FastGrid::detect(img, cell_grid2d, qt)
  int detect(const unsigned char* img, int pitch, int w, int h, const std::vector<svs_fast_cell>& cells,
             std::vector<int>* xy, std::vector<int>* cell_off) {
    if (!ok_ || svs_fast_set_image(h_, img, pitch, w, h) != SVS_OK) return -1;
    xy->resize(2 * (size_t)max_kp_); cell_off->resize(cells.size() + 1);
    const int n = svs_fast_detect(h_, cells.data(), (int)cells.size(), xy->data(), max_kp_, cell_off->data());
    if (n >= 0) xy->resize(2 * (size_t)(n < max_kp_ ? n : max_kp_));
    return n;
  }

 private:
  svs_fast* h_ = nullptr;
  bool ok_ = false;
  svs_fast_grid_params grid_{};
  std::vector<svs_fast_cell> cells_;
  int max_kp_;
};

// ScaViSLAM::DenseTracker (GPU path)
class DenseTracker {
 public:
  DenseTracker(int w0, int h0, int nlevels = 3, int flags = 0) { ok_ = svs_dt_create(-1, w0, h0, nlevels, flags, &h_) == SVS_OK; }
  ~DenseTracker() { if (h_) svs_dt_destroy(h_); }
  DenseTracker(const DenseTracker&) = delete;
  DenseTracker& operator=(const DenseTracker&) = delete;
  bool valid() const { return ok_; }
  svs_dt* handle() { return h_; }
  // denseTrackingGpu(SE3 * T_cur_from_actkey)
  bool denseTrackingGpu(SE3d* T_cur_from_actkey, svs_dt_stats* stats = nullptr) {
    double T[7];
    memcpy(T, T_cur_from_actkey->q, 32); memcpy(T + 4, T_cur_from_actkey->t, 24);
    if (!ok_ || svs_dt_track(h_, T, stats) != SVS_OK) return false;
    memcpy(T_cur_from_actkey->q, T, 32); memcpy(T_cur_from_actkey->t, T + 4, 24);
    return true;
  }
  // computeDensePointCloudGpu(const SE3 & T_cur_from_actkey)
  bool computeDensePointCloudGpu(const SE3d& T_cur_from_actkey, const svs_cam* level_cams) {
    double T[7];
    memcpy(T, T_cur_from_actkey.q, 32); memcpy(T + 4, T_cur_from_actkey.t, 24);
    return ok_ && svs_dt_compute_point_cloud(h_, T, level_cams) == SVS_OK;
  }
  // GpuTracker::residualImage of one level at T_cur_from_prev (dev_residual_img[l], dense_tracking.cpp:180-188):
  // res_rgba receives w_l * h_l float4
  bool residualImage(int level, const SE3d& T_cur_from_prev, std::vector<float>* res_rgba, int w_l, int h_l) {
    double T[7];
    memcpy(T, T_cur_from_prev.q, 32); memcpy(T + 4, T_cur_from_prev.t, 24);
    res_rgba->resize((size_t)4 * w_l * h_l);
    return ok_ && svs_dt_residual_image(h_, level, T, res_rgba->data()) == SVS_OK;
  }

 private:
  svs_dt* h_ = nullptr;
  bool ok_ = false;
};

// ScaViSLAM::DenseTracker as built without SCAVISLAM_CUDA_SUPPORT (dense_tracking.cpp:222-423)
class DenseTrackerCpuVariant {
 public:
  DenseTrackerCpuVariant(int w, int h, int nlevels) { ok_ = svs_dtc_create(-1, w, h, nlevels, &h_) == SVS_OK; }
  ~DenseTrackerCpuVariant() { if (h_) svs_dtc_destroy(h_); }
  DenseTrackerCpuVariant(const DenseTrackerCpuVariant&) = delete;
  DenseTrackerCpuVariant& operator=(const DenseTrackerCpuVariant&) = delete;
  bool valid() const { return ok_; }
  svs_dtc* handle() { return h_; }
  // computeDensePointCloudCpu(T_cur_from_actkey); cam_vec = one svs_cam per pyramid level
  bool computeDensePointCloudCpu(const SE3d& T, const std::vector<svs_cam>& cam_vec) {
    const double t7[7] = {T.q[0], T.q[1], T.q[2], T.q[3], T.t[0], T.t[1], T.t[2]};
    return ok_ && svs_computeDensePointCloudCpu(h_, t7, cam_vec.data()) == SVS_OK;
  }
  // denseTrackingCpu(&T_cur_from_actkey)
  bool denseTrackingCpu(SE3d* T, const std::vector<svs_cam>& cam_vec, svs_dt_stats* stats = nullptr) {
    double t7[7] = {T->q[0], T->q[1], T->q[2], T->q[3], T->t[0], T->t[1], T->t[2]};
    if (!ok_ || svs_denseTrackingCpu(h_, cam_vec.data(), t7, stats) != SVS_OK) return false;
    for (int k = 0; k < 4; ++k) T->q[k] = t7[k];
    for (int k = 0; k < 3; ++k) T->t[k] = t7[4 + k];
    return true;
  }

 private:
  svs_dtc* h_ = nullptr;
  bool ok_ = false;
};

// ScaViSLAM::GuidedMatcher<StereoCamera>
class GuidedMatcher {
 public:
  GuidedMatcher(const std::vector<svs_match_level>& cam_vec, int max_keyframes = 8, int max_points = 8192,
                int max_keypoints = 65536) {
    ok_ = svs_matcher_create(-1, (int)cam_vec.size(), cam_vec.data(), max_keyframes, max_points, max_keypoints, &h_) == SVS_OK;
  }
  ~GuidedMatcher() { if (h_) svs_matcher_destroy(h_); }
  GuidedMatcher(const GuidedMatcher&) = delete;
  GuidedMatcher& operator=(const GuidedMatcher&) = delete;
  bool valid() const { return ok_; }
  svs_matcher* handle() { return h_; }
  // feature_tree of one pyramid level = the corners FastGrid::detect* left on the device
  bool setFeatureTree(int level, FastGrid& fast_grid) { return ok_ && svs_matcher_set_features_from_fast(h_, level, fast_grid.handle()) == SVS_OK; }
  // match(keyframe_map, T_cur_from_actkey, cur_frame, feature_tree, cam_vec, actkey_id, vertex_map, ap_map,
  //       SEARCHRADIUS, thr_mean, thr_std, track_data): frames/keyframes/features are set on the handle first
  int match(const double T_cur_from_actkey[7], const double T_actkey_from_w[7], const std::vector<svs_match_point>& ap_map,
            int SEARCHRADIUS, int thr_mean, int thr_std, std::vector<svs_match_result>* track_data) {
    track_data->resize(ap_map.size());
    if (!ok_) return -1;
    return svs_match(h_, T_cur_from_actkey, T_actkey_from_w, ap_map.data(), (int)ap_map.size(), SEARCHRADIUS, thr_mean,
                     thr_std, track_data->data());
  }

 private:
  svs_matcher* h_ = nullptr;
  bool ok_ = false;
};

// ScaViSLAM::PoseOptimizerParams (pose_optimizer.h:38-58)
struct PoseOptimizerParams : svs_pose_params {
  PoseOptimizerParams(bool robust_kernel_ = true, double kernel_param_ = 1, int num_iter_ = 50, double initial_mu_ = -1) {
    robust_kernel = robust_kernel_; kernel_param = kernel_param_; num_iter = num_iter_; initial_mu = initial_mu_;
    tau = 0.00001;
  }
};

// ScaViSLAM::OptimizerStatistics (pose_optimizer.h:60-98)
struct OptimizerStatistics : svs_pose_stats {
  double rmse() const { return num_obs > 0 ? std::sqrt(chi2 / num_obs) : 0.; }
};

// ScaViSLAM::BA_SE3_XYZ_STEREO = PoseOptimizer<SE3,6,IdObs<3>,3> (pose_optimizer.h:495)
class BA_SE3_XYZ_STEREO {
 public:
  explicit BA_SE3_XYZ_STEREO(int max_obs = 16384) { ok_ = svs_pose_create(-1, max_obs, &h_) == SVS_OK; }
  ~BA_SE3_XYZ_STEREO() { if (h_) svs_pose_destroy(h_); }
  BA_SE3_XYZ_STEREO(const BA_SE3_XYZ_STEREO&) = delete;
  BA_SE3_XYZ_STEREO& operator=(const BA_SE3_XYZ_STEREO&) = delete;
  bool valid() const { return ok_; }
  // calcFastMotionOnly(obs_list, prediction(cam), ba_params, &frame, &point_list); throws where the reference does
  OptimizerStatistics calcFastMotionOnly(const std::vector<int>& obs_point_id, const std::vector<double>& obs_uvu,
                                         const svs_cam& cam, const PoseOptimizerParams& ba_params, SE3d* frame,
                                         const std::vector<double>& point_list_xyz) {
    OptimizerStatistics st{};
    double T[7] = {frame->q[0], frame->q[1], frame->q[2], frame->q[3], frame->t[0], frame->t[1], frame->t[2]};
    const int rc = ok_ ? svs_calcFastMotionOnly(h_, (int)obs_point_id.size(), obs_point_id.data(), obs_uvu.data(),
                                                (int)(point_list_xyz.size() / 3), point_list_xyz.data(), &cam, &ba_params,
                                                T, &st)
                       : SVS_ERR_NOGPU;
    if (rc != SVS_OK) throw std::runtime_error(ok_ ? svs_pose_last_error(h_) : "no CUDA device");
    for (int k = 0; k < 4; ++k) frame->q[k] = T[k];
    for (int k = 0; k < 3; ++k) frame->t[k] = T[4 + k];
    return st;
  }
  // the same on the TrackData the matcher left on the device
  OptimizerStatistics calcFastMotionOnly(GuidedMatcher& matcher, const svs_cam& cam, const PoseOptimizerParams& ba_params,
                                         SE3d* frame) {
    OptimizerStatistics st{};
    double T[7] = {frame->q[0], frame->q[1], frame->q[2], frame->q[3], frame->t[0], frame->t[1], frame->t[2]};
    const int rc = ok_ ? svs_calcFastMotionOnly_matched(h_, matcher.handle(), &cam, &ba_params, T, &st) : SVS_ERR_NOGPU;
    if (rc != SVS_OK) throw std::runtime_error(ok_ ? svs_pose_last_error(h_) : "no CUDA device");
    for (int k = 0; k < 4; ++k) frame->q[k] = T[k];
    for (int k = 0; k < 3; ++k) frame->t[k] = T[4 + k];
    return st;
  }

 private:
  svs_pose* h_ = nullptr;
  bool ok_ = false;
};

// FrameGrabber::preprocessing (frame_grabber.cpp:287-336): pyramids and derivative images, kept on the device
class FramePreprocessor {
 public:
  FramePreprocessor(int w, int h, int nlevels) { ok_ = svs_prep_create(-1, w, h, nlevels, &h_) == SVS_OK; }
  ~FramePreprocessor() { if (h_) svs_prep_destroy(h_); }
  FramePreprocessor(const FramePreprocessor&) = delete;
  FramePreprocessor& operator=(const FramePreprocessor&) = delete;
  bool valid() const { return ok_; }
  svs_prep* handle() { return h_; }
  bool preprocessing(const unsigned char* left, int pitch) { return ok_ && svs_prep_process(h_, left, pitch) == SVS_OK; }
  // device pointers of one level: hand them to svs_fast_set_image_device / svs_dt_set_images_device / ...
  struct Level { int w, h, pitch_u8, stride_f32; const unsigned char* u8; const float *f32, *dx, *dy; };
  bool level(int l, Level* out) {
    return ok_ && svs_prep_level(h_, l, &out->w, &out->h, &out->u8, &out->pitch_u8, &out->f32, &out->dx, &out->dy,
                                 &out->stride_f32) == SVS_OK;
  }

 private:
  svs_prep* h_ = nullptr;
  bool ok_ = false;
};

// SlamGraph::computeConstraint (slam_graph.cpp:785-846) for a batch of pose pairs
class ConstraintBuilder {
 public:
  ConstraintBuilder() { ok_ = svs_constraints_create(-1, &h_) == SVS_OK; }
  ~ConstraintBuilder() { if (h_) svs_constraints_destroy(h_); }
  ConstraintBuilder(const ConstraintBuilder&) = delete;
  ConstraintBuilder& operator=(const ConstraintBuilder&) = delete;
  bool valid() const { return ok_; }
  const char* last_error() const { return svs_constraints_last_error(h_); }
  // poses [P][7]; feature tables as CSR (ascending point ids); per point its anchor pose index and xyz_anchor
  bool computeConstraints(const std::vector<double>& T_me_from_world, const std::vector<int>& feat_ptr,
                          const std::vector<int>& feat_point, const std::vector<int>& point_anchor,
                          const std::vector<double>& xyz_anchor, const std::vector<int>& v1, const std::vector<int>& v2,
                          std::vector<double>* T_1_from_2, std::vector<double>* Lambda, std::vector<int>* visibility_strength) {
    const int n = (int)v1.size();
    T_1_from_2->resize(7 * (size_t)n); Lambda->resize(36 * (size_t)n); visibility_strength->resize(n);
    return ok_ && v2.size() == v1.size() &&
           svs_computeConstraint_batch(h_, (int)(T_me_from_world.size() / 7), T_me_from_world.data(), feat_ptr.data(),
                                       feat_point.data(), (int)point_anchor.size(), point_anchor.data(), xyz_anchor.data(), n,
                                       v1.data(), v2.data(), T_1_from_2->data(), Lambda->data(),
                                       visibility_strength->data()) == SVS_OK;
  }

 private:
  svs_constraints* h_ = nullptr;
  bool ok_ = false;
};

// The part of SlamGraph the optimiser reads, kept on the device; copyDataToG2o as kernels (slam_graph.cpp:907-1032)
class DeviceMap {
 public:
  DeviceMap() { ok_ = svs_map_create(-1, &h_) == SVS_OK; }
  ~DeviceMap() { if (h_) svs_map_destroy(h_); }
  DeviceMap(const DeviceMap&) = delete;
  DeviceMap& operator=(const DeviceMap&) = delete;
  bool valid() const { return ok_; }
  svs_map* handle() { return h_; }
  const char* last_error() const { return svs_map_last_error(h_); }
  // vertex_table_ / point_table_ / feature tables as flat arrays (slam_graph.hpp:65-137): poses [V][7], per point its
  // anchor vertex and xyz_anchor, vis_set + feature_table as CSR over the points (centre (u,v,u_r) and pyramid level)
  bool set(const std::vector<double>& T_me_from_world, const std::vector<int>& point_anchor, const std::vector<double>& xyz_anchor,
           const std::vector<int>& vis_ptr, const std::vector<int>& vis_pose, const std::vector<double>& feat_center,
           const std::vector<int>& feat_level) {
    V_ = (int)(T_me_from_world.size() / 7); Np_ = (int)point_anchor.size();
    return ok_ && svs_map_set(h_, V_, T_me_from_world.data(), Np_, point_anchor.data(), xyz_anchor.data(), vis_ptr.data(),
                              vis_pose.data(), feat_center.data(), feat_level.data()) == SVS_OK;
  }
  // copyDataToG2o (slam_graph.cpp:985-1032) into `graph`'s bundle adjuster; returns the number of edges, < 0 on error
  int copyDataToG2o(svs_ba* ba, const std::vector<int>& window_vertex, const std::vector<int>& active_point, const svs_cam& cam,
                    const std::vector<int>& c_i = {}, const std::vector<int>& c_j = {}, const std::vector<double>& c_T = {},
                    const std::vector<double>& c_Lambda = {}) {
    int E = 0;
    const int rc = ok_ ? svs_ba_set_problem_from_map(ba, h_, (int)window_vertex.size(), window_vertex.data(), nullptr,
                                                     (int)active_point.size(), active_point.data(), (int)c_i.size(), c_i.data(),
                                                     c_j.data(), c_T.data(), c_Lambda.data(), &cam, &E)
                       : SVS_ERR_NOGPU;
    return rc == SVS_OK ? E : rc;
  }
  // restoreDataFromG2o (slam_graph.cpp:1037-1058), device to device
  bool restoreDataFromG2o(svs_ba* ba) { return ok_ && svs_map_absorb(h_, ba) == SVS_OK; }
  bool updatePoses(const std::vector<int>& vertex, const std::vector<double>& T) {
    return ok_ && svs_map_update_poses(h_, (int)vertex.size(), vertex.data(), T.data()) == SVS_OK;
  }
  bool updatePoints(const std::vector<int>& point, const std::vector<double>& xyz_anchor) {
    return ok_ && svs_map_update_points(h_, (int)point.size(), point.data(), xyz_anchor.data()) == SVS_OK;
  }
  // the pose graph: per vertex its neighbours, strongest first (Vertex::neighbor_ids_ordered_by_strength), and per
  // directed entry the marginalised constraint of the edge table (both payload vectors may be empty)
  bool setGraph(const std::vector<int>& nbr_ptr, const std::vector<int>& nbr_id, const std::vector<double>& nbr_T = {},
                const std::vector<double>& nbr_Lambda = {}) {
    nn_ = (int)nbr_id.size();
    return ok_ && svs_map_set_graph(h_, nbr_ptr.data(), nbr_id.data(), nbr_T.empty() ? nullptr : nbr_T.data(),
                                    nbr_Lambda.empty() ? nullptr : nbr_Lambda.data()) == SVS_OK;
  }
  // computeInitialDoubleWin + computeActivePointsAndExtendOuterWindow + the pair loop of copyContraintsToG2o
  // (slam_graph.cpp:556-663, 938-981): what prepareForOptimization (:290-311) hands to copyDataToG2o
  struct DoubleWindow {
    std::vector<int> window_vertex, active_point, c_i, c_j;
    std::vector<unsigned char> inner;
    std::vector<double> c_T, c_Lambda;
  };
  bool computeDoubleWindow(int root_id, int inner_window_size, int double_window_size, DoubleWindow* w) {
    if (!ok_) return false;
    const int capC = nn_ > 0 ? nn_ : 1;
    w->window_vertex.resize(V_); w->inner.resize(V_); w->active_point.resize(Np_ > 0 ? Np_ : 1);
    w->c_i.resize(capC); w->c_j.resize(capC); w->c_T.resize(7 * (size_t)capC); w->c_Lambda.resize(36 * (size_t)capC);
    int P = 0, L = 0, C = 0;
    if (svs_map_select_window(h_, root_id, inner_window_size, double_window_size, V_, &P, w->window_vertex.data(), w->inner.data(),
                              (int)w->active_point.size(), &L, w->active_point.data(), capC, &C, w->c_i.data(), w->c_j.data(),
                              w->c_T.data(), w->c_Lambda.data()) != SVS_OK)
      return false;
    w->window_vertex.resize(P); w->inner.resize(P); w->active_point.resize(L);
    w->c_i.resize(C); w->c_j.resize(C); w->c_T.resize(7 * (size_t)C); w->c_Lambda.resize(36 * (size_t)C);
    return true;
  }
  // addKeyframe (slam_graph.cpp:144-186): returns the index of the new vertex, < 0 on error
  int addKeyframe(int oldkey_id, const double T_newkey_from_oldkey[7], const std::vector<int>& new_anchor,
                  const std::vector<double>& new_xyz_anchor, const std::vector<double>& new_anchor_center,
                  const std::vector<int>& new_anchor_level, const std::vector<double>& new_center, const std::vector<int>& new_level,
                  const std::vector<int>& track_point, const std::vector<double>& track_center, const std::vector<int>& track_level) {
    if (!ok_) return SVS_ERR_NOGPU;
    int v = -1, q = -1;
    const int rc = svs_map_add_keyframe(h_, oldkey_id, T_newkey_from_oldkey, (int)new_anchor.size(), new_anchor.data(),
                                        new_xyz_anchor.data(), new_anchor_center.data(), new_anchor_level.data(), new_center.data(),
                                        new_level.data(), (int)track_point.size(), track_point.data(), track_center.data(),
                                        track_level.data(), &v, &q);
    if (rc != SVS_OK) return rc;
    V_ += 1; Np_ += (int)new_anchor.size(); nn_ = 0;
    return v;
  }
  bool get(std::vector<double>* T_me_from_world, std::vector<double>* xyz_anchor) {
    T_me_from_world->resize(7 * (size_t)V_); xyz_anchor->resize(3 * (size_t)(Np_ > 0 ? Np_ : 1));
    const bool r = ok_ && svs_map_get(h_, T_me_from_world->data(), xyz_anchor->data()) == SVS_OK;
    xyz_anchor->resize(3 * (size_t)Np_);
    return r;
  }

 private:
  svs_map* h_ = nullptr;
  bool ok_ = false;
  int V_ = 0, Np_ = 0, nn_ = 0;
};

}  // namespace svs
#endif
