/* pose_oracle.h -- CPU restatement of ScaViSLAM's motion-only Levenberg-Marquardt
 * (PoseOptimizer<SE3,6,IdObs<3>,3>::calcFastMotionOnly).  TEST INFRASTRUCTURE ONLY: nothing under
 * scavislam_b200/ may include, link or call this.  PARITY UNPINNED: the reference ships no test or
 * golden vector for this function; it is pinned by finite differences and ground-truth recovery
 * (tests/test_pose_oracle.py). */
#ifndef SVS_POSE_ORACLE_H
#define SVS_POSE_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  double initial_chi2, chi2, max_err;
  int num_obs;
  int iterations; /* accepted steps */
  int trials;     /* solves */
  int nan_error;  /* the reference throws "Res is NaN!" */
} opo_stats;

/* SE3XYZ_STEREO::map (transformations.h:445-449) */
void opo_map(const double cam[4], const double T[7], const double xyz[3], double uvu[3]);
/* SE3XYZ_STEREO::frameJac (transformations.h:417-443), row-major 3x6 */
void opo_frame_jac(const double cam[4], const double T[7], const double xyz[3], double J[18]);
/* calcFastMotionOnly (pose_optimizer.h:135-298); cam = (f, px, py, baseline); T updated in place */
void opo_calc_fast_motion_only(int n, const int *obs_point_id, const double *obs_uvu, const double *point_xyz,
                               const double cam[4], int robust_kernel, double kernel_param, int num_iter,
                               double initial_mu, double tau, double T[7], opo_stats *st);

#ifdef __cplusplus
}
#endif
#endif
