/* oracle/dt_oracle.h -- CPU restatement of the dense tracker (GPU semantics); test infrastructure only. */
#ifndef SVS_DT_ORACLE_H
#define SVS_DT_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

#define ODT_MAX_LEVELS 8
typedef struct {
  int w, h;
  int stride;             /* floats per row of prev/cur/dx/dy */
  int cloud_stride;       /* float4 per row of the reference point cloud */
  float f, px, py;        /* GpuIntrinsics of this pyramid level */
  const float *prev;      /* previous frame, float image */
  const float *cur, *dx, *dy;
  const float *cloud;     /* float4 per pixel: xyz in the active keyframe, w > 0 if valid */
} odt_level;

typedef struct {
  double chi2[ODT_MAX_LEVELS];
  int passes[ODT_MAX_LEVELS];
} odt_stats;

void odt_pose_to_m34(const double T[7], float m[12]);
void odt_pass(const odt_level *L, const double T[7], int exact_bilinear, double *chi2, double H21[21],
              double b6[6], int *n_valid);
void odt_track(const odt_level *levels, int nlevels, double T[7], int exact_bilinear, odt_stats *st);
void odt_point_cloud(const float TQ[16], const float *disp, int width, int height, int stride_in,
                     int stride_out, int factor, float *cloud);
void odt_make_TQ(const double T_cur_from_actkey[7], double f, double px, double py, double b, float TQ[16]);

/* ---- the tracker the reference builds WITHOUT SCAVISLAM_CUDA_SUPPORT:
 * DenseTracker::denseTrackingCpu / computeDensePointCloudCpu (dense_tracking.cpp:222-423).
 * Every 4th pixel in u and v, residual clamped to +-0.1, exact software bilinear taps, border test
 * isInFrame(uv, 2), H is NOT damped, a level ends on the first rejected step (the reference repeats the
 * identical trial once more and stops).  Sums are FP64 here (the reference: sequential FP32). */
typedef struct {
  int w, h;                 /* size of this pyramid level (the point grid is w/4 x h/4) */
  int stride;               /* floats per row of cur/dx/dy */
  int pitch_u8;             /* bytes per row of prev_u8 */
  double f, px, py, b;      /* StereoCamera of this level (frame_grabber-impl.cpp:50-59) */
  const unsigned char *prev_u8;   /* frame_data_.prev_left().pyr_uint8[level] */
  const float *cur, *dx, *dy;     /* pyr_float32, pyr_float32_dx, pyr_float32_dy */
  const float *cloud;             /* float4 per grid point, (w/4) per row */
} odtc_level;

/* computeDensePointCloudCpu for one level: disp = level-0 float disparity, cloud = (h/4) x (w/4) float4 */
void odtc_point_cloud(const double T_cur_from_actkey[7], double f, double px, double py, double b, const float *disp,
                      int disp_stride, int level, int w, int h, float *cloud);
/* one sweep at pose T: clamped chi2, H (upper triangle, 21) and Jres (6) */
void odtc_pass(const odtc_level *L, const double T[7], double *chi2, double H21[21], double Jres[6], int *n_valid);
/* denseTrackingCpu over levels nlevels-1 .. 0; T updated in place */
void odtc_track(const odtc_level *levels, int nlevels, double T[7], odt_stats *st);
#ifdef __cplusplus
}
#endif
#endif
