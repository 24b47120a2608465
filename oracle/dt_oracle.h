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
#ifdef __cplusplus
}
#endif
#endif
