/* oracle/match_oracle.h -- CPU restatement of GuidedMatcher<StereoCamera>::match; test infrastructure only. */
#ifndef SVS_MATCH_ORACLE_H
#define SVS_MATCH_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

#define OMATCH_MAX_LEVELS 4
typedef struct omatch_tree omatch_tree;

typedef struct { int w, h; double f, px, py; } omatch_level;      /* cam_vec[level] */

typedef struct {                                                    /* cur_frame + feature_tree */
  omatch_level levels[OMATCH_MAX_LEVELS];
  const unsigned char *pyr[OMATCH_MAX_LEVELS];
  int pitch[OMATCH_MAX_LEVELS];
  const float *disp;                                                 /* level-0 disparity */
  int disp_pitch;                                                    /* floats per row */
  const omatch_tree *trees[OMATCH_MAX_LEVELS];
} omatch_frame;

typedef struct {                                                    /* keyframe_map / vertex_map entry */
  double T_me_from_w[7];
  const unsigned char *pyr[OMATCH_MAX_LEVELS];
  int pitch[OMATCH_MAX_LEVELS];
} omatch_keyframe;

typedef struct {                                                    /* CandidatePoint<3> */
  int keyframe;               /* index of the anchor keyframe, -1 = not in vertex_map */
  int anchor_level;
  double xyz_anchor[3];
  double anchor_obs_pyr[2];
} omatch_point;

typedef struct {
  int predicted, textured, matched;
  int n_candidates;
  int index;                  /* quadtree content of the best candidate, -1 = none */
  int min_dist;
  int uv_pyr[2];
  double obs[3];              /* (u, v, u_right) at level 0 */
  double xyz_actkey[3];
} omatch_result;

omatch_tree *omatch_tree_build(int width, int height, const int *xy, const int *content, int n);
void omatch_tree_free(omatch_tree *t);
int omatch_tree_query(const omatch_tree *t, int x, int y, int w, int h, int *out_xyc, int max_out);
void omatch_warp_affine(const unsigned char *frame, int pitch, const omatch_level *L, const double T_c2_from_c1[7],
                        double depth, const double key_uv[2], int halfpatch, unsigned char *patch);
int omatch_znssd(const unsigned char *key8x8, const unsigned char *cur, int cur_pitch, int sumA, int sumAA);
int omatch_match(const omatch_frame *cur, const omatch_keyframe *keyframes, int nkf,
                 const double T_cur_from_actkey[7], const double T_actkey_from_w[7],
                 const omatch_point *pts, int n, int search_radius, int thr_mean, int thr_std,
                 omatch_result *res);
#ifdef __cplusplus
}
#endif
#endif
