/*
 * oracle/match_oracle.c -- CPU restatement of ScaViSLAM's guided patch matcher.
 * TEST INFRASTRUCTURE ONLY (see ba_oracle.h).  PARITY UNPINNED (no reference vectors; the
 * VisionTools helpers LinearCamera::map/unmap/isInFrame and zeroFromPyr_* are absent and
 * restated from their usage).
 *
 * Follows
 *   scavislam/matcher.cpp:42-74    matchPatchZeroMeanSSD (literal integer formula)
 *   scavislam/matcher.cpp:77-96    computePatchScores
 *   scavislam/matcher.cpp:98-142   computePrediction
 *   scavislam/matcher.cpp:144-181  matchCandidates
 *   scavislam/matcher.cpp:183-216  returnBestMatch (sub-pixel refinement is compiled out, :249-308)
 *   scavislam/matcher.cpp:312-398  GuidedMatcher::match
 *   scavislam/matcher.cpp:403-458  warpAffinve
 *   scavislam/matcher-impl.cpp:32-51  createObervation<StereoCamera>
 *   scavislam/maths_utils.cpp:37-44   interpolateDisparity
 *   scavislam/quadtree.h:511-545, 615-710  point quadtree insert / query (candidate order)
 */
#include "match_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>



/* SE3 helpers local to this translation unit: it is compiled with -ffp-contract=off so that the
 * double-precision chain prediction -> affine warp -> uint8 truncation is bit-identical with the
 * CUDA path (compiled with -fmad=false); same formulas as ba_oracle.c / se3_dev.cuh. */
static void m_quat_to_R(const double q[4], double R[9]) {
  const double x = q[0], y = q[1], z = q[2], w = q[3];
  const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w;
  const double txx = tx * x, txy = ty * x, txz = tz * x;
  const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
  R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}
static void m_mat3_vec(const double R[9], const double x[3], double y[3]) {
  y[0] = R[0] * x[0] + R[1] * x[1] + R[2] * x[2];
  y[1] = R[3] * x[0] + R[4] * x[1] + R[5] * x[2];
  y[2] = R[6] * x[0] + R[7] * x[1] + R[8] * x[2];
}
static void m_se3_act(const double A[7], const double x[3], double y[3]) {
  double R[9];
  m_quat_to_R(A, R);
  m_mat3_vec(R, x, y);
  y[0] += A[4]; y[1] += A[5]; y[2] += A[6];
}
static void m_se3_mul(const double A[7], const double B[7], double AB[7]) {
  double R[9], t[3], q[4];
  m_quat_to_R(A, R);
  m_mat3_vec(R, B + 4, t);
  const double ax = A[0], ay = A[1], az = A[2], aw = A[3];
  const double bx = B[0], by = B[1], bz = B[2], bw = B[3];
  q[3] = aw * bw - ax * bx - ay * by - az * bz;
  q[0] = aw * bx + ax * bw + ay * bz - az * by;
  q[1] = aw * by + ay * bw + az * bx - ax * bz;
  q[2] = aw * bz + az * bw + ax * by - ay * bx;
  const double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  AB[0] = q[0] / n; AB[1] = q[1] / n; AB[2] = q[2] / n; AB[3] = q[3] / n;
  AB[4] = A[4] + t[0]; AB[5] = A[5] + t[1]; AB[6] = A[6] + t[2];
}
static void m_se3_inv(const double A[7], double Ai[7]) {
  const double q[4] = {-A[0], -A[1], -A[2], A[3]};
  const double mt[3] = {-A[4], -A[5], -A[6]};
  double R[9], t[3];
  m_quat_to_R(q, R);
  m_mat3_vec(R, mt, t);
  Ai[0] = q[0]; Ai[1] = q[1]; Ai[2] = q[2]; Ai[3] = q[3];
  Ai[4] = t[0]; Ai[5] = t[1]; Ai[6] = t[2];
}

/* ---------------------------------------------------------------- quadtree (capacity 1 leaves) */

typedef struct qnode {
  double x, y, w, h;
  struct qnode *ch[4]; /* xy, xY, Xy, XY */
  int has_children, empty;
  double px, py;
  int content;
} qnode;

struct omatch_tree {
  qnode *root;
  double delta;
};

static qnode *qnew(double x, double y, double w, double h) {
  qnode *n = calloc(1, sizeof *n);
  n->x = x; n->y = y; n->w = w; n->h = h; n->empty = 1;
  return n;
}

static void qfree(qnode *n) {
  if (!n) return;
  if (n->has_children) for (int i = 0; i < 4; ++i) qfree(n->ch[i]);
  free(n);
}

static int qinsert(qnode *n, double px, double py, int content, double delta);

static int qchildren_insert(qnode *n, double px, double py, int content, double delta) {
  /* quadtree.h:511-545 */
  const double rel_x = 1 - (n->x + n->w - px) / n->w;
  const double rel_y = 1 - (n->y + n->h - py) / n->h;
  if (rel_x < 0.5 && rel_y < 0.5) return qinsert(n->ch[0], px, py, content, delta);
  else if (rel_x >= 0.5 && rel_y < 0.5) return qinsert(n->ch[2], px, py, content, delta);
  else if (rel_x < 0.5 && rel_y >= 0.5) return qinsert(n->ch[1], px, py, content, delta);
  return qinsert(n->ch[3], px, py, content, delta);
}

static int qinsert(qnode *n, double px, double py, int content, double delta) {
  /* quadtree.h:615-676 */
  if (!n->has_children) {
    if (n->empty) {
      n->px = px; n->py = py; n->content = content; n->empty = 0;
      return 1;
    }
    if (hypot(n->px - px, n->py - py) < delta) return 0;
    const double x0 = n->x, x1 = n->x + n->w * 0.5, y0 = n->y, y1 = n->y + n->h * 0.5;
    const double w = n->w * 0.5, h = n->h * 0.5;
    n->ch[0] = qnew(x0, y0, w, h);
    n->ch[1] = qnew(x0, y1, w, h);
    n->ch[2] = qnew(x1, y0, w, h);
    n->ch[3] = qnew(x1, y1, w, h);
    n->has_children = 1;
    qchildren_insert(n, n->px, n->py, n->content, delta);
    return qchildren_insert(n, px, py, content, delta);
  }
  return qchildren_insert(n, px, py, content, delta);
}

static int rect_contains(double x, double y, double w, double h, double px, double py) {
  return x <= px && px < x + w && y <= py && py < y + h; /* cv::Rect_::contains */
}

static int rect_intersects(const qnode *a, double x, double y, double w, double h) {
  /* quadtree.h:548-560 */
  if (a->y + a->h <= y) return 0;
  if (a->y >= y + h) return 0;
  if (a->x + a->w <= x) return 0;
  if (a->x >= x + w) return 0;
  return 1;
}

typedef struct { int x, y, content; } qhit;

static void qquery(const qnode *n, double x, double y, double w, double h, qhit *out, int *nout, int max_out) {
  /* quadtree.h:679-710 */
  if (!n->has_children) {
    if (!n->empty && rect_contains(x, y, w, h, n->px, n->py)) {
      if (*nout < max_out) { out[*nout].x = (int)n->px; out[*nout].y = (int)n->py; out[*nout].content = n->content; }
      ++*nout;
    }
    return;
  }
  for (int i = 0; i < 4; ++i)
    if (rect_intersects(n->ch[i], x, y, w, h)) qquery(n->ch[i], x, y, w, h, out, nout, max_out);
}

omatch_tree *omatch_tree_build(int width, int height, const int *xy, const int *content, int n) {
  /* stereo_frontend.cpp:668-671: QuadTree<int>(Rectangle(0,0,width,height), 1) */
  omatch_tree *t = malloc(sizeof *t);
  t->root = qnew(0, 0, width, height);
  t->delta = 1;
  for (int i = 0; i < n; ++i) qinsert(t->root, xy[2 * i], xy[2 * i + 1], content ? content[i] : i, t->delta);
  return t;
}

void omatch_tree_free(omatch_tree *t) {
  if (!t) return;
  qfree(t->root);
  free(t);
}

int omatch_tree_query(const omatch_tree *t, int x, int y, int w, int h, int *out_xyc, int max_out) {
  qhit *tmp = malloc(sizeof(qhit) * (size_t)(max_out > 0 ? max_out : 1));
  int n = 0;
  qquery(t->root, x, y, w, h, tmp, &n, max_out);
  for (int i = 0; i < n && i < max_out; ++i) { out_xyc[3 * i] = tmp[i].x; out_xyc[3 * i + 1] = tmp[i].y; out_xyc[3 * i + 2] = tmp[i].content; }
  free(tmp);
  return n;
}

/* ---------------------------------------------------------------- matcher */

static int in_frame(const omatch_level *L, int u, int v, int border) {
  /* VisionTools LinearCamera::isInFrame (restated) */
  return u >= border && u < L->w - border && v >= border && v < L->h - border;
}

static void cam_map(const omatch_level *L, const double xyz[3], double uv[2]) {
  uv[0] = L->f * (xyz[0] / xyz[2]) + L->px;
  uv[1] = L->f * (xyz[1] / xyz[2]) + L->py;
}

/* matcher.cpp:403-458 -- 10x10 uint8 patch, row-major patch[iy*10+ix] */
void omatch_warp_affine(const unsigned char *frame, int pitch, const omatch_level *L, const double T_c2_from_c1[7],
                        double depth, const double key_uv[2], int halfpatch, unsigned char *patch) {
  double p[3], q[3], f0[2], fu[2], fv[2];
  const double offs[3][2] = {{0, 0}, {1, 0}, {0, 1}};
  double *outs[3] = {f0, fu, fv};
  for (int k = 0; k < 3; ++k) {
    const double ux = (key_uv[0] + offs[k][0] - L->px) / L->f, uy = (key_uv[1] + offs[k][1] - L->py) / L->f;
    p[0] = depth * ux; p[1] = depth * uy; p[2] = depth * 1.;
    m_se3_act(T_c2_from_c1, p, q);
    cam_map(L, q, outs[k]);
  }
  const double a00 = fu[0] - f0[0], a01 = fu[1] - f0[1], a10 = fv[0] - f0[0], a11 = fv[1] - f0[1];
  const double det = a00 * a11 - a01 * a10;
  const double idet = 1. / det;   /* Eigen 2x2 inverse: adjugate / determinant */
  const double i00 = a11 * idet, i01 = -a01 * idet, i10 = -a10 * idet, i11 = a00 * idet;
  const int ps = halfpatch * 2;
  for (int ix = 0; ix < ps; ++ix)
    for (int iy = 0; iy < ps; ++iy) {
      const double dx = ix - halfpatch, dy = iy - halfpatch;
      const double rx = (i00 * dx + i01 * dy) + key_uv[0];
      const double ry = (i10 * dx + i11 * dy) + key_uv[1];
      const double x = floor(rx), y = floor(ry);
      unsigned char val;
      if (x < 0 || y < 0 || x + 1 >= L->w || y + 1 >= L->h) {
        val = 0;
      } else {
        const double sx = rx - x, sy = ry - y;
        const double wx0 = 1 - sx, wx1 = sx, wy0 = 1 - sy, wy1 = sy;
        const int xi = (int)x, yi = (int)y;
        const double v00 = frame[yi * pitch + xi], v01 = frame[(yi + 1) * pitch + xi];
        const double v10 = frame[yi * pitch + xi + 1], v11 = frame[(yi + 1) * pitch + xi + 1];
        const double s = (wx0 * wy0) * v00 + (wx0 * wy1) * v01 + (wx1 * wy0) * v10 + (wx1 * wy1) * v11;
        val = (unsigned char)(s < 255. ? s : 255.);
      }
      patch[iy * ps + ix] = val;
    }
}

int omatch_znssd(const unsigned char *key8x8, const unsigned char *cur, int cur_pitch, int sumA, int sumAA) {
  /* matcher.cpp:42-74, literal formula with truncating int division */
  unsigned sumB = 0, sumBB = 0, sumAB = 0;
  for (int r = 0; r < 8; ++r)
    for (int c = 0; c < 8; ++c) {
      const unsigned b = cur[r * cur_pitch + c];
      sumB += b; sumBB += b * b; sumAB += b * key8x8[r * 8 + c];
    }
  const int iB = (int)sumB, iBB = (int)sumBB, iAB = (int)sumAB;
  return sumAA - 2 * iAB - iBB - (sumA * sumA - 2 * sumA * iB - iB * iB) / 64;
}

/* GuidedMatcher<StereoCamera>::match (matcher.cpp:312-398) over n candidate points */
int omatch_match(const omatch_frame *cur, const omatch_keyframe *keyframes, int nkf,
                 const double T_cur_from_actkey[7], const double T_actkey_from_w[7],
                 const omatch_point *pts, int n, int search_radius, int thr_mean, int thr_std,
                 omatch_result *res) {
  double T_w_from_actkey[7], T_cur_from_w[7];
  m_se3_inv(T_actkey_from_w, T_w_from_actkey);
  m_se3_mul(T_cur_from_actkey, T_actkey_from_w, T_cur_from_w);
  int nmatched = 0;
  int *cand = malloc(sizeof(int) * 3 * 4096);
  for (int i = 0; i < n; ++i) {
    const omatch_point *ap = &pts[i];
    omatch_result *r = &res[i];
    memset(r, 0, sizeof *r);
    r->index = -1;
    if (ap->keyframe < 0 || ap->keyframe >= nkf) continue;   /* vertex_map.find(anchor_id) == end */
    const omatch_keyframe *kf = &keyframes[ap->keyframe];
    const int lv = ap->anchor_level;
    const omatch_level *L = &cur->levels[lv];
    /* computePrediction (matcher.cpp:98-142) */
    double Tai[7], T_cur_from_anchor[7], xyz_cur[3], uv_pyr[2];
    m_se3_inv(kf->T_me_from_w, Tai);
    m_se3_mul(T_cur_from_w, Tai, T_cur_from_anchor);
    m_se3_act(T_cur_from_anchor, ap->xyz_anchor, xyz_cur);
    cam_map(L, xyz_cur, uv_pyr);
    if (!in_frame(L, (int)ap->anchor_obs_pyr[0], (int)ap->anchor_obs_pyr[1], 4)) continue;
    const double depth_cur = 1. / xyz_cur[2], depth_anchor = 1. / ap->xyz_anchor[2];
    if (depth_cur > depth_anchor * 3 || depth_anchor > depth_cur * 3) continue;
    r->predicted = 1;
    const int ui = (int)uv_pyr[0], vi = (int)uv_pyr[1];
    const int D = search_radius * 2 + 1;
    int nc = omatch_tree_query(cur->trees[lv], ui - search_radius, vi - search_radius, D, D, cand, 4096);
    if (nc > 4096) nc = 4096;
    /* warpAffinve + key patch scores */
    unsigned char patch10[100], key[64];
    omatch_warp_affine(kf->pyr[lv], kf->pitch[lv], L, T_cur_from_anchor, ap->xyz_anchor[2], ap->anchor_obs_pyr, 5, patch10);
    int sumA = 0, sumAA = 0;
    for (int rr = 0; rr < 8; ++rr)
      for (int cc = 0; cc < 8; ++cc) {
        const int v = patch10[(rr + 1) * 10 + cc + 1];
        key[rr * 8 + cc] = (unsigned char)v; sumA += v; sumAA += v * v;
      }
    if (sumA * sumA - sumAA < (int)(thr_std * thr_std * 64)) continue;
    r->textured = 1;
    /* matchCandidates (matcher.cpp:144-181) */
    int min_dist = thr_mean * thr_mean * 64, index = -1, bu = 0, bv = 0;
    for (int c = 0; c < nc; ++c) {
      const int cu = cand[3 * c], cv = cand[3 * c + 1];
      if (!in_frame(L, cu, cv, 4 + 2)) continue;
      const int z = omatch_znssd(key, cur->pyr[lv] + (size_t)(cv - 4) * cur->pitch[lv] + (cu - 4), cur->pitch[lv], sumA, sumAA);
      if (z < min_dist) { min_dist = z; index = cand[3 * c + 2]; bu = cu; bv = cv; }
    }
    r->n_candidates = nc;
    if (index < 0) continue;
    r->index = index; r->min_dist = min_dist; r->uv_pyr[0] = bu; r->uv_pyr[1] = bv;
    /* createObervation (matcher-impl.cpp:32-51), interpolateDisparity (maths_utils.cpp:37-44) */
    const double inv_factor = 1. / (double)(1 << lv);
    const double dd = (double)cur->disp[(size_t)(bv << lv) * cur->disp_pitch + (bu << lv)] * inv_factor;
    if (!(dd > 0)) continue;
    const double s = (double)(1 << lv);   /* zeroFromPyr_3d (restated: x * 2^level) */
    r->obs[0] = (double)(float)bu * s; r->obs[1] = (double)(float)bv * s; r->obs[2] = ((double)(float)bu - dd) * s;
    /* xyz_actkey = (T_anchorkey_from_w * T_w_from_actkey)^-1 * xyz_anchor */
    double Tak[7], Taki[7];
    m_se3_mul(kf->T_me_from_w, T_w_from_actkey, Tak);
    m_se3_inv(Tak, Taki);
    m_se3_act(Taki, ap->xyz_anchor, r->xyz_actkey);
    r->matched = 1;
    ++nmatched;
  }
  free(cand);
  return nmatched;
}
