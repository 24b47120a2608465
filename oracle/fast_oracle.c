/*
 * oracle/fast_oracle.c -- CPU restatement of ScaViSLAM's grid FAST detector.
 * TEST INFRASTRUCTURE ONLY (see ba_oracle.h for the rules).
 *
 * Follows scavislam/fast_grid.cpp:23-58 (cell layout), :60-83 (FastGrid::detect) and
 * :86-152 (FastGrid::detectAdaptively).  The FAST-9/16 segment test itself lives in
 * OpenCV 2.4.2 (cv::FastFeatureDetector(threshold, nonmaxSuppression=false), absent here);
 * its published definition is restated in ofast_is_corner and pinned in tests/ against
 * Python cv2 4.13's FastFeatureDetector (same test, same raster order).
 */
#include "fast_oracle.h"

#include <stdlib.h>
#include <string.h>

/* Bresenham circle of radius 3, OpenCV order */
static const int kDx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
static const int kDy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

/* p is a corner iff >= 9 contiguous circle pixels are all > p + t or all < p - t (strict) */
int ofast_is_corner(const unsigned char *img, int pitch, int x, int y, int t) {
  const int v = img[y * pitch + x];
  int br = 0, dk = 0; /* 16-bit masks */
  for (int k = 0; k < 16; ++k) {
    const int c = img[(y + kDy[k]) * pitch + x + kDx[k]];
    if (c > v + t) br |= 1 << k;
    if (c < v - t) dk |= 1 << k;
  }
  for (int pass = 0; pass < 2; ++pass) {
    const unsigned m = pass ? (unsigned)dk : (unsigned)br;
    const unsigned xx = m | (m << 16);
    int run = 0;
    for (int k = 0; k < 25; ++k) {
      if (xx & (1u << k)) { if (++run >= 9) return 1; } else run = 0;
    }
  }
  return 0;
}

/* max threshold at which (x,y) is still a corner, -1 if it is not a corner even at t = 0 */
int ofast_score(const unsigned char *img, int pitch, int x, int y) {
  int lo = -1;
  for (int t = 0; t < 256; ++t) {
    if (ofast_is_corner(img, pitch, x, y, t)) lo = t; else break;
  }
  return lo;
}

/* cv::FAST on the ROI [u0,u1) x [v0,v1): rows 3..rows-4, cols 3..cols-4 of the ROI, raster order */
int ofast_detect_roi(const unsigned char *img, int pitch, int u0, int u1, int v0, int v1, int thr,
                     int *out_xy, int max_out) {
  int n = 0;
  for (int y = v0 + 3; y < v1 - 3; ++y)
    for (int x = u0 + 3; x < u1 - 3; ++x)
      if (ofast_is_corner(img, pitch, x, y, thr)) {
        if (n < max_out) { out_xy[2 * n] = x; out_xy[2 * n + 1] = y; }
        ++n;
      }
  return n;
}

/* fast_grid.cpp:23-58 */
void ofast_grid_init(ofast_grid *g, int img_w, int img_h, int num_features_per_cell, int boundary_per_cell,
                     int fast_thr, int grid_w, int grid_h, int fast_min, int fast_max) {
  g->grid_w = grid_w; g->grid_h = grid_h;
  g->fast_min = fast_min; g->fast_max = fast_max;
  g->min_inner = (int)(num_features_per_cell - boundary_per_cell * 0.33);
  g->min_outer = num_features_per_cell - boundary_per_cell;
  g->max_inner = (int)(num_features_per_cell + boundary_per_cell * 0.33);
  g->max_outer = num_features_per_cell + boundary_per_cell;
  const int cw = img_w / grid_w, ch = img_h / grid_h;
  for (int j = 0; j < grid_h; ++j)
    for (int i = 0; i < grid_w; ++i) {
      ofast_cell *c = &g->cells[j * grid_w + i];
      c->u0 = i * cw; c->u1 = i * cw + cw;
      c->v0 = j * ch; c->v1 = j * ch + ch;
      c->thr = fast_thr;
    }
}

/* fast_grid.cpp:60-83.  out_xy grouped by cell in (j, i) order; cell_off[ncells+1]. */
int ofast_detect(const unsigned char *img, int pitch, const ofast_cell *cells, int ncells,
                 int *out_xy, int max_out, int *cell_off) {
  int n = 0;
  for (int c = 0; c < ncells; ++c) {
    cell_off[c] = n;
    const int room = max_out - n > 0 ? max_out - n : 0;
    n += ofast_detect_roi(img, pitch, cells[c].u0, cells[c].u1, cells[c].v0, cells[c].v1, cells[c].thr,
                          out_xy + 2 * (n < max_out ? n : 0), n < max_out ? room : 0);
  }
  cell_off[ncells] = n;
  return n;
}

/* fast_grid.cpp:86-152, literally (including prev_thr / prev_prev_thr being shared by the
 * cells of one grid row). */
int ofast_detect_adaptively(const unsigned char *img, int pitch, ofast_grid *g, int trials,
                            int *out_xy, int max_out, int *cell_off) {
  int n = 0;
  int *tmp = malloc(sizeof(int) * 2 * (size_t)(max_out > 0 ? max_out : 1));
  for (int j = 0; j < g->grid_h; ++j) {
    int prev_thr = -1, prev_prev_thr = -2;
    for (int i = 0; i < g->grid_w; ++i) {
      ofast_cell *c = &g->cells[j * g->grid_w + i];
      int nd = 0;
      for (int trial = 0; trial < trials; ++trial) {
        nd = ofast_detect_roi(img, pitch, c->u0, c->u1, c->v0, c->v1, c->thr, tmp, max_out);
        if (prev_prev_thr == c->thr) {
          c->thr = (c->thr + prev_prev_thr) / 2;
          break;
        }
        prev_prev_thr = prev_thr;
        prev_thr = c->thr;
        if (nd < g->min_inner) {
          if (c->thr <= g->fast_min) break;
          --c->thr;
          if (nd < g->min_outer) {
            if (c->thr <= g->fast_min) break;
            --c->thr;
            continue;
          }
        } else if (nd > g->max_inner) {
          if (c->thr >= g->fast_max) break;
          ++c->thr;
          if (nd > g->max_outer) {
            if (c->thr >= g->fast_max) break;
            ++c->thr;
            continue;
          }
        }
        break;
      }
      cell_off[j * g->grid_w + i] = n;
      if (trials <= 0) nd = 0;
      for (int k = 0; k < nd; ++k) {
        if (n < max_out && k < max_out) { out_xy[2 * n] = tmp[2 * k]; out_xy[2 * n + 1] = tmp[2 * k + 1]; }
        ++n;
      }
    }
  }
  cell_off[g->grid_w * g->grid_h] = n;
  free(tmp);
  return n;
}
