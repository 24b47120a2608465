/* oracle/fast_oracle.h -- CPU restatement of FastGrid (fast_grid.cpp); test infrastructure only. */
#ifndef SVS_FAST_ORACLE_H
#define SVS_FAST_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct { int u0, u1, v0, v1, thr; } ofast_cell;   /* cv::Range urange, vrange + fast_thr (keyframes.h:30-43) */
#define OFAST_MAX_CELLS 64
typedef struct {
  int grid_w, grid_h, fast_min, fast_max;
  int min_inner, min_outer, max_inner, max_outer;
  ofast_cell cells[OFAST_MAX_CELLS];
} ofast_grid;

int ofast_is_corner(const unsigned char *img, int pitch, int x, int y, int t);
int ofast_score(const unsigned char *img, int pitch, int x, int y);
int ofast_detect_roi(const unsigned char *img, int pitch, int u0, int u1, int v0, int v1, int thr,
                     int *out_xy, int max_out);
void ofast_grid_init(ofast_grid *g, int img_w, int img_h, int num_features_per_cell, int boundary_per_cell,
                     int fast_thr, int grid_w, int grid_h, int fast_min, int fast_max);
int ofast_detect(const unsigned char *img, int pitch, const ofast_cell *cells, int ncells,
                 int *out_xy, int max_out, int *cell_off);
int ofast_detect_adaptively(const unsigned char *img, int pitch, ofast_grid *g, int trials,
                            int *out_xy, int max_out, int *cell_off);
#ifdef __cplusplus
}
#endif
#endif
