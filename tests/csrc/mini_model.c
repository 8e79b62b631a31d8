/*
 * mini_model.c -- test driver: a C program that calls samodel() the way the reference's
 * run_model_lee_semi_analytical() does (model/bam.c:3103-3241): grids allocated one malloc per row, a scene table,
 * the DEPTHS grid passed by value, ten caller-owned output grids. Links the drop-in (libsamodel_b200.so) instead of
 * model/samodel.c. Used by tests/test_host_shim.py to check the failure behaviour on a machine without a CUDA
 * device: the reference's convention, printf + exit(1) (model/common.h:62-67), and no CPU fallback.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "photic_abi.h"

void samodel(photic_scene scene_data[], photic_geogrid gridded_data[], int *scene_indexes, int nscenes,
             photic_bool empirical_depth_present, photic_geogrid empirical_depths, int n_smoothing_radius,
             int n_spatial, int n_bottoms, float **depth, float **depth_sigma, float **model_error,
             float **bottom_albedo, float **bottom_sand, float **bottom_seagrass, float **bottom_coral, float **K_min,
             float **bottom_type, float **index_optical_depth, float pagesize, int background, int linewidth);

static float **grid(int nrows, int ncols, float v) {
  float **g = (float **)malloc(nrows * sizeof(float *));
  for (int i = 0; i < nrows; i++) {
    g[i] = (float *)malloc(ncols * sizeof(float));
    for (int j = 0; j < ncols; j++) g[i][j] = v;
  }
  return g;
}

int main(void) {
  enum { R = 4, C = 5, NS = 2 };
  static photic_scene scenes[NS];
  static photic_geogrid grids[NS * 4 + 1];
  static const int wl[4] = {443, 482, 561, 655};
  static const float rrs[4] = {0.0065f, 0.0082f, 0.0071f, 0.0016f};
  int idx[NS] = {0, 1};
  for (int s = 0; s < NS; s++) {
    snprintf(scenes[s].scene_name, sizeof(scenes[s].scene_name), "date%d", s);
    scenes[s].n_bands = 4; scenes[s].nrows = R; scenes[s].ncols = C;
    scenes[s].theta_v = 0.0; scenes[s].theta_w = 25.0 + 3.0 * s; scenes[s].H_tide = 0.1 * s;
    for (int b = 0; b < 4; b++) {
      photic_geogrid *g = &grids[4 * s + b];
      scenes[s].band_indexes[b] = 4 * s + b; scenes[s].wavelengths[b] = wl[b]; scenes[s].R_sigma[b] = 1.0e-4;
      g->nrows = R; g->ncols = C; g->cellsize = 30.0f; g->nodata_value = -9999.0f;
      g->array = grid(R, C, rrs[b] * (1.0f + 0.01f * s));
    }
  }
  photic_geogrid *pr = &grids[NS * 4];
  pr->nrows = R; pr->ncols = C; pr->cellsize = 30.0f; pr->nodata_value = -9999.0f; pr->array = grid(R, C, -6.5f);
  float **out[10];
  for (int k = 0; k < 10; k++) out[k] = grid(R, C, 7.0f);
  samodel(scenes, grids, idx, NS, 1, *pr, 1, 2, 3, out[0], out[1], out[2], out[3], out[4], out[5], out[6], out[7],
          out[8], out[9], 8.0f, 0, 1);
  printf("\nmini_model: depth[1][2] = %.6f, sigma[1][2] = %.6f\n", out[0][1][2], out[1][1][2]);
  return 0;
}
