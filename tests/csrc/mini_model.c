/*
 * mini_model.c -- test driver: a C program that calls samodel() the way the reference's
 * run_model_lee_semi_analytical() does (model/bam.c:3103-3241): grids allocated one malloc per row, a scene table,
 * the DEPTHS grid passed by value, ten caller-owned output grids, and a write_nc() in the program (model/nc.c:14; here
 * a stand-in that logs the file names and a checksum of every grid it is handed). Links the drop-in
 * (libsamodel_b200.so) instead of model/samodel.c. With -DPHOTIC_REFERENCE_TREE the types are the reference's own
 * (model/common.h through the shim's include path); otherwise photic_abi.h, the ABI mirror.
 *
 *   mini_model                         constant grids (tests the failure behaviour on a machine without a CUDA device:
 *                                      the reference's convention, printf + exit(1), and no CPU fallback)
 *   mini_model in.bin out.bin log.txt  in.bin: int32 R, C, NS, then NS*4 planes [R][C] float32, then the DEPTHS grid;
 *                                      out.bin: the ten output grids in samodel()'s argument order; log.txt: one line
 *                                      per write_nc call: file name, ncols, nrows, spval, sum of the grid
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifdef PHOTIC_REFERENCE_TREE
#include "samodel.h"
typedef geogrid photic_geogrid;
typedef scene photic_scene;
typedef bool photic_bool;
#else
#include "photic_abi.h"
void samodel(photic_scene scene_data[], photic_geogrid gridded_data[], int *scene_indexes, int nscenes,
             photic_bool empirical_depth_present, photic_geogrid empirical_depths, int n_smoothing_radius,
             int n_spatial, int n_bottoms, float **depth, float **depth_sigma, float **model_error,
             float **bottom_albedo, float **bottom_sand, float **bottom_seagrass, float **bottom_coral, float **K_min,
             float **bottom_type, float **index_optical_depth, float pagesize, int background, int linewidth);
#endif

static FILE *g_log = NULL;

#ifndef MINI_NO_WRITE_NC
void write_nc(char *file, float **grid, int ncols, int nrows, float *lons, float *lats, double spval) {
  double sum = 0.0;
  for (int i = 0; i < nrows; i++)
    for (int j = 0; j < ncols; j++) sum += (double)grid[i][j];
  if (g_log) fprintf(g_log, "%s %d %d %.1f %.17g %.9g %.9g\n", file, ncols, nrows, spval, sum, (double)lons[ncols - 1], (double)lats[nrows - 1]);
}
#endif

static float **grid(int nrows, int ncols, float v) {
  float **g = (float **)malloc(nrows * sizeof(float *));
  for (int i = 0; i < nrows; i++) {
    g[i] = (float *)malloc(ncols * sizeof(float));
    for (int j = 0; j < ncols; j++) g[i][j] = v;
  }
  return g;
}

int main(int argc, char **argv) {
  enum { MAXNS = 16 };
  static photic_scene scenes[MAXNS];
  static photic_geogrid grids[MAXNS * 4 + 1];
  static const int wl[4] = {443, 482, 561, 655};
  static const float rrs[4] = {0.0065f, 0.0082f, 0.0071f, 0.0016f};
  int idx[MAXNS], R = 4, C = 5, NS = 2;
  FILE *in = NULL;
  if (argc >= 4) {
    int hdr[3];
    in = fopen(argv[1], "rb");
    if (!in || fread(hdr, sizeof(int), 3, in) != 3) { printf("mini_model: cannot read %s\n", argv[1]); return 2; }
    R = hdr[0]; C = hdr[1]; NS = hdr[2];
    g_log = fopen(argv[3], "w");
  }
  for (int s = 0; s < NS; s++) {
    idx[s] = s;
    snprintf(scenes[s].scene_name, sizeof(scenes[s].scene_name), "  date%d \t", s); /* trim()med for the file names */
    scenes[s].n_bands = 4; scenes[s].nrows = R; scenes[s].ncols = C;
    scenes[s].theta_v = 0.0; scenes[s].theta_w = 25.0 + 3.0 * s; scenes[s].H_tide = 0.1 * s;
    for (int b = 0; b < 4; b++) {
      photic_geogrid *g = &grids[4 * s + b];
      scenes[s].band_indexes[b] = 4 * s + b; scenes[s].wavelengths[b] = wl[b]; scenes[s].R_sigma[b] = 1.0e-4;
      g->nrows = R; g->ncols = C; g->cellsize = 30.0f; g->wlon = 100.0f; g->slat = -20.0f; g->nodata_value = -9999.0f;
      g->array = grid(R, C, rrs[b] * (1.0f + 0.01f * s));
      if (in)
        for (int i = 0; i < R; i++)
          if (fread(g->array[i], sizeof(float), C, in) != (size_t)C) return 2;
    }
  }
  photic_geogrid *pr = &grids[NS * 4];
  pr->nrows = R; pr->ncols = C; pr->cellsize = 30.0f; pr->nodata_value = -9999.0f; pr->array = grid(R, C, -6.5f);
  if (in) {
    for (int i = 0; i < R; i++)
      if (fread(pr->array[i], sizeof(float), C, in) != (size_t)C) return 2;
    fclose(in);
  }
  float **out[10];
  for (int k = 0; k < 10; k++) out[k] = grid(R, C, 7.0f);
  samodel(scenes, grids, idx, NS, 1, *pr, 1, 2, 3, out[0], out[1], out[2], out[3], out[4], out[5], out[6], out[7],
          out[8], out[9], 8.0f, 0, 1);
  printf("\nmini_model: depth[1][2] = %.6f, sigma[1][2] = %.6f, scene 0 is now named '%s'\n", out[0][1][2], out[1][1][2],
         scenes[0].scene_name);
  if (argc >= 4) {
    FILE *o = fopen(argv[2], "wb");
    for (int k = 0; k < 10; k++)
      for (int i = 0; i < R; i++) fwrite(out[k][i], sizeof(float), C, o);
    fclose(o);
    if (g_log) fclose(g_log);
  }
  return 0;
}
