/*
 * photic_abi.h -- ABI mirror of the two reference structs that cross the samodel() boundary.
 *
 * When samodel_b200.c is compiled inside the reference tree (-DPHOTIC_REFERENCE_TREE) it includes the
 * reference's own "samodel.h" and this file is not used. Stand-alone (this repo's tests), these
 * declarations reproduce the memory layout of `geogrid` (model/common.h:69-84) and `scene`
 * (model/common.h:194-218) so that a caller built against the reference headers can call the shim.
 * tests/test_host_shim.py checks sizeof/offsetof against the real headers where they are present.
 * Note: the reference's `bool` is `typedef int bool` (model/common.h:58), not _Bool.
 */
#ifndef PHOTIC_ABI_H_
#define PHOTIC_ABI_H_

#define PHOTIC_MAX_STRING_LEN 2048 /* MAX_STRING_LEN, model/common.h:47 */
#define PHOTIC_MAX_GRIDS 265       /* MAX_GRIDS,      model/common.h:48 */

typedef int photic_bool;

typedef struct photic_geogrid {
  int nrows, ncols;
  float cellsize, wlon, slat, elon, nlat;
  float nodata_value;
  float lambda, theta_v, theta_w;
  float **array; /* one malloc per row (model/common.c:562-572) */
} photic_geogrid;

typedef struct photic_scene {
  char scene_name[PHOTIC_MAX_STRING_LEN];
  int n_bands, nrows, ncols;
  int band_indexes[PHOTIC_MAX_GRIDS]; /* indices into gridded_data[] */
  int wavelengths[PHOTIC_MAX_GRIDS];  /* integer nm */
  double theta_w, theta_v, H_tide;
  double R_inf[PHOTIC_MAX_GRIDS], R_sigma[PHOTIC_MAX_GRIDS], K[PHOTIC_MAX_GRIDS], K_sigma[PHOTIC_MAX_GRIDS];
  double ratio_min, ratio_max, slope_min, slope_max;
  photic_bool pgx_present;
  int ps_p, ps_g, ps_x;
} photic_scene;

#endif
