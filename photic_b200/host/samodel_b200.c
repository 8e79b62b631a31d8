/*
 * samodel_b200.c -- drop-in replacement for model/samodel.c's entry point.
 *
 * Exports `samodel` with the reference's exact C signature (model/samodel.h:8-19; sole call site
 * model/bam.c:3236-3241) and forwards the inversion to libphotic_b200.so (include/photic_b200.h).
 * In the reference tree: compile this file INSTEAD of samodel.c with -DPHOTIC_REFERENCE_TREE and link
 * -lphotic_b200 (tests/test_host_shim.py builds exactly that against /root/reference/model).
 * NetCDF output, the REPL and graphics stay on the host, unchanged.
 *
 * What it does, in the order samodel() does it:
 *   1. reads the scene fields samodel() reads (bands, int wavelengths, angles, tide: samodel.c:389-548), checks that
 *      every grid of the call has the shape of the first one, and records every grid's own nodata_value (the
 *      reference tests band k against gridded_data[k].nodata_value: samodel.c:683, 941, 2999-3003)
 *   2. phb_invert_rows(): the caller's `float **` grids (one allocation per row, common.c:562-572) go to the device(s)
 *      row by row through the library's pinned staging ring and the results come back the same way -- no packed copy of
 *      the scene on the host. One device by default; PHOTIC_B200_DEVICES=all|"0,1,.." spreads the rows over several
 *      GPUs inside this one process (equal row bands; the devices share the work at run time over NVLink)
 *   3. depth_sigma from phb_depth_sigma_rows() (samodel.c:1376-1477; the reference seeds it with time(NULL), here
 *      PHOTIC_B200_SIGMA_SEED can fix the seed)
 *   4. writes the per-scene K/P/G/X grids and the ten model grids through the host's write_nc, the file names and the
 *      order of samodel.c:1513-1687, when the host program provides write_nc
 * Errors: the reference has no error channel (printf + exit(1), common.h:62-67); so does this shim.
 */
#ifdef PHOTIC_REFERENCE_TREE
#include "samodel.h"
typedef geogrid photic_geogrid;
typedef scene photic_scene;
typedef bool photic_bool;
#define PHOTIC_MAX_STRING_LEN MAX_STRING_LEN /* model/common.h:47 */
#else
#include "photic_abi.h"
#endif

#include <ctype.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "photic_b200.h"

/* host NetCDF writer of the reference (model/nc.c:14); optional at link time */
extern void write_nc(char *file, float **grid, int ncols, int nrows, float *lons, float *lats, double spval)
    __attribute__((weak));

static phb_ctx *g_ctx = NULL;             /* context of the first device (also runs the depth-error phase) */
static phb_ctx *g_ctxs[64];
static int g_n_ctx = 0;

/* Devices: PHOTIC_B200_DEVICES=all | "0,2,5" (one row band per device, one process), else
 * PHOTIC_B200_DEVICE=<ordinal> (default 0). Contexts are created once and kept for the next MODEL command. */
static void open_devices(void) {
  const char *many = getenv("PHOTIC_B200_DEVICES"), *one = getenv("PHOTIC_B200_DEVICE");
  int ids[64], n = 0, k, rc;
  if (g_n_ctx > 0) return;
  if (many && strcmp(many, "all") == 0) {
    const int cnt = phb_device_count();
    for (k = 0; k < cnt && k < 64; k++) ids[n++] = k;
  } else if (many && *many) {
    const char *p = many;
    while (*p && n < 64) {
      char *end;
      const long v = strtol(p, &end, 10);
      if (end == p) break;
      ids[n++] = (int)v;
      p = (*end == ',') ? end + 1 : end;
    }
  }
  if (n == 0) ids[n++] = one ? atoi(one) : 0;
  for (k = 0; k < n; k++) {
    rc = phb_ctx_create(ids[k], &g_ctxs[k]);
    if (rc) {
      printf("\n\nERROR: photic_b200: cannot open CUDA device %d (there is no CPU fallback): %s\n\n", ids[k],
             phb_error_string(rc));
      exit(1);
    }
  }
  g_n_ctx = n;
  g_ctx = g_ctxs[0];
}

/* Releases the device contexts; the next samodel() call opens them again (and re-reads the environment). */
void samodel_b200_shutdown(void) {
  int k;
  for (k = 0; k < g_n_ctx; k++) phb_ctx_destroy(g_ctxs[k]);
  g_n_ctx = 0;
  g_ctx = NULL;
}

static void die(const char *what, int rc) {
  printf("\n\nERROR: photic_b200: %s: %s\n\n", what, phb_error_string(rc));
  exit(1);
}

static void *must(void *p) {
  if (p == NULL) { printf("\n\nERROR: Out of memory.\n\n"); exit(1); }
  return p;
}

/* a grid held the way the reference holds its own (allocate_float_array_2d, common.c:562-572) */
static float **grid_alloc(int nrows, int ncols) {
  float **g = (float **)must(malloc((size_t)nrows * sizeof(float *)));
  int r;
  for (r = 0; r < nrows; r++) g[r] = (float *)must(calloc((size_t)ncols, sizeof(float)));
  return g;
}

static void grid_free(float **g, int nrows) {
  int r;
  for (r = 0; r < nrows; r++) free(g[r]);
  free(g);
}

#ifndef PHOTIC_REFERENCE_TREE
/* trim(), common.c:715-750: leading and trailing white space removed IN PLACE (the reference passes scene_name
 * itself, samodel.c:1516). In the reference tree the reference's own trim() is linked. */
static char *trim(char *str) {
  char *front = str, *end;
  size_t len;
  if (str == NULL || str[0] == '\0') return str;
  while (isspace((unsigned char)*front)) front++;
  len = strlen(front);
  memmove(str, front, len + 1);
  end = str + len;
  while (end > str && isspace((unsigned char)end[-1])) end--;
  *end = '\0';
  return str;
}
#endif

void samodel(photic_scene scene_data[], photic_geogrid gridded_data[], int *scene_indexes, int nscenes,
             photic_bool empirical_depth_present, photic_geogrid empirical_depths, int n_smoothing_radius,
             int n_spatial, int n_bottoms, float **depth, float **depth_sigma, float **model_error,
             float **bottom_albedo, float **bottom_sand, float **bottom_seagrass, float **bottom_coral, float **K_min,
             float **bottom_type, float **index_optical_depth, float pagesize, int background, int linewidth) {
  phb_scene_desc d;
  phb_row_outputs out;
  phb_stats st;
  const float *const *plane_rows[PHB_MAX_SCENES * PHB_MAX_BANDS];
  float **Kg[PHB_MAX_SCENES * PHB_MAX_BANDS], **Pg[PHB_MAX_SCENES], **Gg[PHB_MAX_SCENES], **Xg[PHB_MAX_SCENES];
  float *const *Kp[PHB_MAX_SCENES * PHB_MAX_BANDS], *const *Pp[PHB_MAX_SCENES], *const *Gp[PHB_MAX_SCENES],
      *const *Xp[PHB_MAX_SCENES];
  int s, b, g = 0, k, rc, nrows, ncols, max_bands = 0, r, c, nodata_differs = 0;
  (void)pagesize; (void)background; (void)linewidth;

  if (nscenes < 1 || nscenes > PHB_MAX_SCENES) die("number of scenes", PHB_EINVAL);
  memset(&d, 0, sizeof(d));
  nrows = scene_data[scene_indexes[0]].nrows;
  ncols = scene_data[scene_indexes[0]].ncols;
  printf("\nn_smoothing_radius = %d\nn_spatial = %d, \nn_bottoms = %d\n", n_smoothing_radius, n_spatial, n_bottoms);
  printf("\nnrows,ncols = %d,%d\n", nrows, ncols);
  d.n_scenes = nscenes;
  d.nodata = gridded_data[scene_data[scene_indexes[0]].band_indexes[0]].nodata_value;
  for (s = 0; s < nscenes; s++) {
    const photic_scene *sc = &scene_data[scene_indexes[s]];
    if (sc->n_bands < 2 || sc->n_bands > PHB_MAX_BANDS) die("bands per scene", PHB_EINVAL);
    d.n_bands[s] = sc->n_bands;
    if (sc->n_bands > max_bands) max_bands = sc->n_bands;
    d.theta_view[s] = sc->theta_v;
    d.theta_sun[s] = sc->theta_w;
    d.h_tide[s] = sc->H_tide;
    for (b = 0; b < sc->n_bands; b++, g++) {
      const photic_geogrid *gr = &gridded_data[sc->band_indexes[b]];
      /* the reference indexes every grid with the first scene's nrows / ncols (samodel.c:672-695): a smaller grid
       * would be read out of bounds there; here it is an error */
      if (gr->nrows != nrows || gr->ncols != ncols) {
        printf("\n\nERROR: photic_b200: grid %d (scene %d, band %d) is %d x %d, the first grid is %d x %d.\n\n",
               sc->band_indexes[b], s, b, gr->nrows, gr->ncols, nrows, ncols);
        exit(1);
      }
      d.wavelengths[s][b] = sc->wavelengths[b];
      d.r_sigma[s][b] = sc->R_sigma[b];
      d.nodata_band[s][b] = gr->nodata_value;
      if (gr->nodata_value != d.nodata) nodata_differs = 1;
      plane_rows[g] = (const float *const *)gr->array;
    }
  }
  d.nodata_per_band = nodata_differs;
  d.n_smoothing_radius = n_smoothing_radius;
  d.n_spatial = n_spatial;
  d.n_bottoms = n_bottoms;
  d.nrows = nrows;
  d.ncols = ncols;
  d.prior_present = empirical_depth_present ? 1 : 0;
  d.prior_nodata = empirical_depths.nodata_value;
  if (empirical_depth_present && (empirical_depths.nrows != nrows || empirical_depths.ncols != ncols)) {
    printf("\n\nERROR: photic_b200: the DEPTHS grid is %d x %d, the scene grids are %d x %d.\n\n", empirical_depths.nrows,
           empirical_depths.ncols, nrows, ncols);
    exit(1);
  }

  /* per-scene result grids, as samodel.c:622-630 allocates them: K per band, P, G, X (D stays out: DELTA 0) */
  for (s = 0; s < nscenes; s++) {
    for (b = 0; b < max_bands; b++) { Kg[s * max_bands + b] = grid_alloc(nrows, ncols); Kp[s * max_bands + b] = Kg[s * max_bands + b]; }
    Pg[s] = grid_alloc(nrows, ncols); Gg[s] = grid_alloc(nrows, ncols); Xg[s] = grid_alloc(nrows, ncols);
    Pp[s] = Pg[s]; Gp[s] = Gg[s]; Xp[s] = Xg[s];
  }
  memset(&out, 0, sizeof(out));
  out.depth = depth; out.model_error = model_error; out.bottom_albedo = bottom_albedo; out.bottom_sand = bottom_sand;
  out.bottom_seagrass = bottom_seagrass; out.bottom_coral = bottom_coral; out.K_min = K_min; out.bottom_type = bottom_type;
  out.index_optical_depth = index_optical_depth;
  out.K = Kp; out.P = Pp; out.G = Gp; out.X = Xp;

  open_devices();
  rc = phb_invert_rows(g_ctxs, g_n_ctx, &d, plane_rows, empirical_depth_present ? (const float *const *)empirical_depths.array : NULL,
                       &out, &st, NULL, NULL);
  if (rc) die("inversion failed", rc);
  printf("\nNumber of optically shallow pixels = %lld\n", (long long)st.n_valid);
  printf("\nGPU inversion on %d device(s): %.1f ms (%.0f px/sec), mean iterations = %.0f, diverged = %.2f (%%)\n", g_n_ctx, st.ms_solve,
         st.n_valid / (st.ms_solve * 1e-3 + 1e-12), st.n_valid ? (double)st.n_evals / st.n_valid : 0.0,
         st.n_valid ? 100.0 * (double)(st.n_valid - st.n_converged) / st.n_valid : 0.0);

  /* depth-error estimate, samodel.c:1376-1477. The reference seeds rand() with time(NULL) (samodel.c:371);
   * PHOTIC_B200_SIGMA_SEED fixes the seed. The trials run as the reference runs them -- ONE hot-start chain through
   * all depth intervals (samodel.c:1394-1456; inherently serial, ~6 ms per trial) -- unless
   * PHOTIC_B200_SIGMA_CHAIN=interval asks for one chain per depth interval, all intervals in parallel (~40x faster; the
   * chain then restarts cold at every interval, which the reference does not do). */
  if (empirical_depth_present) {
    const char *e_seed = getenv("PHOTIC_B200_SIGMA_SEED"), *e_chain = getenv("PHOTIC_B200_SIGMA_CHAIN");
    const unsigned seed = e_seed ? (unsigned)strtoul(e_seed, NULL, 10) : (unsigned)time(NULL);
    const int chain = (e_chain && strcmp(e_chain, "interval") == 0) ? PHB_SIGMA_CHAIN_PER_INTERVAL : PHB_SIGMA_CHAIN_REFERENCE;
    phb_stats st2;
    int32_t n_int = 0;
    printf("\nComputing depth error estimates...");
    rc = phb_depth_sigma_rows(g_ctx, &d, plane_rows, (const float *const *)empirical_depths.array, (const float *const *)depth,
                              seed, 128, chain, PHB_SIGMA_MAX_INTERVALS, depth_sigma, NULL, &n_int, NULL, &st2);
    if (rc) die("depth error estimate failed", rc);
    printf("...finished (%d depth intervals, %lld trials, %.1f ms).\n", (int)n_int, (long long)st2.n_valid, st2.ms_solve);
  } else {
    /* without a DEPTHS grid the reference still runs the trials, every one from h_empirical = 5 (samodel.c:1448) and
     * hot-started from whatever pixel its own loop visited last; that chain has no defined first link here */
    printf("\nDepth error estimates need a DEPTHS grid (see include/photic_b200.h): H_sigma left at 0.\n");
    for (r = 0; r < nrows; r++)
      for (c = 0; c < ncols; c++) depth_sigma[r][c] = 0.0f;
  }

  /* file side effects of samodel.c:1494-1687, through the host's own NetCDF writer when present: per scene
   * <scene>_K_{coastal,blue,green,red}.nc (always four: md->K[scene][0..3], samodel.c:1153-1156), _P, _G, _X
   * (_D only under DELTA, which is 0: samodel.c:1586-1596), then the ten modelled_*.nc in the reference's order */
  if (write_nc) {
    const photic_geogrid *g0 = &gridded_data[scene_data[scene_indexes[0]].band_indexes[0]];
    float *lons = (float *)must(malloc(ncols * sizeof(float))), *lats = (float *)must(malloc(nrows * sizeof(float)));
    static const char *band_name[4] = {"coastal", "blue", "green", "red"};
    float **zero = NULL; /* K of a band the scene does not have */
    char file_name[PHOTIC_MAX_STRING_LEN + 64];
    struct { const char *name; float **grid; } model[10] = {
        {"H", depth}, {"H_sigma", depth_sigma}, {"error", model_error}, {"albedo", bottom_albedo},
        {"bottom_sand", bottom_sand}, {"bottom_seagrass", bottom_seagrass}, {"bottom_coral", bottom_coral},
        {"min_K", K_min}, {"bottom_type", bottom_type}, {"index_optical_depth", index_optical_depth}};
    printf("\nWriting model data to file...");
    for (c = 0; c < ncols; c++) lons[c] = g0->wlon + ((float)c) * g0->cellsize;
    for (r = 0; r < nrows; r++) lats[r] = g0->slat + ((float)r) * g0->cellsize;
    for (s = 0; s < nscenes; s++) {
      char *nm = trim(scene_data[scene_indexes[s]].scene_name);
      for (b = 0; b < 4; b++) {
        float **kg;
        if (b < d.n_bands[s]) kg = Kg[s * max_bands + b];
        else { if (!zero) zero = grid_alloc(nrows, ncols); kg = zero; }
        snprintf(file_name, sizeof(file_name), "%s_K_%s.nc", nm, band_name[b]);
        printf("\nWriting file %s...", file_name);
        write_nc(file_name, kg, ncols, nrows, lons, lats, g0->nodata_value);
      }
      snprintf(file_name, sizeof(file_name), "%s_P.nc", nm);
      printf("\nWriting file %s...", file_name);
      write_nc(file_name, Pg[s], ncols, nrows, lons, lats, g0->nodata_value);
      snprintf(file_name, sizeof(file_name), "%s_G.nc", nm);
      printf("\nWriting file %s...", file_name);
      write_nc(file_name, Gg[s], ncols, nrows, lons, lats, g0->nodata_value);
      snprintf(file_name, sizeof(file_name), "%s_X.nc", nm);
      printf("\nWriting file %s...", file_name);
      write_nc(file_name, Xg[s], ncols, nrows, lons, lats, g0->nodata_value);
    }
    for (k = 0; k < 10; k++) {
      snprintf(file_name, sizeof(file_name), "modelled_%s.nc", model[k].name);
      printf("\nWriting file %s...", file_name);
      write_nc(file_name, model[k].grid, ncols, nrows, lons, lats, g0->nodata_value);
    }
    printf("\n... finished.\n");
    if (zero) grid_free(zero, nrows);
    free(lons); free(lats);
  }

  for (s = 0; s < nscenes; s++) {
    for (b = 0; b < max_bands; b++) grid_free(Kg[s * max_bands + b], nrows);
    grid_free(Pg[s], nrows); grid_free(Gg[s], nrows); grid_free(Xg[s], nrows);
  }
}
