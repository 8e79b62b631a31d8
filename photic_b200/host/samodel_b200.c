/*
 * samodel_b200.c -- drop-in replacement for model/samodel.c's entry point.
 *
 * Exports `samodel` with the reference's exact C signature (model/samodel.h:8-19; sole call site
 * model/bam.c:3236-3241) and forwards the inversion to libphotic_b200.so (include/photic_b200.h).
 * In the reference tree: compile this file INSTEAD of samodel.c with -DPHOTIC_REFERENCE_TREE and link
 * -lphotic_b200. NetCDF output, the REPL and graphics stay on the host, unchanged.
 *
 * What it does, in the order samodel() does it:
 *   1. reads the scene fields samodel() reads (bands, int wavelengths, angles, tide: samodel.c:389-548)
 *   2. packs the float** row-pointer grids into contiguous planes (the device wants [plane][row][col])
 *   3. phb_invert_host() -- or, with PHOTIC_B200_DEVICES=all|"0,1,..", phb_invert_host_multi(): cost-balanced
 *      row bands over the GPUs of the box, one host thread per device inside this one process -- validity
 *      scan, per-pixel cold-start inversion on the GPU, output defaults
 *   4. scatters the 9 result planes back into the caller's float** grids (depth already negated,
 *      samodel.c:1486-1490); depth_sigma comes from phb_depth_sigma_host() (samodel.c:1376-1477; the
 *      reference seeds it with time(NULL), here PHOTIC_B200_SIGMA_SEED can fix the seed)
 *   5. writes the per-scene K/P/G/X grids and the ten model grids through the host's write_nc, exactly
 *      the file names of samodel.c:1513-1687, when the host program provides write_nc
 * Errors: the reference has no error channel (printf + exit(1), common.h:62-67); so does this shim.
 */
#ifdef PHOTIC_REFERENCE_TREE
#include "samodel.h"
typedef geogrid photic_geogrid;
typedef scene photic_scene;
typedef bool photic_bool;
#else
#include "photic_abi.h"
#endif

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

/* Devices: PHOTIC_B200_DEVICES=all | "0,2,5" (one row band per device, one process: phb_invert_host_multi),
 * else PHOTIC_B200_DEVICE=<ordinal> (default 0). Contexts are created once and kept for the next MODEL command. */
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

static float *pack_rows(float **rows, int nrows, int ncols) {
  float *p = (float *)malloc((size_t)nrows * ncols * sizeof(float));
  int r;
  if (p == NULL) { printf("\n\nERROR: Out of memory.\n\n"); exit(1); }
  for (r = 0; r < nrows; r++) memcpy(p + (size_t)r * ncols, rows[r], (size_t)ncols * sizeof(float));
  return p;
}

static void unpack_rows(const float *p, float **rows, int nrows, int ncols) {
  int r;
  for (r = 0; r < nrows; r++) memcpy(rows[r], p + (size_t)r * ncols, (size_t)ncols * sizeof(float));
}

static float **row_view(float *p, int nrows, int ncols) {
  float **v = (float **)malloc((size_t)nrows * sizeof(float *));
  int r;
  for (r = 0; r < nrows; r++) v[r] = p + (size_t)r * ncols;
  return v;
}

void samodel(photic_scene scene_data[], photic_geogrid gridded_data[], int *scene_indexes, int nscenes,
             photic_bool empirical_depth_present, photic_geogrid empirical_depths, int n_smoothing_radius,
             int n_spatial, int n_bottoms, float **depth, float **depth_sigma, float **model_error,
             float **bottom_albedo, float **bottom_sand, float **bottom_seagrass, float **bottom_coral, float **K_min,
             float **bottom_type, float **index_optical_depth, float pagesize, int background, int linewidth) {
  phb_scene_desc d;
  phb_outputs out;
  phb_stats st;
  const float *planes[PHB_MAX_SCENES * PHB_MAX_BANDS];
  float *packed[PHB_MAX_SCENES * PHB_MAX_BANDS], *prior = NULL, *res[9], *Kp, *Pp, *Gp, *Xp;
  float **grids9[9];
  int s, b, g = 0, k, rc, nrows, ncols, max_bands = 0, r, c;
  size_t px;
  (void)pagesize; (void)background; (void)linewidth;

  if (nscenes < 1 || nscenes > PHB_MAX_SCENES) die("number of scenes", PHB_EINVAL);
  memset(&d, 0, sizeof(d));
  nrows = scene_data[scene_indexes[0]].nrows;
  ncols = scene_data[scene_indexes[0]].ncols;
  px = (size_t)nrows * ncols;
  printf("\nn_smoothing_radius = %d\nn_spatial = %d, \nn_bottoms = %d\n", n_smoothing_radius, n_spatial, n_bottoms);
  printf("\nnrows,ncols = %d,%d\n", nrows, ncols);
  d.n_scenes = nscenes;
  for (s = 0; s < nscenes; s++) {
    const photic_scene *sc = &scene_data[scene_indexes[s]];
    if (sc->n_bands < 2 || sc->n_bands > PHB_MAX_BANDS) die("bands per scene", PHB_EINVAL);
    d.n_bands[s] = sc->n_bands;
    if (sc->n_bands > max_bands) max_bands = sc->n_bands;
    d.theta_view[s] = sc->theta_v;
    d.theta_sun[s] = sc->theta_w;
    d.h_tide[s] = sc->H_tide;
    for (b = 0; b < sc->n_bands; b++, g++) {
      d.wavelengths[s][b] = sc->wavelengths[b];
      d.r_sigma[s][b] = sc->R_sigma[b];
      packed[g] = pack_rows(gridded_data[sc->band_indexes[b]].array, nrows, ncols);
      planes[g] = packed[g];
    }
  }
  d.n_smoothing_radius = n_smoothing_radius;
  d.n_spatial = n_spatial;
  d.n_bottoms = n_bottoms;
  d.nrows = nrows;
  d.ncols = ncols;
  d.nodata = gridded_data[scene_data[scene_indexes[0]].band_indexes[0]].nodata_value;
  d.prior_present = empirical_depth_present ? 1 : 0;
  d.prior_nodata = empirical_depths.nodata_value;
  if (empirical_depth_present) prior = pack_rows(empirical_depths.array, nrows, ncols);

  memset(&out, 0, sizeof(out));
  for (k = 0; k < 9; k++) {
    res[k] = (float *)malloc(px * sizeof(float));
    if (res[k] == NULL) { printf("\n\nERROR: Out of memory.\n\n"); exit(1); }
  }
  out.depth = res[0]; out.model_error = res[1]; out.bottom_albedo = res[2]; out.bottom_sand = res[3];
  out.bottom_seagrass = res[4]; out.bottom_coral = res[5]; out.K_min = res[6]; out.bottom_type = res[7];
  out.index_optical_depth = res[8];
  Kp = (float *)malloc(px * sizeof(float) * nscenes * max_bands);
  Pp = (float *)malloc(px * sizeof(float) * nscenes);
  Gp = (float *)malloc(px * sizeof(float) * nscenes);
  Xp = (float *)malloc(px * sizeof(float) * nscenes);
  if (!Kp || !Pp || !Gp || !Xp) { printf("\n\nERROR: Out of memory.\n\n"); exit(1); }
  out.K = Kp; out.P = Pp; out.G = Gp; out.X = Xp;

  open_devices();
  if (g_n_ctx > 1) rc = phb_invert_host_multi(g_ctxs, g_n_ctx, &d, planes, prior, &out, &st, NULL, NULL);
  else rc = phb_invert_host(g_ctx, &d, planes, prior, 0, nrows, &out, &st);
  if (rc) die("inversion failed", rc);
  printf("\nNumber of optically shallow pixels = %lld\n", (long long)st.n_valid);
  printf("\nGPU inversion on %d device(s): %.1f ms (%.0f px/sec), mean iterations = %.0f, diverged = %.2f (%%)\n", g_n_ctx, st.ms_solve,
         st.n_valid / (st.ms_solve * 1e-3 + 1e-12), st.n_valid ? (double)st.n_evals / st.n_valid : 0.0,
         st.n_valid ? 100.0 * (double)(st.n_valid - st.n_converged) / st.n_valid : 0.0);

  grids9[0] = depth; grids9[1] = model_error; grids9[2] = bottom_albedo; grids9[3] = bottom_sand;
  grids9[4] = bottom_seagrass; grids9[5] = bottom_coral; grids9[6] = K_min; grids9[7] = bottom_type;
  grids9[8] = index_optical_depth;
  for (k = 0; k < 9; k++) unpack_rows(res[k], grids9[k], nrows, ncols);
  /* depth-error estimate, samodel.c:1376-1477. The reference seeds rand() with time(NULL) (samodel.c:371);
   * PHOTIC_B200_SIGMA_SEED fixes the seed, PHOTIC_B200_SIGMA_CHAIN=reference runs the trials as one serial
   * hot-start chain exactly as the reference does (default: one chain per depth interval, in parallel). */
  {
    float *sig = (float *)malloc(px * sizeof(float));
    if (sig == NULL) { printf("\n\nERROR: Out of memory.\n\n"); exit(1); }
    memset(sig, 0, px * sizeof(float));
    if (empirical_depth_present) {
      const char *e_seed = getenv("PHOTIC_B200_SIGMA_SEED"), *e_chain = getenv("PHOTIC_B200_SIGMA_CHAIN");
      const unsigned seed = e_seed ? (unsigned)strtoul(e_seed, NULL, 10) : (unsigned)time(NULL);
      const int chain = (e_chain && strcmp(e_chain, "reference") == 0) ? PHB_SIGMA_CHAIN_REFERENCE : PHB_SIGMA_CHAIN_PER_INTERVAL;
      phb_stats st2;
      int32_t n_int = 0;
      printf("\nComputing depth error estimates...");
      rc = phb_depth_sigma_host(g_ctx, &d, planes, prior, res[0], seed, 128, chain, PHB_SIGMA_MAX_INTERVALS, sig, NULL,
                                &n_int, NULL, &st2);
      if (rc) die("depth error estimate failed", rc);
      printf("...finished (%d depth intervals, %lld trials, %.1f ms).\n", (int)n_int, (long long)st2.n_valid, st2.ms_solve);
    } else {
      printf("\nDepth error estimates need a DEPTHS grid (see include/photic_b200.h): H_sigma left at 0.\n");
    }
    unpack_rows(sig, depth_sigma, nrows, ncols);
    free(sig);
  }

  /* file side effects of samodel.c:1494-1687, through the host's own NetCDF writer when present */
  if (write_nc) {
    const photic_geogrid *g0 = &gridded_data[scene_data[scene_indexes[0]].band_indexes[0]];
    float *lons = (float *)malloc(ncols * sizeof(float)), *lats = (float *)malloc(nrows * sizeof(float));
    static const char *band_name[4] = {"coastal", "blue", "green", "red"};
    static const char *grid_name[10] = {"H", "error", "albedo", "bottom_sand", "bottom_seagrass", "bottom_coral",
                                        "min_K", "bottom_type", "index_optical_depth", "H_sigma"};
    char file_name[PHOTIC_MAX_STRING_LEN + 64];
    printf("\nWriting model data to file...");
    for (c = 0; c < ncols; c++) lons[c] = g0->wlon + ((float)c) * g0->cellsize;
    for (r = 0; r < nrows; r++) lats[r] = g0->slat + ((float)r) * g0->cellsize;
    for (s = 0; s < nscenes; s++) {
      const char *nm = scene_data[scene_indexes[s]].scene_name;
      float **v;
      for (b = 0; b < d.n_bands[s] && b < 4; b++) {
        snprintf(file_name, sizeof(file_name), "%s_K_%s.nc", nm, band_name[b]);
        v = row_view(Kp + ((size_t)s * max_bands + b) * px, nrows, ncols);
        write_nc(file_name, v, ncols, nrows, lons, lats, g0->nodata_value);
        free(v);
      }
      snprintf(file_name, sizeof(file_name), "%s_P.nc", nm);
      v = row_view(Pp + (size_t)s * px, nrows, ncols); write_nc(file_name, v, ncols, nrows, lons, lats, g0->nodata_value); free(v);
      snprintf(file_name, sizeof(file_name), "%s_G.nc", nm);
      v = row_view(Gp + (size_t)s * px, nrows, ncols); write_nc(file_name, v, ncols, nrows, lons, lats, g0->nodata_value); free(v);
      snprintf(file_name, sizeof(file_name), "%s_X.nc", nm);
      v = row_view(Xp + (size_t)s * px, nrows, ncols); write_nc(file_name, v, ncols, nrows, lons, lats, g0->nodata_value); free(v);
    }
    for (k = 0; k < 9; k++) {
      snprintf(file_name, sizeof(file_name), "modelled_%s.nc", grid_name[k]);
      write_nc(file_name, grids9[k], ncols, nrows, lons, lats, g0->nodata_value);
    }
    snprintf(file_name, sizeof(file_name), "modelled_%s.nc", grid_name[9]);
    write_nc(file_name, depth_sigma, ncols, nrows, lons, lats, g0->nodata_value);
    printf("\n... finished.\n");
    free(lons); free(lats);
  }

  for (k = 0; k < g; k++) free(packed[k]);
  for (k = 0; k < 9; k++) free(res[k]);
  free(Kp); free(Pp); free(Gp); free(Xp);
  if (prior) free(prior);
}
