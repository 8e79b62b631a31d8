/*
 * ref_harness.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * A small driver, written for this repo, that is compiled TOGETHER WITH the
 * UNMODIFIED reference sources where they lie (/root/reference/model/samodel.c,
 * asa047.c, common.c) into oracle/_ref/libphotic_ref.so (recipe: oracle/Makefile).
 * It exposes the reference's own functions through a flat C ABI that ctypes can
 * call, so that
 *   (1) the plain-C restatement in oracle/photic_oracle.c can be pinned against
 *       the real reference (known-answer dumps -> tests/golden/), and
 *   (2) the reference's CPU implementation can be timed as the CPU baseline.
 *
 * What is driven (all symbols below are the reference's own, from samodel.h):
 *   extract_Rrs_data  samodel.c:2957     samodel_optimise  samodel.c:1768
 *   samodel_error     samodel.c:2432     samodel_Rrs       samodel.c:2846
 *   nelmin            asa047.c:10        interp_1d         common.c:298
 *   random_in_range   common.c:527       frand2            common.c:220
 *   array_max2        common.c:1240      vec_mean2_double / vec_stddev_double  common.c:877,924
 *
 * The per-pixel "cold start" path (extract_Rrs_data -> h_empirical rule of
 * samodel.c:960-976 -> start_at_previous=false -> samodel_optimise) is the
 * well-defined parity oracle: samodel() itself carries a LUT
 * and hot-start state from pixel to pixel and is schedule dependent.
 *
 * The model_data set-up below follows what samodel.c:375-643 does, rewritten here
 * because that code is inlined in samodel() and cannot be called on its own.
 */
#include "samodel.h"
/* secchi.h has no include guard and only adds prototypes on top of common.h: declare what is driven */
void Kd_LS8(float **coastal, float **blue, float **green, float **red, float **kd, int nrows, int ncols,
            float coastal_spv, float blue_spv, float green_spv, float red_spv, float theta_s);
void secchi_disk_depth(float **coastal, float **blue, float **green, float **red, float **zsd, int nrows, int ncols,
                       float coastal_spv, float blue_spv, float green_spv, float red_spv, float theta_s);
#if _OPENMP
#include <omp.h>
#endif

/* reference tables (non-static globals in samodel.c:103-286) */
extern int n_ref_wlens;
extern double ref_wlens[], aw_Pope_Fry1997[], bbw_Morel_1974[], a0_Lee[], a1_Lee[];
extern double bottom_type_sand[], bottom_type_seagrass[], bottom_type_coral[],
    bottom_type_macrophytes[], bottom_type_dark_sediment[], bottom_type_coral_sand[],
    bottom_type_green_algae[], bottom_type_red_algae[];

/* nc.c is not compiled (needs libnetcdf): samodel() calls write_nc at the end. */
void write_nc(char *file, float **grid, int ncols, int nrows, float *lons, float *lats,
              double spval) {
  (void)file; (void)grid; (void)ncols; (void)nrows; (void)lons; (void)lats; (void)spval;
}
/* samodelgraphics.c is not compiled (PGPLOT); only referenced when PLOTTING != 0. */
void samodel_graphics(model_data *md, float pagesize, int linewidth) {
  (void)md; (void)pagesize; (void)linewidth;
}

typedef struct {
  int nscenes, maxb;
  const int *n_bands;      /* [nscenes] */
  const int *wavelengths;  /* [nscenes*maxb], int as in scene.wavelengths (common.h:200) */
  const double *theta_v, *theta_w, *h_tide; /* [nscenes], degrees / metres */
  const double *r_sigma;   /* [nscenes*maxb] */
  int n_smooth, n_spatial, n_bottoms;
} ref_cfg;

static double *bottom_table(int k) {
  switch (k) { /* order of samodel.c:464-478 */
    case 0: return bottom_type_sand;
    case 1: return bottom_type_seagrass;
    case 2: return bottom_type_coral;
    case 3: return bottom_type_macrophytes;
    case 4: return bottom_type_dark_sediment;
    case 5: return bottom_type_coral_sand;
    case 6: return bottom_type_green_algae;
    default: return bottom_type_red_algae;
  }
}

/* Builds a model_data the way samodel.c:375-643 does (calloc'd: SURVEY 8c UB note). */
static model_data *make_md(const ref_cfg *c) {
  int s, b, k, nr;
  model_data *md = (model_data *)calloc(1, sizeof(model_data));
  nr = (int)pow(2 * c->n_spatial + 1, 2); /* samodel.c:379 */
  md->max_n_regions = nr;
  md->n_scenes = c->nscenes;
  md->n_bands = (int *)malloc(c->nscenes * sizeof(int));
  md->n_raw_bands = (int *)malloc(c->nscenes * sizeof(int));
  md->max_n_bands = 0;
  for (s = 0; s < c->nscenes; s++) {
    md->n_bands[s] = md->n_raw_bands[s] = c->n_bands[s];
    if (c->n_bands[s] > md->max_n_bands) md->max_n_bands = c->n_bands[s];
  }
  allocate_double_array_3d(&md->Rrs_measured, nr, c->nscenes, md->max_n_bands);
  allocate_double_array_3d(&md->Rrs_modelled, nr, c->nscenes, md->max_n_bands);
  allocate_double_array_3d(&md->rrs_modelled, nr, c->nscenes, md->max_n_bands);
  allocate_double_array_3d(&md->rrs_bottom, nr, c->nscenes, md->max_n_bands);
  allocate_double_array_3d(&md->rho, nr, c->nscenes, md->max_n_bands);
  allocate_double_array_3d(&md->rrs_dp, nr, c->nscenes, md->max_n_bands);
  allocate_double_array_2d(&md->wavelengths, c->nscenes, md->max_n_bands);
  allocate_double_array_2d(&md->raw_wavelengths, c->nscenes, md->max_n_bands);
  allocate_double_array_2d(&md->a_0, c->nscenes, md->max_n_bands);
  allocate_double_array_2d(&md->a_1, c->nscenes, md->max_n_bands);
  allocate_double_array_2d(&md->a_w, c->nscenes, md->max_n_bands);
  allocate_double_array_2d(&md->b_bw, c->nscenes, md->max_n_bands);
  allocate_double_array_2d(&md->K, c->nscenes, md->max_n_bands);
  allocate_double_array_2d(&md->rrs_noise, c->nscenes, md->max_n_bands);
  for (s = 0; s < c->nscenes; s++)
    for (b = 0; b < md->max_n_bands; b++) md->K[s][b] = 0.0;
  md->theta_view = (double *)malloc(c->nscenes * sizeof(double));
  md->theta_sun = (double *)malloc(c->nscenes * sizeof(double));
  md->sec_theta_view = (double *)malloc(c->nscenes * sizeof(double));
  md->sec_theta_sun = (double *)malloc(c->nscenes * sizeof(double));
  md->H_tide = (double *)malloc(c->nscenes * sizeof(double));
  md->n_bottoms = c->n_bottoms;
  allocate_double_array_2d(&md->Rrs440, nr, c->nscenes);
  allocate_double_array_2d(&md->Rrs490, nr, c->nscenes);
  allocate_double_array_2d(&md->Rrs550, nr, c->nscenes);
  allocate_double_array_2d(&md->Rrs640, nr, c->nscenes);
  allocate_double_array_2d(&md->Rrs750, nr, c->nscenes);
  md->a_w640 = interp_1d(ref_wlens, aw_Pope_Fry1997, n_ref_wlens, 640.0);
  for (s = 0; s < c->nscenes; s++) {
    md->H_tide[s] = c->h_tide[s];
    md->theta_view[s] = c->theta_v[s];
    md->theta_view[s] *= PI / 180.0; /* exactly as samodel.c:538 (NOT theta*PI/180: different rounding) */
    md->sec_theta_view[s] = 1.0 / cos(md->theta_view[s]);
    md->theta_sun[s] = c->theta_w[s];
    md->theta_sun[s] *= PI / 180.0; /* samodel.c:546 */
    md->sec_theta_sun[s] = 1.0 / cos(md->theta_sun[s]);
    for (b = 0; b < c->n_bands[s]; b++) {
      double w = (double)c->wavelengths[s * c->maxb + b];
      md->wavelengths[s][b] = md->raw_wavelengths[s][b] = w;
      md->a_0[s][b] = interp_1d(ref_wlens, a0_Lee, n_ref_wlens, w);
      md->a_1[s][b] = interp_1d(ref_wlens, a1_Lee, n_ref_wlens, w);
      md->b_bw[s][b] = interp_1d(ref_wlens, bbw_Morel_1974, n_ref_wlens, w);
      md->a_w[s][b] = interp_1d(ref_wlens, aw_Pope_Fry1997, n_ref_wlens, w);
      md->rrs_noise[s][b] = c->r_sigma[s * c->maxb + b];
    }
  }
  allocate_double_array_3d(&md->bottom_reflectance, c->n_bottoms, c->nscenes, md->max_n_bands);
  for (k = 0; k < c->n_bottoms; k++)
    for (s = 0; s < c->nscenes; s++)
      for (b = 0; b < c->n_bands[s]; b++)
        md->bottom_reflectance[k][s][b] =
            interp_1d(ref_wlens, bottom_table(k), n_ref_wlens, md->wavelengths[s][b]);
  md->P = (double *)calloc(c->nscenes, sizeof(double));
  md->G = (double *)calloc(c->nscenes, sizeof(double));
  md->X = (double *)calloc(c->nscenes, sizeof(double));
  md->D = (double *)calloc(c->nscenes, sizeof(double));
  md->prev = (double *)calloc(4 * c->nscenes, sizeof(double));
  md->start_at_previous = false;
  md->model_error = 100.0;
  return md;
}

/* restart counter of the wrapped nelmin (end of this file) and where ref_invert_pixels reports it */
void ref_nelmin_counters_reset(void);
int ref_nelmin_restarts(void);
static int *g_numres_out = NULL;
/* buf: [npix] ints filled by the next ref_invert_pixels calls with nelmin's restart count (asa047.c:493), summed
 * over the H starts of the pixel; NULL switches it off */
void ref_set_numres_out(int *buf) { g_numres_out = buf; }

/* Number of doubles per pixel record written by ref_invert_pixels. */
int ref_record_len(int nscenes, int maxb) { return 16 + nscenes * maxb + 3 * nscenes; }

/* Dump of the interpolated per-(scene,band) tables, to pin the restatement's tables.
 * out: [nscenes*maxb*(4 + n_bottoms)] = a0,a1,aw,bbw,bottom[0..] ; out2: a_w640, then sec_v[s], sec_w[s] */
int ref_tables(int nscenes, int maxb, const int *n_bands, const int *wavelengths, const double *theta_v,
               const double *theta_w, const double *h_tide, const double *r_sigma, int n_bottoms,
               double *out, double *out2) {
  ref_cfg c = {nscenes, maxb, n_bands, wavelengths, theta_v, theta_w, h_tide, r_sigma, 1, 2, n_bottoms};
  model_data *md = make_md(&c);
  int s, b, k, o = 0;
  for (s = 0; s < nscenes; s++)
    for (b = 0; b < n_bands[s]; b++) {
      double *p = out + (size_t)(s * maxb + b) * (4 + n_bottoms);
      p[0] = md->a_0[s][b]; p[1] = md->a_1[s][b]; p[2] = md->a_w[s][b]; p[3] = md->b_bw[s][b];
      for (k = 0; k < n_bottoms; k++) p[4 + k] = md->bottom_reflectance[k][s][b];
    }
  out2[o++] = md->a_w640;
  for (s = 0; s < nscenes; s++) out2[o++] = md->sec_theta_view[s];
  for (s = 0; s < nscenes; s++) out2[o++] = md->sec_theta_sun[s];
  return 0;
}

/*
 * Per-pixel cold-start inversion of a list of pixels with the reference's own
 * extract_Rrs_data + samodel_optimise.
 *   planes: scene-major, band-minor, each [nrows][ncols] float32, contiguous.
 *   prior : DEPTHS grid (negative-down metres, samodel.c:960-967) or NULL.
 *   rec   : [npix][ref_record_len] doubles; status[npix]: 1 inverted, 0 skipped (nodata / no regions)
 */
int ref_invert_pixels(int nscenes, int maxb, const int *n_bands, const int *wavelengths,
                      const double *theta_v, const double *theta_w, const double *h_tide,
                      const double *r_sigma, int n_smooth, int n_spatial, int n_bottoms, int nrows,
                      int ncols, const float *planes, float nodata, const float *prior,
                      float prior_nodata, int npix, const int *pix_i, const int *pix_j, double *rec,
                      int *status, int *converged, int *n_iterations, int nthreads) {
  ref_cfg c = {nscenes, maxb, n_bands, wavelengths, theta_v, theta_w, h_tide, r_sigma,
               n_smooth, n_spatial, n_bottoms};
  int s, b, g, r, ngrids = 0, reclen = ref_record_len(nscenes, maxb);
  scene *sc = (scene *)calloc(nscenes, sizeof(scene));
  int *scene_indexes = (int *)malloc(nscenes * sizeof(int));
  geogrid *grids;
  for (s = 0; s < nscenes; s++) ngrids += n_bands[s];
  grids = (geogrid *)calloc(ngrids, sizeof(geogrid));
  g = 0;
  for (s = 0; s < nscenes; s++) {
    scene_indexes[s] = s;
    snprintf(sc[s].scene_name, 64, "scene%d", s);
    sc[s].n_bands = n_bands[s];
    sc[s].nrows = nrows;
    sc[s].ncols = ncols;
    sc[s].theta_v = theta_v[s];
    sc[s].theta_w = theta_w[s];
    sc[s].H_tide = h_tide[s];
    for (b = 0; b < n_bands[s]; b++) {
      sc[s].band_indexes[b] = g;
      sc[s].wavelengths[b] = wavelengths[s * maxb + b];
      sc[s].R_sigma[b] = r_sigma[s * maxb + b];
      grids[g].nrows = nrows;
      grids[g].ncols = ncols;
      grids[g].nodata_value = nodata;
      grids[g].array = (float **)malloc(nrows * sizeof(float *));
      for (r = 0; r < nrows; r++)
        grids[g].array[r] = (float *)(planes + ((size_t)g * nrows + r) * ncols);
      g++;
    }
  }
#if _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel
#endif
  {
    model_data *md = make_md(&c);
    int p;
#if _OPENMP
#pragma omp for schedule(dynamic)
#endif
    for (p = 0; p < npix; p++) {
      int i = pix_i[p], j = pix_j[p], k, ks, kb, o;
      double *R = rec + (size_t)p * reclen;
      bool nd = false;
      float refl;
      for (k = 0; k < reclen; k++) R[k] = 0.0;
      status[p] = 0; converged[p] = 0; n_iterations[p] = 0;
      /* validity rule of samodel.c:933-947 */
      for (ks = 0; ks < nscenes && !nd; ks++)
        for (kb = 0; kb < n_bands[ks]; kb++) {
          refl = grids[sc[ks].band_indexes[kb]].array[i][j];
          if (approx_equal(refl, nodata, 1.0e-6) || refl < 0.0) { nd = true; break; }
        }
      if (nd) continue;
      md->i = i; md->j = j;
      extract_Rrs_data(i, j, sc, grids, scene_indexes, nscenes, n_spatial, n_smooth, nrows, ncols, md, 0.0);
      if (md->n_regions == 0) continue;
      /* h_empirical rule of samodel.c:960-976 */
      if (prior != NULL) {
        float e = prior[(size_t)i * ncols + j];
        if (!approx_equal(e, prior_nodata, 1.0e-6)) {
          md->empirical_depth_present = true;
          md->h_empirical = (e > -1.0) ? 1.0 : fabs(e);
        } else {
          md->empirical_depth_present = false;
          md->h_empirical = 0.0; /* pinned: the reference leaves it stale */
        }
      } else {
        md->empirical_depth_present = false;
        md->h_empirical = 0.0; /* pinned: uninitialised in the reference */
      }
      md->start_at_previous = false;
      md->n_bottoms = n_bottoms;
      ref_nelmin_counters_reset();
      samodel_optimise(md);
      if (g_numres_out) g_numres_out[p] = ref_nelmin_restarts();
      status[p] = 1;
      converged[p] = md->converged ? 1 : 0;
      n_iterations[p] = md->n_iterations;
      R[0] = md->depth; R[1] = md->Rrs_error; R[2] = md->bottom_albedo;
      R[3] = md->B_type_percent[0]; R[4] = md->B_type_percent[1]; R[5] = md->B_type_percent[2];
      R[6] = md->K_min; R[7] = md->index_optical_depth; R[8] = (double)md->bottom_type;
      R[9] = md->model_error; R[10] = md->depth_error; R[11] = md->bottom_error; R[12] = md->K_error;
      R[13] = (double)md->n_regions; R[14] = (double)md->origin; R[15] = md->h_empirical;
      o = 16;
      for (ks = 0; ks < nscenes; ks++)
        for (kb = 0; kb < maxb; kb++) R[o++] = (kb < n_bands[ks]) ? md->K[ks][kb] : 0.0;
      for (ks = 0; ks < nscenes; ks++) { R[o++] = md->P[ks]; R[o++] = md->G[ks]; R[o++] = md->X[ks]; }
    }
  }
  for (g = 0; g < ngrids; g++) free(grids[g].array);
  free(grids); free(sc); free(scene_indexes);
  return 0;
}

/*
 * Known-answer dump of samodel_error (and, through it, samodel_Rrs) on caller-chosen
 * parameter vectors. rrs_measured: [n_regions][nscenes][maxb]; params: [nvec][nparams].
 * out: [nvec][6] = model_error, Rrs_error, depth_error, bottom_error, K_error, bottom_albedo
 * out_Rrs (nullable): [nvec][n_regions][nscenes][maxb] modelled Rrs; out_K: [nvec][nscenes][maxb]
 */
int ref_error_kat(int nscenes, int maxb, const int *n_bands, const int *wavelengths,
                  const double *theta_v, const double *theta_w, const double *h_tide,
                  const double *r_sigma, int n_bottoms_active, int n_regions, int origin,
                  const double *rrs_measured, int nparams, int nvec, const double *params,
                  double *out, double *out_Rrs, double *out_K) {
  ref_cfg c = {nscenes, maxb, n_bands, wavelengths, theta_v, theta_w, h_tide, r_sigma, 1, 2,
               n_bottoms_active};
  model_data *md = make_md(&c);
  int r, s, b, v, k;
  md->n_regions = n_regions;
  md->origin = origin;
  md->n_bottoms = n_bottoms_active;
  md->n_params = nparams;
  for (k = 0; k < n_bottoms_active; k++) md->bottom_type_indexes[k] = k;
  for (r = 0; r < n_regions; r++)
    for (s = 0; s < nscenes; s++) {
      for (b = 0; b < n_bands[s]; b++)
        md->Rrs_measured[r][s][b] = rrs_measured[((size_t)r * nscenes + s) * maxb + b];
      /* samodel.c:1785-1799 */
      md->Rrs440[r][s] = interp_1d(md->wavelengths[s], md->Rrs_measured[r][s], n_bands[s], 440.0);
      md->Rrs490[r][s] = interp_1d(md->wavelengths[s], md->Rrs_measured[r][s], n_bands[s], 490.0);
      md->Rrs550[r][s] = interp_1d(md->wavelengths[s], md->Rrs_measured[r][s], n_bands[s], 550.0);
      md->Rrs640[r][s] = interp_1d(md->wavelengths[s], md->Rrs_measured[r][s], n_bands[s], 640.0);
      md->Rrs750[r][s] = interp_1d(md->wavelengths[s], md->Rrs_measured[r][s], n_bands[s], 750.0);
      if (md->Rrs440[r][s] < 0.0) md->Rrs440[r][s] = 0.0001;
    }
  for (v = 0; v < nvec; v++) {
    double *o = out + (size_t)v * 6;
    double *x = (double *)malloc(nparams * sizeof(double));
    memcpy(x, params + (size_t)v * nparams, nparams * sizeof(double));
    o[0] = samodel_error(x, md);
    o[1] = md->Rrs_error; o[2] = md->depth_error; o[3] = md->bottom_error; o[4] = md->K_error;
    o[5] = md->bottom_albedo;
    if (out_Rrs)
      for (r = 0; r < n_regions; r++)
        for (s = 0; s < nscenes; s++)
          for (b = 0; b < maxb; b++)
            out_Rrs[(((size_t)v * n_regions + r) * nscenes + s) * maxb + b] =
                (b < n_bands[s]) ? md->Rrs_modelled[r][s][b] : 0.0;
    if (out_K)
      for (s = 0; s < nscenes; s++)
        for (b = 0; b < maxb; b++)
          out_K[((size_t)v * nscenes + s) * maxb + b] = (b < n_bands[s]) ? md->K[s][b] : 0.0;
    free(x);
  }
  return 0;
}

/* interp_1d (common.c:298) known answers */
double ref_interp_1d(const double *X, const double *Y, int n, double x) {
  return interp_1d((double *)X, (double *)Y, n, x);
}

/* approx_equal (common.c:392) known answers: float arguments */
int ref_approx_equal(float a, float b, float eps) { return approx_equal(a, b, eps) ? 1 : 0; }

/* --- nelmin (asa047.c:10) known answers on analytic test functions -------------------- */
static double tf_rosenbrock(double x[], model_data *md) {
  int n = md->n_params, i;
  double f = 0.0;
  for (i = 0; i + 1 < n; i++) {
    double a = x[i + 1] - x[i] * x[i], b = 1.0 - x[i];
    f += 100.0 * a * a + b * b;
  }
  return f;
}
static double tf_quartic(double x[], model_data *md) { /* Powell-like, plateaus exercise shrink */
  int n = md->n_params, i;
  double f = 0.0;
  for (i = 0; i < n; i++) {
    double d = x[i] - 0.5 * (double)(i + 1);
    f += d * d * d * d + 0.1 * fabs(d);
  }
  return f;
}
static double tf_steps(double x[], model_data *md) { /* piecewise constant: exact ties + factorial test */
  int n = md->n_params, i;
  double f = 0.0;
  for (i = 0; i < n; i++) f += floor(fabs(x[i]) * 4.0) * 0.25 + 0.01 * x[i] * x[i];
  return f;
}
int ref_nelmin_kat(int fn_id, int n, const double *start, const double *step, double reqmin, int konvge,
                   int kcount, double *xmin, double *ynewlo, int *icount, int *numres, int *ifault) {
  model_data md;
  double *s = (double *)malloc(n * sizeof(double)), *st = (double *)malloc(n * sizeof(double));
  double (*fn)(double[], model_data *) =
      fn_id == 0 ? tf_rosenbrock : (fn_id == 1 ? tf_quartic : tf_steps);
  memset(&md, 0, sizeof(md));
  md.n_params = n;
  memcpy(s, start, n * sizeof(double));
  memcpy(st, step, n * sizeof(double));
  *icount = 0; *numres = 0; *ifault = 0;
  nelmin(fn, &md, n, s, xmin, ynewlo, reqmin, st, konvge, kcount, icount, numres, ifault);
  free(s); free(st);
  return 0;
}

/*
 * samodel() exactly as shipped (LUT + hot start + depth-sigma Monte-Carlo), for the
 * "reference runs as-is" timing of config 1 only. Output is schedule dependent.
 * outputs: 10 planes [nrows][ncols] in the order of samodel.h:14-16.
 */
int ref_samodel_as_is(int nscenes, int maxb, const int *n_bands, const int *wavelengths,
                      const double *theta_v, const double *theta_w, const double *h_tide,
                      const double *r_sigma, int n_smooth, int n_spatial, int n_bottoms, int nrows,
                      int ncols, const float *planes, float nodata, const float *prior,
                      float prior_nodata, float *outputs) {
  int s, b, g = 0, r, k, ngrids = 0;
  static scene sc[MAX_SCENES];
  int scene_indexes[MAX_SCENES];
  geogrid *grids, pg;
  float **outp[10];
  for (s = 0; s < nscenes; s++) ngrids += n_bands[s];
  grids = (geogrid *)calloc(ngrids + 1, sizeof(geogrid));
  for (s = 0; s < nscenes; s++) {
    memset(&sc[s], 0, sizeof(scene));
    scene_indexes[s] = s;
    snprintf(sc[s].scene_name, 64, "scene%d", s);
    sc[s].n_bands = n_bands[s]; sc[s].nrows = nrows; sc[s].ncols = ncols;
    sc[s].theta_v = theta_v[s]; sc[s].theta_w = theta_w[s]; sc[s].H_tide = h_tide[s];
    for (b = 0; b < n_bands[s]; b++) {
      sc[s].band_indexes[b] = g;
      sc[s].wavelengths[b] = wavelengths[s * maxb + b];
      sc[s].R_sigma[b] = r_sigma[s * maxb + b];
      grids[g].nrows = nrows; grids[g].ncols = ncols; grids[g].nodata_value = nodata;
      grids[g].cellsize = 1.0f;
      grids[g].array = (float **)malloc(nrows * sizeof(float *));
      for (r = 0; r < nrows; r++) grids[g].array[r] = (float *)(planes + ((size_t)g * nrows + r) * ncols);
      g++;
    }
  }
  memset(&pg, 0, sizeof(pg));
  if (prior) {
    pg.nrows = nrows; pg.ncols = ncols; pg.nodata_value = prior_nodata;
    pg.array = (float **)malloc(nrows * sizeof(float *));
    for (r = 0; r < nrows; r++) pg.array[r] = (float *)(prior + (size_t)r * ncols);
  }
  for (k = 0; k < 10; k++) {
    outp[k] = (float **)malloc(nrows * sizeof(float *));
    for (r = 0; r < nrows; r++) outp[k][r] = outputs + ((size_t)k * nrows + r) * ncols;
  }
  samodel(sc, grids, scene_indexes, nscenes, prior ? true : false, pg, n_smooth, n_spatial, n_bottoms,
          outp[0], outp[1], outp[2], outp[3], outp[4], outp[5], outp[6], outp[7], outp[8], outp[9],
          8.0f, 0, 1);
  for (k = 0; k < 10; k++) free(outp[k]);
  for (g = 0; g < ngrids; g++) free(grids[g].array);
  if (prior) free(pg.array);
  free(grids);
  return 0;
}

/*
 * The depth-error phase of samodel() (samodel.c:1376-1477), driven with the reference's own functions:
 * random_in_range / frand2 (libc rand()), extract_Rrs_data with the drawn n_sigma, samodel_optimise
 * (hot-started from the previous trial once the first one has run), array_max2, vec_mean2_double,
 * vec_stddev_double. The control flow of those hundred lines is restated here because it is inlined in
 * samodel(); the reference seeds rand() with time(NULL) (samodel.c:371), here the seed is an argument.
 * rand() is used nowhere else in samodel(), so this equals the reference run with srand(seed).
 *   depth        [nrows*ncols] POSITIVE depths as the pixel loop leaves them (0 where nothing was inverted)
 *   n_samples    128 in the reference (samodel.c:1378); smaller values keep tests short
 *   chain_mode   0: one hot-start chain over all trials (the reference); 1: the chain restarts cold at
 *                every depth interval (the parallel variant of the B200 path)
 *   table        [n_intervals] sigma per 0.25 m interval;  trials (nullable) [n_intervals*n_samples]
 *   depth_sigma  (nullable) [nrows*ncols]
 * Requires a DEPTHS prior: without one a hot-started trial would start from md->depth_prev, which the
 * reference leaves at whatever its LUT / hot-start bookkeeping wrote last (order dependent).
 */
int ref_depth_sigma(int nscenes, int maxb, const int *n_bands, const int *wavelengths, const double *theta_v,
                    const double *theta_w, const double *h_tide, const double *r_sigma, int n_smooth,
                    int n_spatial, int n_bottoms, int nrows, int ncols, const float *planes, float nodata,
                    const float *prior, float prior_nodata, const float *depth_in, unsigned seed, int n_samples,
                    int chain_mode, int max_intervals, double *table, int *n_intervals_out, double *trials,
                    float *depth_sigma) {
  ref_cfg c = {nscenes, maxb, n_bands, wavelengths, theta_v, theta_w, h_tide, r_sigma, n_smooth, n_spatial, n_bottoms};
  int s, b, g = 0, r, ngrids = 0, i = 0, j = 0, k_sample, k_depth, k_trial, n_trials, n_depth_intervals;
  double d, depth_interval, max_depth_reached, *trial_depths, mean_trial_depths, n_sigma;
  bool within_interval;
  scene *sc = (scene *)calloc(nscenes, sizeof(scene));
  int *scene_indexes = (int *)malloc(nscenes * sizeof(int));
  geogrid *grids;
  float **depth;
  model_data *md;
  if (prior == NULL) return 1;
  for (s = 0; s < nscenes; s++) ngrids += n_bands[s];
  grids = (geogrid *)calloc(ngrids, sizeof(geogrid));
  for (s = 0; s < nscenes; s++) {
    scene_indexes[s] = s;
    sc[s].n_bands = n_bands[s]; sc[s].nrows = nrows; sc[s].ncols = ncols;
    sc[s].theta_v = theta_v[s]; sc[s].theta_w = theta_w[s]; sc[s].H_tide = h_tide[s];
    for (b = 0; b < n_bands[s]; b++) {
      sc[s].band_indexes[b] = g;
      sc[s].wavelengths[b] = wavelengths[s * maxb + b];
      sc[s].R_sigma[b] = r_sigma[s * maxb + b];
      grids[g].nrows = nrows; grids[g].ncols = ncols; grids[g].nodata_value = nodata;
      grids[g].array = (float **)malloc(nrows * sizeof(float *));
      for (r = 0; r < nrows; r++) grids[g].array[r] = (float *)(planes + ((size_t)g * nrows + r) * ncols);
      g++;
    }
  }
  depth = (float **)malloc(nrows * sizeof(float *));
  for (r = 0; r < nrows; r++) depth[r] = (float *)(depth_in + (size_t)r * ncols);
  md = make_md(&c);

  srand(seed); /* samodel.c:371 has time(NULL) */
  n_trials = (int)sqrt(nrows * ncols);                                            /* samodel.c:1381 */
  depth_interval = 0.25;
  max_depth_reached = array_max2(depth, nrows, ncols, 0.0);                       /* samodel.c:1384-1387 */
  max_depth_reached = depth_interval * ((int)max_depth_reached / depth_interval);
  max_depth_reached = MIN(max_depth_reached, 30.0);
  n_depth_intervals = (int)max_depth_reached / depth_interval;
  if (n_depth_intervals > max_intervals) { n_depth_intervals = max_intervals; max_depth_reached = depth_interval * max_intervals; }
  if (n_depth_intervals < 0) n_depth_intervals = 0;
  trial_depths = (double *)malloc(n_samples * sizeof(double));
  md->start_at_previous = false;
  k_depth = 0;
  for (d = 0.0; d < max_depth_reached; d += depth_interval) {                     /* samodel.c:1396-1461 */
    if (chain_mode == 1) md->start_at_previous = false;
    for (k_sample = 0; k_sample < n_samples; k_sample++) {
      within_interval = false;
      for (k_trial = 0; k_trial < n_trials; k_trial++) {
        i = random_in_range(0, nrows);
        j = random_in_range(0, ncols);
        if (depth[i][j] > d && depth[i][j] < d + depth_interval) { within_interval = true; break; }
      }
      if (!within_interval) { trial_depths[k_sample] = 0.0; continue; }
      n_sigma = frand2(1.0);
      extract_Rrs_data(i, j, sc, grids, scene_indexes, nscenes, n_spatial, n_smooth, nrows, ncols, md, n_sigma);
      if (md->n_regions == 0) continue;
      if (!approx_equal(prior[(size_t)i * ncols + j], prior_nodata, 1.0e-6)) {
        md->empirical_depth_present = true;
        if (prior[(size_t)i * ncols + j] > -1.0) md->h_empirical = 1.0;
        else md->h_empirical = fabs(prior[(size_t)i * ncols + j]);
      } else { trial_depths[k_sample] = 0.0; continue; }
      md->n_bottoms = n_bottoms;
      samodel_optimise(md);
      md->start_at_previous = true;
      trial_depths[k_sample] = md->depth;
    }
    if (trials) for (k_sample = 0; k_sample < n_samples; k_sample++) trials[(size_t)k_depth * n_samples + k_sample] = trial_depths[k_sample];
    mean_trial_depths = vec_mean2_double(trial_depths, n_samples, 0.0);
    table[k_depth++] = vec_stddev_double(trial_depths, n_samples, 0.0, mean_trial_depths);
  }
  *n_intervals_out = k_depth;
  if (depth_sigma) {                                                              /* samodel.c:1463-1477 */
    for (i = 0; i < nrows; i++)
      for (j = 0; j < ncols; j++) {
        depth_sigma[(size_t)i * ncols + j] = 0.0;
        if (depth[i][j] > 0.0) {
          k_depth = 0;
          for (d = 0.0; d < max_depth_reached; d += depth_interval) {
            if (depth[i][j] > d && depth[i][j] <= d + depth_interval) { depth_sigma[(size_t)i * ncols + j] = table[k_depth]; break; }
            k_depth++;
          }
        }
      }
  }
  free(trial_depths); free(depth);
  for (g = 0; g < ngrids; g++) free(grids[g].array);
  free(grids); free(sc); free(scene_indexes);
  return 0;
}

/*
 * MODEL Lee_Kd_LS8 / Lee_Secchi_LS8 (bam.c:3250-3610): the reference's raster functions Kd_LS8 (secchi.c:13) and
 * secchi_disk_depth (secchi.c:59) on four Landsat-8 reflectance planes. mode 0: Kd (minimum diffuse attenuation),
 * mode 1: Secchi-disk depth. spv: the four planes' nodata values (coastal, blue, green, red).
 */
int ref_lee_ls8(int mode, int nrows, int ncols, const float *coastal, const float *blue, const float *green,
                const float *red, const float *spv, float theta_s, float *out) {
  float **p[5];
  const float *src[5] = {coastal, blue, green, red, out};
  int k, r;
  for (k = 0; k < 5; k++) {
    p[k] = (float **)malloc(nrows * sizeof(float *));
    for (r = 0; r < nrows; r++) p[k][r] = (float *)(src[k] + (size_t)r * ncols);
  }
  if (mode == 0) Kd_LS8(p[0], p[1], p[2], p[3], p[4], nrows, ncols, spv[0], spv[1], spv[2], spv[3], theta_s);
  else secchi_disk_depth(p[0], p[1], p[2], p[3], p[4], nrows, ncols, spv[0], spv[1], spv[2], spv[3], theta_s);
  for (k = 0; k < 5; k++) free(p[k]);
  return 0;
}

/*
 * COMPUTE K (bam.c:2362-2392): the reference's own jerlov.c (compiled where it lies), driven as bam.c drives it.
 * Outputs are zeroed first: the reference leaves them untouched on its `return false` paths.
 */
bool jerlov(float wlen_i, float wlen_j, float Lsmi, float Lsmj, float *Li, float *Lj, int npoints, float *ki, float *kj,
            float *m, float *c, float *r, float *water_type, float manual_ratio);
float compute_k(float water_type, float wlen);
void compute_k_from_jerlov(float water_type, float *alphas, int *spectral_indexes, float *wavelengths, int nspec);
bool compute_k_from_ratio(float ratio, float wlen_i, float wlen_j, float *water_type, float *k, float *wavelengths,
                          float n_wlens);

int ref_jerlov_fit(float wlen_i, float wlen_j, float lsm_i, float lsm_j, const float *Li, const float *Lj, int npoints,
                   float manual_ratio, float *out6) {
  int k;
  for (k = 0; k < 6; k++) out6[k] = 0.0f;
  return jerlov(wlen_i, wlen_j, lsm_i, lsm_j, (float *)Li, (float *)Lj, npoints, &out6[0], &out6[1], &out6[2], &out6[3],
                &out6[4], &out6[5], manual_ratio) ? 1 : 0;
}

void ref_jerlov_k(float water_type, const float *wavelengths, int n, float *k) {
  /* through compute_k_from_jerlov with the identity index map, as bam.c:2379 calls it */
  int i, *idx = (int *)malloc((n > 0 ? n : 1) * sizeof(int));
  for (i = 0; i < n; i++) { idx[i] = i; k[i] = 0.0f; }
  compute_k_from_jerlov(water_type, k, idx, (float *)wavelengths, n);
  free(idx);
}

int ref_jerlov_k_from_ratio(float ratio, float wlen_i, float wlen_j, const float *wavelengths, int n, float *water_type,
                            float *k) {
  int i;
  *water_type = 0.0f;
  for (i = 0; i < n; i++) k[i] = 0.0f;
  return compute_k_from_ratio(ratio, wlen_i, wlen_j, water_type, k, (float *)wavelengths, (float)n) ? 1 : 0;
}

/* ------------------------------------------------------------------------------------------
 * REFINE: the reference's own run_refine() (refine.c:12-302, compiled where it lies). It is a REPL verb that reads
 * its arguments from the tokenised command line `parsed[]` and its grids from the `gridded_data[]` / `grid_names[]`
 * globals of common.h, so this driver writes the command the REPL would have tokenised
 *   REFINE IN in OUT out [LAND land SHALLOW shallow] [SCRAP a b] [CLIP a b] [SCALE a b SHAPE s] [LINEAR m c] [POWER a b]
 * (the order refine.c:118-213 parses them in) and registers the input grids in slots 0..2. Numbers are printed with
 * %.9g, which atof() + the float assignment of refine.c turns back into exactly the float passed in.
 * flags / args: layout of include/photic_b200.h (PHB_REFINE_*), the same as pho_refine in oracle/photic_oracle.c.
 * ---------------------------------------------------------------------------------------- */
void run_refine(void);

static void ref_register_grid(int slot, const char *name, int nrows, int ncols, const float *data, float nodata) {
  int r;
  strcpy(grid_names[slot], name);
  allocated_grids[slot] = true;
  gridded_data[slot].nrows = nrows;
  gridded_data[slot].ncols = ncols;
  gridded_data[slot].nodata_value = nodata;
  gridded_data[slot].array = (float **)malloc(nrows * sizeof(float *));
  for (r = 0; r < nrows; r++) gridded_data[slot].array[r] = (float *)(data + (size_t)r * ncols);
}

int ref_refine(int nrows, int ncols, const float *in, float nodata, const float *land, float land_nodata,
               const float *shallow, float shallow_nodata, int flags, const float *args, float *out) {
  int k = 0, n, r, used = 1, ps_out = -1;
  memset(parsed, 0, sizeof(parsed));
  for (n = 0; n < MAX_GRIDS; n++) { allocated_grids[n] = false; grid_names[n][0] = '\0'; }
  ref_register_grid(0, "in", nrows, ncols, in, nodata);
  if (land) { ref_register_grid(used, "land", nrows, ncols, land, land_nodata); used++; }
  if (shallow) { ref_register_grid(used, "shallow", nrows, ncols, shallow, shallow_nodata); used++; }
  n_grids_in_use = used;
#define TOK(s) strcpy(parsed[k++], s)
#define NUM(v) snprintf(parsed[k++], MAX_ARG_SIZE, "%.9g", (double)(v))
  TOK("REFINE"); TOK("IN"); TOK("in"); TOK("OUT"); TOK("out");
  if (land) { TOK("LAND"); TOK("land"); }
  if (shallow) { TOK("SHALLOW"); TOK("shallow"); }
  if (flags & 8) { TOK("SCRAP"); NUM(args[7]); NUM(args[8]); }
  if (flags & 1) { TOK("CLIP"); NUM(args[0]); NUM(args[1]); }
  if (flags & 2) { TOK("SCALE"); NUM(args[2]); NUM(args[3]); TOK("SHAPE"); NUM(args[4]); }
  if (flags & 4) { TOK("LINEAR"); NUM(args[5]); NUM(args[6]); }
  if (flags & 16) { TOK("POWER"); NUM(args[9]); NUM(args[10]); }
#undef TOK
#undef NUM
  run_refine();
  for (n = 0; n < MAX_GRIDS; n++)
    if (strcmp(grid_names[n], "out") == 0) ps_out = n;
  if (ps_out < 0) return 1;
  for (r = 0; r < nrows; r++) memcpy(out + (size_t)r * ncols, gridded_data[ps_out].array[r], ncols * sizeof(float));
  free_float_array_2d(gridded_data[ps_out].array, nrows);
  for (n = 0; n < used; n++) free(gridded_data[n].array);
  for (n = 0; n < MAX_GRIDS; n++) { allocated_grids[n] = false; grid_names[n][0] = '\0'; gridded_data[n].array = NULL; }
  n_grids_in_use = 0;
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * nelmin's restart counter. samodel_optimise_one_bottom_combination keeps `numres` in a local
 * (samodel.c:2146, 2373) and drops it; to see it without touching the reference sources the library is linked
 * with -Wl,--wrap=nelmin: samodel.o's calls to nelmin arrive here, go on to the real nelmin (asa047.c:10), and
 * the restart count of every call is added up per thread. ref_invert_pixels2 reports the total over the H starts
 * of a pixel next to the usual record.
 * ---------------------------------------------------------------------------------------- */
void __real_nelmin(double fn(double x[], model_data *md), model_data *md, int n, double start[], double xmin[],
                   double *ynewlo, double reqmin, double step[], int konvge, int kcount, int *icount, int *numres,
                   int *ifault);
static __thread int tl_numres_total = 0, tl_nelmin_calls = 0;
void __wrap_nelmin(double fn(double x[], model_data *md), model_data *md, int n, double start[], double xmin[],
                   double *ynewlo, double reqmin, double step[], int konvge, int kcount, int *icount, int *numres,
                   int *ifault) {
  __real_nelmin(fn, md, n, start, xmin, ynewlo, reqmin, step, konvge, kcount, icount, numres, ifault);
  tl_numres_total += *numres;
  tl_nelmin_calls += 1;
}
void ref_nelmin_counters_reset(void) { tl_numres_total = 0; tl_nelmin_calls = 0; }
int ref_nelmin_restarts(void) { return tl_numres_total; }
int ref_nelmin_calls(void) { return tl_nelmin_calls; }
