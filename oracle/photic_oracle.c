/*
 * photic_oracle.c -- TEST INFRASTRUCTURE ONLY (see photic_oracle.h).
 *
 * Plain-C restatement of the reference's per-pixel cold-start inversion, written from the
 * algorithm with flat arrays instead of the reference's
 * pointer-to-pointer model_data. Every function cites the reference lines it restates
 * (paths relative to /root/reference/model/). Floating-point operation ORDER follows the
 * reference statement by statement, because the Nelder-Mead decision sequence is what the
 * parity claim rests on; compile with -ffp-contract=off (oracle/Makefile).
 *
 * Pinned bit-for-bit against the compiled reference: tests/test_oracle_vs_golden.py.
 */
#include "photic_oracle.h"

#include <float.h>
#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#if _OPENMP
#include <omp.h>
#endif

#include "../include/photic_spectra.h"

#define PHO_PI 3.141592653589793 /* common.h:19 */

/* what-if PHO_VARIANT_LIBM_JITTER: a pure function of the argument bits nudges the result of
 * exp/log/pow by -1/0/+1 ulp, emulating "another correctly-working libm" (e.g. CUDA's). */
static int g_jitter = 0;
static double jit(double y, double x) {
  union { double d; unsigned long long u; } a, b;
  unsigned long long h;
  if (!g_jitter) return y;
  a.d = x; b.d = y;
  h = a.u * 0x9E3779B97F4A7C15ULL; h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ULL; h ^= h >> 32;
  switch (h % 4) { case 0: b.u += 1; break; case 1: b.u -= 1; break; default: break; }
  return b.d;
}
#define exp(x) jit(exp(x), (x))
#define log(x) jit(log(x), (x))
#define pow(x, y) jit(pow((x), (y)), (x) * 1.37 + (y))
#define PHO_BIG 1.0e10           /* common.h:21 */

/* ------------------------------------------------------------------------------------------
 * scene-level constants (what samodel.c:505-618 derives once per run)
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  int n_scenes, max_bands, n_bottoms, n_spatial, n_smooth;
  int n_bands[PHO_MAX_SCENES];
  double lambda[PHO_MAX_SCENES][PHO_MAX_BANDS];
  double a0[PHO_MAX_SCENES][PHO_MAX_BANDS], a1[PHO_MAX_SCENES][PHO_MAX_BANDS];
  double aw[PHO_MAX_SCENES][PHO_MAX_BANDS], bbw[PHO_MAX_SCENES][PHO_MAX_BANDS];
  double bottom[PHO_MAX_BOTTOMS][PHO_MAX_SCENES][PHO_MAX_BANDS];
  double sec_view[PHO_MAX_SCENES], sec_sun[PHO_MAX_SCENES], tide[PHO_MAX_SCENES];
  double aw640;
} pho_model;

/* per-pixel state: the slice of model_data (common.h:116-190) the hot path reads and writes */
typedef struct {
  int n_regions, origin, nb; /* nb = bottoms active for this pixel (1 or n_bottoms) */
  int n_params;
  int prior_present;
  double h_prior;
  double meas[PHO_MAX_REGIONS][PHO_MAX_SCENES][PHO_MAX_BANDS];
  double r440[PHO_MAX_REGIONS][PHO_MAX_SCENES], r490[PHO_MAX_REGIONS][PHO_MAX_SCENES];
  double r550[PHO_MAX_REGIONS][PHO_MAX_SCENES], r640[PHO_MAX_REGIONS][PHO_MAX_SCENES];
  /* side effects of the last forward-model call */
  double modelled[PHO_MAX_REGIONS][PHO_MAX_SCENES][PHO_MAX_BANDS]; /* Rrs_modelled */
  double rrs_mod[PHO_MAX_REGIONS][PHO_MAX_SCENES][PHO_MAX_BANDS];  /* rrs_modelled */
  double rrs_bot[PHO_MAX_REGIONS][PHO_MAX_SCENES][PHO_MAX_BANDS];  /* rrs_bottom   */
  double K[PHO_MAX_SCENES][PHO_MAX_BANDS];
  double bottom_albedo;
  double e_rrs, e_depth, e_bottom, e_K, e_model;
  /* results */
  double depth, K_min, iod, pct[PHO_MAX_BOTTOMS], P[PHO_MAX_SCENES], G[PHO_MAX_SCENES], X[PHO_MAX_SCENES];
  int bottom_type, converged, n_evals, n_iters, n_restarts;
  int variant;
  int hot;                         /* md->start_at_previous */
  double prev[3 * PHO_MAX_SCENES]; /* md->prev: |P|,|G|,|X| (x100) of the previous optimum, samodel.c:2086-2097 */
  const pho_model *m;
} pho_pixel;

/* common.c:392 -- arguments are FLOATS; the products are evaluated in double. */
int pho_approx_equal(float a, float b, float eps) {
  float d = a - b;
  double fa = fabs((double)a), fb = fabs((double)b);
  return fabs((double)d) <= (fa < fb ? fb : fa) * (double)eps;
}
static int approx_le(float a, float b, float eps) { return a < b || pho_approx_equal(a, b, eps); } /* common.c:400 */
static int approx_ge(float a, float b, float eps) { return a > b || pho_approx_equal(a, b, eps); } /* common.c:404 */

/* common.c:298-333: linear interpolation with float-typed bracketing tests and linear
 * extrapolation outside the table. */
double pho_interp_1d(const double *X, const double *Y, int n, double x) {
  double x0 = 0.0, x1 = 1.0, y0 = 0.0, y1 = 0.0, alpha;
  int i;
  if (pho_approx_equal((float)x, (float)X[0], 1.0e-5f)) return Y[0];
  if (pho_approx_equal((float)x, (float)X[n - 1], 1.0e-5f)) return Y[n - 1];
  if (X[0] < X[n - 1] && x < X[0]) {
    x0 = X[0]; x1 = X[1]; y0 = Y[0]; y1 = Y[1];
  } else if (X[n - 1] > X[0] && x > X[n - 1]) {
    x0 = X[n - 2]; x1 = X[n - 1]; y0 = Y[n - 2]; y1 = Y[n - 1];
  } else {
    for (i = 0; i < n - 1; i++) {
      float lo = (float)X[i], hi = (float)X[i + 1], xf = (float)x;
      if ((approx_le(lo, xf, 1.0e-5f) && approx_ge(hi, xf, 1.0e-5f)) ||
          (approx_ge(lo, xf, 1.0e-5f) && approx_le(hi, xf, 1.0e-5f))) {
        x0 = X[i]; x1 = X[i + 1]; y0 = Y[i]; y1 = Y[i + 1];
        break;
      }
    }
  }
  alpha = (x - x0) / (x1 - x0);
  return y0 * (1.0 - alpha) + y1 * alpha;
}

/* samodel.c:505-618 */
static void model_init(pho_model *m, int nscenes, int maxb, const int *n_bands, const int *wavelengths,
                       const double *theta_v, const double *theta_w, const double *h_tide, int n_smooth,
                       int n_spatial, int n_bottoms) {
  double grid[PH_SPEC_N];
  int i, s, b, k;
  memset(m, 0, sizeof(*m));
  for (i = 0; i < PH_SPEC_N; i++) grid[i] = PH_SPEC_LAMBDA0 + PH_SPEC_DLAMBDA * (double)i;
  m->n_scenes = nscenes; m->n_bottoms = n_bottoms; m->n_spatial = n_spatial; m->n_smooth = n_smooth;
  m->max_bands = 0;
  m->aw640 = pho_interp_1d(grid, PH_SPEC_AW, PH_SPEC_N, 640.0);
  for (s = 0; s < nscenes; s++) {
    double tv = theta_v[s], tw = theta_w[s];
    m->n_bands[s] = n_bands[s];
    if (n_bands[s] > m->max_bands) m->max_bands = n_bands[s];
    m->tide[s] = h_tide[s];
    tv *= PHO_PI / 180.0; /* samodel.c:538 */
    m->sec_view[s] = 1.0 / cos(tv);
    tw *= PHO_PI / 180.0;
    m->sec_sun[s] = 1.0 / cos(tw);
    for (b = 0; b < n_bands[s]; b++) {
      double w = (double)wavelengths[s * maxb + b];
      m->lambda[s][b] = w;
      m->a0[s][b] = pho_interp_1d(grid, PH_SPEC_A0, PH_SPEC_N, w);
      m->a1[s][b] = pho_interp_1d(grid, PH_SPEC_A1, PH_SPEC_N, w);
      m->bbw[s][b] = pho_interp_1d(grid, PH_SPEC_BBW, PH_SPEC_N, w);
      m->aw[s][b] = pho_interp_1d(grid, PH_SPEC_AW, PH_SPEC_N, w);
      for (k = 0; k < n_bottoms; k++) m->bottom[k][s][b] = pho_interp_1d(grid, PH_SPEC_BOTTOM[k], PH_SPEC_N, w);
    }
  }
}

/* common.c:232-276: box mean of radius (smoothing_radius-1), edge clamped, nodata skipped; float math. */
static float smoothed_sample(const float *plane, int i, int j, int nrows, int ncols, int radius, float nodata) {
  float acc = 0.0f, cnt = 0.0f;
  int di, dj;
  for (di = 1 - radius; di < radius; di++) {
    int ii = i + di < 0 ? 0 : (i + di > nrows - 1 ? nrows - 1 : i + di);
    for (dj = 1 - radius; dj < radius; dj++) {
      int jj = j + dj < 0 ? 0 : (j + dj > ncols - 1 ? ncols - 1 : j + dj);
      float v = plane[(size_t)ii * ncols + jj];
      if (!pho_approx_equal(v, nodata, 1.0e-6f)) {
        acc += v;
        cnt += 1.0f;
      }
    }
  }
  if ((double)cnt < 0.5) return nodata;
  return acc / cnt;
}

/* samodel.c:2957-3027: gather the (2*n_spatial-1)^2 neighbourhood; n_sigma * R_sigma[band] is added to every
 * sample when n_sigma != 0 (the depth-error trials, samodel.c:3004-3006). */
static void extract_region_noisy(pho_pixel *px, const float *planes, float nodata, int nrows, int ncols, int i, int j,
                                 float n_sigma, const double *r_sigma, int maxb) {
  const pho_model *m = px->m;
  int nsp = m->n_spatial == 0 ? 1 : m->n_spatial, di, dj, s, b, kr = 0;
  for (di = 1 - nsp; di < nsp; di++) {
    int ii = i + di < 0 ? 0 : (i + di > nrows - 1 ? nrows - 1 : i + di);
    for (dj = 1 - nsp; dj < nsp; dj++) {
      int jj = j + dj < 0 ? 0 : (j + dj > ncols - 1 ? ncols - 1 : j + dj);
      int missing = 0, g = 0;
      for (s = 0; s < m->n_scenes && !missing; s++)
        for (b = 0; b < m->n_bands[s]; b++, g++) {
          float v = smoothed_sample(planes + (size_t)g * nrows * ncols, ii, jj, nrows, ncols, m->n_smooth, nodata);
          if (pho_approx_equal(v, nodata, 1.0e-6f)) { missing = 1; break; }
          px->meas[kr][s][b] = (double)v;
          if (!pho_approx_equal(n_sigma, 0.0f, 1.0e-6f)) px->meas[kr][s][b] += n_sigma * r_sigma[s * maxb + b];
        }
      if (i == ii && j == jj) px->origin = kr; /* samodel.c:3016: last clamped match wins */
      if (!missing) kr++;
    }
  }
  px->n_regions = kr;
}

static void extract_region(pho_pixel *px, const float *planes, float nodata, int nrows, int ncols, int i, int j) {
  extract_region_noisy(px, planes, nodata, nrows, ncols, i, j, 0.0f, NULL, 0);
}

/* ------------------------------------------------------------------------------------------
 * forward model + objective
 * ---------------------------------------------------------------------------------------- */

/* samodel.c:2846-2949 (DELTA == 0): one (region, scene, band) of the Lee/HOPE model. */
static void forward_Rrs(pho_pixel *px, double P, double G, double X, const double *B, const double *q, double H,
                        int s, int b, int r) {
  const pho_model *m = px->m;
  const double S = 0.015;
  double rho = 0.0, a_phi, a_g, a, chi, Y, b_p, bb, u, K, rrs_dp, DuC, DuB, M, rrs_C, rrs_B, rrs;
  int k;
  px->bottom_albedo = 0.0;
  for (k = 0; k < px->nb; k++) {
    px->bottom_albedo += q[k] * B[k];
    rho += q[k] * B[k] * m->bottom[k][s][b]; /* bottom_type_indexes[k] == k on this path */
  }
  a_phi = (m->a0[s][b] + m->a1[s][b] * log(fabs(P))) * fabs(P);
  a_g = fabs(G) * exp(-S * (m->lambda[s][b] - 440.0));
  a = m->aw[s][b] + a_phi + a_g;
  chi = px->r440[r][s] / px->r490[r][s];
  Y = 3.44 * (1.0 - 3.17 * exp(-2.01 * chi));
  if (Y < 0.0) Y = 0.0;
  if (Y > 2.5) Y = 2.5;
  b_p = X * pow(440.0 / m->lambda[s][b], Y);
  bb = m->bbw[s][b] + b_p;
  u = bb / (a + bb);
  K = a + bb;
  if (K < 0.0) K = 0.0;
  if (K > 2.5) K = 2.5;
  px->K[s][b] = K;
  rrs_dp = (0.084 + 0.170 * u) * u;
  DuC = 1.03 * sqrt(1.0 + 2.4 * u);
  DuB = 1.04 * sqrt(1.0 + 5.4 * u);
  M = m->sec_sun[s] + DuC * m->sec_view[s];
  rrs_C = rrs_dp * (1.0 - exp(-M * K * H));
  M = m->sec_sun[s] + DuB * m->sec_view[s];
  rrs_B = rho / PHO_PI * exp(-M * K * H);
  px->rrs_bot[r][s][b] = rrs_B;
  rrs = rrs_C + rrs_B;
  px->rrs_mod[r][s][b] = rrs;
  px->modelled[r][s][b] = 0.5 * rrs / (1.0 - 1.5 * rrs) + 0.0;
}

/* samodel.c:2432-2759 (SAM == 0, DELTA == 0): the weighted objective. */
static double objective(const double *x, pho_pixel *px) {
  const pho_model *m = px->m;
  const int Nr = px->n_regions, Nb = px->nb, off = Nr + 2 * Nb * Nr;
  double B[PHO_MAX_BOTTOMS], q[PHO_MAX_BOTTOMS];
  double err = 0.0, tot = 0.0, q_sum, mean_meas, e_rrs, e_spec;
  double e_depth = 0.0, depth_mean = 0.0, n_out = 0.0, depth_thr;
  double e_bottom = 0.0, bottom_thr, bottom_total = 0.0, bottom_mean = 0.0;
  double e_K = 0.0, K_min = 0.0, H;
  const double min_mean_K = 0.275, min_min_K = 0.185;
  int r, s, b, k, kb, ntot = 0;
  double part[32];
  int tsum = (px->variant & PHO_VARIANT_TREE_SUM) != 0, t = 0;
  if (tsum) memset(part, 0, sizeof(part));

  for (r = 0; r < Nr; r++)
    for (s = 0; s < m->n_scenes; s++) {
      double Hh = fabs(x[r]);
      double P = 0.01 * fabs(x[off + 3 * s]);
      double G = 0.01 * fabs(x[off + 1 + 3 * s]);
      double X = 0.01 * fabs(x[off + 2 + 3 * s]);
      for (k = 0; k < Nb; k++) B[k] = 0.01 * fabs(x[Nr + r * Nb + k]);
      q_sum = 0.0;
      for (k = 0; k < Nb; k++) {
        q[k] = fabs(x[Nr + Nr * Nb + r * Nb + k]);
        q_sum += q[k];
      }
      for (k = 0; k < Nb; k++) q[k] /= q_sum;
      for (b = 0; b < m->n_bands[s]; b++) {
        double d;
        forward_Rrs(px, P, G, X, B, q, Hh, s, b, r);
        d = px->modelled[r][s][b] - px->meas[r][s][b];
        if (tsum) part[t & 31] += d * d; else err += d * d; /* pow(d, 2.0) */
        t++;
        tot += px->meas[r][s][b];
        ntot++;
      }
    }
  if (tsum) { /* what-if: the GPU's 32 strided partial sums combined by an xor butterfly */
    int o, l;
    for (o = 16; o >= 1; o >>= 1) {
      double nx[32];
      for (l = 0; l < 32; l++) nx[l] = part[l] + part[l ^ o];
      memcpy(part, nx, sizeof(part));
    }
    err = part[0];
  }
  mean_meas = tot / ((double)ntot);
  e_rrs = 100.0 * sqrt(err / ((double)ntot)) / mean_meas;
  px->e_rrs = e_rrs;
  e_spec = e_rrs * 1.0; /* SAM error == 1 (samodel.c:2587-2594) */

  /* depth continuity over the region, samodel.c:2596-2629 */
  for (r = 0; r < Nr; r++) depth_mean += fabs(x[r]);
  depth_mean /= (double)Nr;
  if (depth_mean < 4.0) depth_thr = 0.4;
  else if (depth_mean < 8.0) depth_thr = 0.2;
  else if (depth_mean < 12.0) depth_thr = 0.1;
  else depth_thr = 0.05;
  for (r = 0; r < Nr; r++)
    if (fabs(x[r]) < (1.0 - depth_thr) * depth_mean || fabs(x[r]) > (1.0 + depth_thr) * depth_mean) {
      double d = fabs(x[r]) - depth_mean;
      e_depth += d * d;
      n_out += 1.0;
    }
  if (n_out > 0.5) e_depth = 100.0 * sqrt(e_depth / n_out) / depth_mean;
  px->e_depth = e_depth;

  /* bottom continuity over the region, samodel.c:2631-2692 */
  if (depth_mean < 5.0) bottom_thr = 0.25;
  else if (depth_mean < 10.0) bottom_thr = 0.1;
  else if (depth_mean < 15.0) bottom_thr = 0.05;
  else bottom_thr = 0.01;
  n_out = 0.0;
  for (k = 0; k < Nb; k++) {
    bottom_mean = 0.0;
    for (r = 0; r < Nr; r++) {
      q_sum = 0.0;
      for (kb = 0; kb < Nb; kb++) q_sum += fabs(x[Nr + Nr * Nb + r * Nb + kb]);
      bottom_mean += fabs(x[Nr + r * Nb + k]) * fabs(x[Nr + Nr * Nb + r * Nb + k]) / q_sum;
    }
    bottom_mean /= (double)Nr;
    bottom_total += bottom_mean;
    for (r = 0; r < Nr; r++) {
      double bot;
      q_sum = 0.0;
      for (kb = 0; kb < Nb; kb++) q_sum += fabs(x[Nr + Nr * Nb + r * Nb + kb]);
      bot = fabs(x[Nr + r * Nb + k]) * fabs(x[Nr + Nr * Nb + r * Nb + k]) / q_sum;
      if (bot < (1.0 - bottom_thr) * bottom_mean || bot > (1.0 + bottom_thr) * bottom_mean) {
        double d = bot - bottom_mean;
        e_bottom += d * d;
        n_out += 1.0;
      }
    }
  }
  if (n_out > 0.5) {
    bottom_mean = bottom_total / ((double)Nb);
    e_bottom = 100.0 * sqrt(e_bottom / n_out) / bottom_mean;
  }
  px->e_bottom = e_bottom;

  /* K penalties, samodel.c:2694-2732. K[s][b] is what the LAST region left behind. */
  H = fabs(x[px->origin]);
  for (s = 0; s < m->n_scenes; s++) {
    const double t2 = 0.5 * (1.5 * min_min_K + 0.5 * min_mean_K);
    const double t3 = 0.5 * (1.25 * min_min_K + 0.75 * min_mean_K);
    const double t5 = 0.5 * (1.75 * min_min_K + 0.25 * min_mean_K);
    K_min = 1.0e4;
    for (b = 0; b < m->max_bands; b++)
      if (!pho_approx_equal((float)px->K[s][b], 0.0f, 1.0e-6f) && px->K[s][b] < K_min) K_min = px->K[s][b];
    if (H < 1.0 && K_min < min_mean_K) {
      double d = 1.0 / (0.01 + K_min) - 1.0 / (0.01 + min_min_K);
      e_K += 100.0 * (d * d);
    } else if (H < 2.0 && K_min < t2) {
      double d = 1.0 / (0.01 + K_min) - 1.0 / (0.01 + t2);
      e_K += 100.0 * (d * d);
    } else if (H < 3.0 && K_min < t3) {
      double d = 1.0 / (0.01 + K_min) - 1.0 / (0.01 + t3);
      e_K += 100.0 * (d * d);
    } else if (H < 4.0 && K_min < t2) {
      double d = 1.0 / (0.01 + K_min) - 1.0 / (0.01 + t2);
      e_K += 100.0 * (d * d);
    } else if (H < 5.0 && K_min < t5) {
      double d = 1.0 / (0.01 + K_min) - 1.0 / (0.01 + t5);
      e_K += 100.0 * (d * d);
    }
  }
  if (K_min > 0.7) { /* last scene's K_min only */
    double d = 4.0 * (K_min - 0.7);
    e_K += 100.0 * (d * d);
  }
  px->e_K = e_K;

  px->e_model = (80.0 * e_spec + 15.0 * e_depth + 10.0 * e_bottom + 15.0 * e_K) / (80.0 + 15.0 + 10.0 + 15.0);
  return px->e_model;
}

/* ------------------------------------------------------------------------------------------
 * Nelder-Mead, O'Neill AS 47 as modified by the reference (asa047.c:10-502)
 * ---------------------------------------------------------------------------------------- */
typedef double (*pho_fn)(const double *x, void *ctx);

typedef struct {
  int n;
  double *p;   /* vertex j at p[j*n .. j*n+n-1], j = 0..n */
  double *y;   /* n+1 values */
  double *sum; /* running vertex sum (PHO_VARIANT_INCR_CENTROID only) */
} simplex;

static int arg_lowest(const simplex *sx, double *ylo) { /* first minimum, asa047.c:201-211 */
  int i, ilo = 0;
  *ylo = sx->y[0];
  for (i = 1; i <= sx->n; i++)
    if (sx->y[i] < *ylo) { *ylo = sx->y[i]; ilo = i; }
  return ilo;
}

static void resum(simplex *sx) {
  int i, j, n = sx->n;
  for (i = 0; i < n; i++) {
    double z = 0.0;
    for (j = 0; j <= n; j++) z = z + sx->p[i + j * n];
    sx->sum[i] = z;
  }
}

static void replace_vertex(simplex *sx, int j, const double *v, double fv, int incr) {
  int i, n = sx->n;
  if (incr)
    for (i = 0; i < n; i++) sx->sum[i] = (sx->sum[i] - sx->p[i + j * n]) + v[i];
  for (i = 0; i < n; i++) sx->p[i + j * n] = v[i];
  sx->y[j] = fv;
}

/* Returns through the same out-parameters as the reference. start[] is clobbered on restart. */
static void nelder_mead(pho_fn fn, void *ctx, int n, double *start, double *xmin, double *ynewlo, double reqmin,
                        const double *step, int konvge, int kcount, int *icount, int *numres, int *ifault,
                        int *iters, int variant) {
  const double ccoeff = 0.5, ecoeff = 2.0, rcoeff = 1.0, eps = 1.0e-6, rscale = 10.0;
  const int incr = (variant & PHO_VARIANT_INCR_CENTROID) != 0;
  const int nn = n + 1;
  const double dn = (double)n, dnn = (double)nn, rq = reqmin * dn;
  double del = 1.0, ylo, ystar, y2star, z, x;
  double *pstar, *p2star, *pbar;
  simplex sx;
  int i, j, ihi, ilo, jcount, l;
  long int zr, yrnewlo;

  if (iters) *iters = 0;
  if (reqmin <= 0.0 || n < 1 || konvge < 1) { *ifault = 1; return; }
  sx.n = n;
  sx.p = (double *)malloc(sizeof(double) * n * nn);
  sx.y = (double *)malloc(sizeof(double) * nn);
  sx.sum = (double *)malloc(sizeof(double) * n);
  pstar = (double *)malloc(sizeof(double) * n);
  p2star = (double *)malloc(sizeof(double) * n);
  pbar = (double *)malloc(sizeof(double) * n);
  *icount = 0;
  *numres = 0;
  jcount = konvge;

  for (;;) { /* initial or restarted simplex, asa047.c:174-211 */
    for (i = 0; i < n; i++) sx.p[i + n * n] = start[i];
    sx.y[n] = fn(start, ctx);
    *icount += 1;
    for (j = 0; j < n; j++) {
      x = start[j];
      start[j] = start[j] + step[j] * del;
      for (i = 0; i < n; i++) sx.p[i + j * n] = start[i];
      sx.y[j] = fn(start, ctx);
      *icount += 1;
      start[j] = x;
    }
    ilo = arg_lowest(&sx, &ylo);
    if (incr) resum(&sx);

    for (;;) { /* asa047.c:215-435 */
      if (kcount <= *icount) break;
      *ynewlo = sx.y[0];
      ihi = 0;
      for (i = 1; i < nn; i++)
        if (*ynewlo < sx.y[i]) { *ynewlo = sx.y[i]; ihi = i; }
      if (iters) *iters += 1;
      /* centroid of all vertices but ihi: sum everything in vertex order, then subtract */
      if (incr) {
        for (i = 0; i < n; i++) pbar[i] = (sx.sum[i] - sx.p[i + ihi * n]) / dn;
      } else {
        for (i = 0; i < n; i++) {
          z = 0.0;
          for (j = 0; j < nn; j++) z = z + sx.p[i + j * n];
          z = z - sx.p[i + ihi * n];
          pbar[i] = z / dn;
        }
      }
      for (i = 0; i < n; i++) pstar[i] = pbar[i] + rcoeff * (pbar[i] - sx.p[i + ihi * n]);
      ystar = fn(pstar, ctx);
      *icount += 1;
      if (ystar < ylo) { /* expansion */
        for (i = 0; i < n; i++) p2star[i] = pbar[i] + ecoeff * (pstar[i] - pbar[i]);
        y2star = fn(p2star, ctx);
        *icount += 1;
        if (ystar < y2star) replace_vertex(&sx, ihi, pstar, ystar, incr);
        else replace_vertex(&sx, ihi, p2star, y2star, incr);
      } else {
        l = 0;
        for (i = 0; i < nn; i++)
          if (ystar < sx.y[i]) l++;
        if (1 < l) {
          replace_vertex(&sx, ihi, pstar, ystar, incr);
        } else if (l == 0) { /* contraction on the y[ihi] side */
          for (i = 0; i < n; i++) p2star[i] = pbar[i] + ccoeff * (sx.p[i + ihi * n] - pbar[i]);
          y2star = fn(p2star, ctx);
          *icount += 1;
          if (sx.y[ihi] < y2star) { /* shrink everything towards the best vertex */
            for (j = 0; j < nn; j++) {
              for (i = 0; i < n; i++) {
                sx.p[i + j * n] = (sx.p[i + j * n] + sx.p[i + ilo * n]) * 0.5;
                xmin[i] = sx.p[i + j * n];
              }
              sx.y[j] = fn(xmin, ctx);
              *icount += 1;
            }
            ilo = arg_lowest(&sx, &ylo);
            if (incr) resum(&sx);
            continue; /* jcount is NOT decremented on this path */
          }
          replace_vertex(&sx, ihi, p2star, y2star, incr);
        } else { /* l == 1: contraction on the reflection side */
          for (i = 0; i < n; i++) p2star[i] = pbar[i] + ccoeff * (pstar[i] - pbar[i]);
          y2star = fn(p2star, ctx);
          *icount += 1;
          if (y2star <= ystar) replace_vertex(&sx, ihi, p2star, y2star, incr);
          else replace_vertex(&sx, ihi, pstar, ystar, incr);
        }
      }
      if (sx.y[ihi] < ylo) { ylo = sx.y[ihi]; ilo = ihi; }
      jcount -= 1;
      if (0 < jcount) continue;
      if (*icount <= kcount) { /* variance test every konvge iterations, asa047.c:411-434 */
        jcount = konvge;
        z = 0.0;
        for (i = 0; i < nn; i++) z = z + sx.y[i];
        x = z / dnn;
        z = 0.0;
        for (i = 0; i < nn; i++) z = z + (sx.y[i] - x) * (sx.y[i] - x);
        if (z <= rq) break;
        if (incr) resum(&sx); /* bound the drift of the running sum */
      }
    }

    /* factorial test that ylo is a local minimum, asa047.c:440-484 */
    for (i = 0; i < n; i++) xmin[i] = sx.p[i + ilo * n];
    *ynewlo = sx.y[ilo];
    yrnewlo = (long int)(rscale * sx.y[ilo]);
    if (kcount < *icount) { *ifault = 2; break; }
    *ifault = 0;
    for (i = 0; i < n; i++) {
      del = step[i] * eps;
      xmin[i] = xmin[i] + del;
      z = fn(xmin, ctx);
      zr = (long int)(rscale * z);
      *icount += 1;
      if (zr < yrnewlo) { *ifault = 2; break; }
      xmin[i] = xmin[i] - del - del;
      z = fn(xmin, ctx);
      zr = (long int)(rscale * z);
      *icount += 1;
      if (zr < yrnewlo) { *ifault = 2; break; }
      xmin[i] = xmin[i] + del;
    }
    if (*ifault == 0) break;
    for (i = 0; i < n; i++) start[i] = xmin[i]; /* restart from the (perturbed) xmin */
    del = eps;
    *numres += 1;
  }
  free(sx.p); free(sx.y); free(sx.sum); free(pstar); free(p2star); free(pbar);
}

static double objective_cb(const double *x, void *ctx) { return objective(x, (pho_pixel *)ctx); }

/* ------------------------------------------------------------------------------------------
 * per-pixel driver
 * ---------------------------------------------------------------------------------------- */

/* where pho_invert_pixels reports nelmin's restart count (asa047.c:493) summed over the H starts of a pixel; same
 * switch as ref_set_numres_out in oracle/ref_harness.c */
static int *g_numres_out = NULL;
void pho_set_numres_out(int *buf) { g_numres_out = buf; }

/* samodel.c:2119-2427 */
static void optimise_one_combination(pho_pixel *px, double *params) {
  const pho_model *m = px->m;
  const int Nr = px->n_regions, Nb = px->nb, n = px->n_params, off = Nr + 2 * Nb * Nr;
  const double reqmin = 1.0e-2;
  const int konvge = 100, kcount = 5000;
  static const double h_slow[8] = {40.0, 30.0, 20.0, 15.0, 10.0, 7.5, 2.5, 1.0};
  double h_quick[1];
  const double *h_start;
  double start[PHO_MAX_PARAMS], step[PHO_MAX_PARAMS], xmin[PHO_MAX_PARAMS], best[PHO_MAX_PARAMS];
  double lowest = 1.0e4, error, mean490, mean550, mean640;
  int n_h, kh, r, s, k, i, icount, numres, ifault, iters;

  memset(best, 0, sizeof(best));
  h_quick[0] = px->h_prior;
  if (px->prior_present) { n_h = 1; h_start = h_quick; } else { n_h = 8; h_start = h_slow; }
  if (px->hot) n_h = 1; /* samodel.c:2235-2241 (callers guarantee a prior: depth_prev is not modelled) */

  for (kh = 0; kh < n_h; kh++) {
    for (r = 0; r < Nr; r++) { start[r] = h_start[kh]; step[r] = 1.25 * start[r]; }
    mean490 = 0.0;
    for (r = 0; r < Nr; r++)
      for (s = 0; s < m->n_scenes; s++) mean490 += px->r490[r][s];
    mean490 /= (double)Nr * m->n_scenes;
    for (k = 0; k < Nb; k++)
      for (r = 0; r < Nr; r++) {
        start[Nr + r * Nb + k] = 100.0 * 4.0 * mean490;
        step[Nr + r * Nb + k] = 1.5 * start[Nr + r * Nb + k];
      }
    for (k = 0; k < Nb; k++)
      for (r = 0; r < Nr; r++) {
        start[Nr + Nb * Nr + r * Nb + k] = 1.0;
        step[Nr + Nb * Nr + r * Nb + k] = 0.5 * start[Nr + Nb * Nr + r * Nb + k];
      }
    for (s = 0; s < m->n_scenes; s++) {
      mean490 = 0.0; mean550 = 0.0; mean640 = 0.0;
      for (r = 0; r < Nr; r++) { mean490 += px->r490[r][s]; mean550 += px->r550[r][s]; mean640 += px->r640[r][s]; }
      mean490 /= (double)Nr; mean550 /= (double)Nr; mean640 /= (double)Nr;
      start[off + 3 * s] = 100.0 * 0.072 * pow(mean490 / mean550, -1.7);
      start[off + 1 + 3 * s] = 1.5 * start[off + 3 * s];
      start[off + 2 + 3 * s] = 100.0 * 30.0 * m->aw640 * mean640;
      if (px->hot) { /* samodel.c:2286-2309: P, G, X of the previous optimum */
        start[off + 3 * s] = px->prev[3 * s];
        start[off + 1 + 3 * s] = px->prev[3 * s + 1];
        start[off + 2 + 3 * s] = px->prev[3 * s + 2];
      }
      step[off + 3 * s] = 2.0 * start[off + 3 * s];
      step[off + 1 + 3 * s] = 2.0 * start[off + 1 + 3 * s];
      step[off + 2 + 3 * s] = 2.0 * start[off + 2 + 3 * s];
    }
    error = objective(start, px);
    icount = 0; numres = 0; ifault = 0;
    nelder_mead(objective_cb, px, n, start, xmin, &error, reqmin, step, konvge, kcount, &icount, &numres, &ifault,
                &iters, px->variant);
    px->n_restarts += numres;
    if (error < lowest) {
      lowest = error;
      for (i = 0; i < n; i++) best[i] = xmin[i];
      px->n_evals = icount;
      px->n_iters = iters;
      px->converged = (ifault == 0);
      if (lowest < 2.5 * ((float)m->n_scenes)) break;
    }
  }
  error = objective(best, px); /* recompute the side effects at the optimum, samodel.c:2413 */
  for (i = 0; i < n; i++) params[i] = fabs(best[i]);
}

/* samodel.c:1768-2115 (ALL_BOTTOMS == 1 path) */
static void optimise_pixel(pho_pixel *px) {
  const pho_model *m = px->m;
  double best[PHO_MAX_PARAMS], origin_w, q_sum, largest, nobs, Kmin;
  int Nr = px->n_regions, r, s, b, k, off;

  for (r = 0; r < Nr; r++)
    for (s = 0; s < m->n_scenes; s++) {
      px->r440[r][s] = pho_interp_1d(m->lambda[s], px->meas[r][s], m->n_bands[s], 440.0);
      px->r490[r][s] = pho_interp_1d(m->lambda[s], px->meas[r][s], m->n_bands[s], 490.0);
      px->r550[r][s] = pho_interp_1d(m->lambda[s], px->meas[r][s], m->n_bands[s], 550.0);
      px->r640[r][s] = pho_interp_1d(m->lambda[s], px->meas[r][s], m->n_bands[s], 640.0);
      if (px->r440[r][s] < 0.0) px->r440[r][s] = 0.0001;
    }
  px->nb = (px->h_prior > 8.0) ? 1 : m->n_bottoms; /* samodel.c:1781,1826 */
  px->n_params = Nr + 2 * Nr * px->nb + 3 * m->n_scenes;
  optimise_one_combination(px, best);

  /* K_min: common.c:979-995 over K[n_scenes][max_bands], spval 0, float-typed approx_equal 1e-4 */
  Kmin = PHO_BIG;
  for (s = 0; s < m->n_scenes; s++)
    for (b = 0; b < m->max_bands; b++)
      if (!pho_approx_equal((float)px->K[s][b], 0.0f, 1.0e-4f) && px->K[s][b] < Kmin) Kmin = px->K[s][b];
  px->K_min = (Kmin == PHO_BIG) ? 0.0 : Kmin;

  px->depth = 0.0; /* samodel.c:2000-2017 */
  origin_w = sqrt((double)Nr);
  for (r = 0; r < Nr; r++) {
    double H = fabs(best[r]);
    if (r == px->origin) px->depth += origin_w * H; else px->depth += H;
  }
  px->depth /= origin_w + ((double)Nr) - 1.0;

  for (k = 0; k < PHO_MAX_BOTTOMS; k++) px->pct[k] = 0.0; /* samodel.c:2024-2047 */
  q_sum = 0.0;
  for (k = 0; k < px->nb; k++) q_sum += fabs(best[Nr + Nr * px->nb + px->origin * px->nb + k]);
  for (k = 0; k < px->nb; k++) px->pct[k] = 100.0 * fabs(best[Nr + Nr * px->nb + px->origin * px->nb + k]) / q_sum;
  largest = 0.0;
  for (k = 0; k < px->nb; k++)
    if (px->pct[k] > largest) { largest = px->pct[k]; px->bottom_type = 1 + k; }

  px->iod = 0.0; /* samodel.c:2051-2064 */
  nobs = 0.0;
  for (s = 0; s < m->n_scenes; s++)
    for (r = 0; r < Nr; r++)
      for (b = 0; b < m->n_bands[s]; b++) {
        px->iod += px->rrs_bot[r][s][b] / px->rrs_mod[r][s][b];
        nobs += 1.0;
      }
  px->iod = 100.0 * px->iod / nobs;

  off = Nr + 2 * px->nb * Nr; /* samodel.c:2068-2079 */
  for (s = 0; s < m->n_scenes; s++) {
    px->P[s] = 0.01 * fabs(best[off + 3 * s]);
    px->G[s] = 0.01 * fabs(best[off + 3 * s + 1]);
    px->X[s] = 0.01 * fabs(best[off + 3 * s + 2]);
    px->prev[3 * s] = fabs(best[off + 3 * s]); /* samodel.c:2086-2097 */
    px->prev[3 * s + 1] = fabs(best[off + 3 * s + 1]);
    px->prev[3 * s + 2] = fabs(best[off + 3 * s + 2]);
  }
}

/* ------------------------------------------------------------------------------------------
 * flat C ABI (mirrors oracle/ref_harness.c entry for entry)
 * ---------------------------------------------------------------------------------------- */
int pho_record_len(int nscenes, int maxb) { return 16 + nscenes * maxb + 3 * nscenes; }

int pho_tables(int nscenes, int maxb, const int *n_bands, const int *wavelengths, const double *theta_v,
               const double *theta_w, const double *h_tide, const double *r_sigma, int n_bottoms, double *out,
               double *out2) {
  pho_model m;
  int s, b, k, o = 0;
  (void)r_sigma;
  model_init(&m, nscenes, maxb, n_bands, wavelengths, theta_v, theta_w, h_tide, 1, 2, n_bottoms);
  for (s = 0; s < nscenes; s++)
    for (b = 0; b < n_bands[s]; b++) {
      double *p = out + (size_t)(s * maxb + b) * (4 + n_bottoms);
      p[0] = m.a0[s][b]; p[1] = m.a1[s][b]; p[2] = m.aw[s][b]; p[3] = m.bbw[s][b];
      for (k = 0; k < n_bottoms; k++) p[4 + k] = m.bottom[k][s][b];
    }
  out2[o++] = m.aw640;
  for (s = 0; s < nscenes; s++) out2[o++] = m.sec_view[s];
  for (s = 0; s < nscenes; s++) out2[o++] = m.sec_sun[s];
  return 0;
}

int pho_invert_pixels_variant(int variant, int nscenes, int maxb, const int *n_bands, const int *wavelengths,
                              const double *theta_v, const double *theta_w, const double *h_tide,
                              const double *r_sigma, int n_smooth, int n_spatial, int n_bottoms, int nrows,
                              int ncols, const float *planes, float nodata, const float *prior,
                              float prior_nodata, int npix, const int *pix_i, const int *pix_j, double *rec,
                              int *status, int *converged, int *n_iterations, int *n_iters, int nthreads) {
  pho_model m;
  int reclen = pho_record_len(nscenes, maxb);
  (void)r_sigma;
  if (nscenes > PHO_MAX_SCENES || maxb > PHO_MAX_BANDS || n_bottoms > PHO_MAX_BOTTOMS ||
      (2 * n_spatial - 1) * (2 * n_spatial - 1) > PHO_MAX_REGIONS)
    return 1;
  model_init(&m, nscenes, maxb, n_bands, wavelengths, theta_v, theta_w, h_tide, n_smooth, n_spatial, n_bottoms);
  g_jitter = (variant & PHO_VARIANT_LIBM_JITTER) != 0; /* after model_init: tables stay exact */
#if _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel
#endif
  {
    pho_pixel *px = (pho_pixel *)calloc(1, sizeof(pho_pixel));
    int p;
    px->m = &m;
    px->variant = variant;
#if _OPENMP
#pragma omp for schedule(dynamic)
#endif
    for (p = 0; p < npix; p++) {
      int i = pix_i[p], j = pix_j[p], s, b, g = 0, k, o, bad = 0;
      double *R = rec + (size_t)p * reclen;
      for (k = 0; k < reclen; k++) R[k] = 0.0;
      status[p] = 0; converged[p] = 0; n_iterations[p] = 0;
      if (n_iters) n_iters[p] = 0;
      for (s = 0; s < nscenes && !bad; s++) /* validity rule, samodel.c:933-947 */
        for (b = 0; b < n_bands[s]; b++, g++) {
          float v = planes[((size_t)g * nrows + i) * ncols + j];
          if (pho_approx_equal(v, nodata, 1.0e-6f) || v < 0.0) { bad = 1; break; }
        }
      if (bad) continue;
      memset(px->K, 0, sizeof(px->K));
      extract_region(px, planes, nodata, nrows, ncols, i, j);
      if (px->n_regions == 0) continue;
      px->prior_present = 0; px->h_prior = 0.0; /* samodel.c:960-976 */
      if (prior != NULL) {
        float e = prior[(size_t)i * ncols + j];
        if (!pho_approx_equal(e, prior_nodata, 1.0e-6f)) {
          px->prior_present = 1;
          px->h_prior = (e > -1.0) ? 1.0 : fabs(e);
        }
      }
      px->n_restarts = 0;
      optimise_pixel(px);
      if (g_numres_out) g_numres_out[p] = px->n_restarts;
      status[p] = 1;
      converged[p] = px->converged;
      n_iterations[p] = px->n_evals;
      if (n_iters) n_iters[p] = px->n_iters;
      R[0] = px->depth; R[1] = px->e_rrs; R[2] = px->bottom_albedo;
      R[3] = px->pct[0]; R[4] = px->pct[1]; R[5] = px->pct[2];
      R[6] = px->K_min; R[7] = px->iod; R[8] = (double)px->bottom_type;
      R[9] = px->e_model; R[10] = px->e_depth; R[11] = px->e_bottom; R[12] = px->e_K;
      R[13] = (double)px->n_regions; R[14] = (double)px->origin; R[15] = px->h_prior;
      o = 16;
      for (s = 0; s < nscenes; s++)
        for (b = 0; b < maxb; b++) R[o++] = (b < n_bands[s]) ? px->K[s][b] : 0.0;
      for (s = 0; s < nscenes; s++) { R[o++] = px->P[s]; R[o++] = px->G[s]; R[o++] = px->X[s]; }
    }
    free(px);
  }
  return 0;
}

int pho_invert_pixels(int nscenes, int maxb, const int *n_bands, const int *wavelengths, const double *theta_v,
                      const double *theta_w, const double *h_tide, const double *r_sigma, int n_smooth,
                      int n_spatial, int n_bottoms, int nrows, int ncols, const float *planes, float nodata,
                      const float *prior, float prior_nodata, int npix, const int *pix_i, const int *pix_j,
                      double *rec, int *status, int *converged, int *n_iterations, int nthreads) {
  return pho_invert_pixels_variant(PHO_VARIANT_EXACT, nscenes, maxb, n_bands, wavelengths, theta_v, theta_w, h_tide,
                                   r_sigma, n_smooth, n_spatial, n_bottoms, nrows, ncols, planes, nodata, prior,
                                   prior_nodata, npix, pix_i, pix_j, rec, status, converged, n_iterations, NULL,
                                   nthreads);
}

int pho_error_kat(int nscenes, int maxb, const int *n_bands, const int *wavelengths, const double *theta_v,
                  const double *theta_w, const double *h_tide, const double *r_sigma, int n_bottoms_active,
                  int n_regions, int origin, const double *rrs_measured, int nparams, int nvec,
                  const double *params, double *out, double *out_Rrs, double *out_K) {
  pho_model m;
  pho_pixel *px = (pho_pixel *)calloc(1, sizeof(pho_pixel));
  int r, s, b, v;
  (void)r_sigma;
  model_init(&m, nscenes, maxb, n_bands, wavelengths, theta_v, theta_w, h_tide, 1, 2, n_bottoms_active);
  px->m = &m;
  px->n_regions = n_regions; px->origin = origin; px->nb = n_bottoms_active; px->n_params = nparams;
  for (r = 0; r < n_regions; r++)
    for (s = 0; s < nscenes; s++) {
      for (b = 0; b < n_bands[s]; b++) px->meas[r][s][b] = rrs_measured[((size_t)r * nscenes + s) * maxb + b];
      px->r440[r][s] = pho_interp_1d(m.lambda[s], px->meas[r][s], n_bands[s], 440.0);
      px->r490[r][s] = pho_interp_1d(m.lambda[s], px->meas[r][s], n_bands[s], 490.0);
      px->r550[r][s] = pho_interp_1d(m.lambda[s], px->meas[r][s], n_bands[s], 550.0);
      px->r640[r][s] = pho_interp_1d(m.lambda[s], px->meas[r][s], n_bands[s], 640.0);
      if (px->r440[r][s] < 0.0) px->r440[r][s] = 0.0001;
    }
  for (v = 0; v < nvec; v++) {
    double *o = out + (size_t)v * 6;
    o[0] = objective(params + (size_t)v * nparams, px);
    o[1] = px->e_rrs; o[2] = px->e_depth; o[3] = px->e_bottom; o[4] = px->e_K; o[5] = px->bottom_albedo;
    if (out_Rrs)
      for (r = 0; r < n_regions; r++)
        for (s = 0; s < nscenes; s++)
          for (b = 0; b < maxb; b++)
            out_Rrs[(((size_t)v * n_regions + r) * nscenes + s) * maxb + b] = (b < n_bands[s]) ? px->modelled[r][s][b] : 0.0;
    if (out_K)
      for (s = 0; s < nscenes; s++)
        for (b = 0; b < maxb; b++) out_K[((size_t)v * nscenes + s) * maxb + b] = (b < n_bands[s]) ? px->K[s][b] : 0.0;
  }
  free(px);
  return 0;
}

/* analytic test functions: identical formulas to oracle/ref_harness.c so that both Nelder-Mead
 * implementations can be compared step for step. */
typedef struct { int n; } tf_ctx;
static double tf_rosenbrock(const double *x, void *c) {
  int n = ((tf_ctx *)c)->n, i;
  double f = 0.0;
  for (i = 0; i + 1 < n; i++) {
    double a = x[i + 1] - x[i] * x[i], b = 1.0 - x[i];
    f += 100.0 * a * a + b * b;
  }
  return f;
}
static double tf_quartic(const double *x, void *c) {
  int n = ((tf_ctx *)c)->n, i;
  double f = 0.0;
  for (i = 0; i < n; i++) {
    double d = x[i] - 0.5 * (double)(i + 1);
    f += d * d * d * d + 0.1 * fabs(d);
  }
  return f;
}
static double tf_steps(const double *x, void *c) {
  int n = ((tf_ctx *)c)->n, i;
  double f = 0.0;
  for (i = 0; i < n; i++) f += floor(fabs(x[i]) * 4.0) * 0.25 + 0.01 * x[i] * x[i];
  return f;
}
int pho_nelmin_kat(int fn_id, int n, const double *start, const double *step, double reqmin, int konvge, int kcount,
                   double *xmin, double *ynewlo, int *icount, int *numres, int *ifault) {
  tf_ctx c;
  double *s = (double *)malloc(n * sizeof(double));
  pho_fn fn = fn_id == 0 ? tf_rosenbrock : (fn_id == 1 ? tf_quartic : tf_steps);
  c.n = n;
  memcpy(s, start, n * sizeof(double));
  *icount = 0; *numres = 0; *ifault = 0;
  nelder_mead(fn, &c, n, s, xmin, ynewlo, reqmin, step, konvge, kcount, icount, numres, ifault, NULL, 0);
  free(s);
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * REFINE: point-wise depth remap (refine.c:215-301). Arithmetic types follow the reference:
 * every variable is float, pow()/fabs() promote to double and the result is narrowed back.
 * flags/args layout is the one declared in include/photic_b200.h (PHB_REFINE_*).
 * ---------------------------------------------------------------------------------------- */
#define RF_CLIP 1
#define RF_SCALE 2
#define RF_LINEAR 4
#define RF_SCRAP 8
#define RF_POWER 16
/* args: 0 clip_min 1 clip_max 2 scale_min 3 scale_max 4 shape 5 linear_m 6 linear_c 7 scrap_min 8 scrap_max
 *       9 power_a 10 power_b */
int pho_refine(int nrows, int ncols, const float *in, float nodata, const float *land, float land_nodata,
               const float *shallow, float shallow_nodata, int flags, const float *args, float *out) {
  float oldmin = args[0], oldmax = args[1], dmin = args[2], dmax = args[3], scale = args[4];
  float linear_m = args[5], linear_c = args[6], scrapmin = args[7], scrapmax = args[8];
  float power_a = args[9], power_b = args[10];
  float smin = 0, smax = 0, sca = 0, scb = 0, v, depth, alpha, beta;
  size_t n = (size_t)nrows * ncols, t;
  if (!(flags & RF_CLIP)) { /* common.c:1224-1262: min/max skipping nodata (approx_equal 1e-4) */
    oldmin = (float)PHO_BIG;
    oldmax = (float)-PHO_BIG;
    for (t = 0; t < n; t++) {
      if (pho_approx_equal(in[t], nodata, 1.0e-4f)) continue;
      if (in[t] < oldmin) oldmin = in[t];
      if (in[t] > oldmax) oldmax = in[t];
    }
  }
  if (flags & RF_SCALE) {
    smin = (float)pow(fabs((double)oldmin), (double)scale);
    if (oldmin < 0.0) smin = (float)((double)smin * -1.0);
    smax = (float)pow(fabs((double)oldmax), (double)scale);
    if (oldmax < 0.0) smax = (float)((double)smax * -1.0);
    sca = (oldmax - oldmin) / (smax - smin);
    scb = (oldmax * smin - oldmin * smax) / (smin - smax);
  }
  for (t = 0; t < n; t++) {
    int open = (land == NULL && shallow == NULL) ||
               ((land != NULL && land[t] != land_nodata) && (shallow != NULL && shallow[t] != shallow_nodata));
    if (!open) { out[t] = nodata; continue; }
    depth = in[t];
    if (depth == nodata) { out[t] = nodata; continue; }
    if (flags & RF_CLIP) {
      if (depth < oldmin) depth = oldmin;
      else if (depth > oldmax) depth = oldmax;
    }
    if (flags & RF_LINEAR) depth = linear_m * depth + linear_c;
    if (flags & RF_SCALE) {
      if (scale != 1.0) {
        v = (dmin * oldmax - dmax * oldmin + dmax * depth - dmin * depth) / (oldmax - oldmin);
        depth = (float)((double)sca * pow(fabs((double)v), (double)scale) + (double)scb);
        if (v < 0.0) depth = (float)((double)depth * -1.0);
      } else {
        alpha = (oldmax * dmin - oldmin * dmax) / (oldmax - oldmin);
        beta = (dmax - dmin) / (oldmax - oldmin);
        depth = alpha + beta * depth;
      }
    }
    if (flags & RF_SCRAP) {
      if (depth < scrapmin || depth > scrapmax) { out[t] = nodata; continue; }
    }
    if (flags & RF_POWER) {
      if (depth < 0.0) depth = (float)(-1.0 * (double)power_a * pow(fabs((double)depth), (double)power_b));
      else depth = (float)((double)power_a * pow(fabs((double)depth), (double)power_b));
    }
    out[t] = depth;
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * depth-error estimate, samodel.c:1376-1477 (arguments as oracle/ref_harness.c:ref_depth_sigma)
 * ---------------------------------------------------------------------------------------- */

static int rand_below(int lo, int hi) { /* random_in_range, common.c:527-543: unbiased bucket of libc rand() */
  for (;;) {
    int v = rand(), range = hi - lo, rem = RAND_MAX % range, bucket = RAND_MAX / range;
    if (v == RAND_MAX) continue;
    if (v < RAND_MAX - rem) return lo + v / bucket;
  }
}
static float rand_unit(void) { return ((float)rand()) / ((float)RAND_MAX); } /* frand, common.c:215 */
static float rand_signed(float max) {                                         /* frand2, common.c:220-225 */
  if (rand_unit() < 0.5) return ((float)rand()) / ((float)RAND_MAX / max);
  return -1.0 * ((float)rand()) / ((float)RAND_MAX / max);
}

int pho_depth_sigma(int nscenes, int maxb, const int *n_bands, const int *wavelengths, const double *theta_v,
                    const double *theta_w, const double *h_tide, const double *r_sigma, int n_smooth,
                    int n_spatial, int n_bottoms, int nrows, int ncols, const float *planes, float nodata,
                    const float *prior, float prior_nodata, const float *depth, unsigned seed, int n_samples,
                    int chain_mode, int max_intervals, double *table, int *n_intervals_out, double *trials,
                    float *depth_sigma) {
  pho_model m;
  pho_pixel *px;
  double d, maxd, *td;
  float mx = -PHO_BIG;
  int n_trials, n_int, kd = 0, ks, kt, i = 0, j = 0, c;
  size_t q;
  if (prior == NULL || nscenes > PHO_MAX_SCENES || maxb > PHO_MAX_BANDS || n_bottoms > PHO_MAX_BOTTOMS ||
      (2 * n_spatial - 1) * (2 * n_spatial - 1) > PHO_MAX_REGIONS)
    return 1;
  model_init(&m, nscenes, maxb, n_bands, wavelengths, theta_v, theta_w, h_tide, n_smooth, n_spatial, n_bottoms);
  g_jitter = 0;
  px = (pho_pixel *)calloc(1, sizeof(pho_pixel));
  px->m = &m;
  srand(seed);
  n_trials = (int)sqrt(nrows * ncols);
  for (q = 0; q < (size_t)nrows * ncols; q++) /* array_max2(depth, ., ., 0.0), common.c:1240-1258 */
    if (!pho_approx_equal(depth[q], 0.0f, 1.0e-4f) && depth[q] > mx) mx = depth[q];
  maxd = mx;
  maxd = 0.25 * ((int)maxd / 0.25);
  if (!(maxd < 30.0)) maxd = 30.0;
  n_int = (int)maxd / 0.25;
  if (n_int > max_intervals) { n_int = max_intervals; maxd = 0.25 * max_intervals; }
  td = (double *)malloc(n_samples * sizeof(double));
  px->hot = 0;
  for (d = 0.0; d < maxd; d += 0.25) {
    double total = 0.0, sumdev = 0.0, mean;
    if (chain_mode == 1) px->hot = 0;
    for (ks = 0; ks < n_samples; ks++) {
      int found = 0;
      float ns, e;
      for (kt = 0; kt < n_trials; kt++) {
        i = rand_below(0, nrows);
        j = rand_below(0, ncols);
        if (depth[(size_t)i * ncols + j] > d && depth[(size_t)i * ncols + j] < d + 0.25) { found = 1; break; }
      }
      if (!found) { td[ks] = 0.0; continue; }
      ns = (float)(double)rand_signed(1.0); /* double n_sigma narrowed at the call, samodel.c:1425-1428 */
      memset(px->K, 0, sizeof(px->K));
      extract_region_noisy(px, planes, nodata, nrows, ncols, i, j, ns, r_sigma, maxb);
      if (px->n_regions == 0) continue;
      e = prior[(size_t)i * ncols + j];
      if (pho_approx_equal(e, prior_nodata, 1.0e-6f)) { td[ks] = 0.0; continue; }
      px->prior_present = 1;
      px->h_prior = (e > -1.0) ? 1.0 : fabs(e);
      optimise_pixel(px);
      px->hot = 1;
      td[ks] = px->depth;
    }
    if (trials) for (ks = 0; ks < n_samples; ks++) trials[(size_t)kd * n_samples + ks] = td[ks];
    c = 0; /* vec_mean2_double returns FLOAT (common.c:877-893), vec_stddev_double common.c:924-941 */
    for (ks = 0; ks < n_samples; ks++)
      if (!pho_approx_equal((float)td[ks], 0.0f, 1.0e-4f)) { total += td[ks]; c++; }
    mean = (c == 0) ? 0.0 : (double)(float)(total / ((double)c));
    c = 0;
    for (ks = 0; ks < n_samples; ks++)
      if (!pho_approx_equal((float)td[ks], 0.0f, 1.0e-4f)) { sumdev += (td[ks] - mean) * (td[ks] - mean); c++; }
    table[kd++] = (c == 0) ? 0.0 : sqrt(sumdev / ((double)c));
  }
  *n_intervals_out = kd;
  if (depth_sigma)
    for (q = 0; q < (size_t)nrows * ncols; q++) {
      depth_sigma[q] = 0.0f;
      if (depth[q] > 0.0) {
        int k = 0;
        for (d = 0.0; d < maxd; d += 0.25, k++)
          if (depth[q] > d && depth[q] <= d + 0.25) { depth_sigma[q] = (float)table[k]; break; }
      }
    }
  free(td); free(px);
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * MODEL Lee_Kd_LS8 / Lee_Secchi_LS8: Lee et al. 2016 QAA-style Kd and Secchi depth for Landsat-8
 * (secchi.c:13-252). Arguments as oracle/ref_harness.c:ref_lee_ls8.
 * ---------------------------------------------------------------------------------------- */
#undef exp
#undef log
#undef pow

/* Lee_Kd_LS8, secchi.c:117-252: inputs and outputs are floats, the arithmetic is double */
static void lee_kd_bands(float R443, float R481, float R554, float R656, float theta_s, float *kd) {
  const double g0 = 0.0895, g1 = 0.1247, aw = 0.05866, h0 = -1.146, h1 = -1.366, h2 = 0.469;
  const double bbw[4] = {0.00244761, 0.00171397, 0.000931339, 0.000448682};
  const double lam[4] = {443.0, 481.0, 554.0, 656.0};
  const double m0 = 0.005, m1 = 4.26, m2 = 0.52, m3 = 10.8, gamma = 0.265;
  const float Rrs[4] = {R443, R481, R554, R656};
  double rrs[4], u[4], a[4], bb[4], bbp[4], chi, eta;
  int b;
  for (b = 0; b < 4; b++) {
    rrs[b] = Rrs[b] / (0.52 + 1.7 * Rrs[b]);
    u[b] = (-g0 + sqrt(g0 * g0 + 4.0 * g1 * rrs[b])) / (2.0 * g1);
  }
  chi = log10((rrs[0] + rrs[1]) / (rrs[2] + 5.0 * rrs[3] * rrs[3] / rrs[1]));
  a[2] = aw + pow(10.0, h0 + h1 * chi + h2 * chi * chi);
  bb[2] = (-a[2] * g0 + 2.0 * a[2] * rrs[2] + a[2] * sqrt(g0 * g0 + 4.0 * g1 * rrs[2])) / (2.0 * (g0 + g1 - rrs[2]));
  bbp[2] = (u[2] * a[2]) / (1 - u[2]) - bbw[2];
  eta = 2.0 * (1.0 - 1.2 * exp(-0.9 * rrs[0] / rrs[2]));
  for (b = 0; b < 4; b++) {
    if (b == 2) continue;
    bbp[b] = bbp[2] * pow(554.0 / lam[b], eta);
    a[b] = (1.0 - u[b]) * (bbw[b] + bbp[b]) / u[b];
    bb[b] = (-a[b] * g0 + 2.0 * a[b] * rrs[b] + a[b] * sqrt(g0 * g0 + 4.0 * g1 * rrs[b])) / (2.0 * (g0 + g1 - rrs[b]));
  }
  for (b = 0; b < 4; b++) {
    double kk1 = (1.0 + m0 * theta_s) * a[b];
    double kk2 = m1 * (1.0 - gamma * bbw[b] / bb[b]);
    double kk3 = (1.0 - m2 * exp(-m3 * a[b])) * bb[b];
    kd[b] = kk1 + kk2 * kk3;
  }
}

static float lee_kd_min(const float *k4) { /* secchi.c:45-52, 100-109: Kd(530) = 0.20 Kd_blue + 0.75 Kd_green; vec_min */
  float kd[5], mn;
  int q;
  kd[0] = k4[0]; kd[1] = k4[1]; kd[2] = 0.20 * k4[1] + 0.75 * k4[2]; kd[3] = k4[2]; kd[4] = k4[3];
  mn = kd[0];
  for (q = 0; q < 5; q++) if (kd[q] < mn) mn = kd[q];
  return mn;
}

int pho_lee_ls8(int mode, int nrows, int ncols, const float *coastal, const float *blue, const float *green,
                const float *red, const float *spv, float theta_s, float *out) {
  size_t q, n = (size_t)nrows * ncols;
  for (q = 0; q < n; q++) {
    float k4[4], kmin;
    if (!(coastal[q] != spv[0] && blue[q] != spv[1] && green[q] != spv[2] && red[q] != spv[3])) { out[q] = spv[0]; continue; }
    lee_kd_bands(coastal[q], blue[q], green[q], red[q], theta_s, k4);
    kmin = lee_kd_min(k4);
    if (mode == 0) out[q] = kmin;
    else { /* Lee_secchi_LS8, secchi.c:86-114 */
      float mx = coastal[q];
      if (blue[q] > mx) mx = blue[q];
      if (green[q] > mx) mx = green[q];
      if (red[q] > mx) mx = red[q];
      out[q] = log(fabs(0.14 - mx) / 0.013) / (2.5 * kmin);
    }
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------------
 * int16 scale/offset packing of a grid for NetCDF (row N4): compress_2d model/nc.c:271-320 and
 * decompress_2d model/nc.c:247-266, restated on flat arrays.
 *
 * Quirks kept (nc.c:287-301): the running maximum starts at FLT_MIN (the smallest positive float, not -FLT_MAX) and
 * sits in the `else` of the running-minimum test, so a value that lowers the minimum when it is met -- the first value
 * always does -- never raises the maximum; the range is therefore order dependent. The short is (short)rint(.) of a
 * FLOAT quotient: on x86-64 that is cvttsd2si to 32 bits (INT_MIN for NaN / out of range) and the low 16 bits of it.
 * ------------------------------------------------------------------------------------------------ */
static short pho_to_short_x86(double v) {
  int w;
  if (!(fabs(v) < 2147483648.0)) w = (int)0x80000000u; /* cvttsd2si: indefinite integer */
  else w = (int)v;
  return (short)(unsigned short)((unsigned)w & 0xffffu);
}

int pho_nc_pack(int nrows, int ncols, const float *grid, double spval, short *packed, float *offset_scale, short *missing) {
  const long n = (long)nrows * ncols;
  const float fspv = (float)spval;               /* nc.c:278 */
  float grmin = FLT_MAX, grmax = FLT_MIN;        /* nc.c:286-287 */
  for (long k = 0; k < n; k++) {                 /* nc.c:289-298, row-major */
    const float v = grid[k];
    if (v == fspv) continue;
    if (v < grmin) grmin = v;
    else if (v > grmax) grmax = v;
  }
  const float scale = (grmax - grmin) / ((float)SHRT_MAX); /* nc.c:306 */
  *missing = SHRT_MIN;                           /* nc.c:282 */
  offset_scale[0] = grmin; offset_scale[1] = scale;
  for (long k = 0; k < n; k++) {                 /* nc.c:311-318 */
    const float v = grid[k];
    if (v == fspv) packed[k] = SHRT_MIN;
    else { const float q = (v - grmin) / scale; packed[k] = pho_to_short_x86(rint((double)q)); }
  }
  return 0;
}

int pho_nc_unpack(int nrows, int ncols, const short *packed, float add_offset, float scale_factor, short missing,
                  double spval, float *grid) {
  const long n = (long)nrows * ncols;
  for (long k = 0; k < n; k++) {                 /* nc.c:254-264 */
    if (packed[k] == missing) grid[k] = (float)spval;
    else { const float t = ((float)packed[k]) * scale_factor; grid[k] = t + add_offset; }
  }
  return 0;
}
