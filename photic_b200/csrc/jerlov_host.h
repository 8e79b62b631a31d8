/*
 * jerlov_host.h -- scene-level water-type fit and spectral attenuation coefficients.
 *
 * Reference (paths under /root/reference/model/): jerlov.c -- `jerlov` :75-210, `compute_k_from_ratio` :214-270,
 * `compute_k_from_jerlov` :274-282, `compute_k` :284-316, `estimate_water_type` :322-340, `interp_jerlov_ratio`
 * :343-392, `interp_jerlov_wavelength` :395-425; `linear_fit` common.c:418-450. Called once per COMPUTE K verb
 * (bam.c:2362-2392) on a transect / brightest-pixel line of a few hundred points and on <= MAX_GRIDS wavelengths.
 *
 * This is HOST code on purpose: the result is a handful of scalars per scene (like the band tables of
 * build_model()), the regression sums are sequential FP64 accumulations whose order fixes the bits, and there is
 * no raster to spread over a GPU. Arithmetic types follow the reference expression by expression (float products
 * accumulated in double, (float)numerator / double, double `1.0 - w` next to float `w * k`), compiled without FMA
 * contraction like the reference's x86-64 build, so the results are bit-identical (tests/test_jerlov.py against
 * outputs of the unmodified jerlov.c in oracle/_ref).
 *
 * The table is Jerlov (1976) Table XXVII for the coastal types C3..C9 and Austin & Petzold (1984) for the oceanic
 * types I..III and C1 -- the rows the reference has active (jerlov.c:29-51).
 */
#ifndef PHOTIC_JERLOV_HOST_H_
#define PHOTIC_JERLOV_HOST_H_

#include <math.h>

#include <vector>

namespace phb_jerlov {

constexpr int kTypes = 10; /* OI OIA OIB OII OIII C1 C3 C5 C7 C9 */
constexpr int kWl = 13;    /* 400..700 nm in 25 nm steps */

/* K_d [1/m], one row per water type */
static const float kTable[kTypes][kWl] = {
    /* OI   */ {0.0217f, 0.0185f, 0.0176f, 0.0184f, 0.0280f, 0.0504f, 0.0640f, 0.0931f, 0.2408f, 0.3174f, 0.3559f, 0.4372f, 0.6513f},
    /* OIA  */ {0.0316f, 0.0280f, 0.0257f, 0.0250f, 0.0332f, 0.0545f, 0.0674f, 0.0960f, 0.2437f, 0.3206f, 0.3601f, 0.4410f, 0.6530f},
    /* OIB  */ {0.0438f, 0.0395f, 0.0355f, 0.0330f, 0.0396f, 0.0596f, 0.0715f, 0.0995f, 0.2471f, 0.3245f, 0.3652f, 0.4457f, 0.6550f},
    /* OII  */ {0.0878f, 0.0814f, 0.0714f, 0.0620f, 0.0627f, 0.0779f, 0.0863f, 0.1122f, 0.2595f, 0.3389f, 0.3837f, 0.4626f, 0.6623f},
    /* OIII */ {0.1697f, 0.1594f, 0.1381f, 0.1160f, 0.1056f, 0.1120f, 0.1139f, 0.1359f, 0.2826f, 0.3655f, 0.4181f, 0.4942f, 0.6760f},
    /* C1   */ {0.2516f, 0.2374f, 0.2048f, 0.1700f, 0.1486f, 0.1461f, 0.1415f, 0.1596f, 0.3057f, 0.3922f, 0.4525f, 0.5257f, 0.6896f},
    /* C3   */ {0.78f, 0.54f, 0.39f, 0.29f, 0.22f, 0.2f, 0.19f, 0.21f, 0.33f, 0.4f, 0.46f, 0.56f, 0.71f},
    /* C5   */ {1.1f, 0.78f, 0.56f, 0.43f, 0.36f, 0.31f, 0.3f, 0.33f, 0.4f, 0.48f, 0.54f, 0.65f, 0.8f},
    /* C7   */ {1.6f, 1.2f, 0.89f, 0.71f, 0.58f, 0.49f, 0.46f, 0.46f, 0.48f, 0.54f, 0.63f, 0.78f, 0.92f},
    /* C9   */ {2.4f, 1.9f, 1.6f, 1.23f, 0.99f, 0.78f, 0.63f, 0.58f, 0.6f, 0.65f, 0.76f, 0.92f, 1.1f}};

/* column of the table at wavelength wl, all types (jerlov.c:395-425); false outside 400..700 nm */
inline bool column_at(float wl, float col[kTypes]) {
  if (wl < 400.0f || wl > 700.0f) return false;
  int lo = kWl - 2;
  for (int i = 1; i < kWl; i++)
    if (wl <= (float)(400 + 25 * i)) { lo = i - 1; break; }
  const float alpha = (wl - (float)(400 + 25 * lo)) / (float)25;
  for (int t = 0; t < kTypes; t++) col[t] = (1.0f - alpha) * kTable[t][lo] + alpha * kTable[t][lo + 1];
  return true;
}

/* K at the slope `ratio` between two bracketing water types (jerlov.c:343-392): no extrapolation */
inline bool k_at_ratio(float ratio, const float ratios[kTypes], const float col[kTypes], float *k) {
  const float first = ratios[0], last = ratios[kTypes - 1];
  if ((first > last && ratio > first) || (first < last && ratio < first) || (first < last && ratio > last) ||
      (first > last && ratio < last)) {
    *k = 0.0f;
    return false;
  }
  for (int i = 1; i < kTypes; i++) {
    const float a = ratios[i - 1], b = ratios[i];
    if ((ratio <= a && ratio >= b) || (ratio >= a && ratio <= b)) {
      const float alpha = (ratio - a) / (b - a);
      *k = alpha * col[i - 1] + (1.0f - alpha) * col[i]; /* the reference's weights, as they are (jerlov.c:389) */
      return true;
    }
  }
  return false;
}

/* fractional water-type index from the slope (jerlov.c:322-340): strict inequalities, rising or falling */
inline bool water_type_of(float ratio, const float ratios[kTypes], float *wt) {
  for (int i = 1; i < kTypes; i++) {
    const float a = ratios[i - 1], b = ratios[i];
    if (ratio > a && ratio < b) { *wt = ((float)i - 1.0f) + (ratio - a) / (b - a); return true; }
    if (ratio < a && ratio > b) { *wt = ((float)i - 1.0f) + (ratio - b) / (a - b); return true; }
  }
  return false;
}

/* compute_k (jerlov.c:284-316): linear in the water-type index; `1.0 - w` is a double, `w * k` a float product.
 * The reference reads one element past its table for water_type >= 9 (undefined); here that is an error. */
inline bool k_of_type(float water_type, float wl, float *k) {
  float col[kTypes];
  *k = 0.0f;
  if (!(water_type >= 0.0f) || !(water_type < (float)(kTypes - 1))) return false;
  if (!column_at(wl, col)) return true; /* the reference returns 0.0 for a wavelength outside the table */
  const double fl = floor((double)water_type);
  const int i = (int)fl;
  const float w = (float)((double)water_type - fl);
  *k = (float)((double)col[i] * (1.0 - (double)w) + (double)(w * col[i + 1]));
  return true;
}

/* linear_fit (common.c:418-450): y = m x + b with float products summed in double, in index order */
inline bool line_fit(const float *x, const float *y, int n, float *m, float *b, float *r) {
  double Sx = 0.0, Sy = 0.0, Sxx = 0.0, Sxy = 0.0, Syy = 0.0;
  for (int i = 0; i < n; i++) {
    Sx += (double)x[i];
    Sy += (double)y[i];
    Sxx += (double)(x[i] * x[i]);
    Sxy += (double)(x[i] * y[i]);
    Syy += (double)(y[i] * y[i]);
  }
  const double dn = (double)n;
  const double delta = dn * Sxx - Sx * Sx;
  { /* approx_equal((float)delta, 0.0f, 1e-6), common.c:392 */
    const float d = (float)delta;
    if (fabs((double)d) <= fabs((double)d) * (double)1.0e-6f) { *m = 0.0f; *b = 0.0f; *r = 0.0f; return false; }
  }
  *m = (float)((double)(float)(dn * Sxy - Sx * Sy) / delta);
  *b = (float)((double)(float)(Sy * Sxx - Sx * Sxy) / delta);
  *r = (float)((double)(float)(Sxy - Sx * Sy / dn) / sqrt((Sxx - Sx * Sx / dn) * (Syy - Sy * Sy / dn)));
  return true;
}

struct Fit { float ki, kj, m, c, r, water_type; int n_shallow; };

/* jerlov (jerlov.c:75-210): regression of log(Lj - Lsmj) on log(Li - Lsmi) over the optically shallow points,
 * slope -> K_i, K_j and water type by the two-way table interpolation. */
inline bool fit(float wlen_i, float wlen_j, float lsm_i, float lsm_j, const float *Li, const float *Lj, int npoints,
                float manual_ratio, Fit *out) {
  std::vector<float> Xi, Xj;
  Xi.reserve(npoints > 0 ? npoints : 0);
  Xj.reserve(npoints > 0 ? npoints : 0);
  for (int n = 0; n < npoints; n++) {
    if ((double)Li[n] > (double)lsm_i + 1.0 && (double)Lj[n] > (double)lsm_j + 1.0) {
      Xi.push_back((float)log((double)(Li[n] - lsm_i)));
      Xj.push_back((float)log((double)(Lj[n] - lsm_j)));
    }
  }
  out->n_shallow = (int)Xi.size();
  if (!line_fit(Xi.data(), Xj.data(), out->n_shallow, &out->m, &out->c, &out->r)) return false;
  { /* ! approx_equal(manual_ratio, 0.0, 1e-4): any non-zero manual ratio replaces the slope (jerlov.c:137-139) */
    if (!(fabs((double)manual_ratio) <= fabs((double)manual_ratio) * (double)1.0e-4f)) out->m = manual_ratio;
  }
  float ci[kTypes], cj[kTypes], ratios[kTypes];
  if (!column_at(wlen_i, ci) || !column_at(wlen_j, cj)) return false;
  for (int t = 0; t < kTypes; t++) ratios[t] = cj[t] / ci[t];
  if (!k_at_ratio(out->m, ratios, ci, &out->ki)) return false;
  if (!k_at_ratio(out->m, ratios, cj, &out->kj)) return false;
  return water_type_of(out->m, ratios, &out->water_type);
}

/* compute_k_from_ratio (jerlov.c:214-270): water type from a given slope, then K at every wavelength */
inline bool k_from_ratio(float ratio, float wlen_i, float wlen_j, const float *wavelengths, int n, float *water_type,
                         float *k) {
  float ci[kTypes], cj[kTypes], ratios[kTypes];
  if (!column_at(wlen_i, ci) || !column_at(wlen_j, cj)) return false;
  for (int t = 0; t < kTypes; t++) ratios[t] = cj[t] / ci[t];
  if (!water_type_of(ratio, ratios, water_type)) return false;
  bool ok = true;
  for (int i = 0; i < n; i++) ok = k_of_type(*water_type, wavelengths[i], &k[i]) && ok;
  return ok;
}

}  // namespace phb_jerlov

#endif
