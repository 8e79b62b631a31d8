/*
 * photic_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C CPU restatement of the reference's per-pixel semi-analytical inversion
 * (stblake/photic model/samodel.c, model/asa047.c, model/common.c). It exists to CHECK the
 * CUDA path; it is never imported, linked or executed by the product (photic_b200/).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load it.
 *
 * PARITY PIN: the reference ships no tests, fixtures or golden vectors for this path
 * (SURVEY.md section 4), so the oracle is pinned against outputs of the reference itself:
 * oracle/_ref/libphotic_ref.so is the unmodified reference hot path compiled here
 * (oracle/Makefile), and tests/golden/ holds known-answer dumps generated from it by
 * tests/golden/make_golden.py. tests/test_oracle_vs_golden.py requires bit equality.
 */
#ifndef PHOTIC_ORACLE_H_
#define PHOTIC_ORACLE_H_

#ifdef __cplusplus
extern "C" {
#endif

#define PHO_MAX_SCENES 16
#define PHO_MAX_BANDS 8
#define PHO_MAX_REGIONS 25 /* (2*n_spatial-1)^2 with n_spatial <= 3 */
#define PHO_MAX_BOTTOMS 8
#define PHO_MAX_PARAMS (PHO_MAX_REGIONS * (1 + 2 * PHO_MAX_BOTTOMS) + 3 * PHO_MAX_SCENES)

/* "what-if" arithmetic variants used to measure, on the CPU, how sensitive the optimiser's
 * decision sequence is to the two re-orderings the GPU design makes. */
#define PHO_VARIANT_EXACT 0
#define PHO_VARIANT_TREE_SUM 1      /* error sum: 32 strided partials + butterfly tree     */
#define PHO_VARIANT_INCR_CENTROID 2 /* centroid from a running vertex sum, resummed at checks */
#define PHO_VARIANT_LIBM_JITTER 4   /* exp/log/pow results nudged by a deterministic -1/0/+1 ulp */

int pho_record_len(int nscenes, int maxb);

int pho_tables(int nscenes, int maxb, const int *n_bands, const int *wavelengths, const double *theta_v,
               const double *theta_w, const double *h_tide, const double *r_sigma, int n_bottoms,
               double *out, double *out2);

int pho_invert_pixels(int nscenes, int maxb, const int *n_bands, const int *wavelengths,
                      const double *theta_v, const double *theta_w, const double *h_tide,
                      const double *r_sigma, int n_smooth, int n_spatial, int n_bottoms, int nrows,
                      int ncols, const float *planes, float nodata, const float *prior,
                      float prior_nodata, int npix, const int *pix_i, const int *pix_j, double *rec,
                      int *status, int *converged, int *n_iterations, int nthreads);

/* same as pho_invert_pixels with an arithmetic variant (bit mask of PHO_VARIANT_*) and,
 * optionally, the number of Nelder-Mead iterations per pixel (n_iters, nullable). */
int pho_invert_pixels_variant(int variant, int nscenes, int maxb, const int *n_bands,
                              const int *wavelengths, const double *theta_v, const double *theta_w,
                              const double *h_tide, const double *r_sigma, int n_smooth, int n_spatial,
                              int n_bottoms, int nrows, int ncols, const float *planes, float nodata,
                              const float *prior, float prior_nodata, int npix, const int *pix_i,
                              const int *pix_j, double *rec, int *status, int *converged,
                              int *n_iterations, int *n_iters, int nthreads);

int pho_error_kat(int nscenes, int maxb, const int *n_bands, const int *wavelengths,
                  const double *theta_v, const double *theta_w, const double *h_tide,
                  const double *r_sigma, int n_bottoms_active, int n_regions, int origin,
                  const double *rrs_measured, int nparams, int nvec, const double *params, double *out,
                  double *out_Rrs, double *out_K);

double pho_interp_1d(const double *X, const double *Y, int n, double x);
int pho_approx_equal(float a, float b, float eps);

int pho_nelmin_kat(int fn_id, int n, const double *start, const double *step, double reqmin, int konvge,
                   int kcount, double *xmin, double *ynewlo, int *icount, int *numres, int *ifault);

/* REFINE (model/refine.c:12-302) point-wise remap, flags as in photic_b200.h */
int pho_refine(int nrows, int ncols, const float *in, float nodata, const float *land, float land_nodata,
               const float *shallow, float shallow_nodata, int flags, const float *args, float *out);

/* depth-error estimate of samodel() (samodel.c:1376-1477) with the seed as an argument; see
 * oracle/ref_harness.c:ref_depth_sigma for the arguments */
int pho_depth_sigma(int nscenes, int maxb, const int *n_bands, const int *wavelengths, const double *theta_v,
                    const double *theta_w, const double *h_tide, const double *r_sigma, int n_smooth,
                    int n_spatial, int n_bottoms, int nrows, int ncols, const float *planes, float nodata,
                    const float *prior, float prior_nodata, const float *depth, unsigned seed, int n_samples,
                    int chain_mode, int max_intervals, double *table, int *n_intervals_out, double *trials,
                    float *depth_sigma);

/* MODEL Lee_Kd_LS8 (mode 0) / Lee_Secchi_LS8 (mode 1), secchi.c:13-252; spv = nodata of the four planes */
int pho_lee_ls8(int mode, int nrows, int ncols, const float *coastal, const float *blue, const float *green,
                const float *red, const float *spv, float theta_s, float *out);

#ifdef __cplusplus
}
#endif
#endif
