/*
 * device_model.cuh -- scene-level constants shared by host set-up code and the kernels.
 *
 * ModelConst is what model/samodel.c:505-618 derives once per run (interpolated a0, a1, a_w, b_bw,
 * bottom spectra, secants) plus a few loop-invariant factors the reference recomputes on every
 * forward-model call with identical inputs (exp(-S(lambda-440)), 440/lambda: samodel.c:2891,2903).
 * They are computed ON THE HOST with the host libm, i.e. with the very instructions the reference
 * uses, so hoisting them cannot change a bit of the result.
 */
#ifndef PHOTIC_DEVICE_MODEL_CUH_
#define PHOTIC_DEVICE_MODEL_CUH_

#include <stdint.h>

#include "../../include/photic_b200.h"

namespace phb {

constexpr int kMaxS = PHB_MAX_SCENES;
constexpr int kMaxB = PHB_MAX_BANDS;
constexpr int kMaxSB = kMaxS * kMaxB;
constexpr int kMaxK = PHB_MAX_BOTTOMS;
constexpr int kMaxR = (2 * PHB_MAX_SPATIAL - 1) * (2 * PHB_MAX_SPATIAL - 1);
constexpr int kRecHead = 16; /* leading doubles of a debug record (oracle/ref_harness.c) */

struct ModelConst {
  int n_scenes, SB, n_bottoms, n_spatial, n_smooth, max_bands;
  int nrows, ncols, prior_present;
  float nodata, prior_nodata;
  float nodata_sb[kMaxSB]; /* geogrid.nodata_value of each (scene,band) grid: every band is tested against its OWN grid's
                              value (samodel.c:683, 941, 2999-3003); all equal to `nodata` unless the caller says otherwise */
  int n_bands[kMaxS];
  int sb_begin[kMaxS + 1]; /* first flattened (scene,band) index of scene s */
  int s_of[kMaxSB];        /* scene of a flattened index */
  double a0[kMaxSB], a1[kMaxSB], aw[kMaxSB], bbw[kMaxSB];
  double agexp[kMaxSB];    /* exp(-0.015*(lambda-440))   samodel.c:2891 */
  double ratio440[kMaxSB]; /* 440/lambda                 samodel.c:2903 */
  double bottom[kMaxK][kMaxSB];
  double r_sigma[kMaxSB];  /* scene.R_sigma per (scene,band): depth-error trials only, samodel.c:3005 */
  double sec_view[kMaxS], sec_sun[kMaxS];
  double aw640;
  /* interp_1d (common.c:298) of the measured spectrum at 440/490/550/640 nm: bracket resolved
   * on the host (it depends on the wavelengths only). exact >= 0: return Y[exact]. */
  int ib0[kMaxS][4], ib1[kMaxS][4], iexact[kMaxS][4];
  double ialpha[kMaxS][4], ioma[kMaxS][4]; /* alpha and (1.0 - alpha) */
};

/* Algorithmic FLOPs of one objective evaluation: the reference's literal operation count
 * (SURVEY.md 8d): R = 55 + 5 Nb per samodel_Rrs. Uniform n_bands assumed (max_bands). */
__host__ __device__ inline double flops_eval(int Nr, int Ns, int Nb, int nbands) {
  double R = 55.0 + 5.0 * Nb;
  return (double)Nr * Ns * ((3.0 + 3.0 * Nb) + nbands * (R + 4.0)) + 10.0 * Nr + (double)Nb * Nr * (2.0 * Nb + 11.0) +
         (double)Ns * (nbands + 12.0) + 20.0;
}
__host__ __device__ inline double flops_iter(int n) { return (double)n * n + 9.0 * n; }

}  // namespace phb

#endif
