/*
 * aux_kernels.cuh -- known-answer kernels, REFINE (model/refine.c) and the FP64 peak probe.
 */
#ifndef PHOTIC_AUX_KERNELS_CUH_
#define PHOTIC_AUX_KERNELS_CUH_

#include "invert_kernel.cuh"

namespace phb {

/* ---- known answers: objective on caller-supplied parameter vectors (one warp) ------------------ */
template <int SBP>
__global__ void kat_objective_kernel(const SolveParams p, int nb_active, int n_regions, int origin,
                                     const double *meas, int nvec, const double *params, double *out6, int generic) {
  const ModelConst &M = *p.M;
  stage_cta(p, M, phb_smem);
  __syncthreads();
  Warp w;
  const int lane = threadIdx.x & 31, SB = p.L.SB, Ns = p.L.Ns;
  bind_warp<0, 0>(w, p, p.L, phb_smem, 0, 0);
  Pixel px;
  px.Nr = n_regions; px.Nb = nb_active; px.origin = origin;
  size_pixel(px, lane, SB, Ns, p.L.simplex_doubles, 0);
  for (int t = lane; t < px.T; t += 32) w.meas[t] = meas[t];
  __syncwarp();
  phm::Tables tb;
  tb.exp_tab = w.exp_tab;
  tb.log_tab = p.log_tab;
  tb.pow_tab = p.pow_tab;
  double Bs, Ps, Xs;
  derive_pixel_constants(w, px, M, tb, lane, SB, Ns, Bs, Ps, Xs);
  Side side;
  for (int v = 0; v < nvec; v++) {
    for (int i = lane; i < px.n; i += 32) w.xmin[i] = params[(size_t)v * px.n + i];
    __syncwarp();
    /* the instantiation the solve kernel uses for this substrate count: the compile-time classes for 3 and 1 */
    double e;
    if (!generic && nb_active == 3 && n_regions <= 16) e = objective<3, SBP, true>(w, px, lane, SB, Ns, p.L.NbMax, w.xmin, side);
    else if (!generic && nb_active == 1 && n_regions <= 16) e = objective<1, SBP, true>(w, px, lane, SB, Ns, p.L.NbMax, w.xmin, side);
    else e = objective<0, SBP, true>(w, px, lane, SB, Ns, p.L.NbMax, w.xmin, side);
    if (lane == 0) {
      double *o = out6 + (size_t)v * 6;
      o[0] = e; o[1] = side.e_rrs; o[2] = side.e_depth; o[3] = side.e_bottom; o[4] = side.e_K; o[5] = side.bottom_albedo;
    }
    __syncwarp();
  }
}

/* ---- mapping study: the objective by a TEAM of warps, and a closed-loop evaluation benchmark -------------------
 * north_star asks that the pixel-to-thread mapping be chosen on ncu evidence. The product maps one pixel to one warp
 * (objective()). The alternative -- a pixel on TW co-operating warps ("CTA per pixel" for TW = 16) -- is built here for
 * the part of the work that can be spread at all, the forward-model terms: the (scene,band) and (region,substrate)
 * pre-passes go to different warps, the Nr * SB terms are dealt over 32 * TW lanes, and then ONE warp adds the squared
 * residuals in the reference's order (a sum that cannot be split without changing its bits) and runs the penalties,
 * while the others wait at the team's named barrier. Same operations on the same operands as objective(): the results
 * are bit-identical (tests/test_gpu_parity.py::test_team_objective_equals_warp_objective). eval_bench_kernel runs either
 * form in a closed loop (the next vector depends on the last value, as in the simplex), 16 warps per SM, so the two
 * mappings can be timed and profiled on equal terms (tests/manual/mapping_study.py). */
__device__ __forceinline__ void team_sync(int bar_id, int n_threads) {
#ifndef PHB_HOST_EMU
  asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(n_threads) : "memory");
#endif
}

template <int NB, int SBP, int TW>
__device__ __forceinline__ double objective_team(const Warp &w, const Pixel &px, int lane, int wt, int bar_id, int SB, int Ns,
                                                 int NbMaxRt, const double *__restrict__ x, Side &side, double *xchg) {
  const int Nr = px.Nr, T = px.T, off = px.off;
  const int Nb = NB > 0 ? NB : px.Nb;
  const int NbS = NB > 0 ? NB : NbMaxRt;
  constexpr CtaOff CO = cta_offsets(SBP);
  constexpr WarpOff WO = warp_offsets(SBP);
  const uint64_t *const c_exp = reinterpret_cast<const uint64_t *>(phb_smem + CO.exp);
  const double *const c_bbw = reinterpret_cast<const double *>(phb_smem + CO.bbw);
  const double *const c_secs = reinterpret_cast<const double *>(phb_smem + CO.secs);
  const double *const c_secv = reinterpret_cast<const double *>(phb_smem + CO.secv);
  const double *const c_a0 = reinterpret_cast<const double *>(phb_smem + CO.a0);
  const double *const c_a1 = reinterpret_cast<const double *>(phb_smem + CO.a1);
  const double *const c_aw = reinterpret_cast<const double *>(phb_smem + CO.aw);
  const double *const c_agexp = reinterpret_cast<const double *>(phb_smem + CO.agexp);
  const double *const c_bot = reinterpret_cast<const double *>(phb_smem + CO.bot);
  const int *const c_sof = reinterpret_cast<const int *>(phb_smem + CO.sof);
  unsigned char *const wblk = phb_smem + w.wofs;
  double *const a_sb = reinterpret_cast<double *>(wblk + WO.a);
  double *const X_sb = reinterpret_cast<double *>(wblk + WO.X);
  double *const K_sb = reinterpret_cast<double *>(wblk + WO.K);
  double *const qB = reinterpret_cast<double *>(wblk + WO.qB);
  constexpr int TL = 32 * TW; /* lanes of the team */

  /* pre-passes (objective(): samodel.c:2889-2893 and 2482-2496), one warp each */
  if (wt == 0) {
#pragma unroll 1
    for (int sb = lane; sb < SB; sb += 32) {
      const int s = c_sof[sb];
      const double P = 0.01 * fabs(x[off + 3 * s]);
      const double G = 0.01 * fabs(x[off + 1 + 3 * s]);
      const double a_phi = (c_a0[sb] + c_a1[sb] * phm::log(P, w.log_tab)) * P;
      const double a_g = G * c_agexp[sb];
      a_sb[sb] = c_aw[sb] + a_phi + a_g;
      X_sb[sb] = 0.01 * fabs(x[off + 2 + 3 * s]);
    }
  }
  if (wt == (TW > 1 ? 1 : 0)) {
#pragma unroll 1
    for (int idx = lane; idx < Nr * NbS; idx += 32) {
      const int r = idx / NbS, k = idx - r * NbS;
      double qb = 0.0;
      if (k < Nb) {
        const double *xq = x + Nr + Nr * Nb + r * Nb;
        double q_sum = fabs(xq[0]);
#pragma unroll 1
        for (int kk = 1; kk < Nb; kk++) q_sum += fabs(xq[kk]);
        const double xb = fabs(x[Nr + r * Nb + k]), q = fabs(xq[k]);
        const double xbq = xb * q;
        if (in_fast_range(q_sum) && in_fast_range(q) && in_fast_range(xbq)) {
          const double rs = rcp_refined(q_sum);
          const double q1 = __dmul_rn(q, rs), q2 = __dmul_rn(xbq, rs);
          qb = __fma_rn(rs, __fma_rn(-q_sum, q1, q), q1) * (0.01 * xb);
          w.bq[r * Nb + k] = __fma_rn(rs, __fma_rn(-q_sum, q2, xbq), q2);
        } else {
          qb = div_rn(q, q_sum) * (0.01 * xb);
          w.bq[r * Nb + k] = div_rn(xbq, q_sum);
        }
      }
      qB[idx] = qb;
    }
  }
  team_sync(bar_id, TL);

  /* forward-model terms (samodel.c:2911-2944), term t on team lane t mod TL; no sum here */
  const int Tpad = (T + 31) & ~31;
  {
    const int tl = wt * 32 + lane;
    int r = tl / SB, sb = tl - r * SB;
    const int step_r = TL / SB, step_sb = TL - step_r * SB;
#pragma unroll 1
    for (int t0 = wt * 32; t0 < T; t0 += TL) { /* warp-uniform: a round whose 32 terms are all past the last is skipped */
      const int t = t0 + lane;
      const bool live = t < T;
      const double H = fabs(x[r]);
      const double *qb = qB + r * NbS;
      double rho = qb[0] * c_bot[sb];
      if (NB > 0) {
#pragma unroll
        for (int kb = 1; kb < NB; kb++) rho += qb[kb] * c_bot[kb * SBP + sb];
      } else {
#pragma unroll 1
        for (int kb = 1; kb < NbS; kb++) rho += qb[kb] * c_bot[kb * SBP + sb];
      }
      const double a = a_sb[sb];
      const double bb = c_bbw[sb] + X_sb[sb] * w.powY[t];
      const double secs = c_secs[sb], secv = c_secv[sb];
      const double apb = a + bb;
      bool ok = in_fast_range(bb) && in_fast_range(apb);
      const double u = fast_div(bb, apb);
      double K = apb;
      if (K < 0.0) K = 0.0;
      if (K > 2.5) K = 2.5;
      const double rrs_dp = (kHot[H_084] + kHot[H_170] * u) * u;
      const double DuC = kHot[H_103] * fast_sqrt(1.0 + kHot[H_24] * u);
      const double DuB = kHot[H_104] * fast_sqrt(1.0 + kHot[H_54] * u);
      ok = ok && unit_range(u);
      const double M1 = secs + DuC * secv;
      const double x1 = -M1 * K * H;
      const double M2 = secs + DuB * secv;
      const double x2 = -M2 * K * H;
      ok = ok && exp_arg_in_main_range(x1) && exp_arg_in_main_range(x2);
      const double rrs_C = rrs_dp * (1.0 - exp_main_c(x1, c_exp));
      ok = ok && in_fast_range(rho);
      const double rrs_B = div_by_pi(rho) * exp_main_c(x2, c_exp);
      const double rrs = rrs_C + rrs_B;
      const double num = 0.5 * rrs, den = 1.0 - 1.5 * rrs;
      ok = ok && in_fast_range(num) && in_fast_range(den);
      double Rrs = fast_div(num, den);
      if (!ok && live) Rrs = term_reference<false>(H, rho, a, bb, secs, secv, c_exp);
      const double d = Rrs - w.meas[t];
      w.d2[kD2Zeros + t] = live ? d * d : 0.0; /* t < Tpad: the tables and d2 are padded to whole rounds of 32 */
      if (r == Nr - 1) K_sb[sb] = K;
      r += step_r; sb += step_sb;
      if (sb >= SB) { sb -= SB; r += 1; }
    }
  }
  team_sync(bar_id, TL);

  /* the ordered sum and everything after it: one warp (the others wait below) */
  if (wt == 0) {
    double err = 0.0;
    const double2 *dv = reinterpret_cast<const double2 *>(w.d2 + kD2Zeros);
#pragma unroll 1
    for (int q = 0; q < Tpad; q += 8, dv += 4) {
      const double2 v0 = dv[0], v1 = dv[1], v2 = dv[2], v3 = dv[3];
      err += v0.x; err += v0.y; err += v1.x; err += v1.y; err += v2.x; err += v2.y; err += v3.x; err += v3.y;
    }
    const double e = objective_tail<NB, SBP, false>(w, px, lane, Ns, NbMaxRt, x, err, side);
    if (lane == 0) *xchg = e;
  }
  team_sync(bar_id, TL);
  return *xchg;
}

/* Closed-loop evaluation benchmark: every team (TW = 1: every warp, with the product's objective()) holds one pixel and
 * evaluates the objective `reps` times, each vector derived from the last value. out[2 * team] = sum of the values,
 * out[2 * team + 1] = the first value (checked against the known answers). */
template <int NB, int SBP, int TW>
__global__ void __launch_bounds__(kMaxThreads, 1)
eval_bench_kernel(const SolveParams p, int n_regions, int origin, const double *meas, const double *params, int reps,
                  int same_smsp, int skew_cycles, double *out) {
  const ModelConst &M = *p.M;
  stage_cta(p, M, phb_smem);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5, n_teams = W / TW;
  /* same_smsp: the warps of a team are warp, warp + n_teams, ...: the same scheduler (warp mod 4) when n_teams is a
   * multiple of 4; otherwise consecutive warps, one per scheduler */
  const int team = same_smsp ? warp % n_teams : warp / TW, wt = same_smsp ? warp / n_teams : warp % TW;
  const int bar_id = 1 + team, SB = p.L.SB, Ns = p.L.Ns;
  Warp w;
  bind_warp<0, 0>(w, p, p.L, phb_smem, team, blockIdx.x * n_teams + team);
  Pixel px;
  px.Nr = n_regions; px.Nb = NB; px.origin = origin;
  size_pixel(px, lane, SB, Ns, p.L.simplex_doubles, 0);
  px.mean_meas = 0.0;
  double *xchg = w.rcp + 7;
  if (wt == 0) {
    for (int t = lane; t < px.T; t += 32) w.meas[t] = meas[t];
    for (int t = px.T + lane; t < ((px.T + 31) & ~31); t += 32) { w.meas[t] = 0.0; w.powY[t] = 0.0; }
    __syncwarp();
    phm::Tables tb;
    tb.exp_tab = w.exp_tab; tb.log_tab = p.log_tab; tb.pow_tab = p.pow_tab;
    double Bs, Ps, Xs;
    derive_pixel_constants(w, px, M, tb, lane, SB, Ns, Bs, Ps, Xs);
    for (int i = lane; i < px.n; i += 32) w.xmin[i] = params[i];
    __syncwarp();
  }
  if (TW > 1) team_sync(bar_id, 32 * TW);
  /* skew_cycles > 0: team t starts t * skew_cycles late, so the teams of an SM -- which all run the same code at the same
   * speed -- stay out of phase for the whole run, as the warps of the solve kernel are (different pixels, different
   * evaluation counts); 0: they run in step */
  if (skew_cycles > 0) {
    const long long t0 = clock64(), wait = (long long)team * skew_cycles;
    while (clock64() - t0 < wait) {}
    if (TW > 1) team_sync(bar_id, 32 * TW); else __syncwarp();
  }
  const double x0 = params[0];
  Side side;
  double acc = 0.0, first = 0.0;
#pragma unroll 1
  for (int rep = 0; rep < reps; rep++) {
    double e;
    if (TW == 1) e = objective<NB, SBP, false>(w, px, lane, SB, Ns, p.L.NbMax, w.xmin, side);
    else e = objective_team<NB, SBP, TW>(w, px, lane, wt, bar_id, SB, Ns, p.L.NbMax, w.xmin, side, xchg);
    if (rep == 0) first = e;
    acc += e;
    /* the next vector depends on this value (as the simplex's next vertex does): depth of region 0 nudged */
    if (wt == 0 && lane == 0) w.xmin[0] = x0 * (1.0 + 1.0e-6 * (double)((rep & 7) + 1)) + 1.0e-9 * e;
    if (TW > 1) team_sync(bar_id, 32 * TW); else __syncwarp();
  }
  if (wt == 0 && lane == 0) { out[2 * (blockIdx.x * n_teams + team)] = acc; out[2 * (blockIdx.x * n_teams + team) + 1] = first; }
}

/* ---- known answers: the exact libm port --------------------------------------------------------- */
__global__ void kat_math_kernel(int fn, const double *x, const double *y, long long n, double *out,
                                const unsigned long long *exp_tab, const double *log_tab, const double *pow_tab) {
  phm::Tables tb;
  tb.exp_tab = reinterpret_cast<const uint64_t *>(exp_tab);
  tb.log_tab = log_tab;
  tb.pow_tab = pow_tab;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    double r;
    if (fn == 0) r = phm::exp(x[i], tb.exp_tab);
    else if (fn == 1) r = phm::log(x[i], tb.log_tab);
    else if (fn == 2) r = phm::pow(x[i], y[i], tb);
    else if (fn == 3) r = fast_div(x[i], y[i]);          /* branch-free fast paths ...            */
    else if (fn == 4) r = fast_sqrt(x[i]);
    else if (fn == 5) r = x[i] / y[i];                   /* ... against the ordinary IEEE operators */
    else if (fn == 6) r = sqrt(x[i]);
    else if (fn == 7) r = exp_main_c(x[i], tb.exp_tab);  /* branch-free exp (valid in its main range) */
    else if (fn == 8) r = div_by_pi(x[i]);               /* against x / pi */
    else if (fn == 9) r = in_fast_range(x[i]) ? 1.0 : 0.0;
    else if (fn == 10) r = exp_arg_in_main_range(x[i]) ? 1.0 : 0.0;
    else if (fn == 11) r = unit_range(x[i]) ? 1.0 : 0.0;
    else if (fn == 12) r = phm::log10(x[i], tb.log_tab);
    else if (fn == 13) r = div_by(x[i], y[i], rcp_refined(y[i]), in_fast_range(y[i])); /* guarded division by a fixed divisor */
    else if (fn == 14) r = sqrt_guarded(x[i]);
    else r = div_guarded(x[i], y[i]);                     /* 15 */
    out[i] = r;
  }
}

/* ---- MODEL Lee_Kd_LS8 / Lee_Secchi_LS8, model/secchi.c:13-252 ---------------------------------------
 * Lee et al. (2016) QAA-style diffuse attenuation and Secchi-disk depth from the four Landsat-8 visible bands:
 * closed form per cell, float in / float out, double arithmetic in between (secchi.c:117-252). log10 / pow /
 * exp / log are the exact host-libm ports, the file is compiled without FMA contraction, so a cell equals the
 * CPU's bit for bit. HBM-bound: 16 B in, 4 B out per cell. */
struct LeeParams {
  const float *coastal, *blue, *green, *red;
  float *out;
  long long n;
  float spv[4];
  float theta_s;
  int mode; /* 0: Kd_LS8 (secchi.c:13), 1: secchi_disk_depth (secchi.c:59) */
  const unsigned long long *exp_tab; const double *log_tab; const double *pow_tab;
};

__device__ __forceinline__ void lee_kd_bands(const float *Rrs, float theta_s, const phm::Tables &tb, float *kd) {
  const double g0 = 0.0895, g1 = 0.1247, aw = 0.05866, h0 = -1.146, h1 = -1.366, h2 = 0.469;
  const double bbw[4] = {0.00244761, 0.00171397, 0.000931339, 0.000448682};
  const double ratio[4] = {554.0 / 443.0, 554.0 / 481.0, 1.0, 554.0 / 656.0};
  const double m0 = 0.005, m1 = 4.26, m2 = 0.52, m3 = 10.8, gamma = 0.265;
  double rrs[4], u[4], a[4], bb[4], bbp[4];
#pragma unroll
  for (int b = 0; b < 4; b++) {
    rrs[b] = Rrs[b] / (0.52 + 1.7 * Rrs[b]);
    u[b] = (-g0 + sqrt(g0 * g0 + 4.0 * g1 * rrs[b])) / (2.0 * g1);
  }
  const double chi = phm::log10((rrs[0] + rrs[1]) / (rrs[2] + 5.0 * rrs[3] * rrs[3] / rrs[1]), tb.log_tab);
  a[2] = aw + phm::pow(10.0, h0 + h1 * chi + h2 * chi * chi, tb);
  bb[2] = (-a[2] * g0 + 2.0 * a[2] * rrs[2] + a[2] * sqrt(g0 * g0 + 4.0 * g1 * rrs[2])) / (2.0 * (g0 + g1 - rrs[2]));
  bbp[2] = (u[2] * a[2]) / (1 - u[2]) - bbw[2];
  const double eta = 2.0 * (1.0 - 1.2 * phm::exp(-0.9 * rrs[0] / rrs[2], tb.exp_tab));
#pragma unroll
  for (int b = 0; b < 4; b++) {
    if (b == 2) continue;
    bbp[b] = bbp[2] * phm::pow(ratio[b], eta, tb);
    a[b] = (1.0 - u[b]) * (bbw[b] + bbp[b]) / u[b];
    bb[b] = (-a[b] * g0 + 2.0 * a[b] * rrs[b] + a[b] * sqrt(g0 * g0 + 4.0 * g1 * rrs[b])) / (2.0 * (g0 + g1 - rrs[b]));
  }
#pragma unroll
  for (int b = 0; b < 4; b++) {
    const double kk1 = (1.0 + m0 * theta_s) * a[b];
    const double kk2 = m1 * (1.0 - gamma * bbw[b] / bb[b]);
    const double kk3 = (1.0 - m2 * phm::exp(-m3 * a[b], tb.exp_tab)) * bb[b];
    kd[b] = (float)(kk1 + kk2 * kk3);
  }
}

__global__ void lee_ls8_kernel(const LeeParams p) {
  phm::Tables tb;
  tb.exp_tab = reinterpret_cast<const uint64_t *>(p.exp_tab);
  tb.log_tab = p.log_tab;
  tb.pow_tab = p.pow_tab;
  for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < p.n; q += (long long)gridDim.x * blockDim.x) {
    float R[4] = {__ldg(p.coastal + q), __ldg(p.blue + q), __ldg(p.green + q), __ldg(p.red + q)};
    float res = p.spv[0];
    if (R[0] != p.spv[0] && R[1] != p.spv[1] && R[2] != p.spv[2] && R[3] != p.spv[3]) {
      float k4[4], kd[5];
      lee_kd_bands(R, p.theta_s, tb, k4);
      kd[0] = k4[0]; kd[1] = k4[1]; kd[2] = (float)(0.20 * k4[1] + 0.75 * k4[2]); kd[3] = k4[2]; kd[4] = k4[3]; /* secchi.c:45-49 */
      float kmin = kd[0];
#pragma unroll
      for (int i = 0; i < 5; i++) if (kd[i] < kmin) kmin = kd[i]; /* vec_min, common.c:755 */
      if (p.mode == 0) res = kmin;
      else {
        float mx = R[0];
#pragma unroll
        for (int i = 0; i < 4; i++) if (R[i] > mx) mx = R[i]; /* vec_max, common.c:797 */
        res = (float)(phm::log(fabs(0.14 - mx) / 0.013, tb.log_tab) / (2.5 * kmin)); /* secchi.c:113 */
      }
    }
    p.out[q] = res;
  }
}

/* ---- REFINE, model/refine.c:215-301 --------------------------------------------------------------
 * Point-wise; every variable of the reference is a float, pow()/fabs() promote to double and the
 * result is narrowed back. pow here is the exact host-libm port, so results equal the CPU's. */

__host__ __device__ inline unsigned int float_to_ordered(float f) {
  unsigned int u;
#if defined(__CUDA_ARCH__)
  u = __float_as_uint(f);
#else
  memcpy(&u, &f, 4);
#endif
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ inline float ordered_to_float(unsigned int o) {
  unsigned int u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
#if defined(__CUDA_ARCH__)
  return __uint_as_float(u);
#else
  float f; memcpy(&f, &u, 4); return f;
#endif
}

/* array_min2 / array_max2, common.c:1224-1262 (nodata skipped with approx_equal 1e-4) */
__global__ void refine_minmax_kernel(const float *in, long long n, float nodata, unsigned int *mm) {
  unsigned int lo = 0xffffffffu, hi = 0u;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = __ldg(in + i);
    if (approx_equal_f(v, nodata, 1.0e-4f)) continue;
    if (v != v) continue; /* NaN never wins a < or > comparison */
    const unsigned int o = float_to_ordered(v);
    lo = o < lo ? o : lo;
    hi = o > hi ? o : hi;
  }
  lo = __reduce_min_sync(kFull, lo);
  hi = __reduce_max_sync(kFull, hi);
  if ((threadIdx.x & 31) == 0) { atomicMin(mm + 0, lo); atomicMax(mm + 1, hi); }
}

/* ---- int16 scale/offset packing of a grid for NetCDF (row N4): compress_2d / decompress_2d, model/nc.c:247-320 ----
 * The reference finds the range in ONE sequential pass whose maximum test sits in the `else` of the minimum test and
 * starts at FLT_MIN (nc.c:286-298): a value that lowers the running minimum when it is met never raises the maximum,
 * so the range depends on the ORDER of the cells. Reproduced exactly in three passes over the row-major cells cut into
 * chunks of kPackChunk: (1) the minimum of every chunk, (2) an exclusive prefix-minimum over the chunks (one block),
 * (3) every chunk again, each thread walking 16 consecutive cells from the running minimum that reaches it (exclusive
 * prefix over the threads of the block, seeded with the chunk's carry) and keeping the largest value that did not lower
 * it; then the packing pass. All four are bandwidth-bound: 4 + 4 + 4 B read and 2 B written per cell. */
constexpr int kPackThreads = 256, kPackPer = 16, kPackChunk = kPackThreads * kPackPer;
__device__ __forceinline__ float run_min(float a, float v) { return v < a ? v : a; } /* `if (v < grmin) grmin = v`; NaN never lowers it */

__global__ void nc_chunk_min_kernel(const float *in, long long n, float fspv, float *chunk_min, int n_chunks) {
  __shared__ float wm[kPackThreads / 32];
  for (int ch = blockIdx.x; ch < n_chunks; ch += gridDim.x) {
    const long long base = (long long)ch * kPackChunk;
    float m = 3.402823466e+38f; /* FLT_MAX */
#pragma unroll 4
    for (int q = 0; q < kPackPer; q++) {
      const long long i = base + q * kPackThreads + threadIdx.x;
      if (i < n) { const float v = __ldg(in + i); if (!(v == fspv)) m = run_min(m, v); }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = run_min(m, __shfl_xor_sync(kFull, m, o));
    if ((threadIdx.x & 31) == 0) wm[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = wm[0];
      for (int k = 1; k < kPackThreads / 32; k++) t = run_min(t, wm[k]);
      chunk_min[ch] = t;
    }
    __syncthreads();
  }
}

/* in place: chunk_min[k] <- min(FLT_MAX, chunk_min[0..k-1]); total[0] <- the minimum of everything. One block. */
__global__ void nc_chunk_scan_kernel(float *chunk_min, int n_chunks, float *total) {
  __shared__ float part[1024];
  const int T = blockDim.x, per = (n_chunks + T - 1) / T;
  const int a = threadIdx.x * per, b = a + per < n_chunks ? a + per : n_chunks;
  float m = 3.402823466e+38f;
  for (int k = a; k < b; k++) m = run_min(m, chunk_min[k]);
  part[threadIdx.x] = m;
  __syncthreads();
  if (threadIdx.x == 0) { /* exclusive prefix over <= 1024 partials */
    float run = 3.402823466e+38f;
    for (int k = 0; k < T; k++) { const float v = part[k]; part[k] = run; run = run_min(run, v); }
    total[0] = run;
  }
  __syncthreads();
  float run = part[threadIdx.x];
  for (int k = a; k < b; k++) { const float v = chunk_min[k]; chunk_min[k] = run; run = run_min(run, v); }
}

__global__ void nc_chunk_max_kernel(const float *in, long long n, float fspv, const float *carry, int n_chunks,
                                    unsigned int *grmax_ordered) {
  __shared__ float cell[kPackChunk];
  __shared__ float wpre[kPackThreads / 32];
  float best = 1.175494351e-38f; /* FLT_MIN, nc.c:287 */
  for (int ch = blockIdx.x; ch < n_chunks; ch += gridDim.x) {
    const long long base = (long long)ch * kPackChunk;
    for (int q = 0; q < kPackPer; q++) { /* coalesced load; cells past the end behave like the missing value */
      const long long i = base + q * kPackThreads + threadIdx.x;
      cell[q * kPackThreads + threadIdx.x] = i < n ? __ldg(in + i) : fspv;
    }
    __syncthreads();
    const float *mine = cell + threadIdx.x * kPackPer; /* 16 consecutive cells */
    float m = 3.402823466e+38f;
#pragma unroll
    for (int q = 0; q < kPackPer; q++) { const float v = mine[q]; if (!(v == fspv)) m = run_min(m, v); }
    /* exclusive prefix minimum over the threads of the block, in thread order */
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float inc = m;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const float t = __shfl_up_sync(kFull, inc, o); if (lane >= o) inc = run_min(inc, t); }
    if (lane == 31) wpre[wid] = inc;
    __syncthreads();
    float run = carry[ch];
    for (int k = 0; k < wid; k++) run = run_min(run, wpre[k]);
    const float before = __shfl_up_sync(kFull, inc, 1);
    if (lane > 0) run = run_min(run, before);
    /* the reference's loop body on this thread's cells (nc.c:291-297) */
#pragma unroll
    for (int q = 0; q < kPackPer; q++) {
      const float v = mine[q];
      if (v == fspv) continue;
      if (v < run) run = v;
      else if (v > best) best = v;
    }
    __syncthreads();
  }
  unsigned int o = float_to_ordered(best);
  o = __reduce_max_sync(kFull, o);
  if ((threadIdx.x & 31) == 0) atomicMax(grmax_ordered, o);
}

/* (short) rint(.) as x86-64 does it: cvttsd2si to 32 bits (INT_MIN for NaN / out of range), low 16 bits */
__device__ __forceinline__ short to_short_x86(double v) {
  int w;
  if (!(fabs(v) < 2147483648.0)) w = (int)0x80000000u;
  else w = __double2int_rz(v);
  return (short)(unsigned short)((unsigned)w & 0xffffu);
}

__global__ void nc_pack_kernel(const float *in, long long n, float fspv, float grmin, float scale, short *out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = __ldg(in + i);
    short r;
    if (v == fspv) r = (short)-32768; /* SHRT_MIN, nc.c:282,314 */
    else { const float q = __fdiv_rn(__fsub_rn(v, grmin), scale); r = to_short_x86(rint((double)q)); } /* nc.c:316 */
    out[i] = r;
  }
}

__global__ void nc_unpack_kernel(const short *in, long long n, float add_offset, float scale_factor, short missing,
                                 float fspv, float *out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const short c = in[i];
    float r = fspv;
    if (c != missing) { /* nc.c:259-262 */
      /* NaN results carry x86's bits: an invalid operation (0 * inf, inf - inf: a grid that held infinities has an
       * infinite range) gives the "real indefinite" 0xffc00000, a NaN operand is passed on quieted -- the device's
       * own NaN is 0x7fffffff */
      float t = __fmul_rn((float)c, scale_factor);
      if (t != t) t = (scale_factor != scale_factor) ? __uint_as_float(__float_as_uint(scale_factor) | 0x00400000u) : __uint_as_float(0xffc00000u);
      r = __fadd_rn(t, add_offset);
      if (r != r) r = (t != t) ? t : (add_offset != add_offset) ? __uint_as_float(__float_as_uint(add_offset) | 0x00400000u) : __uint_as_float(0xffc00000u);
    }
    out[i] = r;
  }
}

struct RefineParams {
  const float *in, *land, *shallow;
  float *out;
  long long n;
  float nodata, land_nodata, shallow_nodata;
  int flags;
  float oldmin, oldmax, dmin, dmax, scale, linear_m, linear_c, scrapmin, scrapmax, power_a, power_b;
  float sca, scb;
  const unsigned long long *exp_tab; const double *log_tab; const double *pow_tab;
};

/* host part of refine.c:215-238: the rescale coefficients (uses the host libm pow, as the reference) */
inline RefineParams make_refine_params(int flags, const float *args, const float *minmax) {
  RefineParams r;
  memset(&r, 0, sizeof(r));
  r.flags = flags;
  r.oldmin = args[0]; r.oldmax = args[1]; r.dmin = args[2]; r.dmax = args[3]; r.scale = args[4];
  r.linear_m = args[5]; r.linear_c = args[6]; r.scrapmin = args[7]; r.scrapmax = args[8];
  r.power_a = args[9]; r.power_b = args[10];
  if (!(flags & PHB_REFINE_CLIP)) { r.oldmin = minmax[0]; r.oldmax = minmax[1]; }
  if (flags & PHB_REFINE_SCALE) {
    float smin = (float)pow(fabs((double)r.oldmin), (double)r.scale);
    if (r.oldmin < 0.0) smin = (float)((double)smin * -1.0);
    float smax = (float)pow(fabs((double)r.oldmax), (double)r.scale);
    if (r.oldmax < 0.0) smax = (float)((double)smax * -1.0);
    volatile float num = r.oldmax - r.oldmin, den = smax - smin;
    r.sca = num / den;
    volatile float t0 = r.oldmax * smin, t1 = r.oldmin * smax;
    volatile float num2 = t0 - t1, den2 = smin - smax;
    r.scb = num2 / den2;
  }
  return r;
}

__global__ void refine_kernel(const RefineParams p) {
  phm::Tables tb;
  tb.exp_tab = reinterpret_cast<const uint64_t *>(p.exp_tab);
  tb.log_tab = p.log_tab;
  tb.pow_tab = p.pow_tab;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < p.n; i += (long long)gridDim.x * blockDim.x) {
    const bool open = (p.land == nullptr && p.shallow == nullptr) ||
                      ((p.land != nullptr && p.land[i] != p.land_nodata) &&
                       (p.shallow != nullptr && p.shallow[i] != p.shallow_nodata));
    if (!open) { p.out[i] = p.nodata; continue; }
    float depth = p.in[i];
    if (depth == p.nodata) { p.out[i] = p.nodata; continue; }
    if (p.flags & PHB_REFINE_CLIP) {
      if (depth < p.oldmin) depth = p.oldmin;
      else if (depth > p.oldmax) depth = p.oldmax;
    }
    if (p.flags & PHB_REFINE_LINEAR) depth = __fadd_rn(__fmul_rn(p.linear_m, depth), p.linear_c);
    if (p.flags & PHB_REFINE_SCALE) {
      if (p.scale != 1.0) {
        float v = __fsub_rn(__fmul_rn(p.dmin, p.oldmax), __fmul_rn(p.dmax, p.oldmin));
        v = __fadd_rn(v, __fmul_rn(p.dmax, depth));
        v = __fsub_rn(v, __fmul_rn(p.dmin, depth));
        v = __fdiv_rn(v, __fsub_rn(p.oldmax, p.oldmin));
        depth = (float)__dadd_rn(__dmul_rn((double)p.sca, phm::pow(fabs((double)v), (double)p.scale, tb)), (double)p.scb);
        if (v < 0.0) depth = (float)((double)depth * -1.0);
      } else {
        const float den = __fsub_rn(p.oldmax, p.oldmin);
        const float alpha = __fdiv_rn(__fsub_rn(__fmul_rn(p.oldmax, p.dmin), __fmul_rn(p.oldmin, p.dmax)), den);
        const float beta = __fdiv_rn(__fsub_rn(p.dmax, p.dmin), den);
        depth = __fadd_rn(alpha, __fmul_rn(beta, depth));
      }
    }
    if (p.flags & PHB_REFINE_SCRAP) {
      if (depth < p.scrapmin || depth > p.scrapmax) { p.out[i] = p.nodata; continue; }
    }
    if (p.flags & PHB_REFINE_POWER) {
      const double pw = phm::pow(fabs((double)depth), (double)p.power_b, tb);
      if (depth < 0.0) depth = (float)__dmul_rn(__dmul_rn(-1.0, (double)p.power_a), pw);
      else depth = (float)__dmul_rn((double)p.power_a, pw);
    }
    p.out[i] = depth;
  }
}

/* ---- FP64 pipe peak: kPeakChains independent DFMA chains per thread ------------------------------ */
constexpr int kPeakChains = 8;
__global__ void dfma_peak_kernel(double *sink, int iters) {
  double a[kPeakChains];
  const double b = 1.0000001, c = 1.0e-9 * (threadIdx.x + 1);
#pragma unroll
  for (int k = 0; k < kPeakChains; k++) a[k] = 1.0 + 0.001 * k + 1e-6 * threadIdx.x;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int k = 0; k < kPeakChains; k++) a[k] = __fma_rn(a[k], b, c);
  }
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < kPeakChains; k++) s += a[k];
  if (s == 123.456) sink[threadIdx.x & 1023] = s; /* keep the chains alive */
}

}  // namespace phb

#endif
