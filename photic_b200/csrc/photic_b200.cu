/*
 * photic_b200.cu -- host side of libphotic_b200.so: the C ABI of include/photic_b200.h.
 *
 * Builds the scene-level constants the reference derives in model/samodel.c:505-618 with the HOST
 * libm (so they are the reference's own bits), stages inputs, runs classify -> solve on the GPU,
 * and reads the counters back. No CPU fallback: without a CUDA device every compute entry fails.
 *
 * Compile: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false  (see build.py)
 */
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <unistd.h>

#include <atomic>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/photic_b200.h"
#include "../../include/photic_spectra.h"
#include "device_model.cuh"
#include "invert_kernel.cuh"
#include "aux_kernels.cuh"
#include "jerlov_host.h"

using namespace phb;

namespace {

thread_local std::string g_last_cuda_error;

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) {                                                                       \
      g_last_cuda_error = std::string(#call) + ": " + cudaGetErrorString(e_);                      \
      return PHB_ECUDA;                                                                            \
    }                                                                                              \
  } while (0)

/* interp_1d, common.c:298-333, including its float-typed bracketing tests. When `bracket` is given
 * the resolved bracket is returned instead of the value (used for the per-pixel spectra). */
struct Bracket { int i0, i1, exact; double alpha, oma; };

bool approx_le(float a, float b, float e) { return a < b || approx_equal_f(a, b, e); }
bool approx_ge(float a, float b, float e) { return a > b || approx_equal_f(a, b, e); }

Bracket resolve_bracket(const double *X, int n, double x) {
  Bracket B = {0, n > 1 ? 1 : 0, -1, 0.0, 1.0};
  if (approx_equal_f((float)x, (float)X[0], 1.0e-5f)) { B.exact = 0; return B; }
  if (approx_equal_f((float)x, (float)X[n - 1], 1.0e-5f)) { B.exact = n - 1; return B; }
  if (X[0] < X[n - 1] && x < X[0]) { B.i0 = 0; B.i1 = 1; }
  else if (X[n - 1] > X[0] && x > X[n - 1]) { B.i0 = n - 2; B.i1 = n - 1; }
  else {
    for (int i = 0; i < n - 1; i++) {
      float lo = (float)X[i], hi = (float)X[i + 1], xf = (float)x;
      if ((approx_le(lo, xf, 1.0e-5f) && approx_ge(hi, xf, 1.0e-5f)) ||
          (approx_ge(lo, xf, 1.0e-5f) && approx_le(hi, xf, 1.0e-5f))) { B.i0 = i; B.i1 = i + 1; break; }
    }
  }
  volatile double alpha = (x - X[B.i0]) / (X[B.i1] - X[B.i0]);
  volatile double oma = 1.0 - alpha;
  B.alpha = alpha; B.oma = oma;
  return B;
}

double host_interp_1d(const double *X, const double *Y, int n, double x) {
  Bracket B = resolve_bracket(X, n, x);
  if (B.exact >= 0) return Y[B.exact];
  volatile double t0 = Y[B.i0] * B.oma, t1 = Y[B.i1] * B.alpha; /* no contraction */
  return t0 + t1;
}

int validate(const phb_scene_desc *d) {
  if (!d) return PHB_EINVAL;
  if (d->n_scenes < 1 || d->n_scenes > PHB_MAX_SCENES) return PHB_EINVAL;
  for (int s = 0; s < d->n_scenes; s++)
    if (d->n_bands[s] < 2 || d->n_bands[s] > PHB_MAX_BANDS) return PHB_EINVAL;
  if (d->n_bottoms < 1 || d->n_bottoms > PHB_MAX_BOTTOMS) return PHB_EINVAL;
  if (d->n_spatial < 0 || d->n_spatial > PHB_MAX_SPATIAL) return PHB_EINVAL;
  if (d->n_smoothing_radius < 1 || d->n_smoothing_radius > 8) return PHB_EINVAL;
  if (d->nrows < 1 || d->ncols < 1 || (long long)d->nrows * d->ncols > 2147483647LL) return PHB_EINVAL;
  return PHB_OK;
}

/* samodel.c:505-618 */
void build_model(const phb_scene_desc *d, ModelConst *M) {
  memset(M, 0, sizeof(*M));
  double grid[PH_SPEC_N];
  for (int i = 0; i < PH_SPEC_N; i++) grid[i] = PH_SPEC_LAMBDA0 + PH_SPEC_DLAMBDA * (double)i;
  M->n_scenes = d->n_scenes; M->n_bottoms = d->n_bottoms; M->n_spatial = d->n_spatial;
  M->n_smooth = d->n_smoothing_radius;
  M->nrows = d->nrows; M->ncols = d->ncols; M->nodata = d->nodata;
  M->prior_present = d->prior_present; M->prior_nodata = d->prior_nodata;
  M->aw640 = host_interp_1d(grid, PH_SPEC_AW, PH_SPEC_N, 640.0);
  const double PI_ = 3.141592653589793;
  const double targets[4] = {440.0, 490.0, 550.0, 640.0};
  int sb = 0;
  for (int s = 0; s < d->n_scenes; s++) {
    M->n_bands[s] = d->n_bands[s];
    if (d->n_bands[s] > M->max_bands) M->max_bands = d->n_bands[s];
    M->sb_begin[s] = sb;
    volatile double tv = d->theta_view[s], tw = d->theta_sun[s];
    tv = tv * (PI_ / 180.0); /* samodel.c:538 */
    tw = tw * (PI_ / 180.0);
    M->sec_view[s] = 1.0 / cos(tv);
    M->sec_sun[s] = 1.0 / cos(tw);
    double lam[PHB_MAX_BANDS];
    for (int b = 0; b < d->n_bands[s]; b++, sb++) {
      const double w = (double)d->wavelengths[s][b];
      lam[b] = w;
      M->s_of[sb] = s;
      M->a0[sb] = host_interp_1d(grid, PH_SPEC_A0, PH_SPEC_N, w);
      M->a1[sb] = host_interp_1d(grid, PH_SPEC_A1, PH_SPEC_N, w);
      M->bbw[sb] = host_interp_1d(grid, PH_SPEC_BBW, PH_SPEC_N, w);
      M->aw[sb] = host_interp_1d(grid, PH_SPEC_AW, PH_SPEC_N, w);
      for (int k = 0; k < d->n_bottoms; k++) M->bottom[k][sb] = host_interp_1d(grid, PH_SPEC_BOTTOM[k], PH_SPEC_N, w);
      volatile double arg = -0.015 * (w - 440.0); /* -S*(lambda - 440.0), samodel.c:2891 */
      M->agexp[sb] = exp(arg);
      M->ratio440[sb] = 440.0 / w;
      M->r_sigma[sb] = d->r_sigma[s][b];
      M->nodata_sb[sb] = d->nodata_per_band ? d->nodata_band[s][b] : d->nodata;
    }
    for (int g = 0; g < 4; g++) {
      Bracket B = resolve_bracket(lam, d->n_bands[s], targets[g]);
      M->ib0[s][g] = B.i0; M->ib1[s][g] = B.i1; M->iexact[s][g] = B.exact;
      M->ialpha[s][g] = B.alpha; M->ioma[s][g] = B.oma;
    }
  }
  M->sb_begin[d->n_scenes] = sb;
  M->SB = sb;
}

template <class T>
struct DevBuf {
  T *p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    cudaError_t e = cudaMalloc(&p, n * sizeof(T));
    if (e == cudaSuccess) cap = n;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

constexpr int kMaxViews = 64;          /* bands a solve kernel can take work from (its own + peers) */
constexpr size_t kRingBytes = 16u << 20; /* one slot of the pinned staging ring */
constexpr int kScalars = 8;            /* [0] n class 0 (all substrates) [1] n class 1 (sand only) [2] n both [3] head 0 [4] head 1 */

}  // namespace

struct phb_ctx {
  int device = 0;
  int n_sm = 0;
  size_t smem_optin = 0;
  size_t l2_bytes = 0;
  ModelConst *d_model = nullptr;
  unsigned long long *d_exp_tab = nullptr;
  double *d_log_tab = nullptr, *d_pow_tab = nullptr;
  DevBuf<int> q_shallow, q_deep, queue;  /* queues of phb_invert_device (caller-owned rasters) and the trial arrays */
  int *d_scalars = nullptr;              /* kScalars ints */
  unsigned long long *d_counters = nullptr; /* 4 counters */
  double *d_flops = nullptr;
  BandView *d_views = nullptr;           /* kMaxViews */
  DevBuf<double> slabs;
  DevBuf<float> planes, prior, outs;     /* staging of the depth-error / REFINE / Lee host entries */
  DevBuf<int> nev;
  DevBuf<double> dbg_rec;
  DevBuf<int> dbg_pix, dbg_iters;
  phb_shard *host_shard = nullptr;       /* the band of the host entry points, kept between calls */
  float *ring[2] = {nullptr, nullptr};   /* pinned staging ring (host entry points) */
  cudaEvent_t ring_ev[2] = {nullptr, nullptr};
  cudaEvent_t ev[4];
};

/* A row band resident on one device: one allocation (so one IPC handle) holding the rasters, the work queues and the
 * result planes, and the BandView that describes it in the owner's address space. */
struct phb_shard {
  phb_ctx *ctx = nullptr;
  phb_scene_desc desc;
  ModelConst M;
  int row_begin = 0, row_end = 0, SB = 0, mb = 0, Ns = 0, scene_planes = 0, two_classes = 0;
  unsigned char *base = nullptr;
  size_t bytes = 0;
  BandView view;
  int *scalars = nullptr, *q_shallow = nullptr, *q_deep = nullptr, *q_all = nullptr;
  struct Mapped { long long pid; unsigned long long base; unsigned char *ptr; bool ipc; };
  std::vector<Mapped> mapped;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  const void *resident_src = nullptr; /* the host rasters whose rows the band currently holds (host entry points) */
  /* debug records of the next solve (parity tests; own band only) */
  double *dbg_rec = nullptr; int *dbg_pix = nullptr, *dbg_iters = nullptr; long long dbg_cap = 0;
};

namespace {

/* what a handle carries: where the band lives and its view as OFFSETS into the allocation */
struct HandleData {
  uint32_t magic;
  int32_t device;
  int64_t pid;
  uint64_t base, bytes;
  int32_t ipc_ok, nrows, ncols, two_classes;
  cudaIpcMemHandle_t ipc;
  uint64_t off[8 + 15]; /* planes, prior, queue[2], n_queue[2], head[2], then the 15 members of phb_outputs */
};
static_assert(sizeof(HandleData) <= sizeof(phb_shard_handle), "handle too small");
constexpr uint32_t kHandleMagic = 0x50484253u; /* "PHBS" */
constexpr uint64_t kNullOff = ~0ull;

/* the pointer members of a BandView in a fixed order (out: the 15 members of phb_outputs in declaration order) */
void view_pointers(BandView &v, const void **slots[23]) {
  int k = 0;
  slots[k++] = (const void **)&v.planes; slots[k++] = (const void **)&v.prior;
  slots[k++] = (const void **)&v.queue[0]; slots[k++] = (const void **)&v.queue[1];
  slots[k++] = (const void **)&v.n_queue[0]; slots[k++] = (const void **)&v.n_queue[1];
  slots[k++] = (const void **)&v.head[0]; slots[k++] = (const void **)&v.head[1];
  phb_outputs &o = v.out;
  slots[k++] = (const void **)&o.depth; slots[k++] = (const void **)&o.model_error; slots[k++] = (const void **)&o.bottom_albedo;
  slots[k++] = (const void **)&o.bottom_sand; slots[k++] = (const void **)&o.bottom_seagrass; slots[k++] = (const void **)&o.bottom_coral;
  slots[k++] = (const void **)&o.K_min; slots[k++] = (const void **)&o.bottom_type; slots[k++] = (const void **)&o.index_optical_depth;
  slots[k++] = (const void **)&o.K; slots[k++] = (const void **)&o.P; slots[k++] = (const void **)&o.G; slots[k++] = (const void **)&o.X;
  slots[k++] = (const void **)&o.converged; slots[k++] = (const void **)&o.n_evals;
}

bool use_two_classes(const ModelConst &M) {
  if (M.n_bottoms != 3) return false; /* the default NBOTTOMS has the compile-time instantiations */
  const int nsp = M.n_spatial == 0 ? 1 : M.n_spatial;
  if ((2 * nsp - 1) * (2 * nsp - 1) > 16) return false; /* ... written for up to 16 regions (NSPATIAL <= 2, the default) */
  const char *e = getenv("PHB_ONE_CLASS");
  return !(e && atoi(e) != 0);
}

}  // namespace

extern "C" {

int phb_version(void) { return 200; }

const char *phb_error_string(int code) {
  switch (code) {
    case PHB_OK: return "ok";
    case PHB_EINVAL: return "invalid argument or scene descriptor";
    case PHB_ENODEVICE: return "no usable CUDA device (photic_b200 has no CPU fallback)";
    case PHB_ECUDA: return g_last_cuda_error.c_str();
    case PHB_ENOMEM: return "out of memory";
    case PHB_ENOFIT: return "no Jerlov fit: singular regression or slope / wavelength outside Jerlov's table";
    case PHB_ENOPEER: return "a peer band cannot be mapped on this device (no peer access / IPC refused)";
    default: return "unknown error";
  }
}

int phb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

int phb_band_tables(const phb_scene_desc *desc, double *out_tables, double *out_aux) {
  int rc = validate(desc);
  if (rc) return rc;
  ModelConst M;
  build_model(desc, &M);
  const int stride = 4 + PHB_MAX_BOTTOMS;
  for (int s = 0; s < M.n_scenes; s++)
    for (int b = 0; b < M.n_bands[s]; b++) {
      double *o = out_tables + ((size_t)s * PHB_MAX_BANDS + b) * stride;
      const int sb = M.sb_begin[s] + b;
      o[0] = M.a0[sb]; o[1] = M.a1[sb]; o[2] = M.aw[sb]; o[3] = M.bbw[sb];
      for (int k = 0; k < PHB_MAX_BOTTOMS; k++) o[4 + k] = k < M.n_bottoms ? M.bottom[k][sb] : 0.0;
    }
  int o = 0;
  out_aux[o++] = M.aw640;
  for (int s = 0; s < M.n_scenes; s++) out_aux[o++] = M.sec_view[s];
  for (int s = 0; s < M.n_scenes; s++) out_aux[o++] = M.sec_sun[s];
  return PHB_OK;
}

/* test hook: the scene constants exactly as the kernels receive them (ModelConst, csrc/device_model.cuh), host only */
int64_t phb_debug_model_const(const phb_scene_desc *desc, void *out, int64_t capacity) {
  if (validate(desc) != PHB_OK) return -1;
  if (out && capacity >= (int64_t)sizeof(ModelConst)) build_model(desc, static_cast<ModelConst *>(out));
  return (int64_t)sizeof(ModelConst);
}

static int ctx_init(phb_ctx *c, int device) {
  c->device = device;
  for (int i = 0; i < 4; i++) c->ev[i] = nullptr;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  c->n_sm = prop.multiProcessorCount;
  c->smem_optin = prop.sharedMemPerBlockOptin;
  c->l2_bytes = prop.l2CacheSize;
  static const unsigned long long exp_tab[2 * PHM_N] = PHM_EXP_TAB;
  static const double log_tab[2 * PHM_N] = PHM_LOG_TAB;
  static const double pow_tab[3 * PHM_N] = PHM_POWLOG_TAB;
  CK(cudaMalloc(&c->d_model, sizeof(ModelConst)));
  CK(cudaMalloc(&c->d_exp_tab, sizeof(exp_tab)));
  CK(cudaMalloc(&c->d_log_tab, sizeof(log_tab)));
  CK(cudaMalloc(&c->d_pow_tab, sizeof(pow_tab)));
  CK(cudaMemcpy(c->d_exp_tab, exp_tab, sizeof(exp_tab), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(c->d_log_tab, log_tab, sizeof(log_tab), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(c->d_pow_tab, pow_tab, sizeof(pow_tab), cudaMemcpyHostToDevice));
  CK(cudaMalloc(&c->d_scalars, kScalars * sizeof(int)));
  CK(cudaMalloc(&c->d_counters, 4 * sizeof(unsigned long long)));
  CK(cudaMalloc(&c->d_flops, sizeof(double)));
  CK(cudaMalloc(&c->d_views, kMaxViews * sizeof(BandView)));
  for (int i = 0; i < 4; i++) CK(cudaEventCreate(&c->ev[i]));
  /* kHot[H_RCP_PI]: the refined reciprocal of pi exactly as this device's division fast path builds it */
  rcp_pi_kernel<<<1, 1>>>(reinterpret_cast<double *>(c->d_counters)); /* two doubles of scratch */
  CK(cudaGetLastError());
  CK(cudaMemcpyToSymbol(kHot, c->d_counters, sizeof(double), H_RCP_PI * sizeof(double), cudaMemcpyDeviceToDevice));
  CK(cudaMemcpyToSymbol(kHot, reinterpret_cast<double *>(c->d_counters) + 1, sizeof(double), H_RCP_120 * sizeof(double),
                        cudaMemcpyDeviceToDevice));
  CK(cudaDeviceSynchronize());
  return PHB_OK;
}

int phb_ctx_create(int device, phb_ctx **out) {
  if (!out) return PHB_EINVAL;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) return PHB_ENODEVICE;
  CK(cudaSetDevice(device));
  phb_ctx *c = new phb_ctx();
  const int rc = ctx_init(c, device);
  if (rc) { phb_ctx_destroy(c); return rc; } /* nothing of a half-built context is left behind */
  *out = c;
  return PHB_OK;
}

void phb_ctx_destroy(phb_ctx *c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->host_shard) phb_shard_destroy(c->host_shard);
  cudaFree(c->d_model); cudaFree(c->d_exp_tab); cudaFree(c->d_log_tab); cudaFree(c->d_pow_tab);
  cudaFree(c->d_scalars); cudaFree(c->d_counters); cudaFree(c->d_flops); cudaFree(c->d_views);
  c->q_shallow.release(); c->q_deep.release(); c->queue.release(); c->slabs.release();
  c->planes.release(); c->prior.release(); c->outs.release(); c->nev.release();
  c->dbg_rec.release(); c->dbg_pix.release(); c->dbg_iters.release();
  for (int i = 0; i < 2; i++) { if (c->ring[i]) cudaFreeHost(c->ring[i]); if (c->ring_ev[i]) cudaEventDestroy(c->ring_ev[i]); }
  for (int i = 0; i < 4; i++) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
  delete c;
}

int phb_debug_record_len(const phb_scene_desc *d) {
  if (validate(d)) return 0;
  int mb = 0;
  for (int s = 0; s < d->n_scenes; s++) mb = d->n_bands[s] > mb ? d->n_bands[s] : mb;
  return kRecHead + d->n_scenes * mb + 3 * d->n_scenes;
}


struct LaunchGeom { int W, ctas, smem, regs; };

/* Shared-memory layout, launch geometry and launch of the persistent solve kernel (pixel queues or, trials = true,
 * chains of depth-error trials). sp arrives with the work description filled in (views, classes, trial arrays); the
 * model, layouts, slabs, counters and libm tables are bound here. */
static int launch_solve(phb_ctx *c, const ModelConst &M, SolveParams &sp, bool trials, cudaStream_t st, LaunchGeom *geom) {
  const int nsp = M.n_spatial == 0 ? 1 : M.n_spatial;
  const int NrMax = (2 * nsp - 1) * (2 * nsp - 1);
  sp.L = make_layout(M.SB, M.n_scenes, M.n_bottoms, NrMax);
  const long long simplex_doubles = (long long)(sp.L.nmax + 1) * sp.L.nmax;
  /* kernel instantiation: compile-time substrate count for the default NBOTTOMS 3 (followed, with two pixel classes,
   * by the one-substrate code for the sand-only queue), run-time loop otherwise; compile-time (scene,band) stride 32
   * (up to 8 dates x 4 bands) or the maximum */
  void (*kern)(const SolveParams) = sp.L.SBP == 32 ? solve_kernel<0, 32, false> : solve_kernel<0, kMaxSB, false>;
  if (trials || !use_two_classes(M)) sp.n_classes = 1;
  if (sp.n_classes == 2) kern = sp.L.SBP == 32 ? solve_kernel<3, 32, false> : solve_kernel<3, kMaxSB, false>; /* the default NBOTTOMS */
  /* the Landsat-8 configurations of BASELINE.json -- four bands per scene, 3 x 3 neighbourhoods, 4 / 6 / 8 dates -- have
   * instantiations whose scene count and shared-memory layout are compile-time constants (PHB_CT_LAYOUT=0: off) */
  {
    bool four = true;
    for (int s = 0; s < M.n_scenes; s++) four = four && M.n_bands[s] == 4;
    const char *e = getenv("PHB_CT_LAYOUT");
    if (sp.n_classes == 2 && sp.L.SBP == 32 && four && NrMax == 9 && !(e && atoi(e) == 0)) {
      if (M.n_scenes == 4) kern = solve_kernel<3, 32, false, 4>;
      else if (M.n_scenes == 6) kern = solve_kernel<3, 32, false, 6>;
      else if (M.n_scenes == 8) kern = solve_kernel<3, 32, false, 8>;
    }
  }
  if (trials) kern = sp.L.SBP == 32 ? solve_kernel<0, 32, true> : solve_kernel<0, kMaxSB, true>;
  cudaFuncAttributes fa0;
  CK(cudaFuncGetAttributes(&fa0, kern));
  const int W_reg = fa0.maxThreadsPerBlock / 32;
  /* CTAs per SM (experiment knob PHB_CTAS_PER_SM, default 1): k CTAs of 16 / k warps share the SM's shared memory (1 KB
   * of it is reserved per CTA), registers and tensor memory */
  int cps = 1;
  if (const char *e = getenv("PHB_CTAS_PER_SM")) { int v = atoi(e); if (v == 2 || v == 4) cps = v; }
  const long long smem_cta = ((long long)c->smem_optin + 1024) / cps - 1024;
  const int W_smem = (int)((smem_cta - sp.L.cta_bytes) / sp.L.warp_bytes);
  if (W_smem < 1) return PHB_EINVAL; /* configuration does not fit shared memory */
  /* Occupancy vs simplex residency: each warp's leftover shared memory holds the first rows of its
   * simplex (all of it for sand-only pixels); the rest lives in an L2-resident global slab. Default:
   * 16 warps (4 per scheduler) when they fit; PHB_WARPS_PER_CTA overrides (tuning / profiling). */
  int W = 16 / cps;
  if (const char *e = getenv("PHB_WARPS_PER_CTA")) { int v = atoi(e); if (v >= 1) W = v; }
  if (W > W_reg) W = W_reg;
  if (W > W_smem) W = W_smem;
  if (W < 1) W = 1;
  long long cache_bytes = (smem_cta - sp.L.cta_bytes) / W - sp.L.warp_bytes;
  if (cache_bytes > simplex_doubles * 8) cache_bytes = simplex_doubles * 8;
  if (const char *e = getenv("PHB_SIMPLEX_SMEM_BYTES")) { long long v = atoll(e); if (v >= 0 && v < cache_bytes) cache_bytes = v; }
  add_simplex_cache(sp.L, (int)cache_bytes);
  /* tensor memory (512 columns x 128 lanes per SM, idle on this path) holds the first simplex rows:
   * warps sharing a lane quarter split the CTA's columns */
  sp.tmem_alloc_cols = 512 / cps;
  sp.L.tmem_cols = (sp.tmem_alloc_cols / ((W + 3) / 4)) & ~1;
  if (const char *e = getenv("PHB_TMEM")) { if (atoi(e) == 0) sp.L.tmem_cols = 0; }
  if (sp.n_classes > 1) {
    /* the sand-only class: same CTA-shared block and the same warp stride, a smaller fixed part (n = Nr + 2 Nr + 3 Ns
     * parameters), so more of its -- much smaller -- simplex stays in shared memory */
    sp.L1 = make_layout(M.SB, M.n_scenes, 1, NrMax);
    sp.L1.cta_bytes = sp.L.cta_bytes;
    long long cache1 = (long long)sp.L.warp_bytes - sp.L1.w_simplex;
    const long long simplex1 = (long long)(sp.L1.nmax + 1) * sp.L1.nmax * 8;
    if (cache1 > simplex1) cache1 = simplex1;
    if (const char *e = getenv("PHB_SIMPLEX_SMEM_BYTES")) { long long v = atoll(e); if (v >= 0 && v < cache1) cache1 = v; }
    add_simplex_cache(sp.L1, (int)cache1);
    sp.L1.warp_bytes = sp.L.warp_bytes;
    sp.L1.tmem_cols = sp.L.tmem_cols;
  } else {
    sp.L1 = sp.L;
  }
  int ctas = c->n_sm * cps;
  if (const char *e = getenv("PHB_CTAS")) { int v = atoi(e); if (v >= 1) ctas = v; }
  const size_t smem = (size_t)sp.L.cta_bytes + (size_t)W * sp.L.warp_bytes;
  /* slab depth: the global-tier rows of the deepest simplex of either class (at least one row: nothing is ever empty) */
  int slab_rows = max_global_rows(sp.L, NrMax, M.n_scenes, M.n_bottoms);
  { const int r1 = max_global_rows(sp.L, NrMax, M.n_scenes, 1); if (r1 > slab_rows) slab_rows = r1; }
  if (sp.n_classes > 1) { const int r1 = max_global_rows(sp.L1, NrMax, M.n_scenes, 1); if (r1 > slab_rows) slab_rows = r1; }
  if (slab_rows < 1) slab_rows = 1;
  const long long slab_doubles = slab_doubles_for(sp.L, slab_rows);
  CK(c->slabs.ensure((size_t)ctas * W * slab_doubles));
  sp.M = c->d_model;
  sp.slabs = c->slabs.p; sp.slab_stride = slab_doubles; sp.slab_rows = slab_rows;
  if (const char *e = getenv("PHB_ALIGN")) sp.align_evals = atoi(e) != 0 ? 1 : 0;
  sp.counters = c->d_counters; sp.flops = c->d_flops;
  sp.exp_tab = c->d_exp_tab; sp.log_tab = c->d_log_tab; sp.pow_tab = c->d_pow_tab;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  /* (A persisting-L2 access-policy window over the slabs was tried in round 2 and removed: DRAM write-back went UP 2.4x
   * -- the carve-out takes capacity from everything else and only a random share of the window's lines is kept -- with
   * no change in throughput. What reaches DRAM is one write-back per line stored into the global tier, ~13 GB/s.) */
  kern<<<ctas, W * 32, smem, st>>>(sp);
  {
    cudaError_t le = cudaGetLastError();
    if (le != cudaSuccess) {
      char msg[256];
      snprintf(msg, sizeof(msg), "solve_kernel launch (ctas=%d W=%d smem=%zu regs=%d maxThreads=%d): %s", ctas, W, smem,
               fa0.numRegs, fa0.maxThreadsPerBlock, cudaGetErrorString(le));
      g_last_cuda_error = msg;
      return PHB_ECUDA;
    }
  }
  if (geom) { geom->W = W; geom->ctas = ctas; geom->smem = (int)smem; geom->regs = fa0.numRegs; }
  return PHB_OK;
}

/* classification pre-pass of one band: validity, output defaults, the two work lists (and their concatenation when
 * one generic kernel serves both classes); asynchronous on st. The band's ModelConst must be in c->d_model. */
static int launch_classify(phb_ctx *c, const BandView &v, int row_begin, int row_end, int *q_shallow, int *q_deep, int *q_all,
                           int *scalars, bool two_classes, cudaStream_t st) {
  CK(cudaMemsetAsync(scalars, 0, kScalars * sizeof(int), st));
  ClassifyParams cp;
  cp.M = c->d_model; cp.planes = v.planes; cp.prior = v.prior;
  cp.row_begin = row_begin; cp.row_end = row_end;
  cp.queue_shallow = q_shallow; cp.queue_deep = q_deep;
  cp.n_shallow = scalars + 0; cp.n_deep = scalars + 1;
  cp.out = v.out;
  classify_kernel<<<c->n_sm * 8, 256, 0, st>>>(cp);
  if (!two_classes)
    concat_queue_kernel<<<c->n_sm * 4, 256, 0, st>>>(q_shallow, q_deep, scalars + 0, scalars + 1, q_all, scalars + 2);
  CK(cudaGetLastError());
  return PHB_OK;
}

/* the queue members of a band's view for the one- or two-class kernels */
static void bind_queues(BandView &v, int *q_shallow, int *q_deep, int *q_all, int *scalars, bool two_classes) {
  if (two_classes) {
    v.queue[0] = q_shallow; v.n_queue[0] = scalars + 0; v.head[0] = scalars + 3;
    v.queue[1] = q_deep; v.n_queue[1] = scalars + 1; v.head[1] = scalars + 4;
  } else {
    v.queue[0] = q_all; v.n_queue[0] = scalars + 2; v.head[0] = scalars + 3;
    v.queue[1] = q_all; v.n_queue[1] = scalars + 5; v.head[1] = scalars + 4; /* scalars[5] stays 0: an empty queue */
  }
}

/* solve over a list of views (host copies; [0] = the device's own band) + counters read-back */
static int run_solve(phb_ctx *c, const ModelConst &M, const BandView *views, int n_views, bool two_classes,
                     double *d_rec, int *d_pix, int *d_iters, int reclen, long long dbg_cap, cudaStream_t st,
                     cudaEvent_t e_begin, cudaEvent_t e_end, phb_stats *stats) {
  if (n_views < 1 || n_views > kMaxViews) return PHB_EINVAL;
  CK(cudaMemcpyAsync(c->d_views, views, (size_t)n_views * sizeof(BandView), cudaMemcpyHostToDevice, st));
  CK(cudaMemsetAsync(c->d_counters, 0, 4 * sizeof(unsigned long long), st));
  CK(cudaMemsetAsync(c->d_flops, 0, sizeof(double), st));
  SolveParams sp;
  memset(&sp, 0, sizeof(sp));
  sp.views = c->d_views; sp.n_views = n_views; sp.n_classes = two_classes ? 2 : 1;
  sp.dbg_rec = d_rec; sp.dbg_pix = d_pix; sp.dbg_iters = d_iters; sp.reclen = reclen; sp.dbg_capacity = dbg_cap;
  LaunchGeom geom;
  CK(cudaEventRecord(e_begin, st));
  int rc = launch_solve(c, M, sp, false, st, &geom);
  if (rc) return rc;
  CK(cudaEventRecord(e_end, st));
  if (stats) {
    unsigned long long cnt[4];
    double fl;
    CK(cudaMemcpyAsync(cnt, c->d_counters, sizeof(cnt), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(&fl, c->d_flops, sizeof(fl), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    stats->n_valid = (int64_t)cnt[3];
    stats->n_evals = (int64_t)cnt[0]; stats->n_iters = (int64_t)cnt[1]; stats->n_converged = (int64_t)cnt[2];
    stats->alg_flops = fl;
    CK(cudaEventElapsedTime(&stats->ms_solve, e_begin, e_end));
    stats->warps_per_cta = geom.W; stats->ctas = geom.ctas; stats->smem_bytes = geom.smem;
    stats->regs = geom.regs;
  }
  return PHB_OK;
}

static int invert_device_impl(phb_ctx *c, const phb_scene_desc *desc, const float *d_planes, const float *d_prior,
                              int row_begin, int row_end, const phb_outputs *d_out, cudaStream_t st, phb_stats *stats) {
  if (!c || !d_planes || !d_out) return PHB_EINVAL;
  int rc = validate(desc);
  if (rc) return rc;
  if (row_begin < 0 || row_end > desc->nrows || row_begin >= row_end) return PHB_EINVAL;
  CK(cudaSetDevice(c->device));
  ModelConst M;
  build_model(desc, &M);
  CK(cudaMemcpyAsync(c->d_model, &M, sizeof(M), cudaMemcpyHostToDevice, st));
  const size_t npx = (size_t)(row_end - row_begin) * desc->ncols;
  const bool two = use_two_classes(M);
  CK(c->q_shallow.ensure(npx)); CK(c->q_deep.ensure(npx));
  if (!two) CK(c->queue.ensure(npx));
  BandView v;
  memset(&v, 0, sizeof(v));
  v.planes = d_planes; v.prior = desc->prior_present ? d_prior : nullptr; v.nrows = desc->nrows; v.out = *d_out;
  bind_queues(v, c->q_shallow.p, c->q_deep.p, c->queue.p, c->d_scalars, two);
  CK(cudaEventRecord(c->ev[0], st));
  rc = launch_classify(c, v, row_begin, row_end, c->q_shallow.p, c->q_deep.p, c->queue.p, c->d_scalars, two, st);
  if (rc) return rc;
  CK(cudaEventRecord(c->ev[1], st));
  phb_stats local;
  memset(&local, 0, sizeof(local));
  rc = run_solve(c, M, &v, 1, two, nullptr, nullptr, nullptr, phb_debug_record_len(desc), 0, st, c->ev[2], c->ev[3],
                 stats ? &local : nullptr);
  if (rc) return rc;
  if (stats) {
    int sc[kScalars];
    CK(cudaMemcpy(sc, c->d_scalars, sizeof(sc), cudaMemcpyDeviceToHost));
    local.n_shallow = sc[0];
    CK(cudaEventElapsedTime(&local.ms_classify, c->ev[0], c->ev[1]));
    *stats = local;
  }
  return PHB_OK;
}

int phb_invert_device(phb_ctx *ctx, const phb_scene_desc *desc, const float *d_planes, const float *d_prior,
                      int row_begin, int row_end, const phb_outputs *d_out, void *stream, phb_stats *stats) {
  return invert_device_impl(ctx, desc, d_planes, d_prior, row_begin, row_end, d_out, (cudaStream_t)stream, stats);
}

/* ---- row bands shareable between devices ------------------------------------------------------------------- */

int phb_shard_create(phb_ctx *c, const phb_scene_desc *desc, int row_begin, int row_end, int scene_planes, phb_shard **out) {
  if (!c || !out) return PHB_EINVAL;
  int rc = validate(desc);
  if (rc) return rc;
  if (row_begin < 0 || row_end > desc->nrows || row_begin > row_end) return PHB_EINVAL;
  CK(cudaSetDevice(c->device));
  phb_shard *S = new phb_shard();
  S->ctx = c; S->desc = *desc; S->row_begin = row_begin; S->row_end = row_end; S->scene_planes = scene_planes ? 1 : 0;
  build_model(desc, &S->M);
  S->SB = S->M.SB; S->mb = S->M.max_bands; S->Ns = S->M.n_scenes;
  S->two_classes = use_two_classes(S->M) ? 1 : 0;
  const size_t px = (size_t)desc->nrows * desc->ncols;
  const size_t own = (size_t)(row_end - row_begin) * desc->ncols;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t at = o; o += (bytes + 255) & ~(size_t)255; return at; };
  const size_t o_planes = take(px * S->SB * 4), o_prior = take(px * 4);
  size_t o_out[9];
  for (int k = 0; k < 9; k++) o_out[k] = take(px * 4);
  const size_t o_K = take(scene_planes ? px * 4 * S->Ns * S->mb : 0), o_P = take(scene_planes ? px * 4 * S->Ns : 0),
               o_G = take(scene_planes ? px * 4 * S->Ns : 0), o_X = take(scene_planes ? px * 4 * S->Ns : 0);
  const size_t o_conv = take(px), o_nev = take(px * 4);
  const size_t o_qs = take(own * 4 + 4), o_qd = take(own * 4 + 4), o_qa = take(S->two_classes ? 4 : own * 4 + 4), o_sc = take(kScalars * 4);
  S->bytes = o;
  cudaError_t e = cudaMalloc(&S->base, S->bytes);
  if (e != cudaSuccess) { g_last_cuda_error = std::string("cudaMalloc(band): ") + cudaGetErrorString(e); delete S; return e == cudaErrorMemoryAllocation ? PHB_ENOMEM : PHB_ECUDA; }
  unsigned char *b = S->base;
  BandView &v = S->view;
  memset(&v, 0, sizeof(v));
  v.planes = reinterpret_cast<float *>(b + o_planes);
  v.prior = desc->prior_present ? reinterpret_cast<float *>(b + o_prior) : nullptr;
  v.nrows = desc->nrows;
  float **slots[9] = {&v.out.depth, &v.out.model_error, &v.out.bottom_albedo, &v.out.bottom_sand, &v.out.bottom_seagrass,
                      &v.out.bottom_coral, &v.out.K_min, &v.out.bottom_type, &v.out.index_optical_depth};
  for (int k = 0; k < 9; k++) *slots[k] = reinterpret_cast<float *>(b + o_out[k]);
  if (scene_planes) {
    v.out.K = reinterpret_cast<float *>(b + o_K); v.out.P = reinterpret_cast<float *>(b + o_P);
    v.out.G = reinterpret_cast<float *>(b + o_G); v.out.X = reinterpret_cast<float *>(b + o_X);
  }
  v.out.converged = b + o_conv; v.out.n_evals = reinterpret_cast<int32_t *>(b + o_nev);
  S->q_shallow = reinterpret_cast<int *>(b + o_qs); S->q_deep = reinterpret_cast<int *>(b + o_qd);
  S->q_all = reinterpret_cast<int *>(b + o_qa); S->scalars = reinterpret_cast<int *>(b + o_sc);
  bind_queues(v, S->q_shallow, S->q_deep, S->q_all, S->scalars, S->two_classes != 0);
  for (int k = 0; k < 4; k++) {
    e = cudaEventCreate(&S->ev[k]);
    if (e != cudaSuccess) { g_last_cuda_error = std::string("cudaEventCreate: ") + cudaGetErrorString(e); phb_shard_destroy(S); return PHB_ECUDA; }
  }
  e = cudaMemset(S->scalars, 0, kScalars * sizeof(int));
  if (e != cudaSuccess) { g_last_cuda_error = std::string("cudaMemset: ") + cudaGetErrorString(e); phb_shard_destroy(S); return PHB_ECUDA; }
  *out = S;
  return PHB_OK;
}

void phb_shard_destroy(phb_shard *S) {
  if (!S) return;
  cudaSetDevice(S->ctx->device);
  for (auto &m : S->mapped)
    if (m.ipc) cudaIpcCloseMemHandle(m.ptr);
  if (S->base) cudaFree(S->base);
  for (cudaEvent_t x : S->ev) if (x) cudaEventDestroy(x);
  delete S;
}

int phb_shard_buffers(phb_shard *S, float **d_planes, float **d_prior, phb_outputs *d_out) {
  if (!S) return PHB_EINVAL;
  if (d_planes) *d_planes = const_cast<float *>(S->view.planes);
  if (d_prior) *d_prior = const_cast<float *>(S->view.prior);
  if (d_out) *d_out = S->view.out;
  return PHB_OK;
}

int phb_shard_export(phb_shard *S, phb_shard_handle *h) {
  if (!S || !h) return PHB_EINVAL;
  CK(cudaSetDevice(S->ctx->device));
  memset(h, 0, sizeof(*h));
  HandleData d;
  memset(&d, 0, sizeof(d));
  d.magic = kHandleMagic; d.device = S->ctx->device; d.pid = (int64_t)getpid();
  d.base = (uint64_t)(uintptr_t)S->base; d.bytes = S->bytes;
  d.nrows = S->desc.nrows; d.ncols = S->desc.ncols; d.two_classes = S->two_classes;
  d.ipc_ok = cudaIpcGetMemHandle(&d.ipc, S->base) == cudaSuccess ? 1 : 0; /* only another PROCESS needs it */
  if (!d.ipc_ok) (void)cudaGetLastError();
  BandView v = S->view;
  const void **slots[23];
  view_pointers(v, slots);
  for (int k = 0; k < 23; k++) d.off[k] = *slots[k] ? (uint64_t)((const unsigned char *)*slots[k] - S->base) : kNullOff;
  memcpy(h->bytes, &d, sizeof(d));
  return PHB_OK;
}

/* a peer's band in THIS device's address space */
static int map_peer(phb_shard *S, const phb_shard_handle *h, BandView *out) {
  HandleData d;
  memcpy(&d, h->bytes, sizeof(d));
  if (d.magic != kHandleMagic || d.ncols != S->desc.ncols || d.two_classes != S->two_classes) return PHB_EINVAL;
  unsigned char *ptr = nullptr;
  for (auto &m : S->mapped)
    if (m.pid == d.pid && m.base == d.base) ptr = m.ptr;
  if (!ptr) {
    const int me = S->ctx->device;
    if (d.pid == (int64_t)getpid()) { /* same process: the owner's pointer, peer access between the two devices */
      if (d.device != me) {
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, me, d.device) != cudaSuccess || !can) { (void)cudaGetLastError(); return PHB_ENOPEER; }
        cudaError_t e = cudaDeviceEnablePeerAccess(d.device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { (void)cudaGetLastError(); return PHB_ENOPEER; }
        (void)cudaGetLastError();
      }
      ptr = reinterpret_cast<unsigned char *>((uintptr_t)d.base);
      S->mapped.push_back({d.pid, d.base, ptr, false});
    } else { /* another process: CUDA IPC (enables peer access as needed) */
      if (!d.ipc_ok) return PHB_ENOPEER;
      void *p = nullptr;
      cudaError_t e = cudaIpcOpenMemHandle(&p, d.ipc, cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess) {
        g_last_cuda_error = std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e);
        (void)cudaGetLastError();
        return PHB_ENOPEER;
      }
      ptr = static_cast<unsigned char *>(p);
      S->mapped.push_back({d.pid, d.base, ptr, true});
    }
  }
  BandView v;
  memset(&v, 0, sizeof(v));
  const void **slots[23];
  view_pointers(v, slots);
  for (int k = 0; k < 23; k++) *slots[k] = d.off[k] == kNullOff ? nullptr : (const void *)(ptr + d.off[k]);
  v.nrows = d.nrows;
  *out = v;
  return PHB_OK;
}

int phb_shard_prepare(phb_shard *S, void *stream) {
  if (!S) return PHB_EINVAL;
  phb_ctx *c = S->ctx;
  cudaStream_t st = (cudaStream_t)stream;
  CK(cudaSetDevice(c->device));
  CK(cudaMemcpyAsync(c->d_model, &S->M, sizeof(ModelConst), cudaMemcpyHostToDevice, st));
  CK(cudaEventRecord(S->ev[0], st));
  if (S->row_end > S->row_begin) {
    int rc = launch_classify(c, S->view, S->row_begin, S->row_end, S->q_shallow, S->q_deep, S->q_all, S->scalars,
                             S->two_classes != 0, st);
    if (rc) return rc;
  } else {
    CK(cudaMemsetAsync(S->scalars, 0, kScalars * sizeof(int), st));
  }
  CK(cudaEventRecord(S->ev[1], st));
  return PHB_OK;
}

int64_t phb_shard_valid(phb_shard *S) {
  if (!S) return -1;
  int sc[kScalars];
  if (cudaSetDevice(S->ctx->device) != cudaSuccess) return -1;
  if (cudaMemcpy(sc, S->scalars, sizeof(sc), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  return (int64_t)sc[0] + sc[1];
}

int phb_shard_solve(phb_shard *S, const phb_shard_handle *peers, int n_peers, void *stream, phb_stats *stats) {
  if (!S || n_peers < 0 || n_peers + 1 > kMaxViews || (n_peers > 0 && !peers)) return PHB_EINVAL;
  phb_ctx *c = S->ctx;
  cudaStream_t st = (cudaStream_t)stream;
  CK(cudaSetDevice(c->device));
  std::vector<BandView> views(1 + n_peers);
  views[0] = S->view;
  for (int k = 0; k < n_peers; k++) {
    int rc = map_peer(S, &peers[k], &views[1 + k]);
    if (rc) return rc;
  }
  /* the model of THIS band (prepare of another band on the same context may have replaced it) */
  CK(cudaMemcpyAsync(c->d_model, &S->M, sizeof(ModelConst), cudaMemcpyHostToDevice, st));
  phb_stats local;
  memset(&local, 0, sizeof(local));
  int rc = run_solve(c, S->M, views.data(), 1 + n_peers, S->two_classes != 0, S->dbg_rec, S->dbg_pix, S->dbg_iters,
                     phb_debug_record_len(&S->desc), S->dbg_cap, st, S->ev[2], S->ev[3], stats ? &local : nullptr);
  if (rc) return rc;
  if (stats) {
    int sc[kScalars];
    CK(cudaMemcpy(sc, S->scalars, sizeof(sc), cudaMemcpyDeviceToHost));
    local.n_shallow = sc[0];
    CK(cudaEventElapsedTime(&local.ms_classify, S->ev[0], S->ev[1]));
    *stats = local;
  }
  return PHB_OK;
}

/* ---- host entry points: caller's rasters in host memory ------------------------------------------------------ */

}  /* extern "C" */

namespace {

/* read access to the caller's host rasters by row: contiguous planes or the reference's row pointers */
struct RowSrc {
  const float *const *planes = nullptr;            /* [SB], each [rows][ncols] */
  const float *const *const *plane_rows = nullptr; /* [SB][rows] */
  const float *prior = nullptr;
  const float *const *prior_rows = nullptr;
  int ncols = 0;
  const float *row(int g, long long r) const { return plane_rows ? plane_rows[g][r] : planes[g] + (size_t)r * ncols; }
  const float *prow(long long r) const { return prior_rows ? prior_rows[r] : prior + (size_t)r * ncols; }
  bool has_prior() const { return prior != nullptr || prior_rows != nullptr; }
};

/* write access to the caller's result rasters by row */
struct RowDst {
  const phb_outputs *flat = nullptr;
  const phb_row_outputs *rows = nullptr;
  size_t plane_px = 0; /* cells per plane of the caller's stacked K / P / G / X planes (flat form) */
  int ncols = 0;
  /* scalar grid k (0..8) */
  float *scalar(int k, long long r) const {
    if (rows) {
      float *const *g[9] = {rows->depth, rows->model_error, rows->bottom_albedo, rows->bottom_sand, rows->bottom_seagrass,
                            rows->bottom_coral, rows->K_min, rows->bottom_type, rows->index_optical_depth};
      return g[k] ? g[k][r] : nullptr;
    }
    float *g[9] = {flat->depth, flat->model_error, flat->bottom_albedo, flat->bottom_sand, flat->bottom_seagrass,
                   flat->bottom_coral, flat->K_min, flat->bottom_type, flat->index_optical_depth};
    return g[k] ? g[k] + (size_t)r * ncols : nullptr;
  }
  /* stacked planes: which = 0 K, 1 P, 2 G, 3 X; q = plane within the stack */
  float *stacked(int which, int q, long long r) const {
    if (rows) {
      float *const *const *g[4] = {rows->K, rows->P, rows->G, rows->X};
      return g[which] ? g[which][q][r] : nullptr;
    }
    float *g[4] = {flat->K, flat->P, flat->G, flat->X};
    return g[which] ? g[which] + (size_t)q * plane_px + (size_t)r * ncols : nullptr;
  }
  bool wants_stacked() const { return rows ? (rows->K || rows->P || rows->G || rows->X) : (flat->K || flat->P || flat->G || flat->X); }
};

int ring_open(phb_ctx *c) {
  for (int i = 0; i < 2; i++) {
    if (!c->ring[i]) CK(cudaHostAlloc(&c->ring[i], kRingBytes, cudaHostAllocDefault));
    if (!c->ring_ev[i]) CK(cudaEventCreateWithFlags(&c->ring_ev[i], cudaEventDisableTiming));
  }
  return PHB_OK;
}

/* host rows -> device plane through the pinned ring: rows [r0, r1) of one host raster (row accessor) to dst */
template <class RowFn>
int ring_upload(phb_ctx *c, int &slot, RowFn row_of, long long r0, long long r1, int ncols, float *dst, cudaStream_t st) {
  const long long rows_per = (long long)(kRingBytes / ((size_t)ncols * 4));
  if (rows_per < 1) return PHB_EINVAL;
  for (long long a = r0; a < r1; a += rows_per) {
    const long long b = a + rows_per < r1 ? a + rows_per : r1;
    CK(cudaEventSynchronize(c->ring_ev[slot])); /* the copy that last used this slot has drained */
    float *buf = c->ring[slot];
    for (long long r = a; r < b; r++) memcpy(buf + (size_t)(r - a) * ncols, row_of(r), (size_t)ncols * 4);
    CK(cudaMemcpyAsync(dst + (size_t)(a - r0) * ncols, buf, (size_t)(b - a) * ncols * 4, cudaMemcpyHostToDevice, st));
    CK(cudaEventRecord(c->ring_ev[slot], st));
    slot ^= 1;
  }
  return PHB_OK;
}

/* device plane rows -> host rows through the pinned ring, double buffered: while chunk k is in flight chunk k-1 is
 * scattered into the caller's rows */
template <class RowFn>
int ring_download(phb_ctx *c, RowFn row_of, long long r0, long long r1, int ncols, const void *src, size_t elem, cudaStream_t st) {
  const long long rows_per = (long long)(kRingBytes / ((size_t)ncols * elem));
  if (rows_per < 1) return PHB_EINVAL;
  long long pa[2] = {0, 0}, pb[2] = {0, 0};
  bool pending[2] = {false, false};
  int slot = 0;
  auto drain = [&](int s) -> int {
    if (!pending[s]) return PHB_OK;
    CK(cudaEventSynchronize(c->ring_ev[s]));
    const unsigned char *buf = reinterpret_cast<const unsigned char *>(c->ring[s]);
    for (long long r = pa[s]; r < pb[s]; r++) memcpy(row_of(r), buf + (size_t)(r - pa[s]) * ncols * elem, (size_t)ncols * elem);
    pending[s] = false;
    return PHB_OK;
  };
  for (long long a = r0; a < r1; a += rows_per) {
    const long long b = a + rows_per < r1 ? a + rows_per : r1;
    int rc = drain(slot);
    if (rc) return rc;
    CK(cudaMemcpyAsync(c->ring[slot], static_cast<const unsigned char *>(src) + (size_t)(a - r0) * ncols * elem,
                       (size_t)(b - a) * ncols * elem, cudaMemcpyDeviceToHost, st));
    CK(cudaEventRecord(c->ring_ev[slot], st));
    pa[slot] = a; pb[slot] = b; pending[slot] = true;
    slot ^= 1;
  }
  int rc = drain(slot);
  if (rc) return rc;
  return drain(slot ^ 1);
}

/* the band of a host entry point: kept in the context between calls while its shape stays the same */
int host_band(phb_ctx *c, const phb_scene_desc *dd, int row_begin, int row_end, bool scene_planes, phb_shard **out) {
  phb_shard *S = c->host_shard;
  if (S && (memcmp(&S->desc, dd, sizeof(*dd)) != 0 || S->row_begin != row_begin || S->row_end != row_end ||
            S->scene_planes != (scene_planes ? 1 : 0) || S->two_classes != (use_two_classes(S->M) ? 1 : 0))) {
    phb_shard_destroy(S);
    S = c->host_shard = nullptr;
  }
  if (!S) {
    int rc = phb_shard_create(c, dd, row_begin, row_end, scene_planes ? 1 : 0, &S);
    if (rc) return rc;
    c->host_shard = S;
  }
  *out = S;
  return PHB_OK;
}

/* is [p, p + bytes) page-locked host memory (cudaHostAlloc / cudaHostRegister)? Then the copy engine can read or
 * write it directly and the staging ring is skipped. */
bool is_pinned(const void *p, size_t bytes) {
  if (!p || bytes == 0) return false;
  cudaPointerAttributes a0, a1;
  if (cudaPointerGetAttributes(&a0, p) != cudaSuccess) { (void)cudaGetLastError(); return false; }
  if (cudaPointerGetAttributes(&a1, static_cast<const unsigned char *>(p) + bytes - 1) != cudaSuccess) { (void)cudaGetLastError(); return false; }
  return a0.type == cudaMemoryTypeHost && a1.type == cudaMemoryTypeHost;
}

/* rows [w0, w1) of the caller's rasters -> the band's device rasters */
int band_upload(phb_shard *S, const RowSrc &src, long long w0, cudaStream_t st) {
  phb_ctx *c = S->ctx;
  int rc = ring_open(c);
  if (rc) return rc;
  const int nr = S->desc.nrows, nc = S->desc.ncols;
  const size_t px = (size_t)nr * nc;
  int slot = 0;
  for (int g = 0; g < S->SB; g++) {
    float *dst = const_cast<float *>(S->view.planes) + g * px;
    if (src.planes && is_pinned(src.planes[g] + (size_t)w0 * nc, px * 4)) { /* contiguous and page-locked: one DMA */
      CK(cudaMemcpyAsync(dst, src.planes[g] + (size_t)w0 * nc, px * 4, cudaMemcpyHostToDevice, st));
      continue;
    }
    rc = ring_upload(c, slot, [&](long long r) { return src.row(g, r); }, w0, w0 + nr, nc, dst, st);
    if (rc) return rc;
  }
  if (S->view.prior) {
    if (src.prior && is_pinned(src.prior + (size_t)w0 * nc, px * 4)) {
      CK(cudaMemcpyAsync(const_cast<float *>(S->view.prior), src.prior + (size_t)w0 * nc, px * 4, cudaMemcpyHostToDevice, st));
    } else {
      rc = ring_upload(c, slot, [&](long long r) { return src.prow(r); }, w0, w0 + nr, nc, const_cast<float *>(S->view.prior), st);
      if (rc) return rc;
    }
  }
  return PHB_OK;
}

/* the band's own rows of the result planes -> the caller's rasters (scene rows w0 + [row_begin, row_end)) */
int band_download(phb_shard *S, const RowDst &dst, const phb_outputs *flat_flags, long long w0, cudaStream_t st) {
  phb_ctx *c = S->ctx;
  const int nc = S->desc.ncols;
  const size_t px = (size_t)S->desc.nrows * nc, a = (size_t)S->row_begin * nc;
  const long long r0 = w0 + S->row_begin, r1 = w0 + S->row_end;
  const phb_outputs &d = S->view.out;
  const float *dev9[9] = {d.depth, d.model_error, d.bottom_albedo, d.bottom_sand, d.bottom_seagrass, d.bottom_coral,
                          d.K_min, d.bottom_type, d.index_optical_depth};
  int rc;
  const size_t own = (size_t)(r1 - r0) * nc;
  for (int k = 0; k < 9; k++) {
    if (!dst.scalar(k, r0)) continue;
    if (dst.flat && is_pinned(dst.scalar(k, r0), own * 4)) { /* contiguous and page-locked: one DMA, no staging */
      CK(cudaMemcpyAsync(dst.scalar(k, r0), dev9[k] + a, own * 4, cudaMemcpyDeviceToHost, st));
      continue;
    }
    rc = ring_download(c, [&](long long r) { return (void *)dst.scalar(k, r); }, r0, r1, nc, dev9[k] + a, 4, st);
    if (rc) return rc;
  }
  if (S->scene_planes) {
    const float *devs[4] = {d.K, d.P, d.G, d.X};
    const int count[4] = {S->Ns * S->mb, S->Ns, S->Ns, S->Ns};
    for (int w = 0; w < 4; w++)
      for (int q = 0; q < count[w]; q++) {
        if (!dst.stacked(w, q, r0)) continue;
        rc = ring_download(c, [&](long long r) { return (void *)dst.stacked(w, q, r); }, r0, r1, nc, devs[w] + (size_t)q * px + a, 4, st);
        if (rc) return rc;
      }
  }
  if (flat_flags && flat_flags->converged) {
    uint8_t *h = flat_flags->converged;
    if (is_pinned(h + (size_t)r0 * nc, own)) {
      CK(cudaMemcpyAsync(h + (size_t)r0 * nc, d.converged + a, own, cudaMemcpyDeviceToHost, st));
    } else {
      rc = ring_download(c, [&](long long r) { return (void *)(h + (size_t)r * nc); }, r0, r1, nc, d.converged + a, 1, st);
      if (rc) return rc;
    }
  }
  if (flat_flags && flat_flags->n_evals) {
    int32_t *h = flat_flags->n_evals;
    if (is_pinned(h + (size_t)r0 * nc, own * 4)) {
      CK(cudaMemcpyAsync(h + (size_t)r0 * nc, d.n_evals + a, own * 4, cudaMemcpyDeviceToHost, st));
    } else {
      rc = ring_download(c, [&](long long r) { return (void *)(h + (size_t)r * nc); }, r0, r1, nc, d.n_evals + a, 4, st);
      if (rc) return rc;
    }
  }
  return PHB_OK;
}

int scene_bands(const phb_scene_desc *d) {
  int SB = 0;
  for (int s = 0; s < d->n_scenes; s++) SB += d->n_bands[s];
  return SB;
}

/* one device, whole raster (or a row window of it): upload, classify, solve, download */
int invert_host_one(phb_ctx *c, const phb_scene_desc *desc, const RowSrc &src, int row_begin, int row_end, const RowDst &dst,
                    const phb_outputs *flat_flags, phb_stats *stats, double *rec, int32_t *pix, int32_t *iters, int64_t capacity) {
  if (!c) return PHB_EINVAL;
  int rc = validate(desc);
  if (rc) return rc;
  if (row_begin < 0 || row_end > desc->nrows || row_begin >= row_end) return PHB_EINVAL;
  CK(cudaSetDevice(c->device));
  cudaStream_t st = 0;
  phb_scene_desc dd = *desc;
  dd.prior_present = (desc->prior_present && src.has_prior()) ? 1 : 0;
  phb_shard *S = nullptr;
  rc = host_band(c, &dd, row_begin, row_end, dst.wants_stacked(), &S);
  if (rc) return rc;
  struct Events { /* released on every return path */
    cudaEvent_t e[4] = {nullptr, nullptr, nullptr, nullptr};
    ~Events() { for (cudaEvent_t x : e) if (x) cudaEventDestroy(x); }
  } evs;
  for (int k = 0; k < 4; k++) CK(cudaEventCreate(&evs.e[k]));
  CK(cudaEventRecord(evs.e[0], st));
  S->resident_src = nullptr;
  rc = band_upload(S, src, 0, st);
  if (rc) return rc;
  if (row_begin == 0 && row_end == desc->nrows && src.plane_rows) S->resident_src = (const void *)src.plane_rows;
  CK(cudaEventRecord(evs.e[1], st));
  const int reclen = phb_debug_record_len(desc);
  S->dbg_rec = nullptr; S->dbg_pix = nullptr; S->dbg_iters = nullptr; S->dbg_cap = 0;
  if (rec && capacity > 0) {
    CK(c->dbg_rec.ensure((size_t)capacity * reclen)); CK(c->dbg_pix.ensure(capacity)); CK(c->dbg_iters.ensure(3 * capacity));
    CK(cudaMemsetAsync(c->dbg_pix.p, 0xff, capacity * sizeof(int), st));
    S->dbg_rec = c->dbg_rec.p; S->dbg_pix = c->dbg_pix.p; S->dbg_iters = c->dbg_iters.p; S->dbg_cap = capacity;
  }
  rc = phb_shard_prepare(S, st);
  if (rc) return rc;
  phb_stats local;
  memset(&local, 0, sizeof(local));
  rc = phb_shard_solve(S, nullptr, 0, st, &local);
  S->dbg_rec = nullptr; S->dbg_pix = nullptr; S->dbg_iters = nullptr; S->dbg_cap = 0;
  if (rc) return rc;
  CK(cudaEventRecord(evs.e[2], st));
  rc = band_download(S, dst, flat_flags, 0, st);
  if (rc) return rc;
  if (rec && capacity > 0) {
    CK(cudaMemcpyAsync(rec, c->dbg_rec.p, (size_t)capacity * reclen * sizeof(double), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(pix, c->dbg_pix.p, capacity * sizeof(int), cudaMemcpyDeviceToHost, st));
    if (iters) CK(cudaMemcpyAsync(iters, c->dbg_iters.p, 3 * capacity * sizeof(int), cudaMemcpyDeviceToHost, st));
  }
  CK(cudaEventRecord(evs.e[3], st));
  CK(cudaStreamSynchronize(st));
  CK(cudaEventElapsedTime(&local.ms_h2d, evs.e[0], evs.e[1]));
  CK(cudaEventElapsedTime(&local.ms_d2h, evs.e[2], evs.e[3]));
  if (stats) *stats = local;
  return PHB_OK;
}

}  // namespace

extern "C" {

int phb_invert_host(phb_ctx *ctx, const phb_scene_desc *desc, const float *const *h_planes, const float *h_prior,
                    int row_begin, int row_end, const phb_outputs *h_out, phb_stats *stats) {
  if (!h_planes || !h_out || !desc) return PHB_EINVAL;
  RowSrc src; src.planes = h_planes; src.prior = h_prior; src.ncols = desc->ncols;
  RowDst dst; dst.flat = h_out; dst.ncols = desc->ncols; dst.plane_px = (size_t)desc->nrows * desc->ncols;
  return invert_host_one(ctx, desc, src, row_begin, row_end, dst, h_out, stats, nullptr, nullptr, nullptr, 0);
}

int phb_invert_host_debug(phb_ctx *ctx, const phb_scene_desc *desc, const float *const *h_planes, const float *h_prior,
                          int row_begin, int row_end, const phb_outputs *h_out, double *rec, int32_t *pix,
                          int32_t *n_iters, int64_t capacity, phb_stats *stats) {
  if (!h_planes || !h_out || !desc) return PHB_EINVAL;
  RowSrc src; src.planes = h_planes; src.prior = h_prior; src.ncols = desc->ncols;
  RowDst dst; dst.flat = h_out; dst.ncols = desc->ncols; dst.plane_px = (size_t)desc->nrows * desc->ncols;
  return invert_host_one(ctx, desc, src, row_begin, row_end, dst, h_out, stats, rec, pix, n_iters, capacity);
}

/* ---- one process, several GPUs: row bands over one context per device ----------------------- */

namespace {

/* Relative cost of one inversion by DEPTHS-prior bin: mean evaluation count per bin (CPU oracle, Exmouth- and
 * Pilbara-shaped rasters) times the per-evaluation cost of the pixel class; the same table as
 * photic_b200/sharded.py:row_cost_from_prior (8-band balance 0.77 -> 0.89 on Pilbara). Only the
 * fallback without peer access plans by it; with peer access the devices share the work at run time. */
const float kCostEdges[8] = {2.0f, 4.0f, 6.0f, 8.0f, 12.0f, 16.0f, 24.0f, 32.0f};
const float kCostWeight[9] = {2.4f, 2.9f, 2.6f, 2.2f, 1.0f, 1.1f, 1.05f, 1.35f, 1.4f};
const float kCostNoPrior = 8.0f * 2.4f; /* eight H starts, all substrates (samodel.c:2222-2241) */

/* estimated work of rows [r0, r1): validity as classify_kernel decides it (samodel.c:933-947), weight by prior bin */
void row_costs(const phb_scene_desc *d, int SB, const RowSrc *src, int r0, int r1, double *cost) {
  const int nc = d->ncols;
  std::vector<unsigned char> ok(nc);
  for (int r = r0; r < r1; r++) {
    for (int c = 0; c < nc; c++) ok[c] = 1;
    int g = 0;
    for (int s = 0; s < d->n_scenes; s++)
      for (int b = 0; b < d->n_bands[s]; b++, g++) {
        const float *p = src->row(g, r);
        const float nd = d->nodata_per_band ? d->nodata_band[s][b] : d->nodata;
        for (int c = 0; c < nc; c++) {
          const float v = p[c];
          if (v < 0.0f || approx_equal_f(v, nd, 1.0e-6f)) ok[c] = 0;
        }
      }
    (void)SB;
    const float *pr = (d->prior_present && src->has_prior()) ? src->prow(r) : nullptr;
    double acc = 0.0;
    for (int c = 0; c < nc; c++) {
      if (!ok[c]) continue;
      float w = kCostNoPrior;
      if (pr) {
        const float e = pr[c];
        if (!approx_equal_f(e, d->prior_nodata, 1.0e-6f)) {
          const float h = (e > -1.0f) ? 1.0f : fabsf(e);
          int bin = 0;
          while (bin < 8 && kCostEdges[bin] <= h) bin++;
          w = kCostWeight[bin];
        }
      }
      acc += (double)w;
    }
    cost[r] = acc;
  }
}

/* contiguous bands of near-equal cumulative cost (photic_b200/sharded.py:plan_row_bands) */
void plan_bands(const double *cost, int nrows, int parts, int32_t *edges) {
  double total = 0.0;
  for (int r = 0; r < nrows; r++) total += cost[r];
  edges[0] = 0; edges[parts] = nrows;
  if (!(total > 0.0)) {
    for (int k = 1; k < parts; k++) edges[k] = (int32_t)llround((double)nrows * k / parts);
    return;
  }
  double cum = 0.0;
  int r = 0; /* edges[k] = smallest r with cost of rows [0, r) >= total * k / parts */
  for (int k = 1; k < parts; k++) {
    const double target = total * k / parts;
    while (r < nrows && cum < target) cum += cost[r++];
    edges[k] = r;
  }
}

int plan_row_bands_src(const phb_scene_desc *desc, const RowSrc &src, int n_parts, int32_t *edges, double *row_cost) {
  const int SB = scene_bands(desc), nrows = desc->nrows;
  std::vector<double> cost(nrows, 0.0);
  int nthr = (int)std::thread::hardware_concurrency(); /* the scan is memory-bound host work: use the cores */
  if (nthr < n_parts) nthr = n_parts;
  if (nthr > 32) nthr = 32;
  if (nthr > nrows) nthr = nrows;
  if (nthr < 1) nthr = 1;
  std::vector<std::thread> th;
  for (int t = 0; t < nthr; t++) {
    const int r0 = (int)((long long)nrows * t / nthr), r1 = (int)((long long)nrows * (t + 1) / nthr);
    th.emplace_back(row_costs, desc, SB, &src, r0, r1, cost.data());
  }
  for (auto &t : th) t.join();
  plan_bands(cost.data(), nrows, n_parts, edges);
  if (row_cost) memcpy(row_cost, cost.data(), (size_t)nrows * sizeof(double));
  return PHB_OK;
}

/* every host thread arrives; nobody leaves before the last one is in */
struct HostBarrier {
  std::mutex m; std::condition_variable cv; int n, waiting = 0; unsigned long gen = 0;
  explicit HostBarrier(int n_) : n(n_) {}
  void arrive() {
    std::unique_lock<std::mutex> lk(m);
    const unsigned long g = gen;
    if (++waiting == n) { waiting = 0; gen++; cv.notify_all(); }
    else cv.wait(lk, [&] { return gen != g; });
  }
};

/* can every device of the list read and write every other one's memory? (several contexts on one device: trivially) */
bool all_peers(phb_ctx *const *ctxs, int n) {
  for (int a = 0; a < n; a++)
    for (int b = 0; b < n; b++) {
      if (ctxs[a]->device == ctxs[b]->device) continue;
      int can = 0;
      if (cudaDeviceCanAccessPeer(&can, ctxs[a]->device, ctxs[b]->device) != cudaSuccess || !can) { (void)cudaGetLastError(); return false; }
    }
  return true;
}

int invert_host_many(phb_ctx *const *ctxs, int n_ctx, const phb_scene_desc *desc, const RowSrc &src, const RowDst &dst,
                     const phb_outputs *flat_flags, phb_stats *stats, phb_stats *per_ctx, int32_t *edges_out) {
  if (!ctxs || n_ctx < 1 || n_ctx > kMaxViews) return PHB_EINVAL;
  for (int k = 0; k < n_ctx; k++)
    if (!ctxs[k]) return PHB_EINVAL;
  int rc = validate(desc);
  if (rc) return rc;
  const int nrows = desc->nrows;
  std::vector<int32_t> edges(n_ctx + 1);
  const char *e_ns = getenv("PHB_NO_STEAL");
  const bool steal = n_ctx > 1 && !(e_ns && atoi(e_ns) != 0) && all_peers(ctxs, n_ctx);
  if (steal || n_ctx == 1) {
    for (int k = 0; k <= n_ctx; k++) edges[k] = (int32_t)((long long)nrows * k / n_ctx); /* equal rows: the kernels share the work */
  } else {
    rc = plan_row_bands_src(desc, src, n_ctx, edges.data(), nullptr);
    if (rc) return rc;
  }
  const int nsp = desc->n_spatial == 0 ? 1 : desc->n_spatial; /* samodel.c:2967 */
  const int halo = (nsp - 1) + (desc->n_smoothing_radius - 1);
  std::vector<int> rcs(n_ctx, PHB_OK);
  std::vector<std::string> errs(n_ctx);
  std::vector<phb_stats> sts(n_ctx);
  memset(sts.data(), 0, sizeof(phb_stats) * n_ctx);
  std::vector<phb_shard_handle> handles(n_ctx);
  std::vector<int> active(n_ctx, 0);
  std::vector<int64_t> own_valid(n_ctx, 0);
  std::atomic<bool> failed(false);
  HostBarrier bar(n_ctx);
  auto fail = [&](int k, int code) { rcs[k] = code; if (code == PHB_ECUDA || code == PHB_ENOPEER) errs[k] = g_last_cuda_error; failed.store(true); };
  auto work = [&](int k) {
    phb_ctx *c = ctxs[k];
    const int r0 = edges[k], r1 = edges[k + 1];
    /* the band with its halo rows is a raster of its own: edge clamping then only ever happens at the real
     * edges of the scene, so band + halo == unsharded (tests: row-band shards equal the whole scene) */
    const int w0 = r0 - halo < 0 ? 0 : r0 - halo, w1 = r1 + halo > nrows ? nrows : r1 + halo;
    phb_shard *S = nullptr;
    cudaStream_t st = 0;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    float ms_h2d = 0.0f, ms_d2h = 0.0f;
    if (r1 > r0) {
      int rc2 = PHB_OK;
      do {
        if (cudaSetDevice(c->device) != cudaSuccess) { rc2 = PHB_ECUDA; g_last_cuda_error = "cudaSetDevice"; break; }
        phb_scene_desc dd = *desc;
        dd.nrows = w1 - w0;
        dd.prior_present = (desc->prior_present && src.has_prior()) ? 1 : 0;
        rc2 = host_band(c, &dd, r0 - w0, r1 - w0, dst.wants_stacked(), &S);
        if (rc2) break;
        for (int q = 0; q < 4; q++)
          if (cudaEventCreate(&ev[q]) != cudaSuccess) { rc2 = PHB_ECUDA; g_last_cuda_error = "cudaEventCreate"; break; }
        if (rc2) break;
        cudaEventRecord(ev[0], st);
        rc2 = band_upload(S, src, w0, st);
        if (rc2) break;
        cudaEventRecord(ev[1], st);
        rc2 = phb_shard_prepare(S, st);
        if (rc2) break;
        rc2 = phb_shard_export(S, &handles[k]);
        if (rc2) break;
        if (cudaStreamSynchronize(st) != cudaSuccess) { rc2 = PHB_ECUDA; g_last_cuda_error = "cudaStreamSynchronize after prepare"; break; }
        active[k] = 1;
      } while (0);
      if (rc2) fail(k, rc2);
    }
    bar.arrive(); /* every band is classified and exported */
    if (S && active[k] && !failed.load()) {
      std::vector<phb_shard_handle> peers;
      if (steal)
        for (int q = 1; q < n_ctx; q++) { const int o = (k + q) % n_ctx; if (active[o]) peers.push_back(handles[o]); }
      int rc2 = phb_shard_solve(S, peers.data(), (int)peers.size(), st, &sts[k]);
      if (rc2) fail(k, rc2);
      else own_valid[k] = phb_shard_valid(S);
    }
    bar.arrive(); /* every device is done: results taken over NVLink have landed in their owners' planes */
    if (S && active[k] && !failed.load()) {
      int rc2 = PHB_OK;
      cudaEventRecord(ev[2], st);
      rc2 = band_download(S, dst, flat_flags, w0, st);
      if (!rc2) {
        cudaEventRecord(ev[3], st);
        if (cudaStreamSynchronize(st) != cudaSuccess) { rc2 = PHB_ECUDA; g_last_cuda_error = "cudaStreamSynchronize after download"; }
      }
      if (!rc2) { cudaEventElapsedTime(&ms_h2d, ev[0], ev[1]); cudaEventElapsedTime(&ms_d2h, ev[2], ev[3]); }
      sts[k].ms_h2d = ms_h2d; sts[k].ms_d2h = ms_d2h;
      if (rc2) fail(k, rc2);
    }
    for (cudaEvent_t x : ev) if (x) cudaEventDestroy(x);
  };
  std::vector<std::thread> th;
  for (int k = 1; k < n_ctx; k++) th.emplace_back(work, k);
  work(0);
  for (auto &t : th) t.join();
  for (int k = 0; k < n_ctx; k++)
    if (rcs[k]) { if (!errs[k].empty()) g_last_cuda_error = errs[k]; return rcs[k]; }
  if (stats) {
    memset(stats, 0, sizeof(*stats));
    for (int k = 0; k < n_ctx; k++) {
      const phb_stats &a = sts[k];
      stats->n_valid += a.n_valid; stats->n_shallow += a.n_shallow; stats->n_evals += a.n_evals;
      stats->n_iters += a.n_iters; stats->n_converged += a.n_converged; stats->alg_flops += a.alg_flops;
      stats->ms_classify = a.ms_classify > stats->ms_classify ? a.ms_classify : stats->ms_classify;
      stats->ms_solve = a.ms_solve > stats->ms_solve ? a.ms_solve : stats->ms_solve;
      stats->ms_h2d = a.ms_h2d > stats->ms_h2d ? a.ms_h2d : stats->ms_h2d;
      stats->ms_d2h = a.ms_d2h > stats->ms_d2h ? a.ms_d2h : stats->ms_d2h;
      if (a.ctas) { stats->warps_per_cta = a.warps_per_cta; stats->ctas = a.ctas; stats->smem_bytes = a.smem_bytes; stats->regs = a.regs; }
    }
  }
  if (per_ctx) memcpy(per_ctx, sts.data(), sizeof(phb_stats) * n_ctx);
  if (edges_out) memcpy(edges_out, edges.data(), sizeof(int32_t) * (n_ctx + 1));
  return PHB_OK;
}

}  // namespace

int phb_plan_row_bands(const phb_scene_desc *desc, const float *const *h_planes, const float *h_prior, int n_parts,
                       int32_t *edges, double *row_cost) {
  int rc = validate(desc);
  if (rc) return rc;
  if (!h_planes || !edges || n_parts < 1) return PHB_EINVAL;
  RowSrc src; src.planes = h_planes; src.prior = h_prior; src.ncols = desc->ncols;
  return plan_row_bands_src(desc, src, n_parts, edges, row_cost);
}

int phb_invert_host_multi(phb_ctx *const *ctxs, int n_ctx, const phb_scene_desc *desc, const float *const *h_planes,
                          const float *h_prior, const phb_outputs *h_out, phb_stats *stats, phb_stats *per_ctx,
                          int32_t *edges_out) {
  if (!h_planes || !h_out || !desc) return PHB_EINVAL;
  RowSrc src; src.planes = h_planes; src.prior = h_prior; src.ncols = desc->ncols;
  RowDst dst; dst.flat = h_out; dst.ncols = desc->ncols; dst.plane_px = (size_t)desc->nrows * desc->ncols;
  return invert_host_many(ctxs, n_ctx, desc, src, dst, h_out, stats, per_ctx, edges_out);
}

int phb_invert_rows(phb_ctx *const *ctxs, int n_ctx, const phb_scene_desc *desc, const float *const *const *plane_rows,
                    const float *const *prior_rows, const phb_row_outputs *out, phb_stats *stats, phb_stats *per_ctx,
                    int32_t *edges_out) {
  if (!plane_rows || !out || !desc) return PHB_EINVAL;
  RowSrc src; src.plane_rows = plane_rows; src.prior_rows = prior_rows; src.ncols = desc->ncols;
  RowDst dst; dst.rows = out; dst.ncols = desc->ncols;
  return invert_host_many(ctxs, n_ctx, desc, src, dst, nullptr, stats, per_ctx, edges_out);
}


/* ---- depth-error estimate (samodel.c:1376-1477) ----------------------------------------------------- */

namespace {
/* libc rand() draws exactly as the reference makes them */
int rand_below(int lo, int hi) { /* random_in_range, common.c:527-543 */
  for (;;) {
    const int v = rand(), range = hi - lo, rem = RAND_MAX % range, bucket = RAND_MAX / range;
    if (v == RAND_MAX) continue;
    if (v < RAND_MAX - rem) return lo + v / bucket;
  }
}
float rand_unit() { return ((float)rand()) / ((float)RAND_MAX); } /* frand, common.c:215 */
float rand_signed(float max) {                                     /* frand2, common.c:220-225 */
  if (rand_unit() < 0.5) return ((float)rand()) / ((float)RAND_MAX / max);
  return -1.0 * ((float)rand()) / ((float)RAND_MAX / max);
}
}  // namespace

}  /* extern "C" */

/* rasters by row (RowSrc), the depth plane and the sigma plane by row accessors: serves both host layouts */
template <class DepthRow, class SigmaRow>
static int depth_sigma_impl(phb_ctx *c, const phb_scene_desc *desc, const RowSrc &src, bool may_reuse, DepthRow depth_row,
                            SigmaRow sigma_row, unsigned seed, int n_samples, int chain_mode, int max_intervals, double *table,
                            int32_t *n_intervals, double *trials, phb_stats *stats) {
  if (!c) return PHB_EINVAL;
  int rc = validate(desc);
  if (rc) return rc;
  if (!desc->prior_present || !src.has_prior()) return PHB_EINVAL; /* see the header: hot trials need the DEPTHS prior */
  if (n_samples < 1 || n_samples > 4096) return PHB_EINVAL;
  if (max_intervals < 1 || max_intervals > PHB_SIGMA_MAX_INTERVALS) max_intervals = PHB_SIGMA_MAX_INTERVALS;
  if (chain_mode != PHB_SIGMA_CHAIN_REFERENCE && chain_mode != PHB_SIGMA_CHAIN_PER_INTERVAL) return PHB_EINVAL;
  CK(cudaSetDevice(c->device));
  const int nrows = desc->nrows, ncols = desc->ncols;
  const size_t px = (size_t)nrows * ncols;
  int SB = 0;
  for (int s = 0; s < desc->n_scenes; s++) SB += desc->n_bands[s];

  /* interval table bounds, samodel.c:1381-1387 (depth here is the negated plane: -h_depth is the reference's) */
  const int n_trials = (int)sqrt((double)(nrows * ncols));
  float mx = -1.0e10f; /* array_max2(depth, ., ., 0.0), common.c:1240-1258 */
  for (int r = 0; r < nrows; r++) {
    const float *dr = depth_row(r);
    for (int q = 0; q < ncols; q++) {
      const float dq = -dr[q];
      if (!approx_equal_f(dq, 0.0f, 1.0e-4f) && dq > mx) mx = dq;
    }
  }
  double maxd = mx;
  if (!(maxd > 0.0)) maxd = 0.0; /* nothing inverted: no interval (the reference would index with (int)-1e10) */
  maxd = 0.25 * ((int)maxd / 0.25);
  if (!(maxd < 30.0)) maxd = 30.0;
  int n_int = (int)((int)maxd / 0.25);
  if (n_int > max_intervals) { n_int = max_intervals; maxd = 0.25 * max_intervals; }

  /* the draws, in the reference's order (samodel.c:1396-1446) */
  std::vector<int> t_pix, t_slot, chain_begin;
  std::vector<float> t_nsig;
  std::vector<double> td((size_t)(n_int > 0 ? n_int : 1) * n_samples, 0.0);
  srand(seed);
  int kd = 0;
  chain_begin.push_back(0);
  for (double d = 0.0; d < maxd; d += 0.25, kd++) {
    for (int ks = 0; ks < n_samples; ks++) {
      bool found = false;
      int i = 0, j = 0;
      for (int kt = 0; kt < n_trials; kt++) {
        i = rand_below(0, nrows);
        j = rand_below(0, ncols);
        const float dq = -depth_row(i)[j];
        if (dq > d && dq < d + 0.25) { found = true; break; }
      }
      if (!found) continue;                          /* trial_depths[k_sample] = 0 */
      const float ns = (float)(double)rand_signed(1.0); /* drawn before the prior test, samodel.c:1425 */
      if (approx_equal_f(src.prow(i)[j], desc->prior_nodata, 1.0e-6f)) continue;
      t_pix.push_back(i * ncols + j);
      t_nsig.push_back(ns);
      t_slot.push_back(kd * n_samples + ks);
    }
    if (chain_mode == PHB_SIGMA_CHAIN_PER_INTERVAL && (int)t_pix.size() > chain_begin.back()) chain_begin.push_back((int)t_pix.size());
  }
  if (chain_mode == PHB_SIGMA_CHAIN_REFERENCE && !t_pix.empty()) chain_begin.push_back((int)t_pix.size());
  const int n_tr = (int)t_pix.size(), n_chains = (int)chain_begin.size() - 1;

  phb_stats local;
  memset(&local, 0, sizeof(local));
  if (n_tr > 0) {
    cudaStream_t st = 0;
    /* the rasters on the device: those the inversion of this very raster left in the context's band, else a fresh copy */
    const float *d_planes = nullptr, *d_prior = nullptr;
    {
      phb_scene_desc dd = *desc;
      dd.prior_present = 1;
      phb_shard *S = c->host_shard;
      if (may_reuse && S && S->resident_src != nullptr && memcmp(&S->desc, &dd, sizeof(dd)) == 0 && S->view.prior &&
          S->resident_src == (const void *)src.plane_rows) {
        d_planes = S->view.planes; d_prior = S->view.prior;
        S->resident_src = nullptr; /* good for the call that directly follows the inversion only */
      } else {
        CK(c->planes.ensure(px * SB));
        CK(c->prior.ensure(px));
        rc = ring_open(c);
        if (rc) return rc;
        int slot = 0;
        for (int g = 0; g < SB; g++) {
          rc = ring_upload(c, slot, [&](long long r) { return src.row(g, r); }, 0, nrows, ncols, c->planes.p + g * px, st);
          if (rc) return rc;
        }
        rc = ring_upload(c, slot, [&](long long r) { return src.prow(r); }, 0, nrows, ncols, c->prior.p, st);
        if (rc) return rc;
        d_planes = c->planes.p; d_prior = c->prior.p;
      }
    }
    /* trial arrays: pix | chain_begin as ints, n_sigma as floats, depths as doubles */
    CK(c->queue.ensure((size_t)n_tr + n_chains + 1));
    CK(c->nev.ensure(n_tr));
    CK(c->dbg_rec.ensure(n_tr));
    CK(cudaMemcpyAsync(c->queue.p, t_pix.data(), n_tr * sizeof(int), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(c->queue.p + n_tr, chain_begin.data(), (n_chains + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(c->nev.p, t_nsig.data(), n_tr * sizeof(float), cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(c->dbg_rec.p, 0, n_tr * sizeof(double), st));
    int sc[4] = {0, 0, n_chains, 0};
    CK(cudaMemcpyAsync(c->d_scalars, sc, sizeof(sc), cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(c->d_counters, 0, 4 * sizeof(unsigned long long), st));
    CK(cudaMemsetAsync(c->d_flops, 0, sizeof(double), st));
    ModelConst M;
    build_model(desc, &M);
    CK(cudaMemcpyAsync(c->d_model, &M, sizeof(M), cudaMemcpyHostToDevice, st));
    SolveParams sp;
    memset(&sp, 0, sizeof(sp));
    BandView tv; /* the raster the trials read; the work items are trial chains (chain_begin), not a pixel queue */
    memset(&tv, 0, sizeof(tv));
    tv.planes = d_planes; tv.prior = d_prior; tv.nrows = nrows;
    tv.n_queue[0] = c->d_scalars + 2; tv.head[0] = c->d_scalars + 3;
    CK(cudaMemcpyAsync(c->d_views, &tv, sizeof(tv), cudaMemcpyHostToDevice, st));
    sp.views = c->d_views; sp.n_views = 1; sp.n_classes = 1;
    sp.trial_pix = c->queue.p; sp.chain_begin = c->queue.p + n_tr;
    sp.trial_nsig = reinterpret_cast<const float *>(c->nev.p);
    sp.trial_depth = c->dbg_rec.p;
    sp.reclen = phb_debug_record_len(desc);
    LaunchGeom geom;
    CK(cudaEventRecord(c->ev[2], st));
    rc = launch_solve(c, M, sp, true, st, &geom);
    if (rc) return rc;
    CK(cudaEventRecord(c->ev[3], st));
    std::vector<double> got(n_tr);
    unsigned long long cnt[4];
    double fl;
    CK(cudaMemcpyAsync(got.data(), c->dbg_rec.p, n_tr * sizeof(double), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(cnt, c->d_counters, sizeof(cnt), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(&fl, c->d_flops, sizeof(fl), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    for (int t = 0; t < n_tr; t++) td[t_slot[t]] = got[t];
    local.n_valid = n_tr; local.n_evals = (int64_t)cnt[0]; local.n_iters = (int64_t)cnt[1]; local.n_converged = (int64_t)cnt[2];
    local.alg_flops = fl;
    CK(cudaEventElapsedTime(&local.ms_solve, c->ev[2], c->ev[3]));
    local.warps_per_cta = geom.W; local.ctas = geom.ctas; local.smem_bytes = geom.smem; local.regs = geom.regs;
  }

  /* per-interval statistics: vec_mean2_double (returns FLOAT, common.c:877-893) and vec_stddev_double
   * (common.c:924-941, pow(.,2) of the host libm), spval 0 */
  double tab[PHB_SIGMA_MAX_INTERVALS];
  for (int k = 0; k < n_int; k++) {
    const double *v = td.data() + (size_t)k * n_samples;
    double total = 0.0, sumdev = 0.0;
    int cnt = 0;
    for (int q = 0; q < n_samples; q++)
      if (!approx_equal_f((float)v[q], 0.0f, 1.0e-4f)) { total += v[q]; cnt++; }
    const double mean = cnt == 0 ? 0.0 : (double)(float)(total / ((double)cnt));
    cnt = 0;
    for (int q = 0; q < n_samples; q++)
      if (!approx_equal_f((float)v[q], 0.0f, 1.0e-4f)) { sumdev += pow(v[q] - mean, 2); cnt++; }
    tab[k] = cnt == 0 ? 0.0 : sqrt(sumdev / ((double)cnt));
  }
  /* every cell gets the sigma of its interval (d, d + 0.25], samodel.c:1463-1477 */
  for (int r = 0; r < nrows; r++) {
    const float *dr = depth_row(r);
    float *sr = sigma_row(r);
    for (int q = 0; q < ncols; q++) {
      const float dq = -dr[q];
      float sg = 0.0f;
      if (dq > 0.0) {
        int k = 0;
        for (double d = 0.0; d < maxd; d += 0.25, k++)
          if (dq > d && dq <= d + 0.25) { sg = (float)tab[k]; break; }
      }
      sr[q] = sg;
    }
  }
  if (table) for (int k = 0; k < n_int; k++) table[k] = tab[k];
  if (n_intervals) *n_intervals = n_int;
  if (trials) memcpy(trials, td.data(), (size_t)n_int * n_samples * sizeof(double));
  if (stats) *stats = local;
  return PHB_OK;
}

extern "C" {

int phb_depth_sigma_host(phb_ctx *c, const phb_scene_desc *desc, const float *const *h_planes, const float *h_prior,
                         const float *h_depth, unsigned seed, int n_samples, int chain_mode, int max_intervals,
                         float *h_depth_sigma, double *table, int32_t *n_intervals, double *trials, phb_stats *stats) {
  if (!c || !desc || !h_planes || !h_depth || !h_depth_sigma) return PHB_EINVAL;
  RowSrc src; src.planes = h_planes; src.prior = h_prior; src.ncols = desc->ncols;
  const int nc = desc->ncols;
  return depth_sigma_impl(c, desc, src, false, [=](int r) { return h_depth + (size_t)r * nc; },
                          [=](int r) { return h_depth_sigma + (size_t)r * nc; }, seed, n_samples, chain_mode, max_intervals,
                          table, n_intervals, trials, stats);
}

int phb_depth_sigma_rows(phb_ctx *c, const phb_scene_desc *desc, const float *const *const *plane_rows,
                         const float *const *prior_rows, const float *const *depth_rows, unsigned seed, int n_samples,
                         int chain_mode, int max_intervals, float *const *sigma_rows, double *table, int32_t *n_intervals,
                         double *trials, phb_stats *stats) {
  if (!c || !desc || !plane_rows || !depth_rows || !sigma_rows) return PHB_EINVAL;
  RowSrc src; src.plane_rows = plane_rows; src.prior_rows = prior_rows; src.ncols = desc->ncols;
  return depth_sigma_impl(c, desc, src, true, [=](int r) { return depth_rows[r]; }, [=](int r) { return sigma_rows[r]; }, seed,
                          n_samples, chain_mode, max_intervals, table, n_intervals, trials, stats);
}

/* ---- known-answer hooks ----------------------------------------------------------------------- */

int phb_kat_objective(phb_ctx *c, const phb_scene_desc *desc, int nb_active, int n_regions, int origin,
                      const double *rrs_measured, int nparams, int nvec, const double *params, double *out6) {
  if (!c) return PHB_EINVAL;
  int rc = validate(desc);
  if (rc) return rc;
  CK(cudaSetDevice(c->device));
  ModelConst M;
  phb_scene_desc dd = *desc;
  dd.n_bottoms = nb_active;
  build_model(&dd, &M);
  CK(cudaMemcpy(c->d_model, &M, sizeof(M), cudaMemcpyHostToDevice));
  SolveParams sp;
  memset(&sp, 0, sizeof(sp));
  sp.L = make_layout(M.SB, M.n_scenes, nb_active, n_regions);
  if (nparams != sp.L.nmax) return PHB_EINVAL;
  const long long slab_doubles = (long long)(sp.L.nmax + 1) * sp.L.nmax + sp.L.nmax + sp.L.Tmax + (long long)((sp.L.nmax + 8) / 8 + 1) * sp.L.nmax;
  CK(c->slabs.ensure(slab_doubles));
  sp.M = c->d_model; sp.slabs = c->slabs.p; sp.slab_stride = slab_doubles;
  sp.exp_tab = c->d_exp_tab; sp.log_tab = c->d_log_tab; sp.pow_tab = c->d_pow_tab;
  double *d_meas, *d_par, *d_out;
  const size_t nm = (size_t)n_regions * M.SB;
  /* rrs_measured arrives [n_regions][n_scenes][max_bands]; flatten to [n_regions][SB] */
  std::vector<double> flat(nm);
  for (int r = 0; r < n_regions; r++)
    for (int s = 0; s < M.n_scenes; s++)
      for (int b = 0; b < M.n_bands[s]; b++)
        flat[(size_t)r * M.SB + M.sb_begin[s] + b] = rrs_measured[((size_t)r * M.n_scenes + s) * M.max_bands + b];
  CK(cudaMalloc(&d_meas, nm * 8)); CK(cudaMalloc(&d_par, (size_t)nvec * nparams * 8)); CK(cudaMalloc(&d_out, (size_t)nvec * 6 * 8));
  CK(cudaMemcpy(d_meas, flat.data(), nm * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_par, params, (size_t)nvec * nparams * 8, cudaMemcpyHostToDevice));
  const size_t smem = (size_t)sp.L.cta_bytes + sp.L.warp_bytes;
  void (*kk)(const SolveParams, int, int, int, const double *, int, const double *, double *, int) =
      sp.L.SBP == 32 ? kat_objective_kernel<32> : kat_objective_kernel<kMaxSB>;
  CK(cudaFuncSetAttribute(kk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const char *e_gen = getenv("PHB_ONE_CLASS"); /* the run-time-substrate-count instantiation for every count */
  kk<<<1, 32, smem>>>(sp, nb_active, n_regions, origin, d_meas, nvec, d_par, d_out, (e_gen && atoi(e_gen) != 0) ? 1 : 0);
  CK(cudaGetLastError());
  CK(cudaMemcpy(out6, d_out, (size_t)nvec * 6 * 8, cudaMemcpyDeviceToHost));
  cudaFree(d_meas); cudaFree(d_par); cudaFree(d_out);
  return PHB_OK;
}

/* Mapping study (aux_kernels.cuh:eval_bench_kernel): closed-loop objective evaluations on every SM, 16 warps per SM,
 * one pixel per team of `team_warps` warps (1: the product's warp-per-pixel objective()). */
int phb_eval_bench(phb_ctx *c, const phb_scene_desc *desc, int nb_active, int n_regions, int origin,
                   const double *rrs_measured, int nparams, const double *params, int team_warps, int same_smsp,
                   int skew_cycles, int reps, double *first_value, double *evals_per_s, float *ms_out) {
  if (!c || !rrs_measured || !params || reps < 1 || skew_cycles < 0) return PHB_EINVAL;
  if (nb_active != 1 && nb_active != 3) return PHB_EINVAL;
  if (team_warps != 1 && team_warps != 2 && team_warps != 4 && team_warps != 8) return PHB_EINVAL;
  int rc = validate(desc);
  if (rc) return rc;
  CK(cudaSetDevice(c->device));
  ModelConst M;
  phb_scene_desc dd = *desc;
  dd.n_bottoms = nb_active;
  build_model(&dd, &M);
  if (M.SB > 32 || n_regions < 1 || n_regions > 16) return PHB_EINVAL; /* the compile-time classes of the product */
  CK(cudaMemcpy(c->d_model, &M, sizeof(M), cudaMemcpyHostToDevice));
  SolveParams sp;
  memset(&sp, 0, sizeof(sp));
  sp.L = make_layout(M.SB, M.n_scenes, nb_active, n_regions);
  if (nparams != sp.L.nmax) return PHB_EINVAL;
  sp.M = c->d_model;
  sp.exp_tab = c->d_exp_tab; sp.log_tab = c->d_log_tab; sp.pow_tab = c->d_pow_tab;
  const int W = 16, ctas = c->n_sm, n_teams = ctas * (W / team_warps);
  const size_t nm = (size_t)n_regions * M.SB;
  std::vector<double> flat(nm);
  for (int r = 0; r < n_regions; r++)
    for (int s = 0; s < M.n_scenes; s++)
      for (int b = 0; b < M.n_bands[s]; b++)
        flat[(size_t)r * M.SB + M.sb_begin[s] + b] = rrs_measured[((size_t)r * M.n_scenes + s) * M.max_bands + b];
  struct Scratch { /* released on every return path */
    DevBuf<double> meas, par, out;
    ~Scratch() { meas.release(); par.release(); out.release(); }
  } scratch;
  DevBuf<double> &d_meas = scratch.meas, &d_par = scratch.par, &d_out = scratch.out;
  CK(d_meas.ensure(nm)); CK(d_par.ensure(nparams)); CK(d_out.ensure((size_t)2 * n_teams));
  CK(cudaMemcpy(d_meas.p, flat.data(), nm * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_par.p, params, (size_t)nparams * 8, cudaMemcpyHostToDevice));
  /* the product's shared-memory footprint (one CTA per SM either way) */
  const size_t smem = (size_t)sp.L.cta_bytes + (size_t)W * sp.L.warp_bytes;
  if (smem > c->smem_optin) return PHB_EINVAL;
  typedef void (*bench_fn)(const SolveParams, int, int, const double *, const double *, int, int, int, double *);
  bench_fn kk = nullptr;
#define PHB_PICK(NB_) \
  (team_warps == 1 ? (bench_fn)eval_bench_kernel<NB_, 32, 1> : team_warps == 2 ? (bench_fn)eval_bench_kernel<NB_, 32, 2> : \
   team_warps == 4 ? (bench_fn)eval_bench_kernel<NB_, 32, 4> : (bench_fn)eval_bench_kernel<NB_, 32, 8>)
  kk = nb_active == 3 ? PHB_PICK(3) : PHB_PICK(1);
#undef PHB_PICK
  CK(cudaFuncSetAttribute(kk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0, 0));
  kk<<<ctas, W * 32, smem>>>(sp, n_regions, origin, d_meas.p, d_par.p, reps, same_smsp ? 1 : 0, skew_cycles, d_out.p);
  cudaError_t le = cudaGetLastError();
  CK(cudaEventRecord(e1, 0));
  cudaError_t se = cudaEventSynchronize(e1);
  float ms = 0.0f;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  if (le != cudaSuccess) { g_last_cuda_error = std::string("eval_bench_kernel: ") + cudaGetErrorString(le); return PHB_ECUDA; }
  if (se != cudaSuccess) { g_last_cuda_error = std::string("eval_bench_kernel: ") + cudaGetErrorString(se); return PHB_ECUDA; }
  std::vector<double> out((size_t)2 * n_teams);
  CK(cudaMemcpy(out.data(), d_out.p, out.size() * 8, cudaMemcpyDeviceToHost));
  if (first_value) {
    *first_value = out[1];
    for (int t = 1; t < n_teams; t++) /* every team evaluated the same vector: any difference is a bug */
      if (memcmp(&out[2 * t + 1], &out[1], 8) != 0) return PHB_EINVAL;
  }
  if (ms_out) *ms_out = ms;
  if (evals_per_s) *evals_per_s = (double)n_teams * reps / (ms * 1e-3);
  return PHB_OK;
}

int phb_kat_math(phb_ctx *c, int fn, const double *x, const double *y, int64_t n, double *out) {
  if (!c || !x || !out || n <= 0 || fn < 0 || fn > 15 || ((fn == 2 || fn == 3 || fn == 5 || fn == 13 || fn == 15) && !y)) return PHB_EINVAL;
  CK(cudaSetDevice(c->device));
  double *dx, *dy = nullptr, *dout;
  CK(cudaMalloc(&dx, n * 8)); CK(cudaMalloc(&dout, n * 8));
  CK(cudaMemcpy(dx, x, n * 8, cudaMemcpyHostToDevice));
  if (y) { CK(cudaMalloc(&dy, n * 8)); CK(cudaMemcpy(dy, y, n * 8, cudaMemcpyHostToDevice)); }
  kat_math_kernel<<<c->n_sm * 4, 256>>>(fn, dx, dy, (long long)n, dout, c->d_exp_tab, c->d_log_tab, c->d_pow_tab);
  CK(cudaGetLastError());
  CK(cudaMemcpy(out, dout, n * 8, cudaMemcpyDeviceToHost));
  cudaFree(dx); cudaFree(dout); if (dy) cudaFree(dy);
  return PHB_OK;
}

/* ---- REFINE ------------------------------------------------------------------------------------- */

int phb_refine_minmax_device(phb_ctx *c, const float *d_in, int64_t n, float nodata, float *h_minmax, void *stream) {
  if (!c || !d_in || !h_minmax || n <= 0) return PHB_EINVAL;
  CK(cudaSetDevice(c->device));
  cudaStream_t st = (cudaStream_t)stream;
  unsigned int init[2] = {0xffffffffu, 0u}; /* ordered-uint encodings of +max / -max */
  unsigned int *d_mm = reinterpret_cast<unsigned int *>(c->d_counters);
  CK(cudaMemcpyAsync(d_mm, init, sizeof(init), cudaMemcpyHostToDevice, st));
  refine_minmax_kernel<<<c->n_sm * 4, 256, 0, st>>>(d_in, (long long)n, nodata, d_mm);
  CK(cudaGetLastError());
  unsigned int mm[2];
  CK(cudaMemcpyAsync(mm, d_mm, sizeof(mm), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  h_minmax[0] = mm[0] == 0xffffffffu ? 1.0e10f : ordered_to_float(mm[0]);   /* BIG, common.c:1228 */
  h_minmax[1] = mm[1] == 0u ? -1.0e10f : ordered_to_float(mm[1]);
  return PHB_OK;
}

int phb_refine_device(phb_ctx *c, const float *d_in, float nodata, const float *d_land, float land_nodata,
                      const float *d_shallow, float shallow_nodata, int64_t n, int flags, const float *args,
                      const float *minmax, float *d_out, void *stream) {
  if (!c || !d_in || !d_out || !args || n <= 0) return PHB_EINVAL;
  if (!(flags & PHB_REFINE_CLIP) && !minmax) return PHB_EINVAL;
  CK(cudaSetDevice(c->device));
  RefineParams rp = make_refine_params(flags, args, minmax);
  rp.in = d_in; rp.land = d_land; rp.shallow = d_shallow; rp.out = d_out; rp.n = (long long)n;
  rp.nodata = nodata; rp.land_nodata = land_nodata; rp.shallow_nodata = shallow_nodata;
  rp.exp_tab = c->d_exp_tab; rp.log_tab = c->d_log_tab; rp.pow_tab = c->d_pow_tab;
  refine_kernel<<<c->n_sm * 8, 256, 0, (cudaStream_t)stream>>>(rp);
  CK(cudaGetLastError());
  return PHB_OK;
}

int phb_refine_host(phb_ctx *c, const float *h_in, float nodata, const float *h_land, float land_nodata,
                    const float *h_shallow, float shallow_nodata, int nrows, int ncols, int flags, const float *args,
                    float *h_out) {
  if (!c || !h_in || !h_out || !args || nrows < 1 || ncols < 1) return PHB_EINVAL;
  CK(cudaSetDevice(c->device));
  const size_t n = (size_t)nrows * ncols;
  CK(c->outs.ensure(4 * n));
  float *d_in = c->outs.p, *d_land = d_in + n, *d_sh = d_land + n, *d_out = d_sh + n;
  CK(cudaMemcpy(d_in, h_in, n * 4, cudaMemcpyHostToDevice));
  if (h_land) CK(cudaMemcpy(d_land, h_land, n * 4, cudaMemcpyHostToDevice));
  if (h_shallow) CK(cudaMemcpy(d_sh, h_shallow, n * 4, cudaMemcpyHostToDevice));
  float mm[2] = {0, 0};
  if (!(flags & PHB_REFINE_CLIP)) {
    int rc = phb_refine_minmax_device(c, d_in, (int64_t)n, nodata, mm, nullptr);
    if (rc) return rc;
  }
  int rc = phb_refine_device(c, d_in, nodata, h_land ? d_land : nullptr, land_nodata, h_shallow ? d_sh : nullptr,
                             shallow_nodata, (int64_t)n, flags, args, mm, d_out, nullptr);
  if (rc) return rc;
  CK(cudaMemcpy(h_out, d_out, n * 4, cudaMemcpyDeviceToHost));
  return PHB_OK;
}

/* ---- int16 scale/offset packing (model/nc.c:247-320) ------------------------------------------------------------ */

int phb_nc_pack_device(phb_ctx *c, const float *d_grid, int64_t n, double spval, int16_t *d_packed, float *add_offset,
                       float *scale_factor, int16_t *missing_value, void *stream) {
  if (!c || !d_grid || !d_packed || !add_offset || !scale_factor || n <= 0) return PHB_EINVAL;
  CK(cudaSetDevice(c->device));
  cudaStream_t st = (cudaStream_t)stream;
  const float fspv = (float)spval;
  const long long n_chunks_ll = (n + kPackChunk - 1) / kPackChunk;
  if (n_chunks_ll > 2147483647LL) return PHB_EINVAL;
  const int n_chunks = (int)n_chunks_ll;
  CK(c->outs.ensure((size_t)n_chunks + 4));
  float *chunk = c->outs.p, *d_min = c->outs.p + n_chunks;
  unsigned int *d_max = reinterpret_cast<unsigned int *>(c->outs.p + n_chunks + 1);
  const unsigned int init = float_to_ordered(1.175494351e-38f); /* FLT_MIN, nc.c:287 */
  CK(cudaMemcpyAsync(d_max, &init, sizeof(init), cudaMemcpyHostToDevice, st));
  const int blocks = n_chunks < c->n_sm * 8 ? n_chunks : c->n_sm * 8;
  nc_chunk_min_kernel<<<blocks, kPackThreads, 0, st>>>(d_grid, (long long)n, fspv, chunk, n_chunks);
  nc_chunk_scan_kernel<<<1, 1024, 0, st>>>(chunk, n_chunks, d_min);
  nc_chunk_max_kernel<<<blocks, kPackThreads, 0, st>>>(d_grid, (long long)n, fspv, chunk, n_chunks, d_max);
  CK(cudaGetLastError());
  float grmin;
  unsigned int mx;
  CK(cudaMemcpyAsync(&grmin, d_min, sizeof(float), cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(&mx, d_max, sizeof(mx), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  const float grmax = ordered_to_float(mx);
  volatile float span = grmax - grmin;           /* nc.c:306: float arithmetic */
  volatile float scale = span / 32767.0f;
  *add_offset = grmin; *scale_factor = scale;
  if (missing_value) *missing_value = (int16_t)-32768;
  nc_pack_kernel<<<c->n_sm * 8, 256, 0, st>>>(d_grid, (long long)n, fspv, grmin, scale, d_packed);
  CK(cudaGetLastError());
  return PHB_OK;
}

int phb_nc_unpack_device(phb_ctx *c, const int16_t *d_packed, int64_t n, float add_offset, float scale_factor,
                         int16_t missing_value, double spval, float *d_grid, void *stream) {
  if (!c || !d_packed || !d_grid || n <= 0) return PHB_EINVAL;
  CK(cudaSetDevice(c->device));
  nc_unpack_kernel<<<c->n_sm * 8, 256, 0, (cudaStream_t)stream>>>(d_packed, (long long)n, add_offset, scale_factor, missing_value,
                                                                  (float)spval, d_grid);
  CK(cudaGetLastError());
  return PHB_OK;
}

int phb_nc_pack_host(phb_ctx *c, const float *h_grid, int nrows, int ncols, double spval, int16_t *h_packed,
                     float *add_offset, float *scale_factor, int16_t *missing_value) {
  if (!c || !h_grid || !h_packed || nrows < 1 || ncols < 1) return PHB_EINVAL;
  CK(cudaSetDevice(c->device));
  const size_t n = (size_t)nrows * ncols;
  CK(c->planes.ensure(n + (n + 1) / 2));
  float *d_in = c->planes.p;
  int16_t *d_out = reinterpret_cast<int16_t *>(c->planes.p + n);
  CK(cudaMemcpy(d_in, h_grid, n * 4, cudaMemcpyHostToDevice));
  int rc = phb_nc_pack_device(c, d_in, (int64_t)n, spval, d_out, add_offset, scale_factor, missing_value, nullptr);
  if (rc) return rc;
  CK(cudaMemcpy(h_packed, d_out, n * 2, cudaMemcpyDeviceToHost));
  return PHB_OK;
}

int phb_nc_unpack_host(phb_ctx *c, const int16_t *h_packed, int nrows, int ncols, float add_offset, float scale_factor,
                       int16_t missing_value, double spval, float *h_grid) {
  if (!c || !h_packed || !h_grid || nrows < 1 || ncols < 1) return PHB_EINVAL;
  CK(cudaSetDevice(c->device));
  const size_t n = (size_t)nrows * ncols;
  CK(c->planes.ensure(n + (n + 1) / 2));
  float *d_out = c->planes.p;
  int16_t *d_in = reinterpret_cast<int16_t *>(c->planes.p + n);
  CK(cudaMemcpy(d_in, h_packed, n * 2, cudaMemcpyHostToDevice));
  int rc = phb_nc_unpack_device(c, d_in, (int64_t)n, add_offset, scale_factor, missing_value, spval, d_out, nullptr);
  if (rc) return rc;
  CK(cudaMemcpy(h_grid, d_out, n * 4, cudaMemcpyDeviceToHost));
  return PHB_OK;
}

/* ---- MODEL Lee_Kd_LS8 / Lee_Secchi_LS8 (secchi.c) --------------------------------------------------- */

int phb_lee_ls8_device(phb_ctx *c, int mode, const float *d_coastal, const float *d_blue, const float *d_green,
                       const float *d_red, const float *spv, float theta_s, int64_t n, float *d_out, void *stream) {
  if (!c || !d_coastal || !d_blue || !d_green || !d_red || !spv || !d_out || n <= 0 || (mode != 0 && mode != 1)) return PHB_EINVAL;
  CK(cudaSetDevice(c->device));
  LeeParams lp;
  lp.coastal = d_coastal; lp.blue = d_blue; lp.green = d_green; lp.red = d_red; lp.out = d_out; lp.n = (long long)n;
  for (int k = 0; k < 4; k++) lp.spv[k] = spv[k];
  lp.theta_s = theta_s; lp.mode = mode;
  lp.exp_tab = c->d_exp_tab; lp.log_tab = c->d_log_tab; lp.pow_tab = c->d_pow_tab;
  lee_ls8_kernel<<<c->n_sm * 8, 256, 0, (cudaStream_t)stream>>>(lp);
  CK(cudaGetLastError());
  return PHB_OK;
}

int phb_lee_ls8_host(phb_ctx *c, int mode, const float *h_coastal, const float *h_blue, const float *h_green,
                     const float *h_red, const float *spv, float theta_s, int nrows, int ncols, float *h_out) {
  if (!c || !h_coastal || !h_blue || !h_green || !h_red || !spv || !h_out || nrows < 1 || ncols < 1) return PHB_EINVAL;
  CK(cudaSetDevice(c->device));
  const size_t n = (size_t)nrows * ncols;
  CK(c->outs.ensure(5 * n));
  float *d = c->outs.p;
  const float *src[4] = {h_coastal, h_blue, h_green, h_red};
  for (int k = 0; k < 4; k++) CK(cudaMemcpy(d + k * n, src[k], n * 4, cudaMemcpyHostToDevice));
  int rc = phb_lee_ls8_device(c, mode, d, d + n, d + 2 * n, d + 3 * n, spv, theta_s, (int64_t)n, d + 4 * n, nullptr);
  if (rc) return rc;
  CK(cudaMemcpy(h_out, d + 4 * n, n * 4, cudaMemcpyDeviceToHost));
  return PHB_OK;
}

/* ---- FP64 peak ---------------------------------------------------------------------------------- */

int phb_fp64_peak(phb_ctx *c, double *tflops, float *ms) {
  if (!c || !tflops) return PHB_EINVAL;
  CK(cudaSetDevice(c->device));
  double *d_sink;
  CK(cudaMalloc(&d_sink, sizeof(double) * 1024));
  const int blocks = c->n_sm * 8, threads = 256, iters = 1 << 14;
  dfma_peak_kernel<<<blocks, threads>>>(d_sink, 64); /* warm-up */
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int rep = 0; rep < 5; rep++) {
    CK(cudaEventRecord(c->ev[0]));
    dfma_peak_kernel<<<blocks, threads>>>(d_sink, iters);
    CK(cudaEventRecord(c->ev[1]));
    CK(cudaEventSynchronize(c->ev[1]));
    float t;
    CK(cudaEventElapsedTime(&t, c->ev[0], c->ev[1]));
    if (t < best) best = t;
  }
  CK(cudaGetLastError());
  const double flops = 2.0 * kPeakChains * (double)iters * blocks * threads;
  *tflops = flops / (best * 1e-3) / 1e12;
  if (ms) *ms = best;
  cudaFree(d_sink);
  return PHB_OK;
}

/* ---- Jerlov water type and spectral attenuation (jerlov.c; host-side scene constants, see jerlov_host.h) ---- */

int phb_jerlov_fit(float wlen_i, float wlen_j, float lsm_i, float lsm_j, const float *Li, const float *Lj, int npoints,
                   float manual_ratio, float *out6, int32_t *n_shallow) {
  if (!out6 || npoints < 0 || (npoints > 0 && (!Li || !Lj))) return PHB_EINVAL;
  phb_jerlov::Fit f;
  memset(&f, 0, sizeof(f));
  const bool ok = phb_jerlov::fit(wlen_i, wlen_j, lsm_i, lsm_j, Li, Lj, npoints, manual_ratio, &f);
  out6[0] = f.ki; out6[1] = f.kj; out6[2] = f.m; out6[3] = f.c; out6[4] = f.r; out6[5] = f.water_type;
  if (n_shallow) *n_shallow = f.n_shallow;
  return ok ? PHB_OK : PHB_ENOFIT;
}

int phb_jerlov_k(float water_type, const float *wavelengths, int n, float *k) {
  if (n < 0 || (n > 0 && (!wavelengths || !k))) return PHB_EINVAL;
  int rc = PHB_OK;
  for (int i = 0; i < n; i++)
    if (!phb_jerlov::k_of_type(water_type, wavelengths[i], &k[i])) rc = PHB_EINVAL;
  return rc;
}

int phb_jerlov_k_from_ratio(float ratio, float wlen_i, float wlen_j, const float *wavelengths, int n, float *water_type,
                            float *k) {
  if (!water_type || n < 0 || (n > 0 && (!wavelengths || !k))) return PHB_EINVAL;
  *water_type = 0.0f;
  for (int i = 0; i < n; i++) k[i] = 0.0f;
  return phb_jerlov::k_from_ratio(ratio, wlen_i, wlen_j, wavelengths, n, water_type, k) ? PHB_OK : PHB_ENOFIT;
}

}  /* extern "C" */
