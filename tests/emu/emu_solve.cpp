/*
 * emu_solve.cpp -- TEST INFRASTRUCTURE ONLY: runs the product's solve_kernel SOURCE (photic_b200/csrc/invert_kernel.cuh,
 * compiled by g++ with -DPHB_HOST_EMU against tests/emu/include/cuda_runtime.h) on the CPU, one warp of 32 fibers, so
 * that the kernel's logic -- optimiser state machine, simplex storage tiers and centroid checkpoints, ordered sums,
 * penalties, neighbourhood gather, result derivation -- is checked against the oracle without a GPU
 * (tests/test_kernel_emulation.py). What it does NOT cover is what only the device has: the PTX fast paths of
 * division / square root (pinned by the device known-answer tests), tensor memory, and nvcc's code generation.
 *
 *   g++ -O2 -std=c++17 -mfma -ffp-contract=off -shared -fPIC -DPHB_HOST_EMU -Itests/emu/include emu_solve.cpp
 */
#include <ucontext.h>

#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "cuda_runtime.h"
#include "../../photic_b200/csrc/invert_kernel.cuh"

emu_dim3 threadIdx = {0, 0, 0}, blockIdx = {0, 0, 0}, blockDim = {32, 1, 1}, gridDim = {1, 1, 1};
uint64_t emu_slot[32];
namespace phb { __attribute__((aligned(16))) unsigned char phb_smem[256 * 1024]; }

namespace {

constexpr size_t kStack = 1u << 20;
bool g_done[32];
int g_cur = 0;
void (*g_kernel)(const phb::SolveParams) = nullptr;
const phb::SolveParams *g_params = nullptr;
void (*g_body)() = nullptr; /* what every lane runs */

/* Fiber switch. glibc's swapcontext makes a sigprocmask system call per switch (two thirds of the emulation's run
 * time); on x86-64 a dozen instructions do: callee-saved registers, MXCSR and the x87 control word, stack pointer. */
#if defined(__x86_64__) && !defined(EMU_UCONTEXT)
extern "C" void emu_switch(void **save_sp, void *load_sp);
asm(R"(
.text
.globl emu_switch
.type emu_switch,@function
emu_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    subq $8, %rsp
    stmxcsr (%rsp)
    fnstcw 4(%rsp)
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    ldmxcsr (%rsp)
    fldcw 4(%rsp)
    addq $8, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size emu_switch,.-emu_switch
)");
void *g_main_sp = nullptr, *g_lane_sp[32];
void yield_to_main() { emu_switch(&g_lane_sp[g_cur], g_main_sp); }
void resume_lane(int l) { emu_switch(&g_main_sp, g_lane_sp[l]); }
extern "C" void emu_lane_trampoline() {
  g_body();
  g_done[g_cur] = true;
  yield_to_main();
  abort(); /* a finished lane is never resumed */
}
void make_lane(int l, char *stack) {
  /* frame emu_switch pops: [mxcsr | x87 cw][r15 r14 r13 r12 rbx rbp][return address]; entry with rsp % 16 == 8 */
  uintptr_t top = ((uintptr_t)stack + kStack) & ~(uintptr_t)15;
  uint64_t *sp = reinterpret_cast<uint64_t *>(top) - 9;
  uint32_t csr; uint16_t cw;
  asm volatile("stmxcsr %0" : "=m"(csr));
  asm volatile("fnstcw %0" : "=m"(cw));
  sp[0] = (uint64_t)csr | ((uint64_t)cw << 32);
  for (int k = 1; k <= 6; k++) sp[k] = 0;
  sp[7] = (uint64_t)(uintptr_t)&emu_lane_trampoline;
  sp[8] = 0;
  g_lane_sp[l] = sp;
}
#else
ucontext_t g_main, g_lane[32];
void yield_to_main() { swapcontext(&g_lane[g_cur], &g_main); }
void resume_lane(int l) { swapcontext(&g_main, &g_lane[l]); }
void lane_entry() {
  g_body();
  g_done[g_cur] = true; /* returns to g_main through uc_link */
}
void make_lane(int l, char *stack) {
  getcontext(&g_lane[l]);
  g_lane[l].uc_stack.ss_sp = stack;
  g_lane[l].uc_stack.ss_size = kStack;
  g_lane[l].uc_link = &g_main;
  makecontext(&g_lane[l], lane_entry, 0);
}
#endif

void body_solve() { g_kernel(*g_params); }

/* runs one warp to completion: every pass advances each live lane to its next collective */
void run_warp() {
  std::vector<char> stacks(32 * kStack + 64);
  for (int l = 0; l < 32; l++) {
    make_lane(l, stacks.data() + (size_t)l * kStack);
    g_done[l] = false;
  }
  for (;;) {
    int live = 0;
    for (int l = 0; l < 32; l++) {
      if (g_done[l]) continue;
      live++;
      g_cur = l;
      threadIdx.x = (unsigned)l;
      resume_lane(l);
    }
    if (live == 0) break;
  }
}

/* known answers of the objective: the body of kat_objective_kernel (csrc/aux_kernels.cuh) on the emulated warp */
struct KatArgs { int nb_active, n_regions, origin, nvec; const double *meas, *params; double *out6; } g_kat;
template <int SBP>
void kat_objective_body() {
  using namespace phb;
  const SolveParams &p = *g_params;
  const ModelConst &M = *p.M;
  stage_cta(p, M, phb_smem);
  __syncthreads();
  Warp w;
  const int lane = threadIdx.x & 31, SB = p.L.SB, Ns = p.L.Ns;
  bind_warp<0, 0>(w, p, p.L, phb_smem, 0, 0);
  Pixel px;
  px.Nr = g_kat.n_regions; px.Nb = g_kat.nb_active; px.origin = g_kat.origin;
  size_pixel(px, lane, SB, Ns, p.L.simplex_doubles, 0);
  for (int t = lane; t < px.T; t += 32) w.meas[t] = g_kat.meas[t];
  __syncwarp();
  phm::Tables tb;
  tb.exp_tab = w.exp_tab; tb.log_tab = p.log_tab; tb.pow_tab = p.pow_tab;
  double Bs, Ps, Xs;
  derive_pixel_constants(w, px, M, tb, lane, SB, Ns, Bs, Ps, Xs);
  Side side;
  for (int v = 0; v < g_kat.nvec; v++) {
    for (int i = lane; i < px.n; i += 32) w.xmin[i] = g_kat.params[(size_t)v * px.n + i];
    __syncwarp();
    const bool generic = getenv("PHB_ONE_CLASS") != nullptr;
    double e; /* as kat_objective_kernel: the compile-time classes for 3 and 1 substrates */
    if (!generic && px.Nb == 3 && px.Nr <= 16) e = objective<3, SBP, true>(w, px, lane, SB, Ns, p.L.NbMax, w.xmin, side);
    else if (!generic && px.Nb == 1 && px.Nr <= 16) e = objective<1, SBP, true>(w, px, lane, SB, Ns, p.L.NbMax, w.xmin, side);
    else e = objective<0, SBP, true>(w, px, lane, SB, Ns, p.L.NbMax, w.xmin, side);
    if (lane == 0) {
      double *o = g_kat.out6 + (size_t)v * 6;
      o[0] = e; o[1] = side.e_rrs; o[2] = side.e_depth; o[3] = side.e_bottom; o[4] = side.e_K; o[5] = side.bottom_albedo;
    }
    __syncwarp();
  }
}

/* wires one band (view 0) and its work queues into sp the way the host library does (photic_b200.cu: bind_queues):
 * NBOTTOMS = 3 -> two pixel classes, queue[0] = pixels with all substrates, queue[1] = sand-only pixels (h_prior > 8 m) */
struct EmuBand {
  phb::BandView view;
  std::vector<int> q0, q1;
  int scal[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  bool two = false;
  void wire(const phb::ModelConst &M, const float *planes, const float *prior, const int *queue, int n_queue,
            const phb_outputs &out) {
    using namespace phb;
    memset(&view, 0, sizeof(view));
    view.planes = planes; view.prior = M.prior_present ? prior : nullptr; view.nrows = M.nrows; view.out = out;
    const int nsp_ = M.n_spatial == 0 ? 1 : M.n_spatial;
    two = M.n_bottoms == 3 && (2 * nsp_ - 1) * (2 * nsp_ - 1) <= 16 && getenv("PHB_ONE_CLASS") == nullptr; /* photic_b200.cu: use_two_classes */
    q0.clear(); q1.clear();
    for (int k = 0; k < n_queue; k++) {
      bool deep = false;
      if (two && view.prior != nullptr) {
        const float e = view.prior[queue[k]];
        if (!approx_equal_f(e, M.prior_nodata, 1.0e-6f)) deep = ((e > -1.0) ? 1.0 : fabs((double)e)) > 8.0;
      }
      (deep ? q1 : q0).push_back(queue[k]);
    }
    q0.push_back(0); q1.push_back(0); /* never empty vectors */
    scal[0] = (int)q0.size() - 1; scal[1] = (int)q1.size() - 1; scal[2] = n_queue;
    view.queue[0] = q0.data(); view.n_queue[0] = &scal[0]; view.head[0] = &scal[3];
    view.queue[1] = q1.data(); view.n_queue[1] = &scal[1]; view.head[1] = &scal[4];
  }
  void bind(phb::SolveParams &sp, const phb::ModelConst &M, int NrMax, int simplex_smem_bytes) {
    using namespace phb;
    sp.views = &view; sp.n_views = 1; sp.n_classes = two ? 2 : 1;
    if (two) { /* as launch_solve: same CTA block and warp stride, the left-over of the warp block holds simplex rows */
      sp.L1 = make_layout(M.SB, M.n_scenes, 1, NrMax);
      sp.L1.cta_bytes = sp.L.cta_bytes;
      long long cache1 = (long long)sp.L.warp_bytes - sp.L1.w_simplex;
      const long long simplex1 = (long long)(sp.L1.nmax + 1) * sp.L1.nmax * 8;
      if (cache1 > simplex1) cache1 = simplex1;
      if (cache1 > simplex_smem_bytes) cache1 = simplex_smem_bytes;
      if (cache1 < 0) cache1 = 0;
      add_simplex_cache(sp.L1, (int)cache1);
      sp.L1.warp_bytes = sp.L.warp_bytes;
      sp.L1.tmem_cols = 0;
    } else {
      sp.L1 = sp.L;
    }
  }
};

const uint64_t kExpTab[2 * PHM_N] = PHM_EXP_TAB;
const double kLogTab[2 * PHM_N] = PHM_LOG_TAB;
const double kPowTab[3 * PHM_N] = PHM_POWLOG_TAB;

}  // namespace

void emu_warp_arrive() { yield_to_main(); }

/* the instantiation launch_solve (csrc/photic_b200.cu) picks: compile-time layout for the Landsat-8 configurations */
static void (*pick_solve_kernel(const phb::SolveParams &sp, const phb::ModelConst &M, bool two))(const phb::SolveParams) {
  using namespace phb;
  if (sp.L.SBP != 32) return two ? solve_kernel<3, kMaxSB, false> : solve_kernel<0, kMaxSB, false>;
  bool four = true;
  for (int s = 0; s < M.n_scenes; s++) four = four && M.n_bands[s] == 4;
  const char *e = getenv("PHB_CT_LAYOUT");
  if (two && four && sp.L.NrMax == 9 && !(e && atoi(e) == 0)) {
    if (M.n_scenes == 4) return solve_kernel<3, 32, false, 4>;
    if (M.n_scenes == 6) return solve_kernel<3, 32, false, 6>;
    if (M.n_scenes == 8) return solve_kernel<3, 32, false, 8>;
  }
  return two ? solve_kernel<3, 32, false> : solve_kernel<0, 32, false>;
}

/* guard words behind the (exactly sized) slab: a write past the depth the host gives the slabs is caught here */
static const int kSlabGuard = 64;
static const double kGuardValue = -7.25e77;
static bool slab_guard_intact(const std::vector<double> &slab, long long used) {
  for (int g = 0; g < kSlabGuard; g++)
    if (slab[used + g] != kGuardValue) return false;
  return true;
}

extern "C" {

int64_t emu_model_const_size(void) { return (int64_t)sizeof(phb::ModelConst); }
int emu_record_len(const void *model) {
  const phb::ModelConst &M = *static_cast<const phb::ModelConst *>(model);
  return phb::kRecHead + M.n_scenes * M.max_bands + 3 * M.n_scenes;
}

/*
 * One warp inverts the queued pixels with the product kernel. model: ModelConst bytes (phb_debug_model_const);
 * planes [SB][nrows][ncols], prior [nrows][ncols] or null; queue: linear pixel indices; simplex_smem_bytes: shared
 * memory given to the warp for simplex rows (the rest of the simplex lives in the "global" slab, checkpoints and all);
 * outputs: full-precision records [n_queue][reclen] + pixel index + (evals, converged | iters << 1, nelmin restarts), the nine result
 * planes (out9, [9][nrows][ncols], nullable), converged / n_evals planes (nullable), counters[4], flops.
 */
int emu_invert(const void *model, int64_t model_size, const float *planes, const float *prior, const int *queue,
               int n_queue, int simplex_smem_bytes, double *rec, int *pix, int *iters, float *out9,
               unsigned char *converged, int *n_evals, unsigned long long *counters, double *flops) {
  using namespace phb;
  if (model_size != (int64_t)sizeof(ModelConst)) return 1;
  const ModelConst &M = *static_cast<const ModelConst *>(model);
  const int nsp = M.n_spatial == 0 ? 1 : M.n_spatial;
  const int NrMax = (2 * nsp - 1) * (2 * nsp - 1);
  SolveParams sp;
  memset(&sp, 0, sizeof(sp));
  sp.L = make_layout(M.SB, M.n_scenes, M.n_bottoms, NrMax);
  const long long simplex_doubles = (long long)(sp.L.nmax + 1) * sp.L.nmax;
  long long cache = simplex_smem_bytes;
  if (cache > simplex_doubles * 8) cache = simplex_doubles * 8;
  add_simplex_cache(sp.L, (int)cache);
  sp.L.tmem_cols = 0;
  /* the slab exactly as deep as the host library makes it (launch_solve), followed by guard words */
  int slab_rows = max_global_rows(sp.L, NrMax, M.n_scenes, M.n_bottoms);
  { const int r1 = max_global_rows(sp.L, NrMax, M.n_scenes, 1); if (r1 > slab_rows) slab_rows = r1; }
  if (slab_rows < 1) slab_rows = 1;
  const long long slab_doubles = slab_doubles_for(sp.L, slab_rows);
  sp.slab_rows = slab_rows;
  if ((size_t)sp.L.cta_bytes + (size_t)sp.L.warp_bytes > sizeof(phb_smem)) return 2;
  memset(phb_smem, 0xcd, sizeof(phb_smem)); /* uninitialised shared memory is not zero on the device either */
  std::vector<double> slab(slab_doubles + kSlabGuard, 0.0);
  for (int g = 0; g < kSlabGuard; g++) slab[slab_doubles + g] = kGuardValue;
  unsigned long long cnt[4] = {0, 0, 0, 0};
  double fl = 0.0;
  const size_t px = (size_t)M.nrows * M.ncols;
  sp.M = &M;
  sp.slabs = slab.data(); sp.slab_stride = slab_doubles;
  phb_outputs o;
  memset(&o, 0, sizeof(o));
  if (out9) {
    o.depth = out9; o.model_error = out9 + px; o.bottom_albedo = out9 + 2 * px;
    o.bottom_sand = out9 + 3 * px; o.bottom_seagrass = out9 + 4 * px; o.bottom_coral = out9 + 5 * px;
    o.K_min = out9 + 6 * px; o.bottom_type = out9 + 7 * px; o.index_optical_depth = out9 + 8 * px;
  }
  o.converged = converged; o.n_evals = n_evals;
  EmuBand band;
  band.wire(M, planes, prior, queue, n_queue, o);
  band.bind(sp, M, NrMax, simplex_smem_bytes);
  sp.dbg_rec = rec; sp.dbg_pix = pix; sp.dbg_iters = iters; sp.reclen = emu_record_len(model); sp.dbg_capacity = n_queue;
  sp.counters = cnt; sp.flops = &fl;
  sp.exp_tab = reinterpret_cast<const unsigned long long *>(kExpTab); sp.log_tab = kLogTab; sp.pow_tab = kPowTab;
  g_kernel = pick_solve_kernel(sp, M, band.two);
  g_params = &sp;
  g_body = body_solve;
  run_warp();
  if (counters) memcpy(counters, cnt, sizeof(cnt));
  if (flops) *flops = fl;
  if (!slab_guard_intact(slab, slab_doubles)) return 7; /* the kernel wrote past the slab depth */
  return 0;
}

/*
 * The whole device side of phb_invert_device on a small raster: classify_kernel (validity, defaults, the two
 * work lists), concat_queue_kernel, solve_kernel. The two pre-pass kernels have no warp collectives and run as one
 * thread; rows [row_begin, row_end). out: the nine planes [9][nrows][ncols]; K [Ns][max_bands][px], P/G/X [Ns][px]
 * (nullable); records for the queued pixels in queue order (capacity = the row window's pixel count).
 */
int emu_invert_raster(const void *model, int64_t model_size, const float *planes, const float *prior, int row_begin,
                      int row_end, int simplex_smem_bytes, float *out9, float *K, float *P, float *G, float *X,
                      unsigned char *converged, int *n_evals, double *rec, int *pix, int *iters, int *n_valid,
                      int *n_shallow) {
  using namespace phb;
  if (model_size != (int64_t)sizeof(ModelConst)) return 1;
  const ModelConst &M = *static_cast<const ModelConst *>(model);
  const size_t px = (size_t)M.nrows * M.ncols, win = (size_t)(row_end - row_begin) * M.ncols;
  std::vector<int> q_shallow(win), q_deep(win), queue(win);
  int scal[4] = {0, 0, 0, 0};
  phb_outputs o;
  memset(&o, 0, sizeof(o));
  o.depth = out9; o.model_error = out9 + px; o.bottom_albedo = out9 + 2 * px; o.bottom_sand = out9 + 3 * px;
  o.bottom_seagrass = out9 + 4 * px; o.bottom_coral = out9 + 5 * px; o.K_min = out9 + 6 * px; o.bottom_type = out9 + 7 * px;
  o.index_optical_depth = out9 + 8 * px;
  o.K = K; o.P = P; o.G = G; o.X = X; o.converged = converged; o.n_evals = n_evals;
  ClassifyParams cp;
  cp.M = &M; cp.planes = planes; cp.prior = M.prior_present ? prior : nullptr;
  cp.row_begin = row_begin; cp.row_end = row_end;
  cp.queue_shallow = q_shallow.data(); cp.queue_deep = q_deep.data(); cp.n_shallow = &scal[0]; cp.n_deep = &scal[1];
  cp.out = o;
  threadIdx.x = 0; blockDim.x = 1; gridDim.x = 1; blockIdx.x = 0;
  classify_kernel(cp);
  concat_queue_kernel(q_shallow.data(), q_deep.data(), &scal[0], &scal[1], queue.data(), &scal[2]);
  blockDim.x = 32;
  if (n_valid) *n_valid = scal[2];
  if (n_shallow) *n_shallow = scal[0];
  if (scal[2] == 0) return 0;
  /* solve: as emu_invert, with the planes wired */
  const int nsp = M.n_spatial == 0 ? 1 : M.n_spatial;
  SolveParams sp;
  memset(&sp, 0, sizeof(sp));
  sp.L = make_layout(M.SB, M.n_scenes, M.n_bottoms, (2 * nsp - 1) * (2 * nsp - 1));
  const long long simplex_doubles = (long long)(sp.L.nmax + 1) * sp.L.nmax;
  long long cache = simplex_smem_bytes;
  if (cache > simplex_doubles * 8) cache = simplex_doubles * 8;
  add_simplex_cache(sp.L, (int)cache);
  sp.L.tmem_cols = 0;
  const int NrMax2 = (2 * nsp - 1) * (2 * nsp - 1);
  int slab_rows = max_global_rows(sp.L, NrMax2, M.n_scenes, M.n_bottoms);
  { const int r1 = max_global_rows(sp.L, NrMax2, M.n_scenes, 1); if (r1 > slab_rows) slab_rows = r1; }
  if (slab_rows < 1) slab_rows = 1;
  const long long slab_doubles = slab_doubles_for(sp.L, slab_rows);
  sp.slab_rows = slab_rows;
  if ((size_t)sp.L.cta_bytes + (size_t)sp.L.warp_bytes > sizeof(phb_smem)) return 2;
  memset(phb_smem, 0xcd, sizeof(phb_smem));
  std::vector<double> slab(slab_doubles + kSlabGuard, 0.0);
  for (int g = 0; g < kSlabGuard; g++) slab[slab_doubles + g] = kGuardValue;
  unsigned long long cnt[4] = {0, 0, 0, 0};
  double fl = 0.0;
  sp.M = &M;
  sp.slabs = slab.data(); sp.slab_stride = slab_doubles;
  EmuBand band; /* the device's two lists as classify_kernel left them (their concatenation for the generic kernel) */
  memset(&band.view, 0, sizeof(band.view));
  band.view.planes = planes; band.view.prior = cp.prior; band.view.nrows = M.nrows; band.view.out = o;
  band.two = M.n_bottoms == 3 && (2 * nsp - 1) * (2 * nsp - 1) <= 16 && getenv("PHB_ONE_CLASS") == nullptr;
  if (band.two) {
    band.view.queue[0] = q_shallow.data(); band.view.n_queue[0] = &scal[0]; band.view.head[0] = &band.scal[3];
    band.view.queue[1] = q_deep.data(); band.view.n_queue[1] = &scal[1]; band.view.head[1] = &band.scal[4];
  } else {
    band.view.queue[0] = queue.data(); band.view.n_queue[0] = &scal[2]; band.view.head[0] = &band.scal[3];
    band.view.queue[1] = queue.data(); band.view.n_queue[1] = &band.scal[5]; band.view.head[1] = &band.scal[4];
  }
  band.bind(sp, M, (2 * nsp - 1) * (2 * nsp - 1), simplex_smem_bytes);
  sp.dbg_rec = rec; sp.dbg_pix = pix; sp.dbg_iters = iters; sp.reclen = emu_record_len(model); sp.dbg_capacity = (long long)win;
  sp.counters = cnt; sp.flops = &fl;
  sp.exp_tab = reinterpret_cast<const unsigned long long *>(kExpTab); sp.log_tab = kLogTab; sp.pow_tab = kPowTab;
  g_kernel = pick_solve_kernel(sp, M, band.two);
  g_params = &sp;
  g_body = body_solve;
  run_warp();
  if (!slab_guard_intact(slab, slab_doubles)) return 7; /* the kernel wrote past the slab depth */
  return 0;
}

/* samodel_error on caller-supplied parameter vectors (what phb_kat_objective does on the device): out6 per vector =
 * objective, Rrs error, depth / bottom / K penalties, bottom albedo */
int emu_kat_objective(const void *model, int64_t model_size, int nb_active, int n_regions, int origin, const double *meas,
                      int nvec, const double *params, double *out6) {
  using namespace phb;
  if (model_size != (int64_t)sizeof(ModelConst)) return 1;
  const ModelConst &M = *static_cast<const ModelConst *>(model);
  SolveParams sp;
  memset(&sp, 0, sizeof(sp));
  sp.L = make_layout(M.SB, M.n_scenes, M.n_bottoms, n_regions);
  if ((size_t)sp.L.cta_bytes + (size_t)sp.L.warp_bytes > sizeof(phb_smem)) return 2;
  memset(phb_smem, 0xcd, sizeof(phb_smem));
  const long long slab_doubles = (long long)(sp.L.nmax + 1) * sp.L.nmax + sp.L.nmax + sp.L.Tmax + (long long)((sp.L.nmax + 8) / 8 + 1) * sp.L.nmax;
  std::vector<double> slab(slab_doubles, 0.0);
  sp.M = &M;
  sp.slabs = slab.data(); sp.slab_stride = slab_doubles;
  sp.exp_tab = reinterpret_cast<const unsigned long long *>(kExpTab); sp.log_tab = kLogTab; sp.pow_tab = kPowTab;
  g_kat = KatArgs{nb_active, n_regions, origin, nvec, meas, params, out6};
  g_params = &sp;
  g_body = sp.L.SBP == 32 ? kat_objective_body<32> : kat_objective_body<kMaxSB>;
  run_warp();
  return 0;
}

}  // extern "C"
