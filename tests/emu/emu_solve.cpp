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
ucontext_t g_main, g_lane[32];
bool g_done[32];
int g_cur = 0;
void (*g_kernel)(const phb::SolveParams) = nullptr;
const phb::SolveParams *g_params = nullptr;

void lane_entry() {
  g_kernel(*g_params);
  g_done[g_cur] = true; /* returns to g_main through uc_link */
}

/* runs one warp to completion: every pass advances each live lane to its next collective */
void run_warp() {
  std::vector<char> stacks(32 * kStack);
  for (int l = 0; l < 32; l++) {
    getcontext(&g_lane[l]);
    g_lane[l].uc_stack.ss_sp = stacks.data() + (size_t)l * kStack;
    g_lane[l].uc_stack.ss_size = kStack;
    g_lane[l].uc_link = &g_main;
    makecontext(&g_lane[l], lane_entry, 0);
    g_done[l] = false;
  }
  for (;;) {
    int live = 0;
    for (int l = 0; l < 32; l++) {
      if (g_done[l]) continue;
      live++;
      g_cur = l;
      threadIdx.x = (unsigned)l;
      swapcontext(&g_main, &g_lane[l]);
    }
    if (live == 0) break;
  }
}

const uint64_t kExpTab[2 * PHM_N] = PHM_EXP_TAB;
const double kLogTab[2 * PHM_N] = PHM_LOG_TAB;
const double kPowTab[3 * PHM_N] = PHM_POWLOG_TAB;

}  // namespace

void emu_warp_arrive() { swapcontext(&g_lane[g_cur], &g_main); }

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
 * outputs: full-precision records [n_queue][reclen] + pixel index + (evals, converged | iters << 1), the nine result
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
  const long long slab_doubles = simplex_doubles + sp.L.nmax + sp.L.Tmax + (long long)((sp.L.nmax + 8) / 8 + 1) * sp.L.nmax;
  long long cache = simplex_smem_bytes;
  if (cache > simplex_doubles * 8) cache = simplex_doubles * 8;
  add_simplex_cache(sp.L, (int)cache);
  sp.L.tmem_cols = 0;
  if ((size_t)sp.L.cta_bytes + (size_t)sp.L.warp_bytes > sizeof(phb_smem)) return 2;
  memset(phb_smem, 0xcd, sizeof(phb_smem)); /* uninitialised shared memory is not zero on the device either */
  std::vector<double> slab(slab_doubles, 0.0);
  int n_q = n_queue, head = 0;
  unsigned long long cnt[4] = {0, 0, 0, 0};
  double fl = 0.0;
  const size_t px = (size_t)M.nrows * M.ncols;
  sp.M = &M;
  sp.planes = planes; sp.prior = M.prior_present ? prior : nullptr;
  sp.queue = queue; sp.n_queue = &n_q; sp.head = &head;
  sp.slabs = slab.data(); sp.slab_stride = slab_doubles;
  if (out9) {
    sp.out.depth = out9; sp.out.model_error = out9 + px; sp.out.bottom_albedo = out9 + 2 * px;
    sp.out.bottom_sand = out9 + 3 * px; sp.out.bottom_seagrass = out9 + 4 * px; sp.out.bottom_coral = out9 + 5 * px;
    sp.out.K_min = out9 + 6 * px; sp.out.bottom_type = out9 + 7 * px; sp.out.index_optical_depth = out9 + 8 * px;
  }
  sp.out.converged = converged; sp.out.n_evals = n_evals;
  sp.dbg_rec = rec; sp.dbg_pix = pix; sp.dbg_iters = iters; sp.reclen = emu_record_len(model); sp.dbg_capacity = n_queue;
  sp.counters = cnt; sp.flops = &fl;
  sp.exp_tab = reinterpret_cast<const unsigned long long *>(kExpTab); sp.log_tab = kLogTab; sp.pow_tab = kPowTab;
  if (sp.L.SBP == 32) g_kernel = M.n_bottoms == 3 ? solve_kernel<3, 32, false> : solve_kernel<0, 32, false>;
  else g_kernel = M.n_bottoms == 3 ? solve_kernel<3, kMaxSB, false> : solve_kernel<0, kMaxSB, false>;
  g_params = &sp;
  run_warp();
  if (counters) memcpy(counters, cnt, sizeof(cnt));
  if (flops) *flops = fl;
  return 0;
}

}  // extern "C"
