/*
 * tests/emu/include/cuda_runtime.h -- TEST INFRASTRUCTURE ONLY.
 *
 * A stand-in for the CUDA headers that lets g++ compile photic_b200/csrc/invert_kernel.cuh (with -DPHB_HOST_EMU) so that
 * solve_kernel -- the real kernel source, state machine, storage tiers, ordered sums and all -- can be run on the CPU,
 * one warp at a time, and compared bit for bit with the oracle (tests/test_kernel_emulation.py). It is not a CUDA
 * emulator: it provides exactly what this one kernel uses.
 *
 * Execution model: the 32 lanes of a warp are 32 fibers (ucontext) on one OS thread. A lane runs until it reaches a
 * warp collective (__syncwarp, shuffles, votes, reductions); the scheduler (emu_solve.cpp) then runs the next lane,
 * and resumes everybody once all 32 have arrived -- which is the guarantee the kernel relies on: collectives are
 * only issued from warp-uniform control flow, and lanes communicate through shared memory only across a __syncwarp.
 * Floating point: every __fma_rn / __dmul_rn / ... is one IEEE binary64 operation here as on the device (compile
 * with -mfma -ffp-contract=off).
 */
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __restrict__
#define __shared__
#define __constant__
#define __align__(n) __attribute__((aligned(n)))

struct double2 { double x, y; };
struct emu_dim3 { unsigned x, y, z; };
extern emu_dim3 threadIdx, blockIdx, blockDim, gridDim; /* threadIdx is switched by the scheduler with the running lane */

/* ---- scheduler interface (emu_solve.cpp) ---- */
void emu_warp_arrive();              /* yield until all 32 lanes have arrived */
extern uint64_t emu_slot[32];        /* per-lane exchange slots of the collectives */
static inline int emu_lane() { return (int)(threadIdx.x & 31u); }

static inline void __syncwarp(unsigned = 0xffffffffu) { emu_warp_arrive(); }
static inline void __syncthreads() { emu_warp_arrive(); } /* one warp per CTA in the emulation */

static inline uint64_t emu_exchange(uint64_t v, int src) {
  emu_slot[emu_lane()] = v;
  emu_warp_arrive();
  const uint64_t r = emu_slot[src & 31];
  emu_warp_arrive();
  return r;
}
static inline int __shfl_sync(unsigned, int v, int src) { return (int)(uint32_t)emu_exchange((uint32_t)v, src); }
static inline unsigned __shfl_sync(unsigned, unsigned v, int src) { return (unsigned)emu_exchange(v, src); }
static inline double __shfl_sync(unsigned, double v, int src) {
  uint64_t u; memcpy(&u, &v, 8); u = emu_exchange(u, src); double r; memcpy(&r, &u, 8); return r;
}
static inline double __shfl_xor_sync(unsigned m, double v, int mask) { return __shfl_sync(m, v, emu_lane() ^ mask); }
static inline int __shfl_xor_sync(unsigned m, int v, int mask) { return __shfl_sync(m, v, emu_lane() ^ mask); }

template <class F> static inline uint64_t emu_reduce(uint64_t v, F f) {
  emu_slot[emu_lane()] = v;
  emu_warp_arrive();
  uint64_t r = emu_slot[0];
  for (int i = 1; i < 32; i++) r = f(r, emu_slot[i]);
  emu_warp_arrive();
  return r;
}
static inline unsigned __ballot_sync(unsigned, bool p) {
  return (unsigned)emu_reduce(p ? (1ull << emu_lane()) : 0ull, [](uint64_t a, uint64_t b) { return a | b; });
}
static inline bool __any_sync(unsigned m, bool p) { return __ballot_sync(m, p) != 0u; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline unsigned __reduce_max_sync(unsigned, unsigned v) {
  return (unsigned)emu_reduce(v, [](uint64_t a, uint64_t b) { return a > b ? a : b; });
}
static inline unsigned __reduce_min_sync(unsigned, unsigned v) {
  return (unsigned)emu_reduce(v, [](uint64_t a, uint64_t b) { return a < b ? a : b; });
}
static inline int __reduce_add_sync(unsigned, int v) {
  return (int)(uint32_t)emu_reduce((uint32_t)v, [](uint64_t a, uint64_t b) { return (uint64_t)(uint32_t)(a + b); });
}

/* ---- memory ---- */
template <class T> static inline T __ldg(const T *p) { return *p; }
template <class T> static inline T __ldcg(const T *p) { return *p; }
static inline int atomicAdd(int *p, int v) { const int o = *p; *p = o + v; return o; } /* lanes never run concurrently */
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { const unsigned long long o = *p; *p = o + v; return o; }
static inline double atomicAdd(double *p, double v) { const double o = *p; *p = o + v; return o; }

/* ---- arithmetic: one IEEE operation each ---- */
static inline double __fma_rn(double a, double b, double c) { return __builtin_fma(a, b, c); }
static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
static inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
static inline double __dsub_rn(double a, double b) { volatile double r = a - b; return r; }
static inline int __double2hiint(double v) { uint64_t u; memcpy(&u, &v, 8); return (int)(uint32_t)(u >> 32); }
static inline int __double2loint(double v) { uint64_t u; memcpy(&u, &v, 8); return (int)(uint32_t)u; }
static inline double __hiloint2double(int hi, int lo) {
  const uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo; double v; memcpy(&v, &u, 8); return v;
}
static inline long long __double_as_longlong(double v) { long long u; memcpy(&u, &v, 8); return u; }
static inline double __longlong_as_double(long long u) { double v; memcpy(&v, &u, 8); return v; }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
static inline long long __double2ll_rz(double v) { return (long long)v; } /* callers range-check first (to_long_x86) */
