// Minimal SIMT emulation of the CUDA constructs used by specloss_kernels.cuh, for the CPU
// test-suite only (tests/emu/specloss_emu.cpp).  One emulated warp = 32 host threads in lock
// step: __syncwarp() is a barrier, __shfl_xor_sync() a barrier-protected exchange.
#pragma once
#include <algorithm>
#include <barrier>
#include <cmath>
#include <cstddef>
#include <cstdlib>

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __restrict__
#define __align__(n)

struct float2 { float x, y; };
struct int4 { int x, y, z, w; };
struct int2 { int x, y; };
struct alignas(16) float4 { float x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline float2 __fadd2_rn(float2 a, float2 b) { return float2{a.x + b.x, a.y + b.y}; }
static inline float2 __fmul2_rn(float2 a, float2 b) { return float2{a.x * b.x, a.y * b.y}; }
static inline float2 __ffma2_rn(float2 a, float2 b, float2 c) { return float2{std::fmaf(a.x, b.x, c.x), std::fmaf(a.y, b.y, c.y)}; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float2 make_float2(float x, float y) { return float2{x, y}; }

struct EmuWarp {
  std::barrier<> bar{32};
  float slots[32];
  double dslots[32];
  unsigned votes[32];
};
extern thread_local EmuWarp* emu_warp;
extern thread_local int emu_lane;

// Fault injection for the race-detector's positive control (tests/emu/tsan_race_check.py): with
// SPECLOSS_EMU_SKIP_SYNC=n every lane skips its n-th __syncwarp() of the process (all 32 lanes skip the same one, so the
// barrier count stays balanced); ThreadSanitizer must then report the shared-memory race that barrier was there to prevent.
static inline bool emu_skip_this_sync() {
  static const long skip = [] { const char* e = std::getenv("SPECLOSS_EMU_SKIP_SYNC"); return e ? std::atol(e) : 0L; }();
  if (skip <= 0) return false;
  static thread_local long count = 0;
  return ++count == skip;
}
static inline void __syncwarp() {
  if (emu_skip_this_sync()) return;
  emu_warp->bar.arrive_and_wait();
}
static inline void __syncthreads() { emu_warp->bar.arrive_and_wait(); }   // emulated CTAs are one warp wide
static inline float __shfl_xor_sync(unsigned, float v, int lane_mask) {
  emu_warp->slots[emu_lane] = v;
  emu_warp->bar.arrive_and_wait();
  const float r = emu_warp->slots[emu_lane ^ lane_mask];
  emu_warp->bar.arrive_and_wait();
  return r;
}
static inline double __shfl_xor_sync(unsigned, double v, int lane_mask) {
  emu_warp->dslots[emu_lane] = v;
  emu_warp->bar.arrive_and_wait();
  const double r = emu_warp->dslots[emu_lane ^ lane_mask];
  emu_warp->bar.arrive_and_wait();
  return r;
}
static inline unsigned __shfl_xor_sync(unsigned, unsigned v, int lane_mask) {
  emu_warp->votes[emu_lane] = v;
  emu_warp->bar.arrive_and_wait();
  const unsigned r = emu_warp->votes[emu_lane ^ lane_mask];
  emu_warp->bar.arrive_and_wait();
  return r;
}
static inline unsigned __reduce_max_sync(unsigned mask, unsigned v) {      // over the lanes named in mask (sub-warp groups)
  emu_warp->votes[emu_lane] = v;
  emu_warp->bar.arrive_and_wait();
  unsigned r = 0;
  for (int i = 0; i < 32; ++i) if ((mask >> i) & 1u) r = std::max(r, emu_warp->votes[i]);
  emu_warp->bar.arrive_and_wait();
  return r;
}
static inline unsigned __reduce_add_sync(unsigned mask, unsigned v) {
  emu_warp->votes[emu_lane] = v;
  emu_warp->bar.arrive_and_wait();
  unsigned r = 0;
  for (int i = 0; i < 32; ++i) if ((mask >> i) & 1u) r += emu_warp->votes[i];
  emu_warp->bar.arrive_and_wait();
  return r;
}
static inline unsigned __reduce_min_sync(unsigned mask, unsigned v) {
  emu_warp->votes[emu_lane] = v;
  emu_warp->bar.arrive_and_wait();
  unsigned r = 0xffffffffu;
  for (int i = 0; i < 32; ++i) if ((mask >> i) & 1u) r = std::min(r, emu_warp->votes[i]);
  emu_warp->bar.arrive_and_wait();
  return r;
}
static inline unsigned __ballot_sync(unsigned, bool pred) {
  emu_warp->votes[emu_lane] = pred ? 1u : 0u;
  emu_warp->bar.arrive_and_wait();
  unsigned r = 0;
  for (int i = 0; i < 32; ++i) r |= emu_warp->votes[i] << i;
  emu_warp->bar.arrive_and_wait();
  return r;
}
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
template <typename T> static inline T __ldg(const T* p) { return *p; }
template <typename T> static inline T __ldcg(const T* p) { return *p; }
static inline float rsqrtf(float x) { return 1.0f / std::sqrt(x); }
using std::min;
