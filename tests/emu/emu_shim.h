// emu_shim.h -- host stand-ins for the CUDA builtins the tree kernels use (TEST INFRASTRUCTURE).
// A kernel launch becomes emu::launch(grid, block, body): blocks run one after another; the threads
// of a block are host threads (created once per launch), grouped in warps of 32 that exchange values through a barrier for
// the warp collectives (__shfl_*_sync, __ballot_sync, __match_any_sync, __reduce_*_sync,
// __any_sync, __syncwarp); __syncthreads is a barrier over the block; __shared__ arrays are statics
// (one block is resident at a time).  Kernels without collectives can be launched `serial`.
#pragma once
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#define GH_HOST_EMU 1
#define __launch_bounds__(...)
#include <cuda_runtime.h>
#undef __shared__
#define __shared__ static

struct EmuDim3 { unsigned x = 0, y = 0, z = 0; };
static thread_local EmuDim3 threadIdx, blockIdx, blockDim, gridDim;

namespace emu {
struct Warp {
  std::barrier<> bar;
  uint64_t slot[32];
  explicit Warp(int n) : bar(n) {}
};
static thread_local Warp *t_warp = nullptr;
static thread_local std::barrier<> *t_block = nullptr;
static thread_local int t_lane = 0;
static unsigned g_block_y = 0, g_grid_y = 1;  // row of a 2-D grid being run
static int g_and_flag = 1;                     // __syncthreads_and accumulator

template <class T> static inline uint64_t bits(T v) { static_assert(sizeof(T) <= 8, ""); uint64_t b = 0; std::memcpy(&b, &v, sizeof(T)); return b; }
template <class T> static inline T from(uint64_t b) { T v; std::memcpy(&v, &b, sizeof(T)); return v; }

template <class F> static void launch(unsigned grid, unsigned block, F body, bool serial = false, unsigned grid_y = 1) {
  if (grid_y > 1) {  // 2-D grid: rows one after another
    for (unsigned y = 0; y < grid_y; y++) {
      g_block_y = y; g_grid_y = grid_y;
      launch(grid, block, body, serial, 1);
    }
    g_block_y = 0; g_grid_y = 1;
    return;
  }
  if (serial) {  // kernels without collectives or barriers: plain loops
    for (unsigned b = 0; b < grid; b++)
      for (unsigned t = 0; t < block; t++) {
        threadIdx.x = t; blockIdx.x = b; blockDim.x = block; gridDim.x = grid;
        blockIdx.y = g_block_y; gridDim.y = g_grid_y;
        t_lane = (int)(t & 31);
        body();
      }
    return;
  }
  // `block` host threads live for the whole launch and run the blocks one after another (a block
  // barrier between two blocks: the __shared__ statics are reused)
  const unsigned nw = (block + 31) / 32;
  std::vector<std::unique_ptr<Warp>> warps;
  for (unsigned w = 0; w < nw; w++) warps.emplace_back(new Warp((int)((w + 1) * 32 <= block ? 32 : block - w * 32)));
  std::barrier<> blockbar((std::ptrdiff_t)block);
  std::vector<std::thread> th;
  th.reserve(block);
  for (unsigned t = 0; t < block; t++)
    th.emplace_back([&, t] {
      threadIdx.x = t; blockDim.x = block; gridDim.x = grid;
      blockIdx.y = g_block_y; gridDim.y = g_grid_y;
      t_lane = (int)(t & 31);
      t_warp = warps[t >> 5].get();
      t_block = &blockbar;
      for (unsigned b = 0; b < grid; b++) {
        blockIdx.x = b;
        body();
        blockbar.arrive_and_wait();
      }
    });
  for (auto &x : th) x.join();
}
}  // namespace emu

static inline void __syncthreads() { emu::t_block->arrive_and_wait(); }
static inline int __syncthreads_and(int pred) {
  if (!pred) __atomic_store_n(&emu::g_and_flag, 0, __ATOMIC_RELAXED);
  emu::t_block->arrive_and_wait();
  const int r = __atomic_load_n(&emu::g_and_flag, __ATOMIC_RELAXED);
  emu::t_block->arrive_and_wait();
  if (threadIdx.x == 0) __atomic_store_n(&emu::g_and_flag, 1, __ATOMIC_RELAXED);
  emu::t_block->arrive_and_wait();
  return r;
}
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::t_warp->bar.arrive_and_wait(); }
template <class T> static inline T emu_exchange(T v, int src) {  // value of lane src (own if out of range)
  emu::t_warp->slot[emu::t_lane] = emu::bits(v);
  emu::t_warp->bar.arrive_and_wait();
  const T r = (src >= 0 && src < 32) ? emu::from<T>(emu::t_warp->slot[src]) : v;
  emu::t_warp->bar.arrive_and_wait();
  return r;
}
template <class T> static inline T __shfl_down_sync(unsigned, T v, int d) { return emu_exchange(v, emu::t_lane + d); }
template <class T> static inline T __shfl_up_sync(unsigned, T v, int d) { return emu_exchange(v, emu::t_lane - d >= 0 ? emu::t_lane - d : -1); }
template <class T> static inline T __shfl_sync(unsigned, T v, int src) { return emu_exchange(v, src & 31); }
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m) { return emu_exchange(v, (emu::t_lane ^ m) & 31); }
static inline unsigned __ballot_sync(unsigned, bool pred) {
  emu::t_warp->slot[emu::t_lane] = pred ? 1u : 0u;
  emu::t_warp->bar.arrive_and_wait();
  unsigned m = 0;
  for (int l = 0; l < 32; l++) m |= (unsigned)(emu::t_warp->slot[l] & 1u) << l;
  emu::t_warp->bar.arrive_and_wait();
  return m;
}
static inline bool __any_sync(unsigned mask, bool pred) { return __ballot_sync(mask, pred) != 0u; }
static inline unsigned __match_any_sync(unsigned, unsigned v) {
  emu::t_warp->slot[emu::t_lane] = v;
  emu::t_warp->bar.arrive_and_wait();
  unsigned m = 0;
  for (int l = 0; l < 32; l++) m |= (unsigned)(emu::t_warp->slot[l] == (uint64_t)v) << l;
  emu::t_warp->bar.arrive_and_wait();
  return m;
}
static inline int __reduce_min_sync(unsigned, int v) {
  emu::t_warp->slot[emu::t_lane] = emu::bits(v);
  emu::t_warp->bar.arrive_and_wait();
  int r = emu::from<int>(emu::t_warp->slot[0]);
  for (int l = 1; l < 32; l++) { const int o = emu::from<int>(emu::t_warp->slot[l]); r = o < r ? o : r; }
  emu::t_warp->bar.arrive_and_wait();
  return r;
}
static inline int __reduce_max_sync(unsigned, int v) {
  emu::t_warp->slot[emu::t_lane] = emu::bits(v);
  emu::t_warp->bar.arrive_and_wait();
  int r = emu::from<int>(emu::t_warp->slot[0]);
  for (int l = 1; l < 32; l++) { const int o = emu::from<int>(emu::t_warp->slot[l]); r = o > r ? o : r; }
  emu::t_warp->bar.arrive_and_wait();
  return r;
}
static inline void __threadfence_block() {}
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
static inline int __clzll(long long v) { return v == 0 ? 64 : __builtin_clzll((unsigned long long)v); }
static inline int __float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline float2 __fadd2_rn(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
static inline float2 __fmul2_rn(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
static inline float2 __ffma2_rn(float2 a, float2 b, float2 c) { return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline int atomicAdd(int *p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline unsigned long long atomicMax(unsigned long long *p, unsigned long long v) {
  unsigned long long old = __atomic_load_n(p, __ATOMIC_RELAXED);
  while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
  return old;
}
static inline int atomicMax(int *p, int v) {
  int old = __atomic_load_n(p, __ATOMIC_RELAXED);
  while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
  return old;
}
