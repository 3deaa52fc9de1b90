// walk_emu.cpp -- runs the REAL device code of the tree walks (gravhopper_b200/csrc/walk.cuh) on the
// CPU: test infrastructure only.  Every warp is 32 host threads in lockstep; the warp-collective
// intrinsics (__shfl_down_sync, __ballot_sync, __reduce_*_sync, __any_sync, __syncwarp) exchange
// values through a barrier, __shared__ arrays become statics shared by the 32 threads, and the
// three inline-PTX spots are replaced under GH_HOST_EMU.  One warp (= one CTA, WPC = 1) runs at a
// time.  Built by tests/test_walk_emu.py:  g++ -O1 -std=c++20 -shared -fPIC -pthread.
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#define GH_HOST_EMU 1
#define __launch_bounds__(...)
#include <cuda_runtime.h>

// ---- per-thread CUDA builtins ---------------------------------------------------------------------
struct EmuDim3 { unsigned x = 0, y = 0, z = 0; };
static thread_local EmuDim3 threadIdx, blockIdx, blockDim, gridDim;

struct EmuWarp {
  std::barrier<> bar{32};
  uint64_t slot[32];
};
static EmuWarp *g_warp = nullptr;
static thread_local int t_lane = 0;

template <class T> static inline uint64_t emu_bits(T v) { uint64_t b = 0; std::memcpy(&b, &v, sizeof(T)); return b; }
template <class T> static inline T emu_from(uint64_t b) { T v; std::memcpy(&v, &b, sizeof(T)); return v; }

static inline void __syncwarp(unsigned = 0xffffffffu) { g_warp->bar.arrive_and_wait(); }
template <class T> static inline T __shfl_down_sync(unsigned, T v, int delta) {
  g_warp->slot[t_lane] = emu_bits(v);
  g_warp->bar.arrive_and_wait();
  const int src = t_lane + delta;
  const T r = (src < 32) ? emu_from<T>(g_warp->slot[src]) : v;
  g_warp->bar.arrive_and_wait();
  return r;
}
static inline unsigned __ballot_sync(unsigned, bool pred) {
  g_warp->slot[t_lane] = pred ? 1u : 0u;
  g_warp->bar.arrive_and_wait();
  unsigned m = 0;
  for (int l = 0; l < 32; l++) m |= (unsigned)(g_warp->slot[l] & 1u) << l;
  g_warp->bar.arrive_and_wait();
  return m;
}
static inline bool __any_sync(unsigned mask, bool pred) { return __ballot_sync(mask, pred) != 0u; }
static inline int __reduce_min_sync(unsigned, int v) {
  g_warp->slot[t_lane] = emu_bits(v);
  g_warp->bar.arrive_and_wait();
  int r = emu_from<int>(g_warp->slot[0]);
  for (int l = 1; l < 32; l++) { const int o = emu_from<int>(g_warp->slot[l]); r = o < r ? o : r; }
  g_warp->bar.arrive_and_wait();
  return r;
}
static inline int __reduce_max_sync(unsigned, int v) {
  g_warp->slot[t_lane] = emu_bits(v);
  g_warp->bar.arrive_and_wait();
  int r = emu_from<int>(g_warp->slot[0]);
  for (int l = 1; l < 32; l++) { const int o = emu_from<int>(g_warp->slot[l]); r = o > r ? o : r; }
  g_warp->bar.arrive_and_wait();
  return r;
}
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
static inline int __float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline float2 __fadd2_rn(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
static inline float2 __fmul2_rn(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
static inline float2 __ffma2_rn(float2 a, float2 b, float2 c) { return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) {
  return __atomic_fetch_add(p, v, __ATOMIC_RELAXED);
}
static inline unsigned long long atomicMax(unsigned long long *p, unsigned long long v) {
  unsigned long long old = __atomic_load_n(p, __ATOMIC_RELAXED);
  while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
  return old;
}
#undef __shared__
#define __shared__ static

#include "../../gravhopper_b200/csrc/walk.cuh"

// common.cuh only declares these; nothing on the walk path calls them
namespace gh {
void set_error(const char *, ...) {}
}

// ---- driver ---------------------------------------------------------------------------------------
template <class F> static void run_warps(int64_t nwarps, F body) {
  for (int64_t w = 0; w < nwarps; w++) {
    EmuWarp warp;
    g_warp = &warp;
    std::vector<std::thread> th;
    th.reserve(32);
    for (int l = 0; l < 32; l++)
      th.emplace_back([&, l, w] {
        t_lane = l;
        threadIdx.x = (unsigned)l;
        blockIdx.x = (unsigned)w;
        blockDim.x = 32;
        gridDim.x = (unsigned)nwarps;
        body();
      });
    for (auto &t : th) t.join();
  }
  g_warp = nullptr;
}

extern "C" {

// The fp32 group walk over a pre-order entry array (8 floats per entry, gh::Node<float>), targets =
// the sorted sources themselves.  flags: bit 0 STATS, bit 1 GUARD (eps == 0), bit 2 HYBRID.
int emu_walk_group(const float *nodes, int nentries, const double *sorted4, const int *order, int64_t ni,
                   const double *root, float eps2, double inv_theta2, int list_limit, float kappa,
                   double *acc_out, unsigned long long *stats4, int flags) {
  using namespace gh;
  TargetsView tv;
  tv.sorted = reinterpret_cast<const double4 *>(sorted4);
  tv.pos64 = nullptr;
  tv.pos32 = nullptr;
  tv.order = order;
  tv.order_offset = 0;
  Epilogue ep;
  std::memset(&ep, 0, sizeof(ep));
  ep.mode = EP_ACC;
  ep.acc_out = acc_out;
  c_hybrid_kappa2 = kappa * kappa;
  const Node<float> *nd = reinterpret_cast<const Node<float> *>(nodes);
  const int64_t nwarps = (ni + 31) / 32;
#define EMU_GW(S, G, H) run_warps(nwarps, [&] { walk_group_kernel<1, S, G, H>(nd, nentries, tv, ni, root, eps2, inv_theta2, list_limit, ep, stats4); })
  switch (flags & 7) {
    case 0: EMU_GW(false, false, false); break;
    case 1: EMU_GW(true, false, false); break;
    case 2: EMU_GW(false, true, false); break;
    case 3: EMU_GW(true, true, false); break;
    case 4: EMU_GW(false, false, true); break;
    case 5: EMU_GW(true, false, true); break;
    case 6: EMU_GW(false, true, true); break;
    default: EMU_GW(true, true, true); break;
  }
#undef EMU_GW
  return 0;
}

// The per-target walk (walk_kernel<float>) on the same inputs.
int emu_walk_target(const float *nodes, int nentries, const double *sorted4, const int *order, int64_t ni,
                    const double *root, float eps2, double inv_theta2, double *acc_out,
                    unsigned long long *stats4, int flags) {
  using namespace gh;
  TargetsView tv;
  tv.sorted = reinterpret_cast<const double4 *>(sorted4);
  tv.pos64 = nullptr;
  tv.pos32 = nullptr;
  tv.order = order;
  tv.order_offset = 0;
  Epilogue ep;
  std::memset(&ep, 0, sizeof(ep));
  ep.mode = EP_ACC;
  ep.acc_out = acc_out;
  const Node<float> *nd = reinterpret_cast<const Node<float> *>(nodes);
  const int64_t nwarps = (ni + 31) / 32;
  if (flags & 1)
    run_warps(nwarps, [&] { walk_kernel<float, true, false, false>(nd, nullptr, nentries, tv, ni, root, true, eps2, inv_theta2, ep, stats4); });
  else
    run_warps(nwarps, [&] { walk_kernel<float, false, false, false>(nd, nullptr, nentries, tv, ni, root, true, eps2, inv_theta2, ep, stats4); });
  return 0;
}

// The fp64 per-target walk (walk_kernel<double>): entries are gh::Node<double> (8 doubles, absolute
// coordinates, s2 = side^2/theta^2, -1 for leaves) plus the skip array; targets are arbitrary
// positions (nt,3) in the given order (order == nullptr: identity).
int emu_walk_target64(const double *nodes, const int *skips, int nentries, const double *tpos, const int *order,
                      int64_t ni, const double *root, double eps2, double inv_theta2, double *acc_out,
                      unsigned long long *stats4, int flags) {
  using namespace gh;
  TargetsView tv;
  tv.sorted = nullptr;
  tv.pos64 = tpos;
  tv.pos32 = nullptr;
  tv.order = order;
  tv.order_offset = 0;
  Epilogue ep;
  std::memset(&ep, 0, sizeof(ep));
  ep.mode = EP_ACC;
  ep.acc_out = acc_out;
  const Node<double> *nd = reinterpret_cast<const Node<double> *>(nodes);
  const int64_t nwarps = (ni + 31) / 32;
  if (flags & 2)
    run_warps(nwarps, [&] { walk_kernel<double, true, true, false>(nd, skips, nentries, tv, ni, root, false, eps2, inv_theta2, ep, stats4); });
  else
    run_warps(nwarps, [&] { walk_kernel<double, true, false, false>(nd, skips, nentries, tv, ni, root, false, eps2, inv_theta2, ep, stats4); });
  return 0;
}

}  // extern "C"
