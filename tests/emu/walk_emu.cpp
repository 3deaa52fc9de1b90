// walk_emu.cpp -- runs the REAL device code of the tree walks (gravhopper_b200/csrc/walk.cuh) on the
// CPU: test infrastructure only.  Every warp is 32 host threads in lockstep; the warp-collective
// intrinsics (__shfl_down_sync, __ballot_sync, __reduce_*_sync, __any_sync, __syncwarp) exchange
// values through a barrier, __shared__ arrays become statics shared by the 32 threads, and the
// three inline-PTX spots are replaced under GH_HOST_EMU.  One warp (= one CTA, WPC = 1) runs at a
// time.  Built by tests/test_walk_emu.py:  g++ -O1 -std=c++20 -shared -fPIC -pthread.
#include "emu_shim.h"

#include "../../gravhopper_b200/csrc/walk.cuh"

// common.cuh only declares these; nothing on the walk path calls them
namespace gh {
void set_error(const char *, ...) {}
}

// ---- driver ---------------------------------------------------------------------------------------
template <class F> static void run_warps(int64_t nwarps, F body) { emu::launch((unsigned)nwarps, 32, body); }

extern "C" {

// The fp32 group walk over a pre-order entry array (8 floats per entry, gh::Node<float>), targets =
// the sorted sources themselves.  flags: bit 0 STATS, bit 1 GUARD (eps == 0), bit 2 HYBRID.
int emu_walk_group(const float *nodes, int nentries, const double *sorted4, const int *order, int64_t ni,
                   const double *root, float eps2, double inv_theta2, int list_limit, float kappa,
                   double *acc_out, unsigned long long *stats4, int flags, const int *walkctl) {
  // walkctl (nullable): {overflow flag, index of the root entry}, see walk_kernel
  using namespace gh;
  TargetsView tv;
  std::memset(&tv, 0, sizeof(tv));
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
#define EMU_GW(S, G, H) run_warps(nwarps, [&] { walk_group_kernel<1, S, G, H>(nd, nentries, tv, ni, root, eps2, inv_theta2, list_limit, ep, stats4, walkctl); })
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
  std::memset(&tv, 0, sizeof(tv));
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
                      unsigned long long *stats4, int flags, const double *quad) {
  // quad (nullable): 6 doubles per entry -> the opt-in quadrupole instantiation of the kernel
  using namespace gh;
  TargetsView tv;
  std::memset(&tv, 0, sizeof(tv));
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
  if (quad)
    run_warps(nwarps, [&] { walk_kernel<double, true, false, false, true>(nd, skips, nentries, tv, ni, root, false, eps2, inv_theta2, ep, stats4, nullptr, quad); });
  else if (flags & 2)
    run_warps(nwarps, [&] { walk_kernel<double, true, true, false>(nd, skips, nentries, tv, ni, root, false, eps2, inv_theta2, ep, stats4); });
  else
    run_warps(nwarps, [&] { walk_kernel<double, true, false, false>(nd, skips, nentries, tv, ni, root, false, eps2, inv_theta2, ep, stats4); });
  return 0;
}

}  // extern "C"
