// ic_emu.cpp -- the device-side IC sampler's real source (gravhopper_b200/csrc/ic.cuh) on the CPU:
// TEST INFRASTRUCTURE.  emu_ic_sample mirrors launch_ic of csrc/ic.cu (sampling kernel, then the
// two-stage centring of positions and velocities).
#include "emu_shim.h"

#include "../../gravhopper_b200/csrc/ic.cuh"

namespace gh {
void set_error(const char *, ...) {}
int64_t &launch_counter() { static thread_local int64_t c = 0; return c; }
}  // namespace gh

using namespace gh;

extern "C" int emu_ic_sample(int kind, int64_t n, const double *prm3, const double *tx, const double *ty, int nt,
                             uint64_t seed, double *pos, double *vel, double *mass) {
  if (n <= 0) return 0;
  emu::launch((unsigned)((n + 255) / 256), 256, [&] { ic_kernel(kind, n, prm3[0], prm3[1], prm3[2], tx, ty, nt, seed, pos, vel, mass); }, true);
  int nb = (int)((n + 4095) / 4096);
  if (nb > 256) nb = 256;
  std::vector<double> scratch(3 * 256);
  for (double *arr : {pos, vel}) {
    emu::launch((unsigned)nb, 256, [&] { mean_stage1(arr, n, scratch.data()); });
    emu::launch((unsigned)nb, 256, [&] { mean_stage2_shift(arr, n, scratch.data(), nb); });
  }
  return 0;
}

extern "C" int emu_ic_sample_expdisk(int64_t n, const double *prm4, const double *tR, const double *tcum,
                                     const double *tvphi, const double *tratio, int nt, uint64_t seed, double *pos,
                                     double *vel, double *mass) {
  if (n <= 0) return 0;
  emu::launch((unsigned)((n + 255) / 256), 256, [&] { ic_expdisk_kernel(n, prm4[0], prm4[1], prm4[2], prm4[3], tR, tcum, tvphi, tratio, nt, seed, pos, vel, mass); }, true);
  int nb = (int)((n + 4095) / 4096);
  if (nb > 256) nb = 256;
  std::vector<double> scratch(3 * 256);
  for (double *arr : {pos, vel}) {
    emu::launch((unsigned)nb, 256, [&] { mean_stage1(arr, n, scratch.data()); });
    emu::launch((unsigned)nb, 256, [&] { mean_stage2_shift(arr, n, scratch.data(), nb); });
  }
  return 0;
}
