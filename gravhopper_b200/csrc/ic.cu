// ic.cu -- device-side initial-condition sampling: the same sampling maths as
// the reference's IC.Plummer / IC.Hernquist / IC.TSIS (/root/reference/gravhopper/gravhopper.py
// :1452-1491, :1544-1605, :1378-1398) with a counter-based generator (Philox4x32-10, written out
// below), one independent stream per particle, so that N = 10M ICs are produced in HBM in a few
// milliseconds instead of seconds on the host + an upload.  The host generators in ic_raw.py stay
// the seed-compatible path (numpy default_rng draw order); this path is validated statistically
// against them (tests/test_gpu_ic.py), not bit-wise: the random streams differ by construction.
#include "common.cuh"
#include "ic.cuh"

namespace gh {

int launch_ic(int kind, int64_t n, const double prm[3], const double *tx, const double *ty, int nt,
              uint64_t seed, double *pos, double *vel, double *mass, double *scratch /* 3*256 */,
              cudaStream_t st) {
  if (n <= 0) return GH_OK;
  ic_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(kind, n, prm[0], prm[1], prm[2], tx, ty, nt, seed,
                                                       pos, vel, mass);
  GH_LAUNCH_CHECK();
  int nb = (int)((n + 4095) / 4096);
  if (nb > 256) nb = 256;
  for (double *arr : {pos, vel}) {
    mean_stage1<<<nb, 256, 0, st>>>(arr, n, scratch);
    GH_LAUNCH_CHECK();
    mean_stage2_shift<<<nb, 256, 0, st>>>(arr, n, scratch, nb);
    GH_LAUNCH_CHECK();
  }
  return GH_OK;
}

int launch_ic_expdisk(int64_t n, const double prm[4], const double *tR, const double *tcum,
                      const double *tvphi, const double *tratio, int nt, uint64_t seed, double *pos,
                      double *vel, double *mass, double *scratch /* 3*256 */, cudaStream_t st) {
  if (n <= 0) return GH_OK;
  ic_expdisk_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, prm[0], prm[1], prm[2], prm[3], tR, tcum, tvphi,
                                                               tratio, nt, seed, pos, vel, mass);
  GH_LAUNCH_CHECK();
  int nb = (int)((n + 4095) / 4096);
  if (nb > 256) nb = 256;
  for (double *arr : {pos, vel}) {
    mean_stage1<<<nb, 256, 0, st>>>(arr, n, scratch);
    GH_LAUNCH_CHECK();
    mean_stage2_shift<<<nb, 256, 0, st>>>(arr, n, scratch, nb);
    GH_LAUNCH_CHECK();
  }
  return GH_OK;
}

}  // namespace gh
