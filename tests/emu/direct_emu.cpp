// direct_emu.cpp -- the direct-summation kernels' real source (gravhopper_b200/csrc/direct.cuh) on the
// CPU: TEST INFRASTRUCTURE.  emu_direct mirrors run_f32 / run_f64 of csrc/direct.cu (source split
// chosen by the product's own choose_split, per-chunk partial sums, finalize kernel, epilogue);
// the epilogue can be the plain store (EP_ACC) or the fused kick + drifts of the leapfrog (EP_STEP).
#include "emu_shim.h"

#include "../../gravhopper_b200/csrc/direct.cuh"

namespace gh {
void set_error(const char *, ...) {}
int64_t &launch_counter() { static thread_local int64_t c = 0; return c; }
}  // namespace gh

using namespace gh;

template <int BLOCK, int KI>
static int run32(const float4 *src, int64_t nj, const float4 *tgt, int64_t ni, double eps, const Epilogue &ep,
                 int *info) {
  Split sp = choose_split(ni, nj, BLOCK * KI, BLOCK);
  std::vector<double> partial(sp.S > 1 ? (size_t)3 * sp.S * ni : 0);
  double *pp = sp.S > 1 ? partial.data() : nullptr;
  const float eps2 = (float)(eps * eps);
  info[0] = sp.S; info[1] = (int)sp.itiles;
  if (eps == 0.0)
    emu::launch(sp.itiles, BLOCK, [&] { direct_f32_kernel<BLOCK, KI, true, 0, 1, 4>(src, nj, tgt, ni, eps2, sp.jchunk, pp, ep); }, false, sp.S);
  else
    emu::launch(sp.itiles, BLOCK, [&] { direct_f32_kernel<BLOCK, KI, false, 0, 1, 4>(src, nj, tgt, ni, eps2, sp.jchunk, pp, ep); }, false, sp.S);
  if (pp) emu::launch((unsigned)((ni + 255) / 256), 256, [&] { finalize_kernel(pp, sp.S, ni, ep); }, true);
  return 0;
}

template <int BLOCK, int KI>
static int run64(const double *pos, const double *mass, int64_t nj, const double *tpos, int64_t ni, double eps,
                 const Epilogue &ep, int *info) {
  Split sp = choose_split(ni, nj, BLOCK * KI, BLOCK);
  std::vector<double> partial(sp.S > 1 ? (size_t)3 * sp.S * ni : 0);
  double *pp = sp.S > 1 ? partial.data() : nullptr;
  const double eps2 = eps * eps;
  info[0] = sp.S; info[1] = (int)sp.itiles;
  if (eps == 0.0)
    emu::launch(sp.itiles, BLOCK, [&] { direct_f64_kernel<BLOCK, KI, true>(pos, mass, nj, tpos, ni, eps2, sp.jchunk, pp, ep); }, false, sp.S);
  else
    emu::launch(sp.itiles, BLOCK, [&] { direct_f64_kernel<BLOCK, KI, false>(pos, mass, nj, tpos, ni, eps2, sp.jchunk, pp, ep); }, false, sp.S);
  if (pp) emu::launch((unsigned)((ni + 255) / 256), 256, [&] { finalize_kernel(pp, sp.S, ni, ep); }, true);
  return 0;
}

extern "C" {
// step == 0: acc_out (ni,3) = accelerations (raw units).  step != 0: the fused leapfrog epilogue with
// xhalf, v_in (ni,3), dt -> x_out, v_out, xhalf_next (ni,3).
// prec 64: pos (nj,3), mass (nj), tpos (ni,3).  prec 32: src32 (nj) / tgt32 (ni) float4 (x - origin, m).
// shape: 1128 / 2128 / 4256 (KI, BLOCK) as in launch_direct.  info = {S, itiles}.
int emu_direct(int prec, const void *src, const double *mass, int64_t nj, const void *tgt, int64_t ni, double eps,
               int shape, int step, double *acc_out, const double *xhalf, const double *v_in, double dt,
               double *x_out, double *v_out, double *xhalf_next, int *info) {
  Epilogue ep;
  std::memset(&ep, 0, sizeof(ep));
  if (!step) {
    ep.mode = EP_ACC;
    ep.acc_out = acc_out;
  } else {
    ep.mode = EP_STEP;
    ep.xhalf = xhalf; ep.v_in = v_in; ep.x_out = x_out; ep.v_out = v_out; ep.xhalf_next = xhalf_next;
    ep.dt = dt;
  }
  if (prec == 32) {
    const float4 *s = static_cast<const float4 *>(src), *t = static_cast<const float4 *>(tgt);
    switch (shape) {
      case 4256: return run32<256, 4>(s, nj, t, ni, eps, ep, info);
      case 2128: return run32<128, 2>(s, nj, t, ni, eps, ep, info);
      default: return run32<128, 1>(s, nj, t, ni, eps, ep, info);
    }
  }
  const double *p = static_cast<const double *>(src), *t = static_cast<const double *>(tgt);
  switch (shape) {
    case 2256: return run64<256, 2>(p, mass, nj, t, ni, eps, ep, info);
    case 4128: return run64<128, 4>(p, mass, nj, t, ni, eps, ep, info);
    default: return run64<128, 1>(p, mass, nj, t, ni, eps, ep, info);
  }
}
}

// Device-native analytic potentials: gh::eval_potentials (common.cuh) is what potentials_kernel
// (csrc/direct.cu) calls per particle; out (n,3) in km/s/Myr.  kinds[k] with 8 doubles prm8[k].
extern "C" int emu_potentials(int npot, const int *kinds, const double *prm8, const double *xhalf, int64_t n,
                              double *out) {
  PotentialSet ps;
  std::memset(&ps, 0, sizeof(ps));
  if (npot > GH_MAX_POTENTIALS) return 1;
  ps.n = npot;
  for (int k = 0; k < npot; k++) {
    ps.kind[k] = kinds[k];
    for (int j = 0; j < GH_POT_NPARAM; j++) ps.prm[k][j] = prm8[k * GH_POT_NPARAM + j];
  }
  for (int64_t i = 0; i < n; i++) {
    double a[3] = {0.0, 0.0, 0.0};
    eval_potentials(ps, xhalf + 3 * i, a);
    out[3 * i] = a[0]; out[3 * i + 1] = a[1]; out[3 * i + 2] = a[2];
  }
  return 0;
}
