// common.cuh -- shared declarations of libgravhopper_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include "../../include/gravhopper_b200.h"

namespace gh {

void set_error(const char *fmt, ...);
int64_t &launch_counter();  // per-thread count of kernel launches (bench.py gpu_launches)

#define GH_CUDA(call)                                                                       \
  do {                                                                                      \
    cudaError_t e_ = (call);                                                                \
    if (e_ != cudaSuccess) {                                                                \
      gh::set_error("%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_));     \
      return (e_ == cudaErrorMemoryAllocation) ? GH_ENOMEM : GH_ECUDA;                      \
    }                                                                                       \
  } while (0)

#define GH_TRY(call)                 \
  do {                               \
    int rc_ = (call);                \
    if (rc_ != GH_OK) return rc_;    \
  } while (0)

#define GH_LAUNCH_CHECK()            \
  do {                               \
    gh::launch_counter()++;          \
    GH_CUDA(cudaGetLastError());     \
  } while (0)

// fp32 kernels: eps^2 below this (kpc^2) is treated as 0 with the zero-distance guard (an fp32
// self term m * rsqrt(eps^2)^3 * 0 would otherwise be inf * 0 = NaN; fp64 and the reference stay finite)
#define GH_F32_MIN_EPS2 1e-16f

// Grow-only device scratch buffer.
struct DeviceBuffer {
  void *ptr = nullptr;
  size_t bytes = 0;
  int reserve(size_t need) {
    if (need <= bytes) return GH_OK;
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    bytes = 0;
    size_t want = need + need / 8 + 256;
    GH_CUDA(cudaMalloc(&ptr, want));
    bytes = want;
    return GH_OK;
  }
  void release() {
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    bytes = 0;
  }
  template <class T> T *as() const { return reinterpret_cast<T *>(ptr); }
};

// What a kernel does with a finished target acceleration a_i (raw units, G = 1).
//   EP_ACC : acc_out[i] = a_i                                  (the _jbgrav entry points)
//   EP_STEP: the kick and both drifts of gravhopper.py:414-416 (+ the next step's :409),
//            with the unit factors of jbgrav.py:48; products and sums are rounded separately
//            (__dmul_rn/__dadd_rn) exactly as numpy evaluates the reference expressions.
//   EP_ACC32: acc32_out[i] = (a_i, 0) in fp32 (distributed tree walk: accelerations of the targets a
//            rank walked for other ranks, all-gathered before the owners' EP_STEP)
enum { EP_ACC = 0, EP_STEP = 1, EP_ACC32 = 2 };

// Device-native analytic external potentials: a tiny kernel evaluates them at
// x_half into the step's external-acceleration buffer just before the force kernel, so a run in a
// static background field never leaves the device (replaces the per-step Python callback of
// gravhopper.py:462-473 for these cases; kept out of the force kernels to keep their register
// count, and so their occupancy, unchanged).
// Units: kpc, Msun, km/s; the returned acceleration is in km/s/Myr.
#define GH_MAX_POTENTIALS 4
#define GH_POT_NPARAM 8
struct PotentialSet {
  int n;
  int kind[GH_MAX_POTENTIALS];
  double prm[GH_MAX_POTENTIALS][GH_POT_NPARAM];
};

struct Epilogue {
  int mode;
  // EP_ACC
  double *acc_out;  // (nt,3)
  // EP_ACC32
  float4 *acc32_out;  // (nt)
  // EP_STEP (all indexed by the local target index i)
  const double *xhalf;   // (nt,3) x_half of this step
  const double *v_in;    // (nt,3) v_n
  double *x_out;         // (nt,3) x_{n+1}
  double *v_out;         // (nt,3) v_{n+1}
  double *xhalf_next;    // (nt,3) x_half of step n+1 (f64 source slice or private array)
  float4 *src32_next;    // nullable: (nt) float4 (x_half_next - origin, mass) for the f32 path
  const double *mass;    // (nt) masses of the owned targets (for src32_next.w)
  const double *ext;     // nullable: (nt,3) external acceleration, km/s/Myr
  double dt;
  double origin[3];
};

// kinds and parameter layouts (centre always in prm[1..3], kpc):
//  1 POINTMASS / PLUMMER: prm[0] = M [Msun], prm[4] = softening b [kpc]:  a = -G M d / (r^2+b^2)^1.5
//  2 HERNQUIST:           prm[0] = M, prm[4] = a:                         a = -G M d / (r (r+a)^2)
//  3 NFW:                 prm[0] = M_s = 4 pi rho0 rs^3, prm[4] = rs:     a = -G M_s (ln(1+x) - x/(1+x)) d / r^3
//  4 LOGHALO:             prm[0] = v0 [km/s], prm[4] = rc, prm[5] = q:    a = -v0^2 (x, y, z/q^2) / (rc^2+x^2+y^2+z^2/q^2)
//  5 MIYAMOTO-NAGAI:      prm[0] = M, prm[4] = a, prm[5] = b
__device__ __forceinline__ void eval_potentials(const PotentialSet &ps, const double x[3], double a[3]) {
  for (int k = 0; k < ps.n; k++) {
    const double *q = ps.prm[k];
    const double d0 = x[0] - q[1], d1 = x[1] - q[2], d2 = x[2] - q[3];
    double f0 = 0.0, f1 = 0.0, f2 = 0.0;  // (km/s)^2 / kpc
    switch (ps.kind[k]) {
      case 1: {
        const double s = d0 * d0 + d1 * d1 + d2 * d2 + q[4] * q[4];
        const double w = (s > 0.0) ? -GH_G * q[0] / (s * sqrt(s)) : 0.0;
        f0 = w * d0; f1 = w * d1; f2 = w * d2;
      } break;
      case 2: {
        const double r = sqrt(d0 * d0 + d1 * d1 + d2 * d2);
        const double w = (r > 0.0) ? -GH_G * q[0] / (r * (r + q[4]) * (r + q[4])) : 0.0;
        f0 = w * d0; f1 = w * d1; f2 = w * d2;
      } break;
      case 3: {
        const double r = sqrt(d0 * d0 + d1 * d1 + d2 * d2);
        const double xs = r / q[4];
        const double w = (r > 0.0) ? -GH_G * q[0] * (log1p(xs) - xs / (1.0 + xs)) / (r * r * r) : 0.0;
        f0 = w * d0; f1 = w * d1; f2 = w * d2;
      } break;
      case 4: {
        const double iq2 = 1.0 / (q[5] * q[5]);
        const double w = -q[0] * q[0] / (q[4] * q[4] + d0 * d0 + d1 * d1 + d2 * d2 * iq2);
        f0 = w * d0; f1 = w * d1; f2 = w * d2 * iq2;
      } break;
      case 5: {
        const double zb = sqrt(d2 * d2 + q[5] * q[5]);
        const double az = q[4] + zb;
        const double s = d0 * d0 + d1 * d1 + az * az;
        const double w = -GH_G * q[0] / (s * sqrt(s));
        f0 = w * d0; f1 = w * d1; f2 = (zb > 0.0) ? w * d2 * az / zb : 0.0;
      } break;
      default: break;
    }
    a[0] += f0 * GH_KPC_PER_KMS_MYR;  // 1 (km/s)^2/kpc = K km/s/Myr
    a[1] += f1 * GH_KPC_PER_KMS_MYR;
    a[2] += f2 * GH_KPC_PER_KMS_MYR;
  }
}

__device__ __forceinline__ void apply_epilogue(const Epilogue &ep, int64_t i, double ax, double ay,
                                               double az) {
  if (ep.mode == EP_ACC) {
    ep.acc_out[3 * i + 0] = ax;
    ep.acc_out[3 * i + 1] = ay;
    ep.acc_out[3 * i + 2] = az;
    return;
  }
  if (ep.mode == EP_ACC32) {
    ep.acc32_out[i] = make_float4((float)ax, (float)ay, (float)az, 0.f);
    return;
  }
  double a[3] = {ax, ay, az};
  double xn[3];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    double acc = __dmul_rn(a[k], GH_C_ACC);                               // jbgrav.py:48
    if (ep.ext) acc = __dadd_rn(acc, ep.ext[3 * i + k]);                  // gravhopper.py:457
    double vn = __dadd_rn(ep.v_in[3 * i + k], __dmul_rn(acc, ep.dt));     // :414
    double hd = __dmul_rn(__dmul_rn(__dmul_rn(0.5, vn), ep.dt), GH_KPC_PER_KMS_MYR);
    double x1 = __dadd_rn(ep.xhalf[3 * i + k], hd);                       // :416
    ep.v_out[3 * i + k] = vn;
    ep.x_out[3 * i + k] = x1;
    xn[k] = __dadd_rn(x1, hd);                                            // next step's :409
    ep.xhalf_next[3 * i + k] = xn[k];
  }
  if (ep.src32_next) {
    ep.src32_next[i] = make_float4((float)(xn[0] - ep.origin[0]), (float)(xn[1] - ep.origin[1]),
                                   (float)(xn[2] - ep.origin[2]), (float)ep.mass[i]);
  }
}

// ---- direct summation (direct.cu) ----------------------------------------------------------
struct DirectArgs {
  int prec;  // GH_PREC_F32 / GH_PREC_F64
  // sources: f64 -> pos (nj,3) + mass (nj); f32 -> src32 (nj) float4 (x - origin, mass)
  const double *src_pos;
  const double *src_mass;
  const float4 *src32;
  int64_t nj;
  // targets: f64 -> tgt_pos (ni,3); f32 -> tgt32 (ni) float4 (x - origin, unused)
  const double *tgt_pos;
  const float4 *tgt32;
  int64_t ni;
  double eps;
  Epilogue ep;
  // 1: the sources are known to carry different masses (the uniform-mass tile fast path will not
  // trigger): the fp32 launch then uses 128 threads x 8 targets, measured 71.0 % of the FP32 peak
  // against 68.8 % for the default 256 x 4 at N = 2^20 (profiles/r01_sweep_direct.txt)
  int mixed_mass;
};
// Launches the force kernel (and a finalize kernel when the source range is split).  `ws` is
// scratch for the per-split partial sums.  force_ms_events (nullable): two events recorded
// around the dominant kernel.
int launch_direct(const DirectArgs &a, DeviceBuffer &ws, cudaStream_t stream,
                  cudaEvent_t *force_events);
// pos (n,3) f64 [+ mass] -> float4 (x - origin, m)
int launch_pack32(const double *pos, const double *mass, int64_t n, const double origin[3],
                  float4 *out, cudaStream_t stream);
// potential energy partial sums; out2 = {KE, PE} accumulated with atomics (must be zeroed)
int launch_energy(const double *x, const double *v, const double *m_tgt, int64_t ni,
                  const double *src_pos, const double *src_mass, int64_t nj, int64_t self_offset,
                  double eps, double *out2, cudaStream_t stream);
// ext_out[i] = (ext_in ? ext_in[i] : 0) + sum of native potentials at xhalf[i]   (km/s/Myr)
int launch_potentials(const PotentialSet &ps, const double *xhalf, int64_t n, const double *ext_in,
                      double *ext_out, cudaStream_t stream);
// x_half = x + ((0.5 v) dt) K   (gravhopper.py:409)
int launch_half_drift(const double *x, const double *v, const double *mass, int64_t n, double dt,
                      double *xhalf, float4 *src32, const double origin[3], cudaStream_t stream);

// ---- device-side IC sampling (ic.cu) -------------------------------------------------------
int launch_ic(int kind, int64_t n, const double prm[3], const double *tx, const double *ty, int nt,
              uint64_t seed, double *pos, double *vel, double *mass, double *scratch,
              cudaStream_t stream);

int launch_ic_expdisk(int64_t n, const double prm[4], const double *tR, const double *tcum,
                      const double *tvphi, const double *tratio, int nt, uint64_t seed, double *pos,
                      double *vel, double *mass, double *scratch, cudaStream_t stream);

// ---- tree (tree.cu) -------------------------------------------------------------------------
struct TreeWorkspace;
TreeWorkspace *tree_workspace_create();
void tree_workspace_destroy(TreeWorkspace *);
struct TreeArgs {
  int prec;
  const double *src_pos;   // (nj,3) device
  const double *src_mass;  // (nj) device
  const float4 *src32;     // alternative f32 sources (x - origin, m); then tgt32 are the targets
  const float4 *tgt32;
  int64_t nj;
  const double *tgt_pos;   // (ni,3) device
  int64_t ni;
  bool targets_are_sources;  // tgt_pos == src_pos + 3*tgt_offset: reuse the Morton order
  int64_t tgt_offset;
  double eps, theta;
  Epilogue ep;
  bool want_stats;
  bool sync_check;  // the caller synchronises anyway: check (and repair) an entry-array overflow
  bool coherent;    // a step of a running simulation: the key distribution of the previous build in this
                    // workspace describes this one (splitter sort, bucketsort.cuh)
};
// One evaluation never needs the host to learn a count from the device: the entry array has a
// capacity, the walk ends its chains at the capacity, an overflow raises a device flag.
int launch_tree(const TreeArgs &a, TreeWorkspace *ws, cudaStream_t stream,
                cudaEvent_t *force_events);
// Distributed build (fp32): this rank sorts / scans / emits only the particles of its key range
// into segment [rank * stride, (rank + 1) * stride) of the global entry array.  The caller runs
// phase 0, all-gathers exchange buffer 1, phase 1, all-gathers buffer 2, phase 2, all-gathers
// buffers 3 and 4 (the entries, the sorted particle indices), phase 3 (walk of this rank's share
// of the global Morton order), all-gathers buffer 5 (accelerations), phase 4 (kick and drift of the
// owned particles).  See build.cuh.
struct TreeDist {
  int rank, world;
  int64_t stride;  // entries a rank's segment holds
  int64_t ncap;    // particles a rank's key range may hold (grids, scans and sort tables are sized for
                   // this, not for all N); 0 = N.  More particles than this raise the overflow flag.
  int blk;         // targets are dealt to the ranks in blocks of this many Morton-consecutive particles
                   // of the GLOBAL sorted order (0 = 2048)
};
int launch_tree_phase(const TreeArgs &a, TreeWorkspace *ws, cudaStream_t stream, cudaEvent_t *force_events,
                      const TreeDist *dist, int phase);
int tree_exchange_buffer(TreeWorkspace *ws, int which, void **ptr, int64_t *bytes_per_rank);
int tree_splitters(TreeWorkspace *ws, int world, cudaStream_t stream);
int tree_poll_overflow(TreeWorkspace *ws, int64_t *entries);
void tree_forget_history(TreeWorkspace *ws);  // new state uploaded: the next build is not coherent with the last
int tree_last_stats(TreeWorkspace *ws, int64_t out[8]);
int tree_walk_mode();
void set_tree_walk_mode(int mode);
int tree_quadrupoles();
void set_tree_quadrupoles(int on);
float group_hybrid_kappa();
void set_group_hybrid_kappa(double kappa);

}  // namespace gh
