// engine.cuh -- the device-resident leapfrog engine's state (engine.cu owns the C ABI over it,
// group.cu drives several of them -- one per GPU -- through a step with NCCL collectives).
#pragma once
#include "common.cuh"

using gh::DeviceBuffer;
using gh::PotentialSet;
using gh::TreeWorkspace;
using gh::launch_counter;
using gh::set_error;

static constexpr int GH_RING = 3;

struct gh_engine {
  int device = 0;
  int64_t n = 0, ib = 0, ni = 0;
  int prec = GH_PREC_F64;
  cudaStream_t stream = nullptr, copy_stream = nullptr;
  double *mass = nullptr;             // (n)
  double *x[GH_RING] = {nullptr}, *v[GH_RING] = {nullptr};
  int cur = 0;
  cudaEvent_t copied[GH_RING] = {nullptr};
  bool copy_pending[GH_RING] = {false};
  cudaEvent_t step_done = nullptr;
  bool external_src = false;
  void *src[2] = {nullptr, nullptr};  // f64: double (n,3); f32: float4 (n)
  int scur = 0;
  double *xh_private = nullptr;       // f32: (ni,3) own x_half in float64
  // fp32 coordinates are stored relative to `origin`, which MOVES with the system: origin_vel is the
  // mean velocity at upload (km/s), and every step's epilogue writes the next step's float4 sources
  // relative to origin + origin_vel dt.  Bulk motion (a merger's or flyby's centre-of-mass velocity)
  // then costs no fp32 digits however long the run; all ranks of a group derive the same
  // sequence from the same uploaded values.
  double origin[3] = {0, 0, 0};
  double origin_vel[3] = {0, 0, 0};
  double origin_next[3] = {0, 0, 0};
  double dt_built = 0.0;
  bool uploaded = false, xhalf_valid = false;
  DeviceBuffer ws, ext, ext2;
  TreeWorkspace *tw = nullptr;
  // events around the force kernel of the last FEV_RING steps (gh_engine_last_force_ms,
  // gh_engine_force_ms_mean: the bench's per-kernel time is a mean over the timed steps)
  static constexpr int FEV_RING = 64;
  cudaEvent_t fev[FEV_RING][2] = {};
  int64_t fev_count = 0;
  int64_t launches = 0;
  double *d_energy = nullptr;
  PotentialSet pots;
  bool history_pinned = false;  // the last gh_engine_run page-locked the caller's history arrays
  bool mixed_mass = false;  // the uploaded masses are not all equal (launch-shape hint, direct fp32)
  // distributed tree build (group.cu): bootstrapped by one redundant single-rank build, then
  // every rank builds its key range; stride = entries a rank's segment holds
  bool dist_ready = false;
  int64_t dist_stride = 0;
  int64_t dist_steps = 0;
  static constexpr int MAXENT_RING = 4;
  int *h_maxent = nullptr;                      // pinned [MAXENT_RING]: largest segment fill of a step
  cudaEvent_t maxent_ev[MAXENT_RING] = {};

  double *xhalf_own(int b) const {
    return prec == GH_PREC_F64 ? reinterpret_cast<double *>(src[b]) + 3 * ib : xh_private;
  }
  float4 *src32_own(int b) const {
    return prec == GH_PREC_F32 ? reinterpret_cast<float4 *>(src[b]) + ib : nullptr;
  }
  size_t src_stride() const { return prec == GH_PREC_F64 ? 3 * sizeof(double) : sizeof(float4); }
};

struct LaunchScope {  // attribute this thread's kernel launches to the engine
  gh_engine *e;
  int64_t before;
  explicit LaunchScope(gh_engine *e_) : e(e_), before(launch_counter()) {}
  ~LaunchScope() { e->launches += launch_counter() - before; }
};

#define GH_ENGINE_GUARD(e)                                          \
  if (!(e)) { set_error("null engine"); return GH_EINVAL; }         \
  GH_CUDA(cudaSetDevice((e)->device));                              \
  LaunchScope scope_(e)


namespace gh {
// One DKD step of one engine, split so that group.cu can interleave collectives with the tree's
// phases: engine_step_args builds the force arguments (epilogue included) for the step,
// engine_step_done rotates the state ring.  engine_step_impl = args + force launch + done.
struct StepArgs {
  int algorithm;
  DirectArgs direct;
  TreeArgs tree;
};
int engine_step_args(gh_engine *e, double dt, double eps, double theta, int algorithm, const double *ext_dev,
                     StepArgs *out);
void engine_step_done(gh_engine *e);
int engine_step_impl(gh_engine *e, double dt, double eps, double theta, int algorithm, const double *ext_dev);
int engine_check_tree(gh_engine *e);  // after a sync: did a tree build overflow its entry array?
}  // namespace gh
