// engine.cu -- the C ABI of libgravhopper_b200.so (include/gravhopper_b200.h): the four
// stateless force entry points and the device-resident leapfrog engine.
#include "common.cuh"
#include "engine.cuh"

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

namespace gh {

static thread_local char g_err[512] = "";
static thread_local int64_t g_launches = 0;

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int64_t &launch_counter() { return g_launches; }

static int check_device() {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    cudaGetLastError();
    set_error("no CUDA device available (%s); libgravhopper_b200 has no CPU fallback",
              e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    return GH_ECUDA;
  }
  return GH_OK;
}

// Per-thread, PER-DEVICE scratch of the stateless entry points (grow-only, reused between calls).
// Keyed by the current device: a call with tensors on another GPU gets that GPU's own buffers.
// Calls on different streams of one device share the scratch, so every call first makes its stream
// wait for the event the previous call recorded on its own stream (no two calls ever touch the
// scratch concurrently; calls on one stream are ordered anyway).
struct Stateless {
  DeviceBuffer pos, mass, tpos, acc, src32, tgt32, ws, root, part, ictab, icout;
  TreeWorkspace *tw = nullptr;
  cudaStream_t stream = nullptr;
  cudaEvent_t last_done = nullptr;   // recorded at the end of the previous call ...
  cudaStream_t last_stream = nullptr;  // ... on this stream
  bool last_valid = false;
};
static constexpr int GH_MAX_DEVICES = 64;
struct ThreadState {
  Stateless *dev[GH_MAX_DEVICES] = {};
  int64_t tree_stats[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  bool want_stats = false;
};
static thread_local ThreadState g_ts;
static Stateless *stateless() {
  int d = 0;
  if (cudaGetDevice(&d) != cudaSuccess) { cudaGetLastError(); d = 0; }
  if (d < 0 || d >= GH_MAX_DEVICES) return nullptr;
  if (!g_ts.dev[d]) g_ts.dev[d] = new (std::nothrow) Stateless();
  return g_ts.dev[d];
}
// order this call after the previous one that used the scratch (see above)
static int scratch_acquire(Stateless *s, cudaStream_t st) {
  if (s->last_valid && s->last_stream != st) GH_CUDA(cudaStreamWaitEvent(st, s->last_done, 0));
  return GH_OK;
}
static int scratch_release(Stateless *s, cudaStream_t st) {
  if (!s->last_done) GH_CUDA(cudaEventCreateWithFlags(&s->last_done, cudaEventDisableTiming));
  GH_CUDA(cudaEventRecord(s->last_done, st));
  s->last_stream = st;
  s->last_valid = true;
  return GH_OK;
}

static bool masses_differ(const double *m, int64_t n) {
  for (int64_t i = 1; i < n; i++)
    if (m[i] != m[0]) return true;
  return false;
}

// mean position -> origin[3] on the device (for the f32 packing).  The mean, not the bbox
// midpoint: centrally concentrated systems have outliers at 10^3 scale radii (untruncated
// Plummer), and an origin far from the core costs fp32 digits exactly where pairs are closest.
// Two tiny deterministic kernels (fixed block count, fixed reduction order).
__global__ void origin_stage1(const double *__restrict__ pos, int64_t n, double *__restrict__ part) {
  __shared__ double sh[3][256];
  double sum[3] = {0.0, 0.0, 0.0};
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    for (int k = 0; k < 3; k++) sum[k] += pos[3 * i + k];
  for (int k = 0; k < 3; k++) sh[k][threadIdx.x] = sum[k];
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s)
      for (int k = 0; k < 3; k++) sh[k][threadIdx.x] += sh[k][threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x < 3) part[blockIdx.x * 3 + threadIdx.x] = sh[threadIdx.x][0];
}
__global__ void origin_stage2(const double *__restrict__ part, int nb, int64_t n,
                              double *__restrict__ origin) {
  if (threadIdx.x >= 3) return;
  int k = threadIdx.x;
  double sum = 0.0;
  for (int b = 0; b < nb; b++) sum += part[b * 3 + k];
  origin[k] = sum / (double)n;
}
__global__ void pack32_dev_origin(const double *__restrict__ pos, const double *__restrict__ mass,
                                  int64_t n, const double *__restrict__ origin,
                                  float4 *__restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = make_float4((float)(pos[3 * i] - origin[0]), (float)(pos[3 * i + 1] - origin[1]),
                       (float)(pos[3 * i + 2] - origin[2]), mass ? (float)mass[i] : 0.f);
}

static int force_common(int alg, int prec, const double *pos, const double *mass, int64_t np,
                        const double *fpos, int64_t nf, double eps, double theta, double *acc_out,
                        int mem, void *stream) {
  if (prec != GH_PREC_F32 && prec != GH_PREC_F64) { set_error("prec must be 32 or 64"); return GH_EINVAL; }
  if (mem != GH_MEM_HOST && mem != GH_MEM_DEVICE) { set_error("mem must be GH_MEM_HOST or GH_MEM_DEVICE"); return GH_EINVAL; }
  if (np < 0 || nf < 0) { set_error("negative particle count"); return GH_EINVAL; }
  const bool self = (fpos == nullptr);
  if (self) nf = np;
  if (nf == 0) return GH_OK;
  if (!pos || !mass || !acc_out) { set_error("null pointer argument"); return GH_EINVAL; }
  if (!(eps >= 0.0)) { set_error("eps must be >= 0"); return GH_EINVAL; }
  GH_TRY(check_device());
  Stateless *s = stateless();
  if (!s) { set_error("out of host memory"); return GH_ENOMEM; }
  cudaStream_t st = (cudaStream_t)stream;
  if (mem == GH_MEM_HOST && !st) {
    if (!s->stream) GH_CUDA(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    st = s->stream;
  }
  GH_TRY(scratch_acquire(s, st));
  const double *dpos = pos, *dmass = mass, *dt = fpos;
  double *dacc = acc_out;
  if (mem == GH_MEM_HOST) {
    GH_TRY(s->pos.reserve(sizeof(double) * 3 * (size_t)(np > 0 ? np : 1)));
    GH_TRY(s->mass.reserve(sizeof(double) * (size_t)(np > 0 ? np : 1)));
    GH_TRY(s->acc.reserve(sizeof(double) * 3 * (size_t)nf));
    GH_CUDA(cudaMemcpyAsync(s->pos.ptr, pos, sizeof(double) * 3 * np, cudaMemcpyHostToDevice, st));
    GH_CUDA(cudaMemcpyAsync(s->mass.ptr, mass, sizeof(double) * np, cudaMemcpyHostToDevice, st));
    dpos = s->pos.as<double>();
    dmass = s->mass.as<double>();
    dacc = s->acc.as<double>();
    if (!self) {
      GH_TRY(s->tpos.reserve(sizeof(double) * 3 * (size_t)nf));
      GH_CUDA(cudaMemcpyAsync(s->tpos.ptr, fpos, sizeof(double) * 3 * nf, cudaMemcpyHostToDevice, st));
      dt = s->tpos.as<double>();
    }
  }
  if (self) dt = dpos;

  Epilogue ep;
  memset(&ep, 0, sizeof(ep));
  ep.mode = EP_ACC;
  ep.acc_out = dacc;

  if (np == 0) {
    GH_CUDA(cudaMemsetAsync(dacc, 0, sizeof(double) * 3 * nf, st));
  } else if (alg == GH_ALG_DIRECT) {
    DirectArgs a;
    memset(&a, 0, sizeof(a));
    a.prec = prec;
    a.nj = np;
    a.ni = nf;
    a.eps = eps;
    a.ep = ep;
    a.mixed_mass = (mem == GH_MEM_HOST && prec == GH_PREC_F32 && nf >= 131072) ? masses_differ(mass, np) : 0;
    if (prec == GH_PREC_F64) {
      a.src_pos = dpos;
      a.src_mass = dmass;
      a.tgt_pos = dt;
    } else {
      // fp32 pair maths on coordinates taken relative to the sources' mean position
      GH_TRY(s->src32.reserve(sizeof(float4) * (size_t)np));
      GH_TRY(s->root.reserve(sizeof(double) * 4));
      GH_TRY(s->part.reserve(sizeof(double) * 6 * 256));
      int nb = (int)((np + 2047) / 2048);
      if (nb > 256) nb = 256;
      origin_stage1<<<nb, 256, 0, st>>>(dpos, np, s->part.as<double>());
      GH_LAUNCH_CHECK();
      origin_stage2<<<1, 32, 0, st>>>(s->part.as<double>(), nb, np, s->root.as<double>());
      GH_LAUNCH_CHECK();
      pack32_dev_origin<<<(unsigned)((np + 255) / 256), 256, 0, st>>>(dpos, dmass, np, s->root.as<double>(),
                                                                    s->src32.as<float4>());
      GH_LAUNCH_CHECK();
      a.src32 = s->src32.as<float4>();
      if (self) {
        a.tgt32 = a.src32;
      } else {
        GH_TRY(s->tgt32.reserve(sizeof(float4) * (size_t)nf));
        pack32_dev_origin<<<(unsigned)((nf + 255) / 256), 256, 0, st>>>(dt, nullptr, nf, s->root.as<double>(),
                                                                      s->tgt32.as<float4>());
        GH_LAUNCH_CHECK();
        a.tgt32 = s->tgt32.as<float4>();
      }
    }
    GH_TRY(launch_direct(a, s->ws, st, nullptr));
  } else {
    if (!s->tw) s->tw = tree_workspace_create();
    TreeArgs a;
    memset(&a, 0, sizeof(a));
    a.prec = prec;
    a.src_pos = dpos;
    a.src_mass = dmass;
    a.nj = np;
    a.tgt_pos = dt;
    a.ni = nf;
    a.targets_are_sources = self;
    a.eps = eps;
    a.theta = theta;
    a.ep = ep;
    a.want_stats = g_ts.want_stats;
    a.sync_check = true;  // one synchronisation at the END of the call (none inside the evaluation)
    GH_TRY(launch_tree(a, s->tw, st, nullptr));
    tree_last_stats(s->tw, g_ts.tree_stats);
  }
  if (mem == GH_MEM_HOST) {
    GH_CUDA(cudaMemcpyAsync(acc_out, dacc, sizeof(double) * 3 * nf, cudaMemcpyDeviceToHost, st));
    GH_CUDA(cudaStreamSynchronize(st));
  }
  GH_TRY(scratch_release(s, st));
  return GH_OK;
}

}  // namespace gh

using namespace gh;

// ---------------------------------------------------------------------------------------------
// engine
// ---------------------------------------------------------------------------------------------
extern "C" {

const char *gh_last_error(void) { return g_err; }
int gh_version(void) { return 100; }

int gh_device_count(int *n) {
  if (!n) return GH_EINVAL;
  *n = 0;
  cudaError_t e = cudaGetDeviceCount(n);
  if (e != cudaSuccess || *n <= 0) {
    cudaGetLastError();
    *n = 0;
    set_error("no CUDA device available");
    return GH_ECUDA;
  }
  return GH_OK;
}

int gh_direct_summation(int prec, const double *pos, const double *mass, int64_t np, double eps,
                        double *acc_out, int mem, void *stream) {
  return force_common(GH_ALG_DIRECT, prec, pos, mass, np, nullptr, np, eps, 0.0, acc_out, mem, stream);
}
int gh_direct_summation_position(int prec, const double *pos, const double *mass, int64_t np,
                                 const double *force_pos, int64_t nf, double eps, double *acc_out,
                                 int mem, void *stream) {
  if (!force_pos && nf > 0) { set_error("null force_pos"); return GH_EINVAL; }
  if (nf == 0) return GH_OK;
  return force_common(GH_ALG_DIRECT, prec, pos, mass, np, force_pos, nf, eps, 0.0, acc_out, mem, stream);
}
int gh_tree_force(int prec, const double *pos, const double *mass, int64_t np, double eps,
                  double theta, double *acc_out, int mem, void *stream) {
  if (!(theta >= 0.0)) { set_error("theta must be >= 0"); return GH_EINVAL; }
  return force_common(GH_ALG_TREE, prec, pos, mass, np, nullptr, np, eps, theta, acc_out, mem, stream);
}
int gh_tree_force_position(int prec, const double *pos, const double *mass, int64_t np,
                           const double *force_pos, int64_t nf, double eps, double theta,
                           double *acc_out, int mem, void *stream) {
  if (!(theta >= 0.0)) { set_error("theta must be >= 0"); return GH_EINVAL; }
  if (!force_pos && nf > 0) { set_error("null force_pos"); return GH_EINVAL; }
  if (nf == 0) return GH_OK;
  return force_common(GH_ALG_TREE, prec, pos, mass, np, force_pos, nf, eps, theta, acc_out, mem, stream);
}
int gh_release_thread_scratch(void) {
  int cur = 0;
  const bool have_cur = cudaGetDevice(&cur) == cudaSuccess;
  cudaGetLastError();
  for (int d = 0; d < GH_MAX_DEVICES; d++) {
    Stateless *s = g_ts.dev[d];
    if (!s) continue;
    cudaSetDevice(d);
    if (s->stream) cudaStreamSynchronize(s->stream);
    if (s->last_valid) cudaEventSynchronize(s->last_done);
    DeviceBuffer *all[] = {&s->pos, &s->mass, &s->tpos, &s->acc, &s->src32, &s->tgt32, &s->ws, &s->root,
                           &s->part, &s->ictab, &s->icout};
    for (auto *b : all) b->release();
    if (s->tw) { tree_workspace_destroy(s->tw); s->tw = nullptr; }
    if (s->stream) { cudaStreamDestroy(s->stream); s->stream = nullptr; }
    if (s->last_done) { cudaEventDestroy(s->last_done); s->last_done = nullptr; }
    delete s;
    g_ts.dev[d] = nullptr;
  }
  if (have_cur) cudaSetDevice(cur);
  cudaGetLastError();
  return GH_OK;
}

int gh_host_alloc(void **ptr, int64_t bytes) {
  if (!ptr || bytes <= 0) { set_error("gh_host_alloc: bad arguments"); return GH_EINVAL; }
  *ptr = nullptr;
  GH_TRY(check_device());
  cudaError_t e = cudaHostAlloc(ptr, (size_t)bytes, cudaHostAllocPortable);
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("gh_host_alloc: cudaHostAlloc(%lld bytes): %s", (long long)bytes, cudaGetErrorString(e));
    *ptr = nullptr;
    return GH_ENOMEM;
  }
  return GH_OK;
}
int gh_host_free(void *ptr) {
  if (!ptr) return GH_OK;
  cudaError_t e = cudaFreeHost(ptr);
  if (e != cudaSuccess) { cudaGetLastError(); set_error("gh_host_free: %s", cudaGetErrorString(e)); return GH_ECUDA; }
  return GH_OK;
}

int gh_tree_last_stats(int64_t out[8]) {
  if (!out) return GH_EINVAL;
  for (int k = 0; k < 8; k++) out[k] = g_ts.tree_stats[k];
  return GH_OK;
}
int gh_set_tree_stats(int enable) {
  g_ts.want_stats = enable != 0;
  return GH_OK;
}

int gh_set_tree_walk(int mode) {
  if (mode != GH_WALK_TARGET && mode != GH_WALK_GROUP) { set_error("unknown tree walk mode %d", mode); return GH_EINVAL; }
  set_tree_walk_mode(mode);
  return GH_OK;
}
int gh_get_tree_walk(void) { return tree_walk_mode(); }

int gh_set_tree_quadrupoles(int enable) {
  set_tree_quadrupoles(enable);
  return GH_OK;
}
int gh_get_tree_quadrupoles(void) { return tree_quadrupoles(); }

int gh_set_tree_walk_hybrid(double kappa) {
  if (!(kappa >= 0.0) || kappa > 1.0) { set_error("gh_set_tree_walk_hybrid: kappa must be in [0, 1]"); return GH_EINVAL; }
  set_group_hybrid_kappa(kappa);
  return GH_OK;
}
double gh_get_tree_walk_hybrid(void) { return (double)group_hybrid_kappa(); }

int gh_ic_sample(int kind, int64_t n, const double *params, int nparams, const double *table_x,
                 const double *table_y, int ntable, uint64_t seed, double *pos, double *vel,
                 double *mass, int mem, void *stream) {
  if (kind < 1 || kind > 3) { set_error("unknown IC kind %d (1 Plummer, 2 Hernquist, 3 TSIS)", kind); return GH_EINVAL; }
  if (n < 0 || !params || nparams < 2 || nparams > 3) { set_error("bad IC parameters"); return GH_EINVAL; }
  if (kind != 3 && (!table_x || !table_y || ntable < 2)) { set_error("this IC kind needs an inverse-CDF table"); return GH_EINVAL; }
  if (n == 0) return GH_OK;
  if (!pos || !vel || !mass) { set_error("null output pointer"); return GH_EINVAL; }
  GH_TRY(check_device());
  Stateless *s = stateless();
  if (!s) return GH_ENOMEM;
  cudaStream_t st = (cudaStream_t)stream;
  if (mem == GH_MEM_HOST && !st) {
    if (!s->stream) GH_CUDA(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    st = s->stream;
  }
  GH_TRY(scratch_acquire(s, st));
  double prm[3] = {params[0], params[1], nparams > 2 ? params[2] : 0.0};
  const int nt = (kind == 3) ? 0 : ntable;
  GH_TRY(s->ictab.reserve(sizeof(double) * (2 * (size_t)(nt > 0 ? nt : 1) + 3 * 256)));
  double *tx = s->ictab.as<double>(), *ty = tx + (nt > 0 ? nt : 1), *scratch = ty + (nt > 0 ? nt : 1);
  if (nt > 0) {
    GH_CUDA(cudaMemcpyAsync(tx, table_x, sizeof(double) * nt, cudaMemcpyHostToDevice, st));
    GH_CUDA(cudaMemcpyAsync(ty, table_y, sizeof(double) * nt, cudaMemcpyHostToDevice, st));
  }
  double *dpos = pos, *dvel = vel, *dmass = mass;
  if (mem == GH_MEM_HOST) {
    GH_TRY(s->icout.reserve(sizeof(double) * 7 * (size_t)n));
    dpos = s->icout.as<double>();
    dvel = dpos + 3 * n;
    dmass = dvel + 3 * n;
  }
  GH_TRY(launch_ic(kind, n, prm, tx, ty, nt, seed, dpos, dvel, dmass, scratch, st));
  if (mem == GH_MEM_HOST) {
    GH_CUDA(cudaMemcpyAsync(pos, dpos, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, st));
    GH_CUDA(cudaMemcpyAsync(vel, dvel, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, st));
    GH_CUDA(cudaMemcpyAsync(mass, dmass, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
    GH_CUDA(cudaStreamSynchronize(st));
  }
  GH_TRY(scratch_release(s, st));
  return GH_OK;
}

int gh_ic_sample_expdisk(int64_t n, const double *params4, const double *table_R, const double *table_cum,
                         const double *table_vphi, const double *table_ratio, int ntable, uint64_t seed,
                         double *pos, double *vel, double *mass, int mem, void *stream) {
  if (n < 0 || !params4) { set_error("bad IC parameters"); return GH_EINVAL; }
  if (!table_R || !table_cum || !table_vphi || !table_ratio || ntable < 2) { set_error("expdisk needs its four radial tables"); return GH_EINVAL; }
  if (!(params4[1] > 0.0) || !(params4[2] > 0.0)) { set_error("expdisk: Rd and z0 must be positive"); return GH_EINVAL; }
  if (n == 0) return GH_OK;
  if (!pos || !vel || !mass) { set_error("null output pointer"); return GH_EINVAL; }
  GH_TRY(check_device());
  Stateless *s = stateless();
  if (!s) return GH_ENOMEM;
  cudaStream_t st = (cudaStream_t)stream;
  if (mem == GH_MEM_HOST && !st) {
    if (!s->stream) GH_CUDA(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    st = s->stream;
  }
  GH_TRY(scratch_acquire(s, st));
  const size_t nt = (size_t)ntable;
  GH_TRY(s->ictab.reserve(sizeof(double) * (4 * nt + 3 * 256)));
  double *tR = s->ictab.as<double>(), *tcum = tR + nt, *tvphi = tcum + nt, *tratio = tvphi + nt, *scratch = tratio + nt;
  GH_CUDA(cudaMemcpyAsync(tR, table_R, sizeof(double) * nt, cudaMemcpyHostToDevice, st));
  GH_CUDA(cudaMemcpyAsync(tcum, table_cum, sizeof(double) * nt, cudaMemcpyHostToDevice, st));
  GH_CUDA(cudaMemcpyAsync(tvphi, table_vphi, sizeof(double) * nt, cudaMemcpyHostToDevice, st));
  GH_CUDA(cudaMemcpyAsync(tratio, table_ratio, sizeof(double) * nt, cudaMemcpyHostToDevice, st));
  double *dpos = pos, *dvel = vel, *dmass = mass;
  if (mem == GH_MEM_HOST) {
    GH_TRY(s->icout.reserve(sizeof(double) * 7 * (size_t)n));
    dpos = s->icout.as<double>();
    dvel = dpos + 3 * n;
    dmass = dvel + 3 * n;
  }
  GH_TRY(launch_ic_expdisk(n, params4, tR, tcum, tvphi, tratio, ntable, seed, dpos, dvel, dmass, scratch, st));
  if (mem == GH_MEM_HOST) {
    GH_CUDA(cudaMemcpyAsync(pos, dpos, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, st));
    GH_CUDA(cudaMemcpyAsync(vel, dvel, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, st));
    GH_CUDA(cudaMemcpyAsync(mass, dmass, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
    GH_CUDA(cudaStreamSynchronize(st));
  }
  GH_TRY(scratch_release(s, st));
  return GH_OK;
}

int gh_engine_create(gh_engine **out, int device, int64_t n_total, int64_t i_begin, int64_t i_count,
                     int prec) {
  if (!out) return GH_EINVAL;
  *out = nullptr;
  if (prec != GH_PREC_F32 && prec != GH_PREC_F64) { set_error("prec must be 32 or 64"); return GH_EINVAL; }
  if (n_total <= 0 || i_begin < 0 || i_count <= 0 || i_begin + i_count > n_total) {
    set_error("bad particle ranges n_total=%lld i_begin=%lld i_count=%lld", (long long)n_total,
              (long long)i_begin, (long long)i_count);
    return GH_EINVAL;
  }
  GH_TRY(check_device());
  GH_CUDA(cudaSetDevice(device));
  gh_engine *e = new (std::nothrow) gh_engine();
  if (!e) return GH_ENOMEM;
  e->device = device;
  e->n = n_total;
  e->ib = i_begin;
  e->ni = i_count;
  e->prec = prec;
  memset(&e->pots, 0, sizeof(e->pots));
  auto fail = [&](int rc) { gh_engine_destroy(e); return rc; };
#define E_CUDA(call)                                                                   \
  do {                                                                                 \
    cudaError_t e_ = (call);                                                           \
    if (e_ != cudaSuccess) {                                                           \
      set_error("%s: %s", #call, cudaGetErrorString(e_));                              \
      return fail(e_ == cudaErrorMemoryAllocation ? GH_ENOMEM : GH_ECUDA);             \
    }                                                                                  \
  } while (0)
  E_CUDA(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
  E_CUDA(cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
  E_CUDA(cudaMalloc(&e->mass, sizeof(double) * n_total));
  for (int r = 0; r < GH_RING; r++) {
    E_CUDA(cudaMalloc(&e->x[r], sizeof(double) * 3 * i_count));
    E_CUDA(cudaMalloc(&e->v[r], sizeof(double) * 3 * i_count));
    E_CUDA(cudaEventCreateWithFlags(&e->copied[r], cudaEventDisableTiming));
  }
  E_CUDA(cudaEventCreateWithFlags(&e->step_done, cudaEventDisableTiming));
  for (int r = 0; r < gh_engine::FEV_RING; r++) {
    E_CUDA(cudaEventCreate(&e->fev[r][0]));
    E_CUDA(cudaEventCreate(&e->fev[r][1]));
  }
  for (int b = 0; b < 2; b++) {
    E_CUDA(cudaMalloc(&e->src[b], e->src_stride() * n_total));
    E_CUDA(cudaMemset(e->src[b], 0, e->src_stride() * n_total));
  }
  if (prec == GH_PREC_F32) E_CUDA(cudaMalloc(&e->xh_private, sizeof(double) * 3 * i_count));
  E_CUDA(cudaMalloc(&e->d_energy, sizeof(double) * 2));
  E_CUDA(cudaMallocHost(&e->h_maxent, sizeof(int) * gh_engine::MAXENT_RING));
  for (int r = 0; r < gh_engine::MAXENT_RING; r++) {
    e->h_maxent[r] = 0;
    E_CUDA(cudaEventCreateWithFlags(&e->maxent_ev[r], cudaEventDisableTiming));
  }
#undef E_CUDA
  e->tw = tree_workspace_create();
  *out = e;
  return GH_OK;
}

int gh_engine_destroy(gh_engine *e) {
  if (!e) return GH_OK;
  cudaSetDevice(e->device);
  if (e->stream) cudaStreamSynchronize(e->stream);
  if (e->copy_stream) cudaStreamSynchronize(e->copy_stream);
  cudaFree(e->mass);
  for (int r = 0; r < GH_RING; r++) {
    cudaFree(e->x[r]);
    cudaFree(e->v[r]);
    if (e->copied[r]) cudaEventDestroy(e->copied[r]);
  }
  if (e->step_done) cudaEventDestroy(e->step_done);
  for (int r = 0; r < gh_engine::FEV_RING; r++) {
    if (e->fev[r][0]) cudaEventDestroy(e->fev[r][0]);
    if (e->fev[r][1]) cudaEventDestroy(e->fev[r][1]);
  }
  if (!e->external_src) { cudaFree(e->src[0]); cudaFree(e->src[1]); }
  cudaFree(e->xh_private);
  cudaFree(e->d_energy);
  if (e->h_maxent) cudaFreeHost(e->h_maxent);
  for (int r = 0; r < gh_engine::MAXENT_RING; r++)
    if (e->maxent_ev[r]) cudaEventDestroy(e->maxent_ev[r]);
  e->ws.release();
  e->ext.release();
  e->ext2.release();
  tree_workspace_destroy(e->tw);
  if (e->stream) cudaStreamDestroy(e->stream);
  if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
  delete e;
  return GH_OK;
}

int gh_engine_set_origin(gh_engine *e, const double origin[3]) {
  if (!e || !origin) return GH_EINVAL;
  for (int k = 0; k < 3; k++) e->origin[k] = origin[k];
  e->xhalf_valid = false;
  return GH_OK;
}
int gh_engine_set_origin_velocity(gh_engine *e, const double vel[3]) {
  if (!e || !vel) return GH_EINVAL;
  for (int k = 0; k < 3; k++) e->origin_vel[k] = vel[k];
  return GH_OK;
}

int gh_engine_clear_potentials(gh_engine *e) {
  if (!e) return GH_EINVAL;
  memset(&e->pots, 0, sizeof(e->pots));
  return GH_OK;
}
int gh_engine_add_potential(gh_engine *e, int kind, const double *params, int nparams) {
  if (!e || !params) return GH_EINVAL;
  if (kind < 1 || kind > 5) { set_error("unknown potential kind %d", kind); return GH_EINVAL; }
  if (nparams < 1 || nparams > GH_POT_NPARAM) { set_error("potential takes 1..%d parameters", GH_POT_NPARAM); return GH_EINVAL; }
  if (e->pots.n >= GH_MAX_POTENTIALS) { set_error("at most %d native potentials", GH_MAX_POTENTIALS); return GH_EINVAL; }
  const int k = e->pots.n++;
  e->pots.kind[k] = kind;
  for (int j = 0; j < GH_POT_NPARAM; j++) e->pots.prm[k][j] = (j < nparams) ? params[j] : 0.0;
  if ((kind == 4) && e->pots.prm[k][5] == 0.0) e->pots.prm[k][5] = 1.0;  // q defaults to spherical
  return GH_OK;
}

int gh_engine_upload(gh_engine *e, const double *pos, const double *vel, const double *mass_all) {
  GH_ENGINE_GUARD(e);
  if (!pos || !vel || !mass_all) { set_error("null pointer argument"); return GH_EINVAL; }
  GH_CUDA(cudaStreamSynchronize(e->copy_stream));
  for (int r = 0; r < GH_RING; r++) e->copy_pending[r] = false;
  GH_CUDA(cudaMemcpyAsync(e->x[e->cur], pos, sizeof(double) * 3 * e->ni, cudaMemcpyHostToDevice, e->stream));
  GH_CUDA(cudaMemcpyAsync(e->v[e->cur], vel, sizeof(double) * 3 * e->ni, cudaMemcpyHostToDevice, e->stream));
  GH_CUDA(cudaMemcpyAsync(e->mass, mass_all, sizeof(double) * e->n, cudaMemcpyHostToDevice, e->stream));
  e->mixed_mass = masses_differ(mass_all, e->n);
  GH_CUDA(cudaStreamSynchronize(e->stream));
  e->uploaded = true;
  e->xhalf_valid = false;
  e->dist_ready = false;
  tree_forget_history(e->tw);
  return GH_OK;
}

int gh_engine_upload_device(gh_engine *e, const double *pos, const double *vel, const double *mass_all) {
  GH_ENGINE_GUARD(e);
  if (!pos || !vel || !mass_all) { set_error("null pointer argument"); return GH_EINVAL; }
  GH_CUDA(cudaStreamSynchronize(e->copy_stream));
  for (int r = 0; r < GH_RING; r++) e->copy_pending[r] = false;
  GH_CUDA(cudaMemcpyAsync(e->x[e->cur], pos, sizeof(double) * 3 * e->ni, cudaMemcpyDeviceToDevice, e->stream));
  GH_CUDA(cudaMemcpyAsync(e->v[e->cur], vel, sizeof(double) * 3 * e->ni, cudaMemcpyDeviceToDevice, e->stream));
  GH_CUDA(cudaMemcpyAsync(e->mass, mass_all, sizeof(double) * e->n, cudaMemcpyDeviceToDevice, e->stream));
  GH_CUDA(cudaStreamSynchronize(e->stream));
  e->uploaded = true;
  e->xhalf_valid = false;
  e->dist_ready = false;
  tree_forget_history(e->tw);
  return GH_OK;
}

int gh_engine_bind_sources(gh_engine *e, void *buf0, void *buf1) {
  GH_ENGINE_GUARD(e);
  if (!buf0 || !buf1 || buf0 == buf1) { set_error("need two distinct source buffers"); return GH_EINVAL; }
  GH_CUDA(cudaStreamSynchronize(e->stream));
  if (!e->external_src) { cudaFree(e->src[0]); cudaFree(e->src[1]); }
  e->src[0] = buf0;
  e->src[1] = buf1;
  e->external_src = true;
  e->xhalf_valid = false;
  return GH_OK;
}
int gh_engine_source_index(gh_engine *e, int *idx) {
  if (!e || !idx) return GH_EINVAL;
  *idx = e->scur;
  return GH_OK;
}
int gh_engine_source_stride_bytes(gh_engine *e, int64_t *bytes) {
  if (!e || !bytes) return GH_EINVAL;
  *bytes = (int64_t)e->src_stride();
  return GH_OK;
}

// x_half for the next step from the current state (gravhopper.py:409), into the owned slice of
// the current source buffer.
int gh_engine_prepare(gh_engine *e, double dt) {
  GH_ENGINE_GUARD(e);
  if (!e->uploaded) { set_error("engine has no state: call gh_engine_upload first"); return GH_ESTATE; }
  GH_TRY(launch_half_drift(e->x[e->cur], e->v[e->cur], e->mass + e->ib, e->ni, dt,
                           e->xhalf_own(e->scur), e->src32_own(e->scur), e->origin, e->stream));
  e->dt_built = dt;
  e->xhalf_valid = true;
  return GH_OK;
}

int gh_engine_set_dt(gh_engine *e, double dt) { return gh_engine_prepare(e, dt); }

}  // extern "C"

namespace gh {
int engine_step_args(gh_engine *e, double dt, double eps, double theta, int algorithm, const double *ext_dev,
                     StepArgs *out) {
  if (!e->uploaded) { set_error("engine has no state: call gh_engine_upload first"); return GH_ESTATE; }
  if (algorithm != GH_ALG_DIRECT && algorithm != GH_ALG_TREE) { set_error("unknown algorithm %d", algorithm); return GH_EINVAL; }
  if (!e->xhalf_valid || dt != e->dt_built) {
    if (e->n != e->ni) {
      // multi-GPU: the caller must gh_engine_prepare + all-gather before stepping
      set_error("sharded engine: call gh_engine_prepare(dt) and all-gather the sources before gh_engine_step");
      return GH_ESTATE;
    }
    GH_TRY(gh_engine_prepare(e, dt));
  }
  const int nxt = (e->cur + 1) % GH_RING;
  if (e->copy_pending[nxt]) {
    GH_CUDA(cudaStreamWaitEvent(e->stream, e->copied[nxt], 0));
    e->copy_pending[nxt] = false;
  }
  if (e->pots.n > 0) {  // native potentials -> external-acceleration buffer, on the device
    GH_TRY(e->ext2.reserve(sizeof(double) * 3 * (size_t)e->ni));
    GH_TRY(launch_potentials(e->pots, e->xhalf_own(e->scur), e->ni, ext_dev, e->ext2.as<double>(), e->stream));
    ext_dev = e->ext2.as<double>();
  }
  Epilogue ep;
  memset(&ep, 0, sizeof(ep));
  ep.mode = EP_STEP;
  ep.xhalf = e->xhalf_own(e->scur);
  ep.v_in = e->v[e->cur];
  ep.x_out = e->x[nxt];
  ep.v_out = e->v[nxt];
  ep.xhalf_next = e->xhalf_own(e->scur ^ 1);
  ep.src32_next = e->src32_own(e->scur ^ 1);
  ep.mass = e->mass + e->ib;
  ep.ext = ext_dev;
  ep.dt = dt;
  // the next step's float4 sources are written relative to where the origin will be then
  for (int k = 0; k < 3; k++) {
    e->origin_next[k] = e->origin[k] + e->origin_vel[k] * dt * GH_KPC_PER_KMS_MYR;
    ep.origin[k] = e->origin_next[k];
  }

  // gravhopper.py:449-450 (Np == 1 feels no N-body force) needs no special case: the only source
  // is the target itself and its term is exactly zero in every kernel.
  memset(out, 0, sizeof(*out));
  out->algorithm = algorithm;
  if (algorithm == GH_ALG_DIRECT) {
    DirectArgs &a = out->direct;
    a.prec = e->prec;
    a.nj = e->n;
    a.ni = e->ni;
    a.eps = eps;
    a.ep = ep;
    a.mixed_mass = e->mixed_mass ? 1 : 0;
    if (e->prec == GH_PREC_F64) {
      a.src_pos = reinterpret_cast<const double *>(e->src[e->scur]);
      a.src_mass = e->mass;
      a.tgt_pos = a.src_pos + 3 * e->ib;
    } else {
      a.src32 = reinterpret_cast<const float4 *>(e->src[e->scur]);
      a.tgt32 = a.src32 + e->ib;
    }
  } else {
    TreeArgs &a = out->tree;
    a.prec = e->prec;
    a.nj = e->n;
    a.ni = e->ni;
    a.eps = eps;
    a.theta = theta;
    a.ep = ep;
    a.targets_are_sources = true;
    a.tgt_offset = e->ib;
    a.coherent = true;
    if (e->prec == GH_PREC_F64) {
      a.src_pos = reinterpret_cast<const double *>(e->src[e->scur]);
      a.src_mass = e->mass;
      a.tgt_pos = a.src_pos + 3 * e->ib;
    } else {
      a.src32 = reinterpret_cast<const float4 *>(e->src[e->scur]);
      a.tgt32 = a.src32 + e->ib;
    }
  }
  return GH_OK;
}

void engine_step_done(gh_engine *e) {
  for (int k = 0; k < 3; k++) e->origin[k] = e->origin_next[k];
  e->fev_count++;
  e->cur = (e->cur + 1) % GH_RING;
  e->scur ^= 1;
}

int engine_step_impl(gh_engine *e, double dt, double eps, double theta, int algorithm, const double *ext_dev) {
  StepArgs sa;
  GH_TRY(engine_step_args(e, dt, eps, theta, algorithm, ext_dev, &sa));
  cudaEvent_t *ev = e->fev[e->fev_count % gh_engine::FEV_RING];
  if (algorithm == GH_ALG_DIRECT) GH_TRY(launch_direct(sa.direct, e->ws, e->stream, ev));
  else GH_TRY(launch_tree(sa.tree, e->tw, e->stream, ev));
  engine_step_done(e);
  return GH_OK;
}

// The engine never synchronises inside a tree step, so an entry-array overflow (the walk then
// leaves the state untouched) can only be reported when the host next waits for the stream.
int engine_check_tree(gh_engine *e) {
  int64_t entries = 0;
  if (e->tw && tree_poll_overflow(e->tw, &entries) < 0) {
    set_error("tree: the entry array overflowed during a step (%lld entries needed); the state stopped advancing "
              "there -- upload it again and rerun (capacity grows with the largest count seen)", (long long)entries);
    return GH_ESTATE;
  }
  return GH_OK;
}
}  // namespace gh

extern "C" {

int gh_engine_step(gh_engine *e, double dt, double eps, double theta, int algorithm,
                   const double *ext_acc, int ext_mem) {
  GH_ENGINE_GUARD(e);
  const double *ext_dev = ext_acc;
  if (ext_acc && ext_mem == GH_MEM_HOST) {
    GH_TRY(e->ext.reserve(sizeof(double) * 3 * e->ni));
    GH_CUDA(cudaMemcpyAsync(e->ext.ptr, ext_acc, sizeof(double) * 3 * e->ni, cudaMemcpyHostToDevice, e->stream));
    ext_dev = e->ext.as<double>();
  }
  return engine_step_impl(e, dt, eps, theta, algorithm, ext_dev);
}

int gh_engine_run(gh_engine *e, int64_t nsteps, double dt, double eps, double theta, int algorithm,
                  int64_t snapshot_every, double *pos_hist, double *vel_hist) {
  GH_ENGINE_GUARD(e);
  if (nsteps < 0 || snapshot_every < 0) { set_error("negative step count"); return GH_EINVAL; }
  if (e->n != e->ni) { set_error("gh_engine_run is single-GPU only; drive sharded engines with gh_engine_step"); return GH_ESTATE; }
  if (snapshot_every > 0 && (!pos_hist || !vel_hist)) { set_error("snapshot buffers are null"); return GH_EINVAL; }
  const size_t row = sizeof(double) * 3 * (size_t)e->ni;
  int64_t nsnap = 0;
  if (snapshot_every > 0) nsnap = nsteps / snapshot_every + ((nsteps % snapshot_every) ? 1 : 0);
  bool reg_p = false, reg_v = false;
  if (nsnap > 0) {
    // pin the caller's history arrays so the copies really overlap the steps -- but never more than
    // GH_PIN_HISTORY_MAX bytes (default 4 GiB) at a time: page-locking tens of GB is slow and can
    // starve the host.  Beyond the cap, or when registration fails, the snapshots are ordinary
    // staged copies (still asynchronous to the host thread, no longer overlapping the kernels);
    // gh_engine_history_pinned() reports which it was.
    size_t cap = (size_t)4 << 30;
    if (const char *env = getenv("GH_PIN_HISTORY_MAX")) cap = (size_t)strtoull(env, nullptr, 10);
    if (2 * row * (size_t)nsnap <= cap) {
      reg_p = cudaHostRegister(pos_hist, row * nsnap, cudaHostRegisterDefault) == cudaSuccess;
      reg_v = cudaHostRegister(vel_hist, row * nsnap, cudaHostRegisterDefault) == cudaSuccess;
      cudaGetLastError();
    }
    e->history_pinned = reg_p && reg_v;
  }
  int rc = GH_OK;
  int64_t k = 0;
  for (int64_t s = 1; s <= nsteps && rc == GH_OK; s++) {
    rc = engine_step_impl(e, dt, eps, theta, algorithm, nullptr);
    if (rc != GH_OK) break;
    if (snapshot_every > 0 && (s % snapshot_every == 0 || s == nsteps)) {
      cudaError_t ce = cudaEventRecord(e->step_done, e->stream);
      if (ce == cudaSuccess) ce = cudaStreamWaitEvent(e->copy_stream, e->step_done, 0);
      if (ce == cudaSuccess) ce = cudaMemcpyAsync(pos_hist + 3 * e->ni * k, e->x[e->cur], row, cudaMemcpyDeviceToHost, e->copy_stream);
      if (ce == cudaSuccess) ce = cudaMemcpyAsync(vel_hist + 3 * e->ni * k, e->v[e->cur], row, cudaMemcpyDeviceToHost, e->copy_stream);
      if (ce == cudaSuccess) ce = cudaEventRecord(e->copied[e->cur], e->copy_stream);
      if (ce != cudaSuccess) { set_error("snapshot copy: %s", cudaGetErrorString(ce)); rc = GH_ECUDA; break; }
      e->copy_pending[e->cur] = true;
      k++;
    }
  }
  cudaError_t c1 = cudaStreamSynchronize(e->stream);
  cudaError_t c2 = cudaStreamSynchronize(e->copy_stream);
  for (int r = 0; r < GH_RING; r++) e->copy_pending[r] = false;
  if (reg_p) cudaHostUnregister(pos_hist);
  if (reg_v) cudaHostUnregister(vel_hist);
  if (rc == GH_OK && (c1 != cudaSuccess || c2 != cudaSuccess)) {
    set_error("gh_engine_run: %s", cudaGetErrorString(c1 != cudaSuccess ? c1 : c2));
    rc = GH_ECUDA;
  }
  if (rc == GH_OK && algorithm == GH_ALG_TREE) rc = engine_check_tree(e);
  return rc;
}

int gh_engine_download(gh_engine *e, double *pos, double *vel) {
  GH_ENGINE_GUARD(e);
  if (!e->uploaded) { set_error("engine has no state"); return GH_ESTATE; }
  if (pos) GH_CUDA(cudaMemcpyAsync(pos, e->x[e->cur], sizeof(double) * 3 * e->ni, cudaMemcpyDeviceToHost, e->stream));
  if (vel) GH_CUDA(cudaMemcpyAsync(vel, e->v[e->cur], sizeof(double) * 3 * e->ni, cudaMemcpyDeviceToHost, e->stream));
  GH_CUDA(cudaStreamSynchronize(e->stream));
  return engine_check_tree(e);
}

int gh_engine_download_xhalf(gh_engine *e, double *xhalf) {
  GH_ENGINE_GUARD(e);
  if (!e->uploaded || !e->xhalf_valid) { set_error("x_half not built: call gh_engine_prepare(dt)"); return GH_ESTATE; }
  GH_CUDA(cudaMemcpyAsync(xhalf, e->xhalf_own(e->scur), sizeof(double) * 3 * e->ni, cudaMemcpyDeviceToHost, e->stream));
  GH_CUDA(cudaStreamSynchronize(e->stream));
  return GH_OK;
}

int gh_engine_energy(gh_engine *e, double eps, double out[2]) {
  GH_ENGINE_GUARD(e);
  if (!e->uploaded || !out) { set_error("engine has no state"); return GH_ESTATE; }
  // sources: positions of ALL particles.  Single GPU: the state itself.  Sharded engines only
  // hold x_half of the others, so the diagnostic is defined for single-GPU engines.
  if (e->n != e->ni) { set_error("gh_engine_energy: single-GPU engines only"); return GH_ESTATE; }
  GH_CUDA(cudaMemsetAsync(e->d_energy, 0, 2 * sizeof(double), e->stream));
  GH_TRY(launch_energy(e->x[e->cur], e->v[e->cur], e->mass, e->ni, e->x[e->cur], e->mass, e->n, 0, eps,
                       e->d_energy, e->stream));
  GH_CUDA(cudaMemcpyAsync(out, e->d_energy, 2 * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
  GH_CUDA(cudaStreamSynchronize(e->stream));
  return GH_OK;
}

int gh_engine_synchronize(gh_engine *e) {
  GH_ENGINE_GUARD(e);
  GH_CUDA(cudaStreamSynchronize(e->stream));
  GH_CUDA(cudaStreamSynchronize(e->copy_stream));
  return engine_check_tree(e);
}

int gh_engine_state_ptrs(gh_engine *e, double **pos_dev, double **vel_dev) {
  if (!e) return GH_EINVAL;
  if (pos_dev) *pos_dev = e->x[e->cur];
  if (vel_dev) *vel_dev = e->v[e->cur];
  return GH_OK;
}
int gh_engine_stream(gh_engine *e, void **stream) {
  if (!e || !stream) return GH_EINVAL;
  *stream = (void *)e->stream;
  return GH_OK;
}
int gh_engine_history_pinned(gh_engine *e, int *pinned) {
  if (!e || !pinned) return GH_EINVAL;
  *pinned = e->history_pinned ? 1 : 0;
  return GH_OK;
}
int gh_engine_launch_count(gh_engine *e, int64_t *count) {
  if (!e || !count) return GH_EINVAL;
  *count = e->launches;
  return GH_OK;
}
int gh_engine_last_force_ms(gh_engine *e, float *ms) {
  GH_ENGINE_GUARD(e);
  if (!ms) return GH_EINVAL;
  if (e->fev_count <= 0) { set_error("no step has run"); return GH_ESTATE; }
  cudaEvent_t *ev = e->fev[(e->fev_count - 1) % gh_engine::FEV_RING];
  GH_CUDA(cudaEventSynchronize(ev[1]));
  GH_CUDA(cudaEventElapsedTime(ms, ev[0], ev[1]));
  return GH_OK;
}
int gh_engine_force_ms_mean(gh_engine *e, int last_k, float *mean_ms, int *count) {
  GH_ENGINE_GUARD(e);
  if (!mean_ms || last_k <= 0) { set_error("gh_engine_force_ms_mean: bad arguments"); return GH_EINVAL; }
  if (e->fev_count <= 0) { set_error("no step has run"); return GH_ESTATE; }
  int64_t k = last_k;
  if (k > e->fev_count) k = e->fev_count;
  if (k > gh_engine::FEV_RING) k = gh_engine::FEV_RING;
  double sum = 0.0;
  for (int64_t j = 0; j < k; j++) {
    cudaEvent_t *ev = e->fev[(e->fev_count - 1 - j) % gh_engine::FEV_RING];
    float ms = 0.f;
    GH_CUDA(cudaEventSynchronize(ev[1]));
    GH_CUDA(cudaEventElapsedTime(&ms, ev[0], ev[1]));
    sum += ms;
  }
  *mean_ms = (float)(sum / (double)k);
  if (count) *count = (int)k;
  return GH_OK;
}
int gh_engine_tree_stats(gh_engine *e, int64_t out[8]) {
  if (!e || !out) return GH_EINVAL;
  return tree_last_stats(e->tw, out);
}

}  // extern "C"
