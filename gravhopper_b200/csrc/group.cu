// group.cu -- several engines (one per GPU) stepping one system together: SURVEY 8(e).
//
// Targets are sharded evenly over the ranks; every rank holds all sources.  Per step:
//   direct summation   one in-place NCCL all-gather of the half-drifted positions, then every rank
//                      tiles all sources against its own targets;
//   tree, fp32         the same all-gather, then the DISTRIBUTED build (build.cuh): the Morton key
//                      space is cut into `world` ranges, each rank sorts / scans / emits only its
//                      range into its segment of the global pre-order entry array; two small
//                      all-gathers (boundary keys; cell-end tables, totals, key samples) stitch the
//                      ranges together and one large in-place all-gather shares the entries; every
//                      rank then walks the whole tree for its own targets.  The first tree step
//                      after an upload builds redundantly once (every rank the whole tree): that
//                      gives the initial key ranges and the segment size; afterwards no step
//                      synchronises the host with the device;
//   tree, fp64         all-gather + redundant build (the fp64 tree is the parity path).
//
// One group object drives either all GPUs of a box from one process (gh_group_create_local,
// ncclCommInitAll: `Simulation(devices=N)`) or one GPU of a torchrun-style job
// (gh_group_create_rank, ncclCommInitRank with an id the caller distributes).  NCCL is loaded with
// dlopen (nccl_dl.cuh); nothing here needs torch.
#include "engine.cuh"
#include "nccl_dl.cuh"

#include <dlfcn.h>

#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

namespace gh {

static NcclApi g_nccl;
static bool g_nccl_loaded = false;

const NcclApi *nccl_api(const char *path) {
  if (g_nccl_loaded) return &g_nccl;
  const char *cands[4] = {path, getenv("GH_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  void *h = nullptr;
  for (int k = 0; k < 4 && !h; k++)
    if (cands[k] && cands[k][0]) h = dlopen(cands[k], RTLD_NOW | RTLD_GLOBAL);
  if (!h) { set_error("NCCL not found (dlopen libnccl.so.2: %s); set GH_NCCL_LIB", dlerror()); return nullptr; }
#define GH_SYM(field, name)                                                          \
  do {                                                                               \
    *(void **)(&g_nccl.field) = dlsym(h, name);                                      \
    if (!g_nccl.field) { set_error("NCCL symbol %s missing", name); return nullptr; } \
  } while (0)
  GH_SYM(GetVersion, "ncclGetVersion");
  GH_SYM(GetUniqueId, "ncclGetUniqueId");
  GH_SYM(CommInitRank, "ncclCommInitRank");
  GH_SYM(CommInitAll, "ncclCommInitAll");
  GH_SYM(CommDestroy, "ncclCommDestroy");
  GH_SYM(AllGather, "ncclAllGather");
  GH_SYM(Broadcast, "ncclBroadcast");
  GH_SYM(GroupStart, "ncclGroupStart");
  GH_SYM(GroupEnd, "ncclGroupEnd");
  GH_SYM(GetErrorString, "ncclGetErrorString");
#undef GH_SYM
  g_nccl_loaded = true;
  return &g_nccl;
}

#define GH_NCCL(api, call)                                                                      \
  do {                                                                                          \
    int r_ = (api)->call;                                                                       \
    if (r_ != 0) { set_error("NCCL %s: %s", #call, (api)->GetErrorString(r_)); return GH_ECUDA; } \
  } while (0)

const int *tree_maxent_ptr(TreeWorkspace *w);  // tree.cu

}  // namespace gh

using namespace gh;

static constexpr int GROUP_EVENTS = 11;

struct gh_group {
  int world = 1, rank0 = 0;
  int64_t n = 0;
  int prec = GH_PREC_F32;
  std::vector<gh_engine *> eng;  // local engines: eng[k] is rank rank0 + k
  std::vector<ncclComm_t> comm;
  std::vector<int64_t> begin, count;  // partition over ALL ranks
  const NcclApi *api = nullptr;
  bool even = true;
  bool tree_dist = true;
  cudaEvent_t pev[GROUP_EVENTS] = {};  // phase boundaries of the last step on local engine 0
  bool pev_valid = false;
  bool pev_dist = false;
};

namespace {

struct EngineScope {  // current device + launch accounting for one engine
  gh_engine *e;
  int64_t before;
  explicit EngineScope(gh_engine *e_) : e(e_), before(launch_counter()) { cudaSetDevice(e->device); }
  ~EngineScope() { e->launches += launch_counter() - before; }
};

void make_partition(gh_group *g) {
  g->begin.resize(g->world);
  g->count.resize(g->world);
  const int64_t base = g->n / g->world, rem = g->n % g->world;
  int64_t b = 0;
  for (int r = 0; r < g->world; r++) {
    g->count[r] = base + (r < rem ? 1 : 0);
    g->begin[r] = b;
    b += g->count[r];
  }
  g->even = rem == 0;
}

int64_t round_up(int64_t v, int64_t m) { return (v + m - 1) / m * m; }

// in-place all-gather of `bytes` per rank at ptr + rank * bytes, on every local engine's stream
int all_gather_slots(gh_group *g, int which) {
  if (g->world == 1) return GH_OK;
  GH_NCCL(g->api, GroupStart());
  for (size_t k = 0; k < g->eng.size(); k++) {
    gh_engine *e = g->eng[k];
    cudaSetDevice(e->device);
    void *ptr = nullptr;
    int64_t bytes = 0;
    GH_TRY(tree_exchange_buffer(e->tw, which, &ptr, &bytes));
    GH_NCCL(g->api, AllGather((const char *)ptr + (size_t)(g->rank0 + (int)k) * (size_t)bytes, ptr, (size_t)bytes, 0,
                              g->comm[k], e->stream));
  }
  GH_NCCL(g->api, GroupEnd());
  return GH_OK;
}

// every rank's slice of the current source buffer -> every rank
int all_gather_sources(gh_group *g) {
  if (g->world == 1) return GH_OK;
  GH_NCCL(g->api, GroupStart());
  for (size_t k = 0; k < g->eng.size(); k++) {
    gh_engine *e = g->eng[k];
    cudaSetDevice(e->device);
    char *buf = reinterpret_cast<char *>(e->src[e->scur]);
    const size_t st = e->src_stride();
    if (g->even) {
      GH_NCCL(g->api, AllGather(buf + (size_t)e->ib * st, buf, (size_t)e->ni * st, 0, g->comm[k], e->stream));
    } else {  // n % world != 0: one broadcast per owner
      for (int q = 0; q < g->world; q++)
        GH_NCCL(g->api, Broadcast(buf + (size_t)g->begin[q] * st, buf + (size_t)g->begin[q] * st,
                                  (size_t)g->count[q] * st, 0, q, g->comm[k], e->stream));
    }
  }
  GH_NCCL(g->api, GroupEnd());
  return GH_OK;
}

void mark(gh_group *g, int k) {
  gh_engine *e = g->eng[0];
  cudaSetDevice(e->device);
  cudaEventRecord(g->pev[k], e->stream);
}

int tree_step_distributed(gh_group *g, double dt, double eps, double theta) {
  const size_t nl = g->eng.size();
  gh_engine *e0 = g->eng[0];
  if (!e0->dist_ready) {
    // bootstrap: one redundant build (every rank the whole tree) gives equal-count key ranges and
    // the entry count the segments are sized from -- the only host/device synchronisation
    for (size_t k = 0; k < nl; k++) {
      gh_engine *e = g->eng[k];
      EngineScope sc(e);
      GH_TRY(engine_step_impl(e, dt, eps, theta, GH_ALG_TREE, nullptr));
      GH_TRY(tree_splitters(e->tw, g->world, e->stream));
    }
    for (size_t k = 0; k < nl; k++) {
      gh_engine *e = g->eng[k];
      EngineScope sc(e);
      GH_CUDA(cudaStreamSynchronize(e->stream));
      int64_t entries = 0;
      if (tree_poll_overflow(e->tw, &entries) < 0) return engine_check_tree(e);
      e->dist_stride = round_up(entries / g->world + entries / (6 * g->world) + 4096, 1024);
      e->dist_ready = true;
      e->dist_steps = 0;
    }
    g->pev_dist = false;
    return GH_OK;
  }
  // segment size: grown ahead of need from the largest fill seen two steps ago (every rank reads
  // the same number out of the same gathered records, so all ranks change size in the same step)
  for (size_t k = 0; k < nl; k++) {
    gh_engine *e = g->eng[k];
    if (e->dist_steps >= 2) {
      const int slot = (int)((e->dist_steps - 2) % gh_engine::MAXENT_RING);
      cudaSetDevice(e->device);
      GH_CUDA(cudaEventSynchronize(e->maxent_ev[slot]));
      const int64_t m = e->h_maxent[slot];
      if (m > e->dist_stride) {
        if (m == 0x7fffffff)
          set_error("tree: a rank's key range held more particles than its arrays (1.25x its share): the system "
                    "changed too violently within one step; the state stopped advancing there -- upload it again "
                    "and rerun (GH_TREE_DIST=0 builds redundantly)");
        else
          set_error("tree: a rank's entry segment overflowed (%lld entries, segment %lld); the state stopped "
                    "advancing there -- upload it again and rerun", (long long)m, (long long)e->dist_stride);
        return GH_ESTATE;
      }
      if (100 * m > 95 * e->dist_stride) e->dist_stride = round_up(m + m / 5, 1024);
    }
  }
  std::vector<StepArgs> sa(nl);
  for (size_t k = 0; k < nl; k++) {
    gh_engine *e = g->eng[k];
    EngineScope sc(e);
    GH_TRY(engine_step_args(e, dt, eps, theta, GH_ALG_TREE, nullptr, &sa[k]));
    // a rank's key range holds N / world particles to within the sampling granularity of the
    // splitters (1/64 of a rank's share): arrays, grids and the gathered buffers for 1.25x that
    int64_t ncap = round_up(g->n / g->world + g->n / (4 * g->world) + 4096, 2048);
    if (ncap > g->n) ncap = g->n;
    TreeDist d{g->rank0 + (int)k, g->world, e->dist_stride, ncap, 0};
    GH_TRY(launch_tree_phase(sa[k].tree, e->tw, e->stream, nullptr, &d, 0));
  }
  mark(g, 2);
  GH_TRY(all_gather_slots(g, 1));
  mark(g, 3);
  for (size_t k = 0; k < nl; k++) {
    gh_engine *e = g->eng[k];
    EngineScope sc(e);
    GH_TRY(launch_tree_phase(sa[k].tree, e->tw, e->stream, nullptr, nullptr, 1));
  }
  mark(g, 4);
  GH_TRY(all_gather_slots(g, 2));
  mark(g, 5);
  for (size_t k = 0; k < nl; k++) {
    gh_engine *e = g->eng[k];
    EngineScope sc(e);
    GH_TRY(launch_tree_phase(sa[k].tree, e->tw, e->stream, nullptr, nullptr, 2));
  }
  mark(g, 6);
  GH_TRY(all_gather_slots(g, 3));
  GH_TRY(all_gather_slots(g, 4));
  mark(g, 7);
  for (size_t k = 0; k < nl; k++) {
    gh_engine *e = g->eng[k];
    EngineScope sc(e);
    GH_TRY(launch_tree_phase(sa[k].tree, e->tw, e->stream, e->fev[e->fev_count % gh_engine::FEV_RING], nullptr, 3));
  }
  mark(g, 8);
  GH_TRY(all_gather_slots(g, 5));
  mark(g, 9);
  for (size_t k = 0; k < nl; k++) {
    gh_engine *e = g->eng[k];
    EngineScope sc(e);
    GH_TRY(launch_tree_phase(sa[k].tree, e->tw, e->stream, nullptr, nullptr, 4));
    const int slot = (int)(e->dist_steps % gh_engine::MAXENT_RING);
    GH_CUDA(cudaMemcpyAsync(&e->h_maxent[slot], tree_maxent_ptr(e->tw), sizeof(int), cudaMemcpyDeviceToHost, e->stream));
    GH_CUDA(cudaEventRecord(e->maxent_ev[slot], e->stream));
    e->dist_steps++;
    engine_step_done(e);
  }
  g->pev_dist = true;
  return GH_OK;
}

}  // namespace

extern "C" {

int gh_nccl_version(int *version) {
  if (!version) return GH_EINVAL;
  const NcclApi *api = nccl_api(nullptr);
  if (!api) return GH_ECUDA;
  return api->GetVersion(version) == 0 ? GH_OK : GH_ECUDA;
}

int gh_group_unique_id(void *id128) {
  if (!id128) { set_error("gh_group_unique_id: null"); return GH_EINVAL; }
  const NcclApi *api = nccl_api(nullptr);
  if (!api) return GH_ECUDA;
  ncclUniqueId id;
  GH_NCCL(api, GetUniqueId(&id));
  memcpy(id128, id.internal, 128);
  return GH_OK;
}

int gh_group_destroy(gh_group *g) {
  if (!g) return GH_OK;
  for (size_t k = 0; k < g->eng.size(); k++) {
    if (g->eng[k]) {
      cudaSetDevice(g->eng[k]->device);
      cudaStreamSynchronize(g->eng[k]->stream);
    }
  }
  for (size_t k = 0; k < g->comm.size(); k++)
    if (g->comm[k] && g->api) g->api->CommDestroy(g->comm[k]);
  for (size_t k = 0; k < g->eng.size(); k++) gh_engine_destroy(g->eng[k]);
  for (int k = 0; k < GROUP_EVENTS; k++)
    if (g->pev[k]) cudaEventDestroy(g->pev[k]);
  delete g;
  return GH_OK;
}

static int group_common(gh_group *g, int64_t n_total, int prec) {
  g->n = n_total;
  g->prec = prec;
  if (const char *env = getenv("GH_TREE_DIST")) g->tree_dist = atoi(env) != 0;
  make_partition(g);
  return GH_OK;
}

int gh_group_create_local(gh_group **out, int ndev, const int *devices, int64_t n_total, int prec) {
  if (!out) return GH_EINVAL;
  *out = nullptr;
  if (ndev < 1 || n_total < ndev) { set_error("gh_group_create_local: need 1 <= ndev <= n_total"); return GH_EINVAL; }
  int have = 0;
  if (cudaGetDeviceCount(&have) != cudaSuccess || have < ndev) {
    cudaGetLastError();
    set_error("gh_group_create_local: %d devices requested, %d visible", ndev, have);
    return GH_ECUDA;
  }
  gh_group *g = new (std::nothrow) gh_group();
  if (!g) return GH_ENOMEM;
  g->world = ndev;
  g->rank0 = 0;
  group_common(g, n_total, prec);
  std::vector<int> devs(ndev);
  for (int k = 0; k < ndev; k++) devs[k] = devices ? devices[k] : k;
  g->eng.assign(ndev, nullptr);
  g->comm.assign(ndev, nullptr);
  for (int k = 0; k < ndev; k++) {
    int rc = gh_engine_create(&g->eng[k], devs[k], n_total, g->begin[k], g->count[k], prec);
    if (rc != GH_OK) { gh_group_destroy(g); return rc; }
  }
  if (ndev > 1) {
    g->api = nccl_api(nullptr);
    if (!g->api) { gh_group_destroy(g); return GH_ECUDA; }
    int r = g->api->CommInitAll(g->comm.data(), ndev, devs.data());
    if (r != 0) { set_error("ncclCommInitAll: %s", g->api->GetErrorString(r)); gh_group_destroy(g); return GH_ECUDA; }
  }
  cudaSetDevice(devs[0]);
  for (int k = 0; k < GROUP_EVENTS; k++) cudaEventCreate(&g->pev[k]);
  *out = g;
  return GH_OK;
}

int gh_group_create_rank(gh_group **out, const void *id128, int rank, int world, int device, int64_t n_total,
                         int prec) {
  if (!out) return GH_EINVAL;
  *out = nullptr;
  if (world < 1 || rank < 0 || rank >= world || n_total < world) { set_error("gh_group_create_rank: bad rank/world"); return GH_EINVAL; }
  if (world > 1 && !id128) { set_error("gh_group_create_rank: null id"); return GH_EINVAL; }
  gh_group *g = new (std::nothrow) gh_group();
  if (!g) return GH_ENOMEM;
  g->world = world;
  g->rank0 = rank;
  group_common(g, n_total, prec);
  g->eng.assign(1, nullptr);
  g->comm.assign(1, nullptr);
  int rc = gh_engine_create(&g->eng[0], device, n_total, g->begin[rank], g->count[rank], prec);
  if (rc != GH_OK) { gh_group_destroy(g); return rc; }
  if (world > 1) {
    g->api = nccl_api(nullptr);
    if (!g->api) { gh_group_destroy(g); return GH_ECUDA; }
    ncclUniqueId id;
    memcpy(id.internal, id128, 128);
    cudaSetDevice(device);
    int r = g->api->CommInitRank(&g->comm[0], world, id, rank);
    if (r != 0) { set_error("ncclCommInitRank: %s", g->api->GetErrorString(r)); gh_group_destroy(g); return GH_ECUDA; }
  }
  cudaSetDevice(device);
  for (int k = 0; k < GROUP_EVENTS; k++) cudaEventCreate(&g->pev[k]);
  *out = g;
  return GH_OK;
}

int gh_group_size(gh_group *g, int *nlocal, int *world, int *rank0) {
  if (!g) return GH_EINVAL;
  if (nlocal) *nlocal = (int)g->eng.size();
  if (world) *world = g->world;
  if (rank0) *rank0 = g->rank0;
  return GH_OK;
}

int gh_group_engine(gh_group *g, int k, gh_engine **e, int64_t *begin, int64_t *count) {
  if (!g || k < 0 || k >= (int)g->eng.size()) { set_error("gh_group_engine: bad index"); return GH_EINVAL; }
  if (e) *e = g->eng[k];
  if (begin) *begin = g->begin[g->rank0 + k];
  if (count) *count = g->count[g->rank0 + k];
  return GH_OK;
}

int gh_group_set_tree_distributed(gh_group *g, int enable) {
  if (!g) return GH_EINVAL;
  g->tree_dist = enable != 0;
  for (auto *e : g->eng) e->dist_ready = false;
  return GH_OK;
}

int gh_group_prepare(gh_group *g, double dt) {
  if (!g) return GH_EINVAL;
  for (auto *e : g->eng) GH_TRY(gh_engine_prepare(e, dt));
  return GH_OK;
}

int gh_group_step(gh_group *g, int64_t nsteps, double dt, double eps, double theta, int algorithm) {
  if (!g) { set_error("null group"); return GH_EINVAL; }
  if (nsteps < 0) { set_error("negative step count"); return GH_EINVAL; }
  if (algorithm != GH_ALG_DIRECT && algorithm != GH_ALG_TREE) { set_error("unknown algorithm %d", algorithm); return GH_EINVAL; }
  for (auto *e : g->eng) {
    if (!e->uploaded) { set_error("engine has no state: upload first"); return GH_ESTATE; }
    if (!e->xhalf_valid || dt != e->dt_built) GH_TRY(gh_engine_prepare(e, dt));
  }
  // (quadrupoles, an opt-in accuracy upgrade, use the single-rank build and the per-target walk)
  const bool dist = algorithm == GH_ALG_TREE && g->world > 1 && g->prec == GH_PREC_F32 && g->tree_dist &&
                    !tree_quadrupoles();
  for (int64_t s = 0; s < nsteps; s++) {
    mark(g, 0);
    GH_TRY(all_gather_sources(g));
    mark(g, 1);
    if (dist) {
      GH_TRY(tree_step_distributed(g, dt, eps, theta));
    } else {
      for (auto *e : g->eng) {
        EngineScope sc(e);
        GH_TRY(engine_step_impl(e, dt, eps, theta, algorithm, nullptr));
      }
      g->pev_dist = false;
    }
    mark(g, 10);
    g->pev_valid = true;
  }
  return GH_OK;
}

int gh_group_synchronize(gh_group *g) {
  if (!g) return GH_EINVAL;
  for (auto *e : g->eng) GH_TRY(gh_engine_synchronize(e));
  return GH_OK;
}

/* ms of the last step on local engine 0: [0] source all-gather, [1] build A (bbox, keys, select,
 * sort), [2] exchange 1, [3] build B (levels, scans, moments), [4] exchange 2, [5] stitch + emit,
 * [6] entry + sorted-index all-gathers, [7] walk, [8] acceleration all-gather, [9] owners' kick and
 * drift, [10] whole step.  Non-distributed steps report [0], [10] and the rest 0. */
int gh_group_phase_ms(gh_group *g, float out[11]) {
  if (!g || !out) return GH_EINVAL;
  if (!g->pev_valid) { set_error("no step has run"); return GH_ESTATE; }
  cudaSetDevice(g->eng[0]->device);
  GH_CUDA(cudaEventSynchronize(g->pev[10]));
  for (int k = 0; k < 11; k++) out[k] = 0.f;
  GH_CUDA(cudaEventElapsedTime(&out[0], g->pev[0], g->pev[1]));
  GH_CUDA(cudaEventElapsedTime(&out[10], g->pev[0], g->pev[10]));
  if (g->pev_dist)
    for (int k = 1; k < 10; k++) GH_CUDA(cudaEventElapsedTime(&out[k], g->pev[k], g->pev[k + 1]));
  return GH_OK;
}

}  // extern "C"
