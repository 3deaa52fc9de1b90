// tree.cu -- Barnes-Hut octree on the GPU for sm_100a.  Replaces the pointer octree of
// /root/reference/gravhopper/_jbgrav.c:360-558 and treeforce_workhorse (:737-806).
//
// The tree that is built is THE REFERENCE'S OCTREE, not an approximation of it: same root cube
// (bbox midpoint, side per :764-769 including the eps padding quirk), same child assignment
// (strict p > centre per axis, child centre = centre +- size/4 accumulated level by level in
// the same floating-point order, :441-462,:406), one particle per leaf, a cell wherever two or
// more particles share an octant path (chains of single-child cells included), monopole moments
// (:432-435,:467-483).  The walk applies the reference's opening test per target (:502) and so
// accepts exactly the reference's node set; only the order of floating-point additions differs
// (sequential depth-first instead of child-subtotal), i.e. results agree to rounding (~1e-14).
//
// How it is built (all data-parallel, no pointers, no recursion, no per-node allocation):
//   K3  bbox            two-stage min/max reduction -> root centre and side (device resident)
//   K4  keys            per particle, descend up to 42 levels comparing against the running
//                       cell centre exactly as gravoct_calc_subnode does; 3 bits per level,
//                       z-major so that key order == the reference's branch order (:441-450);
//                       levels 1-21 in `hi`, 22-42 in `lo`
//   K5  sort            stable LSD radix sort of (lo, hi) with the particle index: hand-written,
//                       8 bits per pass, ballot ranking, shared-memory staged coalesced
//                       scatter (sortscan.cuh); a running fp32 simulation sorts with buckets
//                       taken from the previous step's order instead (bucketsort.cuh), same
//                       result bit for bit; no library kernels anywhere in the build
//   K6a common levels   c[p] = number of octant levels shared by sorted neighbours p, p+1.
//                       A cell of level l starts at p  <=>  c[p-1] < l <= c[p]; therefore the
//                       depth-first (pre-order) position of every cell and leaf is a prefix sum
//                       of (cells opened at p) + 1.
//   K7  moments         sources gathered once into Morton order, then a fused, deterministic
//                       three-phase double-double inclusive scan of (m, m x, m y, m z); a cell
//                       covering sorted particles [p, b] has mass = P[b+1] - P[p] (the
//                       double-double difference is exact to ~1e-30, so no cancellation)
//   K6b emit            per particle: walk down its key, write one entry per opened cell
//                       (centre, side, COM, mass, skip = pre-order index after the subtree, found
//                       by galloping search for the end of the key-prefix run) and its leaf entry;
//                       fp32 single-rank builds use the warp-cooperative form (a warp's opened
//                       cells dealt one per lane, cell ends from level-min tables; build.cuh)
//   K8  walk            one warp per 32 Morton-consecutive targets; the warp scans the
//                       pre-order array once, skipping a subtree when every lane has either
//                       accepted the cell or is already past it; each lane applies ITS OWN
//                       opening test and remembers the pre-order index up to which it has
//                       accepted an ancestor.  Entry loads are warp-uniform (one L1 transaction).
//                       The epilogue (store, or fused kick+drift) runs in the same kernel.
// Cells deeper than 42 levels (|dx| < side * 2^-42) are not split: their particles become sibling
// leaves, which is exact for the force and cannot loop forever on coincident particles (the
// reference segfaults there, :401-413).
#include "common.cuh"
#include "sortscan.cuh"
#include "bucketsort.cuh"
#include "walk.cuh"
#include "build.cuh"

#include <climits>
#include <cstdlib>
#include <cstring>

namespace gh {

// ---- workspace + orchestration --------------------------------------------------------------------
// What one evaluation leaves behind for its later phases (the distributed build interleaves the
// phases with collectives the caller issues; a single-rank evaluation runs them back to back).
struct TreePhaseState {
  int64_t n = 0;            // capacity of the sorted arrays: all sources (single rank) or TreeDist::ncap
  int64_t nall = 0;         // all sources
  const uint64_t *shi = nullptr, *slo = nullptr;
  const int *sidx = nullptr;
  const int *ndev = nullptr;  // &ctl->n_local when the count lives on the device
  int levels = 0;
  bool deep = false, dist = false;
  int rank = 0, world = 1;
  int64_t ecap = 0;         // entries per segment (stride)
  int end = 0;              // world * stride
  bool quad = false;        // this evaluation carries quadrupoles (single rank, per-target walk)
  LevelMin lm = LevelMin{{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr}, {0, 0, 0, 0, 0, 0, 0}, 0};  // level-min tables (warp emit)
  int blk = 2048, T = 0;    // distributed walk: block size of the deal, blocks per (rank, range)
  int64_t slots = 0;        // target slots per rank = world * T * blk
};
struct TreeWorkspace {
  DeviceBuffer root, part, hi, lo, hi2, lo2, lo3, idx, idx2, clev, cnt, base, P;
  DeviceBuffer node, skip, misc, thi, tidx, thi2, tidx2, sorted, bsum, scanlv[4], cntlv[4];
  DeviceBuffer ctl, rec1, rec2, keys_all, tilecnt, tileoff, tilelv[4], sidx_all, acc_all;
  DeviceBuffer lmin;  // level-min tables of the common levels (build.cuh)
  DeviceBuffer P2, quad, scanlv2[4];  // opt-in quadrupoles: second-moment prefixes, per-entry tensors
  RadixScratch rs;
  SplitterState ss;  // buckets of the splitter sort (bucketsort.cuh): valid from one coherent step to the next
  TreePhaseState ph;
  int64_t ecap = 0;          // entries the node buffer holds per segment (grow-only)
  int64_t node_slots = 0;    // entries allocated in `node`
  int64_t last_stats[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int *h_pinned = nullptr;  // [0] entries of the last build, [1] maxlevel, [2] overflow flag ; async readback
  unsigned long long *h_stats = nullptr;
  cudaEvent_t readback = nullptr;  // recorded after the async readback of the last evaluation
  bool readback_valid = false;
};

TreeWorkspace *tree_workspace_create() { return new TreeWorkspace(); }
void tree_workspace_destroy(TreeWorkspace *w) {
  if (!w) return;
  DeviceBuffer *all[] = {&w->root, &w->part, &w->hi, &w->lo, &w->hi2, &w->lo2, &w->lo3, &w->idx, &w->idx2,
                         &w->clev, &w->cnt, &w->base, &w->P, &w->node, &w->sorted, &w->bsum, &w->scanlv[0], &w->scanlv[1], &w->scanlv[2], &w->scanlv[3], &w->cntlv[0], &w->cntlv[1], &w->cntlv[2], &w->cntlv[3],
                         &w->skip, &w->misc, &w->thi, &w->tidx, &w->thi2, &w->tidx2, &w->ctl, &w->rec1, &w->rec2,
                         &w->keys_all, &w->tilecnt, &w->tileoff, &w->sidx_all, &w->acc_all, &w->P2, &w->quad, &w->lmin,
                         &w->scanlv2[0], &w->scanlv2[1], &w->scanlv2[2], &w->scanlv2[3], &w->tilelv[0], &w->tilelv[1], &w->tilelv[2], &w->tilelv[3]};
  for (auto *b : all) b->release();
  w->rs.release();
  w->ss.release();
  if (w->h_pinned) cudaFreeHost(w->h_pinned);
  if (w->h_stats) cudaFreeHost(w->h_stats);
  if (w->readback) cudaEventDestroy(w->readback);
  delete w;
}
int tree_last_stats(TreeWorkspace *w, int64_t out[8]) {
  for (int k = 0; k < 8; k++) out[k] = w->last_stats[k];
  return GH_OK;
}

static inline unsigned nblk(int64_t n, int b) { return (unsigned)((n + b - 1) / b); }

// walk selection: GH_WALK_GROUP (default; fp32 only) or GH_WALK_TARGET (the reference's per-target
// criterion; always used in fp64).  Process-wide; GH_TREE_WALK=target|group sets the initial value.
// sort of a running simulation's steps (tree_impl; bucketsort.cuh): 1 = bucket (two partition passes),
// 2 = place (global atomics), 3 = place2 (shared-memory histograms; the default: measured at N = 4M
// build 1.34 -> 1.11 ms against bucket, gpurun_out/ab_*); GH_SORT=classic|bucket|place|place2 overrides
#ifndef GH_SORT_DEFAULT
#define GH_SORT_DEFAULT 3
#endif
// emit of the fp32 tree: 0 = one thread per particle, 1 = warp-cooperative (single-rank builds; the
// default: -0.14 ms of build at N = 4M); GH_EMIT=thread|warp overrides
#ifndef GH_EMIT_DEFAULT
#define GH_EMIT_DEFAULT 1
#endif
#ifndef GH_WALK_HYBRID_DEFAULT
#define GH_WALK_HYBRID_DEFAULT 0.10f
#endif
static int g_walk_mode = -1;
int tree_walk_mode() {
  if (g_walk_mode < 0) {
    int m = GH_WALK_GROUP;
    if (const char *env = getenv("GH_TREE_WALK")) {
      if (!strcmp(env, "target") || !strcmp(env, "0")) m = GH_WALK_TARGET;
    }
    g_walk_mode = m;
  }
  return g_walk_mode;
}
void set_tree_walk_mode(int m) { g_walk_mode = (m == GH_WALK_TARGET) ? GH_WALK_TARGET : GH_WALK_GROUP; }
// list length beyond which a group gives up and runs the per-target scan (GH_WALK_LIST_LIMIT)
static int group_list_limit() {
  static int v = -1;
  if (v < 0) {
    v = 3000;
    if (const char *env = getenv("GH_WALK_LIST_LIMIT")) { int t = atoi(env); if (t >= 32) v = t; }
  }
  return v;
}

// kappa of the hybrid rule (see walk_group_kernel); 0 = off.  Default 0.10: measured on B200 over
// ALL particles of the N = 4,194,304 Hernquist sphere (profiles/r02_hybrid_sweep_N4M.json) it
// re-evaluates 0.10 % of the targets and brings p99.9 / p99.99 / max of the error against direct
// summation to 0.94x / 1.015x / 1.000x the reference tree's (plain group walk: 1.12x at p99.99),
// mean / median / p99 staying 11-13 % BELOW the reference's, for ~2 % of the walk time (0.15:
// 0.30 % of the targets, 1.002x, ~7 %; 0.2: 1.3 %, 1.000x).  GH_WALK_HYBRID=<kappa> sets the initial value (0 = off),
// gh_set_tree_walk_hybrid() changes it.
static float g_hybrid_kappa = -1.f;
float group_hybrid_kappa() {
  if (g_hybrid_kappa < 0.f) {
    float v = GH_WALK_HYBRID_DEFAULT;
    if (const char *env = getenv("GH_WALK_HYBRID")) { v = (float)atof(env); if (!(v > 0.f) || v > 1.f) v = 0.f; }
    g_hybrid_kappa = v;
  }
  return g_hybrid_kappa;
}
void set_group_hybrid_kappa(double k) { g_hybrid_kappa = (k > 0.0 && k <= 1.0) ? (float)k : 0.f; }

// Opt-in accuracy upgrade beyond the reference: traceless quadrupoles of the
// accepted cells.  Off by default (the reference is monopole only, _jbgrav.c:522-524); with it the
// evaluation uses the per-target walk (the reference's accepted node set) and a single-rank build.
// GH_TREE_QUADRUPOLES=1 sets the initial value, gh_set_tree_quadrupoles() changes it.
static int g_quadrupoles = -1;
int tree_quadrupoles() {
  if (g_quadrupoles < 0) {
    const char *env = getenv("GH_TREE_QUADRUPOLES");
    g_quadrupoles = (env && atoi(env) != 0) ? 1 : 0;
  }
  return g_quadrupoles;
}
void set_tree_quadrupoles(int on) { g_quadrupoles = on ? 1 : 0; }

template <class Real> struct GroupWalk {
  static void launch(const Node<Real> *, int, const TargetsView &, int64_t, const double *, float, double,
                     const Epilogue &, unsigned long long *, bool, bool, unsigned, cudaStream_t, const int *) {}
};
template <> struct GroupWalk<float> {
  static void launch(const Node<float> *nodes, int nentries, const TargetsView &tv, int64_t ni,
                     const double *root, float eps2, double inv_theta2, const Epilogue &ep,
                     unsigned long long *dstats, bool stats, bool guard, unsigned blocks32, cudaStream_t st,
                     const int *ovf) {
    const int lim = group_list_limit();
    int wpc = 2;  // warps per CTA (measured at N = 4M: 32/64/128 threads -> 3.21/3.14/3.15 ms); GH_WALK_BLOCK overrides
    if (const char *env = getenv("GH_WALK_BLOCK")) { int v = atoi(env); if (v == 32 || v == 64 || v == 128) wpc = v / 32; }
    const unsigned nb = (blocks32 + wpc - 1) / wpc;
    const float kappa = group_hybrid_kappa();
    const bool hyb = kappa > 0.f;
    if (hyb) {
      const float k2 = kappa * kappa;
      cudaMemcpyToSymbolAsync(c_hybrid_kappa2, &k2, sizeof(float), 0, cudaMemcpyHostToDevice, st);
    }
#define GH_GWALK0(W, STATS, GUARD, HYB) \
  walk_group_kernel<W, STATS, GUARD, HYB><<<nb, 32 * W, 0, st>>>(nodes, nentries, tv, ni, root, eps2, inv_theta2, lim, ep, dstats, ovf)
#define GH_GWALK1(W, STATS, GUARD) do { if (hyb) GH_GWALK0(W, STATS, GUARD, true); else GH_GWALK0(W, STATS, GUARD, false); } while (0)
#define GH_GWALK(STATS, GUARD) do { if (wpc == 1) GH_GWALK1(1, STATS, GUARD); else if (wpc == 2) GH_GWALK1(2, STATS, GUARD); else GH_GWALK1(4, STATS, GUARD); } while (0)
    if (stats) { if (guard) GH_GWALK(true, true); else GH_GWALK(true, false); }
    else { if (guard) GH_GWALK(false, true); else GH_GWALK(false, false); }
#undef GH_GWALK0
#undef GH_GWALK1
#undef GH_GWALK
  }
};

// entries a segment must be able to hold.  Single rank: grow-only, at least 2 n (the reference
// octree has ~1.5 entries per particle; chains of single-child cells above tight pairs add to
// that), raised ahead of time when the last completed build came within 80 % of it.  The count is
// never needed on the host during an evaluation: the walk ends its chains at the CAPACITY, emit
// guards its writes, and an overflow is a flag the host looks at when it next synchronises.
static int64_t entry_capacity(TreeWorkspace *w, int64_t n, bool fp32) {
  int64_t want = 2 * n + 1024;
  if (w->ecap > want) want = w->ecap;
  const int64_t seen = w->h_pinned ? (int64_t)w->h_pinned[0] : 0;  // possibly one evaluation stale
  if (seen > 0 && 10 * seen > 8 * want) want = seen + seen / 2;
  if (fp32 && want >= (1 << SKIP_BITS)) want = (1 << SKIP_BITS) - 1;
  return want;
}

// emit of the fp32 tree, warp-cooperative form (emit32_warp_kernel); false: not this precision
template <class Real> struct Emit32 {
  static bool launch(const double4 *, const uint64_t *, const signed char *, const int *, const void *, int64_t,
                     const double *, void *, int *, BuildCtl *, bool, const LevelMin &, cudaStream_t) { return false; }
};
template <> struct Emit32<float> {
  static bool launch(const double4 *sp, const uint64_t *shi, const signed char *clev, const int *base, const void *P,
                     int64_t n, const double *root, void *nodes, int *maxlevel, BuildCtl *ctl, bool dist,
                     const LevelMin &lm, cudaStream_t st) {
    Entries<float> E{static_cast<Node<float> *>(nodes), nullptr};
    emit32_warp_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(sp, shi, clev, base, static_cast<const D4 *>(P), n, root,
                                                                    E, maxlevel, ctl, dist, lm);
    return true;
  }
};

// 0 = one thread per particle, 1 = warp-cooperative for single-rank fp32 builds, 2 = warp-cooperative
// for every fp32 build (GH_EMIT=thread|warp; the distributed build's use of the warp form is
// validated on the CPU with the kernel source)
static int emit_mode() {
  static const int mode = [] {
    const char *env = getenv("GH_EMIT");
    if (env && !strcmp(env, "thread")) return 0;
    if (env && !strcmp(env, "warp")) return 2;
    return GH_EMIT_DEFAULT;
  }();
  return mode;
}
template <class Real>
static bool use_warp_emit(const TreePhaseState &ph) {
  const int m = emit_mode();
  return sizeof(Real) == 4 && !ph.quad && (m == 2 || (m == 1 && !ph.dist));
}

template <class Src, class Real>
struct TreeRun {
  using Mom = typename MomentOf<Real>::type;

  // ---- phase A: bbox, keys, (select,) sort, (boundary keys) ---------------------------------------
  static int phase_a(const TreeArgs &a, Src src, TreeWorkspace *w, cudaStream_t st, const TreeDist *d) {
    const int64_t n = a.nj;
    if (n > (int64_t)INT_MAX / 48) { set_error("tree: too many particles (%lld)", (long long)n); return GH_EINVAL; }
    TreePhaseState &ph = w->ph;
    ph = TreePhaseState();
    ph.n = n;
    ph.nall = n;
    ph.levels = (sizeof(Real) == 8) ? LEVELS_MAX : LEVELS_HI;
    ph.deep = ph.levels > LEVELS_HI;
    ph.dist = d != nullptr && d->world > 1;
    ph.quad = tree_quadrupoles() != 0;
    if (ph.dist && ph.quad) { set_error("quadrupoles need the single-rank (redundant) tree build"); return GH_EINVAL; }
    if (ph.dist && ph.deep) { set_error("the distributed tree build is fp32 only"); return GH_EINVAL; }
    if (ph.dist && d->world > DIST_MAX_RANKS) { set_error("at most %d ranks", DIST_MAX_RANKS); return GH_EINVAL; }
    ph.rank = ph.dist ? d->rank : 0;
    ph.world = ph.dist ? d->world : 1;
    if (!w->h_pinned) {
      GH_CUDA(cudaMallocHost(&w->h_pinned, 4 * sizeof(int)));
      w->h_pinned[0] = w->h_pinned[1] = w->h_pinned[2] = w->h_pinned[3] = 0;
    }
    if (!w->h_stats) GH_CUDA(cudaMallocHost(&w->h_stats, 4 * sizeof(unsigned long long)));
    if (!w->readback) GH_CUDA(cudaEventCreateWithFlags(&w->readback, cudaEventDisableTiming));
    const bool deep = ph.deep;

    // entry array: capacity known before anything runs (no host round trip inside an evaluation)
    const bool fp32 = sizeof(Real) == 4;
    int64_t ecap = ph.dist ? (int64_t)d->stride : entry_capacity(w, n, fp32);
    if (fp32 && ecap * ph.world >= (1 << SKIP_BITS)) {
      if (!ph.dist && n + 1 < (1 << SKIP_BITS)) ecap = (1 << SKIP_BITS) - 1;
      else { set_error("tree: %lld entries exceed the fp32 node format (2^%d)", (long long)(ecap * ph.world), SKIP_BITS); return GH_EINVAL; }
    }
    ph.ecap = ecap;
    ph.end = (int)(ecap * ph.world);
    w->ecap = ph.dist ? w->ecap : ecap;
    GH_TRY(w->node.reserve(sizeof(Node<Real>) * (size_t)ph.end));
    if (!fp32) GH_TRY(w->skip.reserve(sizeof(int) * (size_t)ph.end));

    GH_TRY(w->root.reserve(sizeof(double) * ROOT_DOUBLES));
    GH_TRY(w->part.reserve(sizeof(double) * 6 * 1024));
    GH_TRY(w->misc.reserve(64));
    GH_TRY(w->ctl.reserve(sizeof(BuildCtl)));
    BuildCtl *ctl = w->ctl.as<BuildCtl>();

    // K3
    const int nb = (int)((n + 256 * 8 - 1) / (256 * 8) < 1024 ? (n + 256 * 8 - 1) / (256 * 8) : 1024);
    double *root = w->root.as<double>();
    bbox_stage1<<<nb, 256, 0, st>>>(src, n, w->part.as<double>());
    GH_LAUNCH_CHECK();
    bbox_stage2<<<1, 256, 0, st>>>(w->part.as<double>(), nb, a.eps, root);
    GH_LAUNCH_CHECK();

    // K4
    GH_TRY(w->hi.reserve(sizeof(uint64_t) * n));
    GH_TRY(w->hi2.reserve(sizeof(uint64_t) * n));
    GH_TRY(w->idx.reserve(sizeof(int) * n));
    GH_TRY(w->idx2.reserve(sizeof(int) * n));
    if (deep) {
      GH_TRY(w->lo.reserve(sizeof(uint64_t) * n));
      GH_TRY(w->lo2.reserve(sizeof(uint64_t) * n));
    }
    uint64_t *hi = w->hi.as<uint64_t>(), *hi2 = w->hi2.as<uint64_t>();
    uint64_t *lo = deep ? w->lo.as<uint64_t>() : nullptr, *lo2 = deep ? w->lo2.as<uint64_t>() : nullptr;
    int *idx = w->idx.as<int>(), *idx2 = w->idx2.as<int>();
    if (ph.dist) {
      // keys of ALL sources (the rank's own targets are among them), then the stable selection of
      // the keys in this rank's range
      GH_TRY(w->keys_all.reserve(sizeof(uint64_t) * n));
      uint64_t *kall = w->keys_all.as<uint64_t>();
      keys_kernel<<<nblk(n, 256), 256, 0, st>>>(src, n, root, ph.levels, kall, (uint64_t *)nullptr, idx2);
      GH_LAUNCH_CHECK();
      const int ntiles = (int)((n + SEL_TILE - 1) / SEL_TILE);
      GH_TRY(w->tilecnt.reserve(sizeof(int) * (size_t)(ntiles + 1)));
      GH_TRY(w->tileoff.reserve(sizeof(int) * (size_t)(ntiles + 1)));
      ctl_init_dist<<<1, 1, 0, st>>>(ctl, ph.rank, ph.world, (int)ecap);
      GH_LAUNCH_CHECK();
      // key ranges of this step = what the previous step derived from the gathered key samples
      // (or the bootstrap's, tree_splitters)
      splitters_advance_kernel<<<1, DIST_MAX_RANKS + 1, 0, st>>>(ctl, ph.world);
      GH_LAUNCH_CHECK();
      select_count_kernel<<<ntiles, SEL_THREADS, 0, st>>>(kall, n, ctl, w->tilecnt.as<int>());
      GH_LAUNCH_CHECK();
      GH_TRY((chunked_scan<int, InArray<int>>(InArray<int>{w->tilecnt.as<int>()}, ntiles, w->tileoff.as<int>(),
                                               w->tilelv, 0, st)));
      // from here on everything is sized for the rank's share (ncap), not for all N
      int64_t ncap = (d->ncap > 0 && d->ncap < n) ? d->ncap : n;
      select_compact_kernel<<<ntiles, SEL_THREADS, 0, st>>>(kall, n, ctl, w->tileoff.as<int>(), ntiles, hi, idx, ncap);
      GH_LAUNCH_CHECK();
      ph.ndev = &ctl->n_local;
      ph.n = ncap;
      ph.blk = d->blk > 0 ? d->blk : 2048;
      if (ph.blk % 32) { set_error("tree: the deal's block size must be a multiple of 32"); return GH_EINVAL; }
      const int64_t nblk = (ncap + ph.blk - 1) / ph.blk;
      ph.T = (int)((nblk + ph.world - 1) / ph.world);
      ph.slots = (int64_t)ph.world * ph.T * ph.blk;
      GH_TRY(w->sidx_all.reserve(sizeof(int) * (size_t)ph.world * (size_t)ncap));
      GH_TRY(w->acc_all.reserve(sizeof(float4) * (size_t)ph.world * (size_t)ph.slots));
    } else {
      keys_kernel<<<nblk(n, 256), 256, 0, st>>>(src, n, root, ph.levels, hi, lo, idx);
      GH_LAUNCH_CHECK();
      ctl_init_single<<<1, 1, 0, st>>>(ctl, (int)n, (int)ecap);
      GH_LAUNCH_CHECK();
    }

    // K5: stable LSD radix sort (sortscan.cuh) over (lo, hi); both ping-pong buffers are clobbered
    bool inB = false;
    if (deep) {
      // pass 1: by lo; pass 2: by hi gathered through the pass-1 order (stable).  The unsorted lo
      // keys are needed again at the end, so keep a copy.
      GH_TRY(w->lo3.reserve(sizeof(uint64_t) * n));
      GH_CUDA(cudaMemcpyAsync(w->lo3.ptr, lo, sizeof(uint64_t) * n, cudaMemcpyDeviceToDevice, st));
      GH_TRY(radix_sort_pairs(lo, idx, lo2, idx2, n, 63, w->rs, st, &inB));
      int *order1 = inB ? idx2 : idx;
      int *other1 = inB ? idx : idx2;
      uint64_t *hs = inB ? lo : lo2;  // free key buffer of pass 1 holds hi gathered in pass-1 order
      gather_u64<<<nblk(n, 256), 256, 0, st>>>(hi, order1, n, hs);
      GH_LAUNCH_CHECK();
      GH_TRY(radix_sort_pairs(hs, order1, hi2, other1, n, 63, w->rs, st, &inB));
      ph.shi = inB ? hi2 : hs;
      ph.sidx = inB ? other1 : order1;
      uint64_t *lsorted = (ph.shi == hi2) ? hs : hi2;  // the key buffer not holding the result
      gather_u64<<<nblk(n, 256), 256, 0, st>>>(w->lo3.as<uint64_t>(), ph.sidx, n, lsorted);
      GH_LAUNCH_CHECK();
      ph.slo = lsorted;
    } else {
      // A running simulation (a.coherent) sorts with the buckets the previous step left behind
      // (bucketsort.cuh: 3 trips through global memory instead of 8); the first step after an
      // upload, stateless calls and small systems use the classic LSD sort.  GH_SORT=classic|bucket.
      static const int sort_mode = [] {  // 0 classic, 1 bucket (two partition passes), 2 place (global atomics), 3 place2 (shared-memory histograms)
        const char *env = getenv("GH_SORT");
        if (env && !strcmp(env, "classic")) return 0;
        if (env && !strcmp(env, "place")) return 2;
        if (env && !strcmp(env, "place2")) return 3;
        if (env && !strcmp(env, "bucket")) return 1;
        return GH_SORT_DEFAULT;
      }();
      // Single rank only: a rank's key range moves by ~1 % of its particles per step, and the keys
      // it gains all land in the first or last bucket, which then takes the slow oversize path
      // (measured at 8 GPUs: 0.7 ms of skew at the next exchange).
      const bool bucket = sort_mode >= 1 && a.coherent && !ph.dist && w->ss.nb >= BS_MIN_BUCKETS && w->ss.cap == ph.n;
      if (bucket && sort_mode == 3 && w->ss.nb <= BP2_MAX_NB) {
        // a couple of CTAs per SM, each with a contiguous chunk of the keys
        int dev = 0, sms = 0;
        GH_CUDA(cudaGetDevice(&dev));
        GH_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        static const int per_sm = [] { const char *e = getenv("GH_BP2_CTAS_PER_SM"); const int v = e ? atoi(e) : 2; return v >= 1 && v <= 4 ? v : 2; }();
        GH_TRY(splitter_place2_sort_pairs(hi, idx, hi2, idx2, ph.n, 63, w->ss, sms * per_sm, st, ph.ndev));
        inB = true;
      } else if (bucket && sort_mode == 2) {
        GH_TRY(splitter_place_sort_pairs(hi, idx, hi2, idx2, ph.n, 63, w->ss, st, ph.ndev));
        inB = true;
      } else if (bucket) {
        GH_TRY(splitter_sort_pairs(hi, idx, hi2, idx2, ph.n, 63, w->rs, w->ss, st, ph.ndev));
        inB = false;
      } else {
        GH_TRY(radix_sort_pairs(hi, idx, hi2, idx2, ph.n, 63, w->rs, st, &inB, ph.ndev));
      }
      ph.shi = inB ? hi2 : hi;
      ph.sidx = inB ? idx2 : idx;
      if (a.coherent && sort_mode >= 1 && !ph.dist) GH_TRY(splitter_refresh(w->ss, ph.shi, ph.n, st, ph.ndev));
      else w->ss.nb = 0;
    }
    if (ph.dist) {
      GH_TRY(w->rec1.reserve(sizeof(RankRec1) * (size_t)ph.world));
      GH_TRY(w->rec2.reserve(sizeof(RankRec2) * (size_t)ph.world));
      rec1_kernel<<<1, 1, 0, st>>>(ph.shi, ctl, w->rec1.as<RankRec1>());
      GH_LAUNCH_CHECK();
    }
    return GH_OK;
  }

  // ---- phase B: common levels, pre-order offsets, Morton gather, moments, (cell-end table) ----------
  static int phase_b(const TreeArgs &a, Src src, TreeWorkspace *w, cudaStream_t st) {
    TreePhaseState &ph = w->ph;
    const int64_t n = ph.n;
    BuildCtl *ctl = w->ctl.as<BuildCtl>();
    const BuildCtl *cd = ph.dist ? ctl : nullptr;
    if (ph.dist) {
      neighbours_kernel<<<1, 1, 0, st>>>(w->rec1.as<RankRec1>(), ctl);
      GH_LAUNCH_CHECK();
    }
    // K6a + pre-order offsets: base[p] = sum_{q<p} (cells opened at q + 1), base[n] = entries
    GH_TRY(w->clev.reserve(n));
    GH_TRY(w->cnt.reserve(sizeof(int) * (n + 1)));
    GH_TRY(w->base.reserve(sizeof(int) * (n + 1)));
    signed char *clev = w->clev.as<signed char>();
    int *cnt = w->cnt.as<int>(), *base = w->base.as<int>();
    // the warp form of the emit finds the cells' ends in level-min tables of the common levels
    // (GH_EMIT_SEARCH=keys keeps the search over the keys)
    static const bool lm_on = [] { const char *e = getenv("GH_EMIT_SEARCH"); return !(e && !strcmp(e, "keys")); }();
    ph.lm.ntab = 0;
    if (lm_on && use_warp_emit<Real>(ph)) {
      size_t off[LM_MAX_TABLES];
      const size_t bytes = lm_layout(n, ph.lm, off);
      GH_TRY(w->lmin.reserve(bytes));
      for (int k = 0; k < ph.lm.ntab; k++) ph.lm.t[k] = w->lmin.as<unsigned char>() + off[k];
      if (ph.lm.ntab > 1) GH_CUDA(cudaMemsetAsync(ph.lm.t[1], 0, bytes - off[1], st));
    }
    levels_kernel<<<nblk(n, 256), 256, 0, st>>>(ph.shi, ph.slo, n, ph.levels, clev, cnt, cd, ph.lm);
    GH_LAUNCH_CHECK();
    if (ph.lm.ntab > 2) {
      levelmin_top_kernel<<<1, 1024, 0, st>>>(ph.lm);
      GH_LAUNCH_CHECK();
    }
    GH_TRY((chunked_scan<int, InArray<int>>(InArray<int>{cnt}, n, base, w->cntlv, 0, st, ph.ndev)));
    if (!ph.dist) GH_CUDA(cudaMemcpyAsync(&w->h_pinned[0], base + n, sizeof(int), cudaMemcpyDeviceToHost, st));

    // K7
    GH_TRY(w->sorted.reserve(sizeof(double4) * (size_t)n));
    double4 *sp = w->sorted.as<double4>();
    gather_sorted_kernel<<<nblk(n, 256), 256, 0, st>>>(src, ph.sidx, n, sp, cd,
                                                       ph.dist ? w->sidx_all.as<int>() + (size_t)ph.rank * (size_t)n : nullptr);
    GH_LAUNCH_CHECK();
    GH_TRY(w->P.reserve(sizeof(Mom) * (size_t)(n + 1)));
    Mom *P = w->P.as<Mom>();
    double *root = w->root.as<double>();
    if (sizeof(Real) == 4) {
      GH_TRY((chunked_scan<D4, InParticlesRel>(InParticlesRel{sp, root}, n, reinterpret_cast<D4 *>(P), w->scanlv, 0, st, ph.ndev)));
    } else {
      GH_TRY((chunked_scan<DD4, InParticles>(InParticles{sp}, n, reinterpret_cast<DD4 *>(P), w->scanlv, 0, st)));
    }
    if (ph.dist) {
      rec2_kernel<<<1, 32, 0, st>>>(ph.shi, base, reinterpret_cast<const D4 *>(P), ctl, w->rec2.as<RankRec2>());
      GH_LAUNCH_CHECK();
    }
    if (ph.quad) {
      GH_TRY(w->P2.reserve(sizeof(D6) * (size_t)(n + 1)));
      GH_TRY((chunked_scan<D6, InSecondRel>(InSecondRel{sp, root}, n, w->P2.as<D6>(), w->scanlv2, 0, st)));
    }
    return GH_OK;
  }

  // ---- phase C: (stitch,) emit ------------------------------------------------------------------------
  static int phase_c(const TreeArgs &a, Src src, TreeWorkspace *w, cudaStream_t st) {
    TreePhaseState &ph = w->ph;
    const int64_t n = ph.n;
    BuildCtl *ctl = w->ctl.as<BuildCtl>();
    if (ph.dist) {
      stitch_kernel<<<1, 1, 0, st>>>(w->rec1.as<RankRec1>(), w->rec2.as<RankRec2>(), ctl, ph.nall);
      GH_LAUNCH_CHECK();
    }
    const bool rel_origin = (sizeof(Real) == 4);  // fp32 entries are stored relative to the root centre
    Entries<Real> E{w->node.as<Node<Real>>(), sizeof(Real) == 4 ? nullptr : w->skip.as<int>()};
    const double inv_theta2 = 1.0 / (a.theta * a.theta);  // theta = 0 -> inf: cells are never accepted
    int *maxlevel = w->misc.as<int>();
    GH_CUDA(cudaMemsetAsync(w->misc.ptr, 0, 64, st));
    if (ph.quad) GH_TRY(w->quad.reserve(sizeof(Real) * 6 * (size_t)ph.end));
    // fp32 entries without quadrupoles: the warp-cooperative form (build.cuh), same array bit for bit
    if (use_warp_emit<Real>(ph) &&
        Emit32<Real>::launch(w->sorted.as<double4>(), ph.shi, w->clev.as<signed char>(), w->base.as<int>(), w->P.ptr, n,
                             w->root.as<double>(), w->node.ptr, maxlevel, ctl, ph.dist, ph.lm, st)) {
      GH_LAUNCH_CHECK();
      return GH_OK;
    }
    emit_kernel<Src, Real><<<nblk(n, 128), 128, 0, st>>>(w->sorted.as<double4>(), ph.shi, ph.slo,
                                                       w->clev.as<signed char>(), w->base.as<int>(),
                                                       w->P.as<Mom>(), n, w->root.as<double>(), rel_origin,
                                                       inv_theta2, E, maxlevel, ctl, ph.dist,
                                                       ph.quad ? w->P2.as<D6>() : nullptr,
                                                       ph.quad ? w->quad.as<Real>() : nullptr);
    GH_LAUNCH_CHECK();
    return GH_OK;
  }

  // ---- phase D: targets in Morton order, walk, epilogue ----------------------------------------------
  static int phase_d(const TreeArgs &a, Src src, const float4 *tgt32, TreeWorkspace *w, cudaStream_t st,
                     cudaEvent_t *ev) {
    TreePhaseState &ph = w->ph;
    const int64_t n = ph.n, ni = a.ni;
    BuildCtl *ctl = w->ctl.as<BuildCtl>();
    double *root = w->root.as<double>();
    const bool rel_origin = (sizeof(Real) == 4);
    Entries<Real> E{w->node.as<Node<Real>>(), sizeof(Real) == 4 ? nullptr : w->skip.as<int>()};
    const double inv_theta2 = 1.0 / (a.theta * a.theta);
    int *maxlevel = w->misc.as<int>();
    unsigned long long *dstats = reinterpret_cast<unsigned long long *>(w->misc.as<char>() + 16);
    const int nentries = ph.end;  // where every chain ends (capacity, see entry_capacity)

    // targets: Morton order.  Self case: the source order restricted to the owned slice is the
    // sorted order itself when the slice is everything; otherwise sort the targets' own keys.
    TargetsView tv;
    tv.sorted = nullptr;
    tv.pos64 = tgt32 ? nullptr : a.tgt_pos;
    tv.pos32 = tgt32;
    tv.order = nullptr;
    tv.order_offset = 0;
    tv.dist_sidx = nullptr;
    tv.dist_counts = nullptr;
    tv.dist_rank = tv.dist_world = tv.dist_ncap = tv.dist_blk = tv.dist_T = 0;
    Epilogue ep = a.ep;
    int64_t nwalk = ni;
    if (ph.dist) {
      // this rank's share of the GLOBAL Morton order (blocks dealt round-robin over the ranks): groups
      // of 32 are spatial neighbours whoever owns them; the accelerations go to this rank's slot of
      // the gathered buffer and the owners apply the kick and the drift (phase 4)
      if (!tgt32) { set_error("the distributed walk needs fp32 engine sources"); return GH_EINVAL; }
      tv.pos32 = a.src32;
      tv.dist_sidx = w->sidx_all.as<int>();
      tv.dist_counts = ctl->counts;
      tv.dist_rank = ph.rank; tv.dist_world = ph.world; tv.dist_ncap = (int)n; tv.dist_blk = ph.blk; tv.dist_T = ph.T;
      memset(&ep, 0, sizeof(ep));
      ep.mode = EP_ACC32;
      ep.acc32_out = w->acc_all.as<float4>() + (size_t)ph.rank * (size_t)ph.slots;
      nwalk = ph.slots;
    } else if (a.targets_are_sources && ni == n) {
      tv.order = ph.sidx;
      tv.sorted = w->sorted.as<double4>();
    } else if (ni > 32) {
      GH_TRY(w->thi.reserve(sizeof(uint64_t) * ni));
      GH_TRY(w->thi2.reserve(sizeof(uint64_t) * ni));
      GH_TRY(w->tidx.reserve(sizeof(int) * ni));
      GH_TRY(w->tidx2.reserve(sizeof(int) * ni));
      if (tgt32) {
        Src32 ts{tgt32};
        keys_kernel<<<nblk(ni, 256), 256, 0, st>>>(ts, ni, root, LEVELS_HI, w->thi.as<uint64_t>(),
                                                  (uint64_t *)nullptr, w->tidx.as<int>());
      } else {
        Src64 ts{a.tgt_pos, nullptr};
        keys_kernel<<<nblk(ni, 256), 256, 0, st>>>(ts, ni, root, LEVELS_HI, w->thi.as<uint64_t>(),
                                                  (uint64_t *)nullptr, w->tidx.as<int>());
      }
      GH_LAUNCH_CHECK();
      bool tinB = false;
      GH_TRY(radix_sort_pairs(w->thi.as<uint64_t>(), w->tidx.as<int>(), w->thi2.as<uint64_t>(),
                              w->tidx2.as<int>(), ni, 63, w->rs, st, &tinB));
      tv.order = tinB ? w->tidx2.as<int>() : w->tidx.as<int>();
    }

    // K8
    Real eps2 = (Real)(a.eps * a.eps);
    // fp32: see GH_F32_MIN_EPS2 (common.cuh)
    const bool tiny_eps = (sizeof(Real) == 4) ? !((float)eps2 >= GH_F32_MIN_EPS2) : (a.eps == 0.0);
    if (tiny_eps) eps2 = (Real)0;
    const int64_t nwarps = (nwalk + 31) / 32;
    int wb = 128;
    if (const char *env = getenv("GH_WALK_BLOCK")) { int v = atoi(env); if (v == 32 || v == 64 || v == 128) wb = v; }
    const bool group = (sizeof(Real) == 4) && tree_walk_mode() == GH_WALK_GROUP && !ph.quad;
    const unsigned blocks = (unsigned)((nwarps + wb / 32 - 1) / (wb / 32));
    if (ev) GH_CUDA(cudaEventRecord(ev[0], st));
    const bool guard = tiny_eps;
    const int *ovf = &ctl->overflow;  // (overflow, first): the walk kernels' `walkctl`
    // L2 prefetch hint of each entry's skip target; GH_WALK_PREFETCH=0/1 overrides
    bool prefetch = false;
    if (const char *env = getenv("GH_WALK_PREFETCH")) prefetch = atoi(env) != 0;
    if (group) {
      GroupWalk<Real>::launch(E.node, nentries, tv, nwalk, root, (float)eps2, inv_theta2, ep, dstats,
                              a.want_stats, guard, (unsigned)nwarps, st, ovf);
    } else {
#define GH_WALK(STATS, GUARD, PF)                                                                      \
  walk_kernel<Real, STATS, GUARD, PF><<<blocks, wb, 0, st>>>(E.node, E.skip, nentries, tv, nwalk, root, \
                                                            rel_origin, eps2, inv_theta2, ep, dstats, ovf)
#define GH_WALKQ(STATS, GUARD)                                                                                 \
  walk_kernel<Real, STATS, GUARD, false, true><<<blocks, wb, 0, st>>>(E.node, E.skip, nentries, tv, nwalk, root, \
                                                                     rel_origin, eps2, inv_theta2, ep, dstats, ovf, \
                                                                     w->quad.as<Real>())
#define GH_WALK2(STATS, GUARD) do { if (ph.quad) GH_WALKQ(STATS, GUARD); else if (prefetch) GH_WALK(STATS, GUARD, true); else GH_WALK(STATS, GUARD, false); } while (0)
      if (a.want_stats) { if (guard) GH_WALK2(true, true); else GH_WALK2(true, false); }
      else { if (guard) GH_WALK2(false, true); else GH_WALK2(false, false); }
#undef GH_WALK2
#undef GH_WALKQ
#undef GH_WALK
    }
    GH_LAUNCH_CHECK();
    if (ev) GH_CUDA(cudaEventRecord(ev[1], st));
    // async readback: deepest level, overflow flag, (distributed: this rank's entry count)
    GH_CUDA(cudaMemcpyAsync(&w->h_pinned[1], maxlevel, sizeof(int), cudaMemcpyDeviceToHost, st));
    GH_CUDA(cudaMemcpyAsync(&w->h_pinned[2], &ctl->overflow, sizeof(int), cudaMemcpyDeviceToHost, st));
    if (ph.dist) GH_CUDA(cudaMemcpyAsync(&w->h_pinned[0], &ctl->nentries, sizeof(int), cudaMemcpyDeviceToHost, st));
    GH_CUDA(cudaEventRecord(w->readback, st));
    w->readback_valid = true;
    if (a.want_stats) {
      GH_CUDA(cudaMemcpyAsync(w->h_stats, dstats, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
      GH_CUDA(cudaStreamSynchronize(st));
      w->last_stats[0] = w->h_pinned[0];
      w->last_stats[1] = w->h_pinned[0] - n;
      w->last_stats[2] = w->h_pinned[1];
      w->last_stats[3] = (int64_t)w->h_stats[0];
      w->last_stats[4] = (int64_t)w->h_stats[1];
      w->last_stats[5] = (int64_t)w->h_stats[2];
      w->last_stats[6] = (int64_t)w->h_stats[3];
      w->last_stats[7] = (int64_t)((ni + 31) / 32);
    }
    return GH_OK;
  }
};

// phase 4 of a distributed step: the owners' kick and drift from the gathered accelerations
static int tree_phase_epilogue(const TreeArgs &a, TreeWorkspace *w, cudaStream_t st) {
  TreePhaseState &ph = w->ph;
  if (!ph.dist) return GH_OK;
  const int64_t tot = (int64_t)ph.world * ph.n;
  GH_TRY(w->tidx.reserve(sizeof(int) * (size_t)a.ni));
  int *inv = w->tidx.as<int>();
  dist_inverse_kernel<<<nblk(tot, 256), 256, 0, st>>>(w->sidx_all.as<int>(), w->ctl.as<BuildCtl>(), ph.world,
                                                     (int)ph.n, a.tgt_offset, a.ni, inv);
  GH_LAUNCH_CHECK();
  dist_epilogue_kernel<<<nblk(a.ni, 256), 256, 0, st>>>(inv, w->ctl.as<BuildCtl>(), w->acc_all.as<float4>(), ph.world,
                                                       (int)ph.n, ph.blk, ph.T, a.ni, a.ep);
  GH_LAUNCH_CHECK();
  return GH_OK;
}

// phase: 0 = A, 1 = B, 2 = C, 3 = D, 4 = distributed epilogue, -1 = all (single rank)
static int tree_dispatch(const TreeArgs &a, TreeWorkspace *w, cudaStream_t st, cudaEvent_t *ev,
                         const TreeDist *d, int phase) {
#define GH_TREE_PHASES(SRC, REAL, src, tgt)                                                    \
  do {                                                                                         \
    if (phase == 0 || phase < 0) GH_TRY((TreeRun<SRC, REAL>::phase_a(a, src, w, st, d)));      \
    if (phase == 1 || phase < 0) GH_TRY((TreeRun<SRC, REAL>::phase_b(a, src, w, st)));         \
    if (phase == 2 || phase < 0) GH_TRY((TreeRun<SRC, REAL>::phase_c(a, src, w, st)));         \
    if (phase == 3 || phase < 0) GH_TRY((TreeRun<SRC, REAL>::phase_d(a, src, tgt, w, st, ev))); \
    return GH_OK;                                                                              \
  } while (0)
  if (phase == 4) return tree_phase_epilogue(a, w, st);
  if (a.prec == GH_PREC_F64) {
    Src64 s{a.src_pos, a.src_mass};
    GH_TREE_PHASES(Src64, double, s, nullptr);
  } else if (a.prec == GH_PREC_F32) {
    if (a.src32) {  // f32 engine: sources and targets are float4 (x - origin, m)
      Src32 s{a.src32};
      GH_TREE_PHASES(Src32, float, s, a.tgt32);
    }
    Src64 s{a.src_pos, a.src_mass};
    GH_TREE_PHASES(Src64, float, s, nullptr);
  }
#undef GH_TREE_PHASES
  set_error("launch_tree: bad precision %d", a.prec);
  return GH_EINVAL;
}

int launch_tree(const TreeArgs &a, TreeWorkspace *w, cudaStream_t st, cudaEvent_t *ev) {
  if (a.ni <= 0 || a.nj <= 0) return GH_OK;
  GH_TRY(tree_dispatch(a, w, st, ev, nullptr, -1));
  if (a.sync_check) {
    // callers that synchronise anyway (the stateless entry points): an overflow of the entry
    // array is repaired here by growing it to the exact need and evaluating again
    GH_CUDA(cudaStreamSynchronize(st));
    if (w->h_pinned[2]) {
      w->ecap = (int64_t)w->h_pinned[0] + (int64_t)w->h_pinned[0] / 8 + 1024;
      GH_TRY(tree_dispatch(a, w, st, ev, nullptr, -1));
      GH_CUDA(cudaStreamSynchronize(st));
      if (w->h_pinned[2]) { set_error("tree: entry array overflow (%d entries)", w->h_pinned[0]); return GH_ENOMEM; }
    }
    if (!a.want_stats) { w->last_stats[0] = w->h_pinned[0]; w->last_stats[1] = w->h_pinned[0] - a.nj; w->last_stats[2] = w->h_pinned[1]; }
  }
  return GH_OK;
}

int launch_tree_phase(const TreeArgs &a, TreeWorkspace *w, cudaStream_t st, cudaEvent_t *ev,
                      const TreeDist *d, int phase) {
  if (a.ni <= 0 || a.nj <= 0) return GH_OK;
  return tree_dispatch(a, w, st, ev, d, phase);
}

// 0: no evaluation finished yet / still running; 1: finished without overflow; -1: overflowed.
// Never blocks.  *entries (nullable) = the entry count of that evaluation.
int tree_poll_overflow(TreeWorkspace *w, int64_t *entries) {
  if (!w || !w->readback_valid) return 0;
  if (cudaEventQuery(w->readback) != cudaSuccess) { cudaGetLastError(); return 0; }
  if (entries) *entries = w->h_pinned[0];
  return w->h_pinned[2] ? -1 : 1;
}

void tree_forget_history(TreeWorkspace *w) { if (w) w->ss.nb = 0; }

const int *tree_maxent_ptr(TreeWorkspace *w) { return &w->ctl.as<BuildCtl>()->maxent; }

int tree_exchange_buffer(TreeWorkspace *w, int which, void **ptr, int64_t *bytes_per_rank) {
  if (!w || !ptr || !bytes_per_rank) return GH_EINVAL;
  switch (which) {
    case 1: *ptr = w->rec1.ptr; *bytes_per_rank = sizeof(RankRec1); return GH_OK;
    case 2: *ptr = w->rec2.ptr; *bytes_per_rank = sizeof(RankRec2); return GH_OK;
    case 3: *ptr = w->node.ptr; *bytes_per_rank = (int64_t)sizeof(Node<float>) * w->ph.ecap; return GH_OK;
    case 4: *ptr = w->sidx_all.ptr; *bytes_per_rank = (int64_t)sizeof(int) * w->ph.n; return GH_OK;
    case 5: *ptr = w->acc_all.ptr; *bytes_per_rank = (int64_t)sizeof(float4) * w->ph.slots; return GH_OK;
    default: return GH_EINVAL;
  }
}

// Bootstrap of the distributed build: equal-count key ranges from the fully sorted keys of the
// single-rank build that just ran in this workspace (every later step derives the next step's
// ranges from the gathered key samples, stitch_kernel).
int tree_splitters(TreeWorkspace *w, int world, cudaStream_t st) {
  if (!w || !w->ctl.ptr || !w->ph.shi || w->ph.dist) { set_error("tree_splitters: needs a finished single-rank build"); return GH_ESTATE; }
  if (world < 1 || world > DIST_MAX_RANKS) { set_error("tree_splitters: 1..%d ranks", DIST_MAX_RANKS); return GH_EINVAL; }
  splitters_from_sorted_kernel<<<1, DIST_MAX_RANKS + 1, 0, st>>>(w->ph.shi, w->ph.n, world, w->ctl.as<BuildCtl>());
  GH_LAUNCH_CHECK();
  return GH_OK;
}

}  // namespace gh
