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
//                       8 bits per pass, MATCH.ANY ranking, shared-memory staged coalesced
//                       scatter (sortscan.cuh); no library kernels anywhere in the build
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
//                       by galloping search for the end of the key-prefix run) and its leaf entry
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
#include "walk.cuh"
#include "build.cuh"

#include <climits>
#include <cstdlib>
#include <cstring>

namespace gh {

// ---- workspace + orchestration --------------------------------------------------------------------
struct TreeWorkspace {
  DeviceBuffer root, part, hi, lo, hi2, lo2, lo3, idx, idx2, clev, cnt, base, P;
  DeviceBuffer node, skip, misc, thi, tidx, thi2, tidx2, sorted, bsum, scanlv[4], cntlv[4];
  RadixScratch rs;
  int64_t last_stats[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int *h_pinned = nullptr;  // [0] nentries, [1] maxlevel ; pinned for async readback
  unsigned long long *h_stats = nullptr;
};

TreeWorkspace *tree_workspace_create() { return new TreeWorkspace(); }
void tree_workspace_destroy(TreeWorkspace *w) {
  if (!w) return;
  DeviceBuffer *all[] = {&w->root, &w->part, &w->hi, &w->lo, &w->hi2, &w->lo2, &w->lo3, &w->idx, &w->idx2,
                         &w->clev, &w->cnt, &w->base, &w->P, &w->node, &w->sorted, &w->bsum, &w->scanlv[0], &w->scanlv[1], &w->scanlv[2], &w->scanlv[3], &w->cntlv[0], &w->cntlv[1], &w->cntlv[2], &w->cntlv[3],
                         &w->skip, &w->misc, &w->thi, &w->tidx, &w->thi2, &w->tidx2};
  for (auto *b : all) b->release();
  w->rs.release();
  if (w->h_pinned) cudaFreeHost(w->h_pinned);
  if (w->h_stats) cudaFreeHost(w->h_stats);
  delete w;
}
int tree_last_stats(TreeWorkspace *w, int64_t out[8]) {
  for (int k = 0; k < 8; k++) out[k] = w->last_stats[k];
  return GH_OK;
}

static inline unsigned nblk(int64_t n, int b) { return (unsigned)((n + b - 1) / b); }

// walk selection: GH_WALK_GROUP (default; fp32 only) or GH_WALK_TARGET (the reference's per-target
// criterion; always used in fp64).  Process-wide; GH_TREE_WALK=target|group sets the initial value.
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

// kappa of the hybrid rule (see walk_group_kernel); 0 = off.  GH_WALK_HYBRID=<kappa> sets the initial
// value, gh_set_tree_walk_hybrid() changes it.
static float g_hybrid_kappa = -1.f;
float group_hybrid_kappa() {
  if (g_hybrid_kappa < 0.f) {
    float v = 0.f;
    if (const char *env = getenv("GH_WALK_HYBRID")) { v = (float)atof(env); if (!(v > 0.f) || v > 1.f) v = 0.f; }
    g_hybrid_kappa = v;
  }
  return g_hybrid_kappa;
}
void set_group_hybrid_kappa(double k) { g_hybrid_kappa = (k > 0.0 && k <= 1.0) ? (float)k : 0.f; }

template <class Real> struct GroupWalk {
  static void launch(const Node<Real> *, int, const TargetsView &, int64_t, const double *, float, double,
                     const Epilogue &, unsigned long long *, bool, bool, unsigned, cudaStream_t) {}
};
template <> struct GroupWalk<float> {
  static void launch(const Node<float> *nodes, int nentries, const TargetsView &tv, int64_t ni,
                     const double *root, float eps2, double inv_theta2, const Epilogue &ep,
                     unsigned long long *dstats, bool stats, bool guard, unsigned blocks32, cudaStream_t st) {
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
  walk_group_kernel<W, STATS, GUARD, HYB><<<nb, 32 * W, 0, st>>>(nodes, nentries, tv, ni, root, eps2, inv_theta2, lim, ep, dstats)
#define GH_GWALK1(W, STATS, GUARD) do { if (hyb) GH_GWALK0(W, STATS, GUARD, true); else GH_GWALK0(W, STATS, GUARD, false); } while (0)
#define GH_GWALK(STATS, GUARD) do { if (wpc == 1) GH_GWALK1(1, STATS, GUARD); else if (wpc == 2) GH_GWALK1(2, STATS, GUARD); else GH_GWALK1(4, STATS, GUARD); } while (0)
    if (stats) { if (guard) GH_GWALK(true, true); else GH_GWALK(true, false); }
    else { if (guard) GH_GWALK(false, true); else GH_GWALK(false, false); }
#undef GH_GWALK0
#undef GH_GWALK1
#undef GH_GWALK
  }
};

template <class Src, class Real>
static int tree_impl(const TreeArgs &a, Src src, const float4 *tgt32, TreeWorkspace *w,
                     cudaStream_t st, cudaEvent_t *ev) {
  const int64_t n = a.nj, ni = a.ni;
  if (n > (int64_t)INT_MAX / 48) { set_error("tree: too many particles (%lld)", (long long)n); return GH_EINVAL; }
  const int levels = (sizeof(Real) == 8) ? LEVELS_MAX : LEVELS_HI;
  const bool deep = levels > LEVELS_HI;
  const bool rel_origin = (sizeof(Real) == 4);  // fp32 entries are stored relative to the root centre
  if (!w->h_pinned) GH_CUDA(cudaMallocHost(&w->h_pinned, 4 * sizeof(int)));
  if (!w->h_stats) GH_CUDA(cudaMallocHost(&w->h_stats, 4 * sizeof(unsigned long long)));

  // K3
  const int nb = (int)((n + 256 * 8 - 1) / (256 * 8) < 1024 ? (n + 256 * 8 - 1) / (256 * 8) : 1024);
  GH_TRY(w->root.reserve(sizeof(double) * ROOT_DOUBLES));
  GH_TRY(w->part.reserve(sizeof(double) * 6 * 1024));
  GH_TRY(w->misc.reserve(64));
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
  keys_kernel<<<nblk(n, 256), 256, 0, st>>>(src, n, root, levels, hi, lo, idx);
  GH_LAUNCH_CHECK();

  // K5: stable LSD radix sort (sortscan.cuh) over (lo, hi); both ping-pong buffers are clobbered
  const uint64_t *shi, *slo = nullptr;
  const int *sidx;
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
    shi = inB ? hi2 : hs;
    sidx = inB ? other1 : order1;
    uint64_t *lsorted = (shi == hi2) ? hs : hi2;  // the key buffer not holding the result
    gather_u64<<<nblk(n, 256), 256, 0, st>>>(w->lo3.as<uint64_t>(), sidx, n, lsorted);
    GH_LAUNCH_CHECK();
    slo = lsorted;
  } else {
    GH_TRY(radix_sort_pairs(hi, idx, hi2, idx2, n, 63, w->rs, st, &inB));
    shi = inB ? hi2 : hi;
    sidx = inB ? idx2 : idx;
  }

  // K6a + pre-order offsets: base[p] = sum_{q<p} (cells opened at q + 1), base[n] = entries
  GH_TRY(w->clev.reserve(n));
  GH_TRY(w->cnt.reserve(sizeof(int) * (n + 1)));
  GH_TRY(w->base.reserve(sizeof(int) * (n + 1)));
  signed char *clev = w->clev.as<signed char>();
  int *cnt = w->cnt.as<int>(), *base = w->base.as<int>();
  levels_kernel<<<nblk(n, 256), 256, 0, st>>>(shi, slo, n, levels, clev, cnt);
  GH_LAUNCH_CHECK();
  GH_TRY((chunked_scan<int, InArray<int>>(InArray<int>{cnt}, n, base, w->cntlv, 0, st)));
  GH_CUDA(cudaMemcpyAsync(&w->h_pinned[0], base + n, sizeof(int), cudaMemcpyDeviceToHost, st));

  // K7
  GH_TRY(w->sorted.reserve(sizeof(double4) * (size_t)n));
  double4 *sp = w->sorted.as<double4>();
  gather_sorted_kernel<<<nblk(n, 256), 256, 0, st>>>(src, sidx, n, sp);
  GH_LAUNCH_CHECK();
  using Mom = typename MomentOf<Real>::type;
  GH_TRY(w->P.reserve(sizeof(Mom) * (size_t)(n + 1)));
  Mom *P = w->P.as<Mom>();
  if (sizeof(Real) == 4) {
    GH_TRY((chunked_scan<D4, InParticlesRel>(InParticlesRel{sp, root}, n, reinterpret_cast<D4 *>(P), w->scanlv, 0, st)));
  } else {
    GH_TRY((chunked_scan<DD4, InParticles>(InParticles{sp}, n, reinterpret_cast<DD4 *>(P), w->scanlv, 0, st)));
  }

  // entries: need the count on the host to size the arrays
  GH_CUDA(cudaStreamSynchronize(st));
  const int nentries = w->h_pinned[0];
  GH_TRY(w->node.reserve(sizeof(Node<Real>) * (size_t)nentries));
  if (sizeof(Real) == 4) {
    if (nentries >= (1 << SKIP_BITS)) { set_error("tree: %d entries exceed the fp32 node format (2^%d)", nentries, SKIP_BITS); return GH_EINVAL; }
  } else {
    GH_TRY(w->skip.reserve(sizeof(int) * (size_t)nentries));
  }
  Entries<Real> E{w->node.as<Node<Real>>(), sizeof(Real) == 4 ? nullptr : w->skip.as<int>()};
  const double inv_theta2 = 1.0 / (a.theta * a.theta);  // theta = 0 -> inf: cells are never accepted
  int *maxlevel = w->misc.as<int>();
  unsigned long long *dstats = reinterpret_cast<unsigned long long *>(w->misc.as<char>() + 16);
  GH_CUDA(cudaMemsetAsync(w->misc.ptr, 0, 64, st));
  emit_kernel<Src, Real><<<nblk(n, 128), 128, 0, st>>>(sp, shi, slo, clev, base, P, n, root, rel_origin,
                                                     inv_theta2, E, maxlevel);
  GH_LAUNCH_CHECK();

  // targets: Morton order.  Self case: the source order restricted to the owned slice is the
  // sorted order itself when the slice is everything; otherwise sort the targets' own keys.
  TargetsView tv;
  tv.sorted = nullptr;
  tv.pos64 = tgt32 ? nullptr : a.tgt_pos;
  tv.pos32 = tgt32;
  tv.order = nullptr;
  tv.order_offset = 0;
  if (a.targets_are_sources && ni == n) {
    tv.order = sidx;
    tv.sorted = sp;
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
  const Real eps2 = (Real)(a.eps * a.eps);
  const int64_t nwarps = (ni + 31) / 32;
  int wb = 128;
  if (const char *env = getenv("GH_WALK_BLOCK")) { int v = atoi(env); if (v == 32 || v == 64 || v == 128) wb = v; }
  const bool group = (sizeof(Real) == 4) && tree_walk_mode() == GH_WALK_GROUP;
  const unsigned blocks = (unsigned)((nwarps + wb / 32 - 1) / (wb / 32));
  if (ev) GH_CUDA(cudaEventRecord(ev[0], st));
  const bool guard = (a.eps == 0.0);
  // L2 prefetch hint of each entry's skip target; GH_WALK_PREFETCH=0/1 overrides
  bool prefetch = false;
  if (const char *env = getenv("GH_WALK_PREFETCH")) prefetch = atoi(env) != 0;
  if (group) {
    GroupWalk<Real>::launch(E.node, nentries, tv, ni, root, (float)eps2, inv_theta2, a.ep, dstats,
                            a.want_stats, guard, (unsigned)nwarps, st);
  } else {
#define GH_WALK(STATS, GUARD, PF)                                                                      \
  walk_kernel<Real, STATS, GUARD, PF><<<blocks, wb, 0, st>>>(E.node, E.skip, nentries, tv, ni, root, \
                                                            rel_origin, eps2, inv_theta2, a.ep, dstats)
#define GH_WALK2(STATS, GUARD) do { if (prefetch) GH_WALK(STATS, GUARD, true); else GH_WALK(STATS, GUARD, false); } while (0)
    if (a.want_stats) { if (guard) GH_WALK2(true, true); else GH_WALK2(true, false); }
    else { if (guard) GH_WALK2(false, true); else GH_WALK2(false, false); }
#undef GH_WALK2
#undef GH_WALK
  }
  GH_LAUNCH_CHECK();
  if (ev) GH_CUDA(cudaEventRecord(ev[1], st));
  GH_CUDA(cudaMemcpyAsync(&w->h_pinned[1], maxlevel, sizeof(int), cudaMemcpyDeviceToHost, st));
  w->last_stats[0] = nentries;
  w->last_stats[1] = nentries - n;
  if (a.want_stats) {
    GH_CUDA(cudaMemcpyAsync(w->h_stats, dstats, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    GH_CUDA(cudaStreamSynchronize(st));
    w->last_stats[2] = w->h_pinned[1];
    w->last_stats[3] = (int64_t)w->h_stats[0];
    w->last_stats[4] = (int64_t)w->h_stats[1];
    w->last_stats[5] = (int64_t)w->h_stats[2];
    w->last_stats[6] = (int64_t)w->h_stats[3];
    w->last_stats[7] = (int64_t)((ni + 31) / 32);
  }
  return GH_OK;
}

int launch_tree(const TreeArgs &a, TreeWorkspace *w, cudaStream_t st, cudaEvent_t *ev) {
  if (a.ni <= 0 || a.nj <= 0) return GH_OK;
  if (a.prec == GH_PREC_F64) {
    Src64 s{a.src_pos, a.src_mass};
    return tree_impl<Src64, double>(a, s, nullptr, w, st, ev);
  } else if (a.prec == GH_PREC_F32) {
    if (a.src32) {  // f32 engine: sources and targets are float4 (x - origin, m)
      Src32 s{a.src32};
      return tree_impl<Src32, float>(a, s, a.tgt32, w, st, ev);
    }
    Src64 s{a.src_pos, a.src_mass};
    return tree_impl<Src64, float>(a, s, nullptr, w, st, ev);
  }
  set_error("launch_tree: bad precision %d", a.prec);
  return GH_EINVAL;
}

}  // namespace gh
