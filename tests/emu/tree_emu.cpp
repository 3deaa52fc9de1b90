// tree_emu.cpp -- the tree BUILD kernels (gravhopper_b200/csrc/build.cuh, sortscan.cuh) and the walk
// kernels (walk.cuh), compiled for the host and run through emu::launch: TEST INFRASTRUCTURE.
// emu_tree_build mirrors the launch sequence of tree_impl (csrc/tree.cu): bbox -> keys -> stable LSD
// radix sort -> common levels -> pre-order offsets -> gather -> moment scan -> emit; the kernels are
// the product's own source, the sequence is restated here (tree.cu's talks to the CUDA runtime).
#include "emu_shim.h"

#include "../../gravhopper_b200/csrc/build.cuh"
#include "../../gravhopper_b200/csrc/bucketsort.cuh"

#include <cstdlib>

namespace gh {
void set_error(const char *, ...) {}
int64_t &launch_counter() { static thread_local int64_t c = 0; return c; }
}  // namespace gh

using namespace gh;

static inline unsigned nblk(int64_t n, int b) { return (unsigned)((n + b - 1) / b); }

// chunked_scan of sortscan.cuh, launch for launch
template <class T, class In> static void emu_scan(In in, int64_t n, T *P, const int *ndev = nullptr) {
  if (n <= 0) return;
  if (n <= SCAN_WARP_ELEMS) {
    emu::launch(1, 32, [&] { scan_phase3<T, In>(in, n, nullptr, P, ndev); });
    return;
  }
  const int64_t nw = (n + SCAN_WARP_ELEMS - 1) / SCAN_WARP_ELEMS;
  const unsigned nsb = (unsigned)((nw * 32 + SCAN_THREADS - 1) / SCAN_THREADS);
  std::vector<T> buf((size_t)(2 * nw + 1));
  T *totals = buf.data(), *prefix = totals + nw;
  emu::launch(nsb, SCAN_THREADS, [&] { scan_phase1<T, In>(in, n, totals, ndev); });
  emu_scan<T, InArray<T>>(InArray<T>{totals}, nw, prefix);
  emu::launch(nsb, SCAN_THREADS, [&] { scan_phase3<T, In>(in, n, prefix, P, ndev); });
}

// radix_sort_pairs of sortscan.cuh, launch for launch
static bool emu_sort(uint64_t *kA, int *vA, uint64_t *kB, int *vB, int64_t n, int nbits, const int *ndev = nullptr) {
  if (n <= 1) return false;
  const int nblocks = (int)((n + RS_TILE - 1) / RS_TILE);
  const int npass = (nbits + 7) / 8;
  std::vector<int> hist((size_t)RS_RADIX * nblocks), gtot((size_t)RS_RADIX * npass, 0);
  uint64_t *kin = kA, *kout = kB;
  int *vin = vA, *vout = vB;
  bool inB = false;
  for (int pass = 0; pass < npass; pass++) {
    const int shift = 8 * pass;
    int *g = gtot.data() + RS_RADIX * pass;
    emu::launch((unsigned)nblocks, RS_THREADS, [&] { rs_hist_kernel(kin, n, shift, hist.data(), nblocks, g, ndev); });
    emu::launch(RS_RADIX, RS_THREADS, [&] { rs_rowscan_kernel(hist.data(), nblocks); });
    emu::launch((unsigned)nblocks, RS_THREADS, [&] { rs_scatter_kernel(kin, vin, kout, vout, n, shift, hist.data(), g, nblocks, ndev); });
    std::swap(kin, kout);
    std::swap(vin, vout);
    inB = !inB;
  }
  return inB;
}

// level-min tables for n particles (tree.cu phase B): storage + the view the kernels take
struct EmuLevelMin {
  std::vector<unsigned char> buf;
  LevelMin lm;
  explicit EmuLevelMin(int64_t n) {
    size_t off[LM_MAX_TABLES];
    const size_t bytes = lm_layout(n, lm, off);
    buf.assign(bytes + 64, 0);
    unsigned char *base = buf.data();
    base += (32 - (reinterpret_cast<uintptr_t>(base) & 31)) & 31;
    for (int k = 0; k < lm.ntab; k++) lm.t[k] = base + off[k];
  }
  void finish() { if (lm.ntab > 2) emu::launch(1, 1024, [&] { levelmin_top_kernel(lm); }); }
};

template <class Real>
static int build_impl(const double *pos, const double *mass, int64_t n, double eps, double theta, void *nodes_out,
                      int *skips_out, int nodes_cap, double *sorted_out, int *order_out, double *root_out,
                      int *info_out, uint64_t *keys_out, int seg_cap, void *quad_out) {
  using Mom = typename MomentOf<Real>::type;
  const int levels = (sizeof(Real) == 8) ? LEVELS_MAX : LEVELS_HI;
  const bool deep = levels > LEVELS_HI;
  const bool rel_origin = (sizeof(Real) == 4);
  Src64 src{pos, mass};
  std::vector<double> root(ROOT_DOUBLES), part(6 * 1024);
  const int nb = (int)((n + 256 * 8 - 1) / (256 * 8) < 1024 ? (n + 256 * 8 - 1) / (256 * 8) : 1024);
  emu::launch((unsigned)nb, 256, [&] { bbox_stage1<Src64>(src, n, part.data()); });
  emu::launch(1, 256, [&] { bbox_stage2(part.data(), nb, eps, root.data()); });

  std::vector<uint64_t> hi(n), hi2(n), lo(deep ? n : 0), lo2(deep ? n : 0), lo3(deep ? n : 0);
  std::vector<int> idx(n), idx2(n);
  uint64_t *lop = deep ? lo.data() : nullptr;
  emu::launch(nblk(n, 256), 256, [&] { keys_kernel<Src64>(src, n, root.data(), levels, hi.data(), lop, idx.data()); }, true);

  const uint64_t *shi, *slo = nullptr;
  const int *sidx;
  if (deep) {
    lo3 = lo;
    bool inB = emu_sort(lo.data(), idx.data(), lo2.data(), idx2.data(), n, 63);
    int *order1 = inB ? idx2.data() : idx.data();
    int *other1 = inB ? idx.data() : idx2.data();
    uint64_t *hs = inB ? lo.data() : lo2.data();
    emu::launch(nblk(n, 256), 256, [&] { gather_u64(hi.data(), order1, n, hs); }, true);
    inB = emu_sort(hs, order1, hi2.data(), other1, n, 63);
    shi = inB ? hi2.data() : hs;
    sidx = inB ? other1 : order1;
    uint64_t *lsorted = (shi == hi2.data()) ? hs : hi2.data();
    emu::launch(nblk(n, 256), 256, [&] { gather_u64(lo3.data(), sidx, n, lsorted); }, true);
    slo = lsorted;
  } else {
    const bool inB = emu_sort(hi.data(), idx.data(), hi2.data(), idx2.data(), n, 63);
    shi = inB ? hi2.data() : hi.data();
    sidx = inB ? idx2.data() : idx.data();
  }

  std::vector<signed char> clev(n);
  std::vector<int> cnt(n + 1), base(n + 1);
  EmuLevelMin elm(n);
  emu::launch(nblk(n, 256), 256, [&] { levels_kernel(shi, slo, n, levels, clev.data(), cnt.data(), nullptr, elm.lm); });
  elm.finish();
  emu_scan<int, InArray<int>>(InArray<int>{cnt.data()}, n, base.data());

  std::vector<double4> sp(n);
  emu::launch(nblk(n, 256), 256, [&] { gather_sorted_kernel<Src64>(src, sidx, n, sp.data()); }, true);
  std::vector<Mom> P(n + 1);
  if (sizeof(Real) == 4)
    emu_scan<D4, InParticlesRel>(InParticlesRel{sp.data(), root.data()}, n, reinterpret_cast<D4 *>(P.data()));
  else
    emu_scan<DD4, InParticles>(InParticles{sp.data()}, n, reinterpret_cast<DD4 *>(P.data()));

  const int nentries = base[n];
  info_out[0] = nentries;
  if (nentries > nodes_cap) return 1;
  Entries<Real> E{reinterpret_cast<Node<Real> *>(nodes_out), sizeof(Real) == 4 ? nullptr : skips_out};
  const double inv_theta2 = 1.0 / (theta * theta);
  int maxlevel = 0;
  // one segment whose capacity is the entry count itself, so that the chain end == entry count
  // (the product sizes the segment generously and ends the chains at the capacity)
  BuildCtl ctl;
  std::memset(&ctl, 0, sizeof(ctl));
  emu::launch(1, 1, [&] { ctl_init_single(&ctl, (int)n, seg_cap > 0 ? seg_cap : nentries); }, true);
  std::vector<D6> P2(quad_out ? n + 1 : 0);
  if (quad_out)  // opt-in quadrupoles: second-moment prefixes, per-entry tensors (6 Reals per entry)
    emu_scan<D6, InSecondRel>(InSecondRel{sp.data(), root.data()}, n, P2.data());
  emu::launch(nblk(n, 128), 128, [&] {
    emit_kernel<Src64, Real>(sp.data(), shi, slo, clev.data(), base.data(), P.data(), n, root.data(), rel_origin,
                             inv_theta2, E, &maxlevel, &ctl, false, quad_out ? P2.data() : nullptr,
                             reinterpret_cast<Real *>(quad_out));
  }, true);
  info_out[1] = maxlevel;
  info_out[2] = ctl.overflow;
  if (sizeof(Real) == 4 && !quad_out) {
    // the warp-cooperative form of the emit (emit32_warp_kernel) must write the same array, bit for bit
    const int filled = ctl.stride < nentries ? ctl.stride : nentries;
    std::vector<Node<float>> alt((size_t)filled + 1);
    std::memset(alt.data(), 0xff, sizeof(Node<float>) * alt.size());
    BuildCtl ctl2;
    std::memset(&ctl2, 0, sizeof(ctl2));
    emu::launch(1, 1, [&] { ctl_init_single(&ctl2, (int)n, seg_cap > 0 ? seg_cap : nentries); }, true);
    int maxlevel2 = 0;
    Entries<float> E2{alt.data(), nullptr};
    for (int tables = 0; tables < 2; tables++) {  // cell ends from the keys, then from the level-min tables
      LevelMin none = elm.lm;
      none.ntab = 0;
      const LevelMin use = tables ? elm.lm : none;
      std::memset(alt.data(), 0xff, sizeof(Node<float>) * alt.size());
      maxlevel2 = 0;
      ctl2.overflow = 0;
      emu::launch(nblk(n, 128), 128, [&] {
        emit32_warp_kernel(sp.data(), shi, clev.data(), base.data(), reinterpret_cast<const D4 *>(P.data()), n, root.data(),
                           E2, &maxlevel2, &ctl2, false, use);
      });
      if (maxlevel2 != maxlevel || ctl2.overflow != ctl.overflow) return 7 + tables;
      if (std::memcmp(alt.data(), nodes_out, sizeof(Node<float>) * (size_t)filled) != 0) return 7 + tables;
    }
  }
  if (keys_out) std::memcpy(keys_out, shi, sizeof(uint64_t) * (size_t)n);
  std::memcpy(sorted_out, sp.data(), sizeof(double4) * (size_t)n);
  std::memcpy(order_out, sidx, sizeof(int) * (size_t)n);
  std::memcpy(root_out, root.data(), sizeof(double) * ROOT_DOUBLES);
  return 0;
}

// splitter_sort_pairs of bucketsort.cuh, launch for launch (result in kA / vA)
static void emu_splitter_sort(uint64_t *kA, int *vA, uint64_t *kB, int *vB, int64_t n, int nbits, const uint64_t *spl,
                              int nb, const int *ndev) {
  const int nblocks = (int)((n + RS_TILE - 1) / RS_TILE);
  std::vector<int> hist((size_t)RS_RADIX * nblocks), gtot((size_t)RS_RADIX * 2, 0), boff(nb + 2);
  const BucketOf bo{spl, nb};
  uint64_t *kin = kA, *kout = kB;
  int *vin = vA, *vout = vB;
  for (int pass = 0; pass < 2; pass++) {
    const BucketDigit dg{bo, 8 * pass};
    int *g = gtot.data() + RS_RADIX * pass;
    emu::launch((unsigned)nblocks, RS_THREADS, [&] { bs_hist_kernel(kin, n, dg, hist.data(), nblocks, g, ndev); });
    emu::launch(RS_RADIX, RS_THREADS, [&] { rs_rowscan_kernel(hist.data(), nblocks); });
    emu::launch((unsigned)nblocks, RS_THREADS, [&] { bs_scatter_kernel(kin, vin, kout, vout, n, dg, hist.data(), g, nblocks, ndev); });
    std::swap(kin, kout);
    std::swap(vin, vout);
  }
  emu::launch((unsigned)((nb + 1 + 255) / 256), 256, [&] { bs_offsets_kernel(kA, n, bo, boff.data(), ndev); }, true);
  emu::launch((unsigned)nb, RS_THREADS, [&] { bs_bucket_kernel(kA, vA, kB, vB, boff.data(), spl, nb, nbits); });
}

// splitter_place_sort_pairs of bucketsort.cuh, launch for launch (result in kB / vB)
static void emu_place_sort(uint64_t *kA, int *vA, uint64_t *kB, int *vB, int64_t n, int nbits, const uint64_t *spl,
                           int nb, const int *ndev) {
  std::vector<unsigned short> bid((size_t)n);
  std::vector<int> count(nb + 1, 0), cursor(nb + 1, 0), boff(nb + 2, 0);
  const unsigned nblocks = (unsigned)((n + BP_THREADS * BP_ROUNDS - 1) / (BP_THREADS * BP_ROUNDS));
  emu::launch(nblocks, BP_THREADS, [&] { bp_count_kernel(kA, n, spl, nb, bid.data(), count.data(), ndev); });
  emu::launch(1, RS_THREADS, [&] { bp_scan_kernel(count.data(), nb, boff.data(), cursor.data()); });
  emu::launch(nblocks, BP_THREADS, [&] { bp_place_kernel(kA, vA, bid.data(), n, cursor.data(), kB, vB, ndev); });
  int vbits = 1;
  while (vbits < 31 && (int64_t(1) << vbits) < n) vbits++;
  emu::launch((unsigned)nb, RS_THREADS, [&] { bp_bucket_kernel(kB, vB, kA, vA, boff.data(), spl, nb, nbits, vbits); });
}

// splitter_place2_sort_pairs of bucketsort.cuh, launch for launch (result in kB / vB)
static void emu_place2_sort(uint64_t *kA, int *vA, uint64_t *kB, int *vB, int64_t n, int nbits, const uint64_t *spl,
                            int nb, int G, const int *ndev) {
  std::vector<unsigned short> bid((size_t)n);
  std::vector<int> ghist((size_t)G * nb, -1), tot(nb + 1, 0), cursor(nb + 1, 0), boff(nb + 2, 0);
  emu::launch((unsigned)G, BP2_THREADS, [&] { bp2_count_kernel(kA, n, spl, nb, bid.data(), ghist.data(), ndev); });
  emu::launch((unsigned)((nb + 31) / 32), 256, [&] { bp2_colscan_kernel(ghist.data(), G, nb, tot.data()); });
  emu::launch(1, RS_THREADS, [&] { bp_scan_kernel(tot.data(), nb, boff.data(), cursor.data()); });
  emu::launch((unsigned)G, BP2_THREADS, [&] { bp2_place_kernel(kA, vA, bid.data(), n, ghist.data(), boff.data(), nb, kB, vB, ndev); });
  int vbits = 1;
  while (vbits < 31 && (int64_t(1) << vbits) < n) vbits++;
  emu::launch((unsigned)nb, RS_THREADS, [&] { bp_bucket_kernel(kB, vB, kA, vA, boff.data(), spl, nb, nbits, vbits); });
}

extern "C" {
// keys[n] (63-bit), vals = 0..n-1.  The splitters are the (b n_spl / nb)-th keys of the sorted array
// `spl_from` (n_spl keys: the same data = fresh splitters, other data = stale ones).  n_real <= n:
// the device-side count.  Outputs the splitter sort's keys / vals; returns 0.
// stats_out (nullable, 4): which way the "place" form's buckets went (g_bp_stats of bucketsort.cuh).
// place == 1: the "place" form with global atomics; place >= 2: its shared-memory form with G = place CTAs.
// place != 0: the "place" form (one counting and one placing pass, compact in-bucket ranking).
int emu_splitter_sort_test(const uint64_t *keys, int64_t n, int64_t n_real, const uint64_t *spl_from, int64_t n_spl,
                           int nb, uint64_t *keys_out, int *vals_out, int place, long long *stats_out) {
  for (int k = 0; k < 4; k++) g_bp_stats[k] = 0;
  std::vector<uint64_t> kA(keys, keys + n), kB(n);
  std::vector<int> vA(n), vB(n);
  for (int64_t i = 0; i < n; i++) vA[i] = (int)i;
  std::vector<uint64_t> spl(nb + 1);
  const int nsp = (int)n_spl;
  emu::launch((unsigned)((nb + 255) / 256), 256, [&] { bs_splitters_kernel(spl_from, n_spl, nb, spl.data(), &nsp); }, true);
  const int nr = (int)n_real;
  if (place) {
    if (place >= 2) emu_place2_sort(kA.data(), vA.data(), kB.data(), vB.data(), n, 63, spl.data(), nb, place, &nr);  // G = place
    else emu_place_sort(kA.data(), vA.data(), kB.data(), vB.data(), n, 63, spl.data(), nb, &nr);
    std::memcpy(keys_out, kB.data(), sizeof(uint64_t) * (size_t)n_real);
    std::memcpy(vals_out, vB.data(), sizeof(int) * (size_t)n_real);
    if (stats_out) for (int k = 0; k < 4; k++) stats_out[k] = g_bp_stats[k];
    return 0;
  }
  emu_splitter_sort(kA.data(), vA.data(), kB.data(), vB.data(), n, 63, spl.data(), nb, &nr);
  std::memcpy(keys_out, kA.data(), sizeof(uint64_t) * (size_t)n_real);
  std::memcpy(vals_out, vA.data(), sizeof(int) * (size_t)n_real);
  return 0;
}

// cell_end (build.cuh) on a given array of common levels: tables 0 and 1 as levels_kernel leaves them
// (built here on the host), the upper tables by levelmin_top_kernel; out[i] = cell_end(p[i], level[i]).
int emu_cell_end_test(const signed char *clev, int64_t n, const int64_t *p, const int *level, int64_t nq, int64_t *out,
                      int *ntab_out) {
  EmuLevelMin elm(n);
  for (int64_t q = 0; q < n; q++) elm.lm.t[0][q] = (unsigned char)(clev[q] + 1);
  if (elm.lm.ntab > 1)
    for (int64_t j = 0; j < elm.lm.n[1]; j++) {
      unsigned m = 255;
      for (int i = 0; i < 32; i++) {
        const int64_t q = 32 * j + i;
        const unsigned v = q < n ? (unsigned)(clev[q] + 1) : 0u;
        m = v < m ? v : m;
      }
      elm.lm.t[1][j] = (unsigned char)m;
    }
  elm.finish();
  *ntab_out = elm.lm.ntab;
  emu::launch((unsigned)((nq + 255) / 256), 256, [&] {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nq) out[i] = cell_end(elm.lm, p[i], level[i], n);
  }, true);
  return 0;
}

// prec 32: Node<float> entries (8 floats each) into nodes_out; prec 64: Node<double> + skips_out.
// info_out = {entries, deepest cell level}.  Returns 1 if nodes_cap is too small (info_out[0] = need).
// info_out = {entries, deepest cell level, overflow flag}; keys_out (nullable): the sorted `hi` keys;
// seg_cap > 0: capacity of the segment (chains end there; entries beyond it raise the overflow flag).
int emu_tree_build(int prec, const double *pos, const double *mass, int64_t n, double eps, double theta,
                   void *nodes_out, int *skips_out, int nodes_cap, double *sorted_out, int *order_out,
                   double *root_out, int *info_out, uint64_t *keys_out, int seg_cap, void *quad_out) {
  // quad_out (nullable): 6 values per entry, the opt-in quadrupole tensors
  if (prec == 32)
    return build_impl<float>(pos, mass, n, eps, theta, nodes_out, skips_out, nodes_cap, sorted_out, order_out,
                             root_out, info_out, keys_out, seg_cap, quad_out);
  return build_impl<double>(pos, mass, n, eps, theta, nodes_out, skips_out, nodes_cap, sorted_out, order_out,
                            root_out, info_out, keys_out, seg_cap, quad_out);
}

// The DISTRIBUTED fp32 build (tree.cu phases A-C with TreeDist) for `world` ranks run one after the
// other on the host; the three all-gathers are shared arrays every rank's kernels write their slot
// of.  split[world+1]: the key ranges (split[0] = 0; the last rank is unbounded above).
// nodes_out: (world * stride, 8) floats, the global virtual-index entry array.
// counts_out[2*world]: {n_local, entries} per rank; split_next_out[world+1]: the next step's
// ranges as rank 0's stitch derived them; sorted_out / order_out: the ranks' sorted particles
// concatenated.  Returns 0, or 2 when a rank raised the overflow flag.
int emu_tree_build_dist(int world, const uint64_t *split, const double *pos, const double *mass, int64_t n,
                        double eps, double theta, float *nodes_out, int stride, double *sorted_out, int *order_out,
                        double *root_out, int *counts_out, uint64_t *split_next_out, int *maxlevel_out,
                        double *acc_out, int blk) {
  // acc_out (nullable, (n,3)): ALSO run the distributed walk -- every rank walks its share of the
  // global Morton order (blocks of `blk` dealt round-robin) into its slot of the gathered
  // acceleration buffer, then every rank's dist_epilogue_kernel stores the accelerations of the
  // particles it owns (contiguous even split of the source indices) into acc_out.
  const int levels = LEVELS_HI;
  Src64 src{pos, mass};
  std::vector<double> root(ROOT_DOUBLES), part(6 * 1024);
  const int nb = (int)((n + 256 * 8 - 1) / (256 * 8) < 1024 ? (n + 256 * 8 - 1) / (256 * 8) : 1024);
  emu::launch((unsigned)nb, 256, [&] { bbox_stage1<Src64>(src, n, part.data()); });
  emu::launch(1, 256, [&] { bbox_stage2(part.data(), nb, eps, root.data()); });
  std::vector<uint64_t> kall(n);
  std::vector<int> dummy(n);
  emu::launch(nblk(n, 256), 256, [&] { keys_kernel<Src64>(src, n, root.data(), levels, kall.data(), (uint64_t *)nullptr, dummy.data()); }, true);

  struct Rank {
    BuildCtl ctl;
    std::vector<uint64_t> hi, hi2;
    std::vector<int> idx, idx2, cnt, base;
    std::vector<signed char> clev;
    std::vector<double4> sp;
    std::vector<D4> P;
    const uint64_t *shi;
    const int *sidx;
    std::unique_ptr<EmuLevelMin> elm;
  };
  std::vector<Rank> R(world);
  std::vector<RankRec1> rec1(world);
  std::vector<RankRec2> rec2(world);
  std::vector<int> sidx_all((size_t)world * n, -1);  // the gathered sorted-index array (ncap = n)
  const int ntiles = (int)((n + SEL_TILE - 1) / SEL_TILE);
  // phase A
  for (int r = 0; r < world; r++) {
    Rank &k = R[r];
    std::memset(&k.ctl, 0, sizeof(k.ctl));
    for (int j = 0; j <= world; j++) k.ctl.split_next[j] = split[j];
    k.hi.assign(n, 0); k.hi2.assign(n, 0); k.idx.assign(n, 0); k.idx2.assign(n, 0);
    emu::launch(1, 1, [&] { ctl_init_dist(&k.ctl, r, world, stride); }, true);
    emu::launch(1, DIST_MAX_RANKS + 1, [&] { splitters_advance_kernel(&k.ctl, world); }, true);
    std::vector<int> tilecnt(ntiles + 1), tileoff(ntiles + 2);
    emu::launch((unsigned)ntiles, SEL_THREADS, [&] { select_count_kernel(kall.data(), n, &k.ctl, tilecnt.data()); });
    emu_scan<int, InArray<int>>(InArray<int>{tilecnt.data()}, ntiles, tileoff.data());
    emu::launch((unsigned)ntiles, SEL_THREADS, [&] { select_compact_kernel(kall.data(), n, &k.ctl, tileoff.data(), ntiles, k.hi.data(), k.idx.data(), n); });
    const bool inB = emu_sort(k.hi.data(), k.idx.data(), k.hi2.data(), k.idx2.data(), n, 63, &k.ctl.n_local);
    k.shi = inB ? k.hi2.data() : k.hi.data();
    k.sidx = inB ? k.idx2.data() : k.idx.data();
    emu::launch(1, 1, [&] { rec1_kernel(k.shi, &k.ctl, rec1.data()); }, true);
  }
  // phase B (after the all-gather of rec1)
  for (int r = 0; r < world; r++) {
    Rank &k = R[r];
    emu::launch(1, 1, [&] { neighbours_kernel(rec1.data(), &k.ctl); }, true);
    k.clev.assign(n, 0); k.cnt.assign(n + 1, 0); k.base.assign(n + 1, 0);
    k.elm.reset(new EmuLevelMin(n));
    emu::launch(nblk(n, 256), 256, [&] { levels_kernel(k.shi, nullptr, n, levels, k.clev.data(), k.cnt.data(), &k.ctl, k.elm->lm); });
    k.elm->finish();
    emu_scan<int, InArray<int>>(InArray<int>{k.cnt.data()}, n, k.base.data(), &k.ctl.n_local);
    k.sp.assign(n, double4{0, 0, 0, 0});
    emu::launch(nblk(n, 256), 256, [&] { gather_sorted_kernel<Src64>(src, k.sidx, n, k.sp.data(), &k.ctl, sidx_all.data() + (size_t)r * n); }, true);
    k.P.assign(n + 1, D4{{0, 0, 0, 0}});
    emu_scan<D4, InParticlesRel>(InParticlesRel{k.sp.data(), root.data()}, n, k.P.data(), &k.ctl.n_local);
    emu::launch(1, 32, [&] { rec2_kernel(k.shi, k.base.data(), k.P.data(), &k.ctl, rec2.data()); });
  }
  // phase C (after the all-gather of rec2): stitch + emit into the rank's segment
  const double inv_theta2 = 1.0 / (theta * theta);
  int maxlevel = 0, rc = 0;
  int64_t ofs = 0;
  Entries<float> E{reinterpret_cast<Node<float> *>(nodes_out), nullptr};
  for (int r = 0; r < world; r++) {
    Rank &k = R[r];
    emu::launch(1, 1, [&] { stitch_kernel(rec1.data(), rec2.data(), &k.ctl, n); }, true);
    emu::launch(nblk(n, 128), 128, [&] {
      emit_kernel<Src64, float>(k.sp.data(), k.shi, nullptr, k.clev.data(), k.base.data(), k.P.data(), n, root.data(),
                                true, inv_theta2, E, &maxlevel, &k.ctl, true);
    }, true);
    {  // the warp-cooperative form writes the same segment
      std::vector<Node<float>> alt((size_t)world * stride);
      std::memset(alt.data(), 0xff, sizeof(Node<float>) * alt.size());
      BuildCtl c2 = k.ctl;
      c2.overflow = 0;
      int ml2 = 0;
      Entries<float> E2{alt.data(), nullptr};
      const int fill = k.ctl.nentries < stride ? k.ctl.nentries : stride;
      for (int tables = 0; tables < 2; tables++) {
        LevelMin use = k.elm->lm;
        if (!tables) use.ntab = 0;
        std::memset(alt.data(), 0xff, sizeof(Node<float>) * alt.size());
        ml2 = 0;
        emu::launch(nblk(n, 128), 128, [&] {
          emit32_warp_kernel(k.sp.data(), k.shi, k.clev.data(), k.base.data(), k.P.data(), n, root.data(), E2, &ml2, &c2, true, use);
        });
        if (std::memcmp(alt.data() + (size_t)r * stride, E.node + (size_t)r * stride, sizeof(Node<float>) * (size_t)fill) != 0) return 7 + tables;
        if (ml2 > maxlevel) return 7 + tables;
      }
    }
    if (k.ctl.overflow) rc = 2;
    counts_out[2 * r] = k.ctl.n_local;
    counts_out[2 * r + 1] = k.ctl.nentries;
    std::memcpy(sorted_out + 4 * ofs, k.sp.data(), sizeof(double4) * (size_t)k.ctl.n_local);
    std::memcpy(order_out + ofs, k.sidx, sizeof(int) * (size_t)k.ctl.n_local);
    ofs += k.ctl.n_local;
  }
  if (acc_out) {
    const int T = (int)(((n + blk - 1) / blk + world - 1) / world);
    const int64_t slots = (int64_t)world * T * blk;
    std::vector<float4> acc_all((size_t)world * slots, float4{0, 0, 0, 0});
    std::vector<float4> pos32(n);
    for (int64_t i = 0; i < n; i++)
      pos32[i] = make_float4((float)pos[3 * i], (float)pos[3 * i + 1], (float)pos[3 * i + 2], (float)mass[i]);
    const float eps2 = (float)(eps * eps);
    for (int r = 0; r < world; r++) {  // walk
      TargetsView tv;
      std::memset(&tv, 0, sizeof(tv));
      tv.pos32 = pos32.data();
      tv.dist_sidx = sidx_all.data();
      tv.dist_counts = R[r].ctl.counts;
      tv.dist_rank = r; tv.dist_world = world; tv.dist_ncap = (int)n; tv.dist_blk = blk; tv.dist_T = T;
      Epilogue ep;
      std::memset(&ep, 0, sizeof(ep));
      ep.mode = EP_ACC32;
      ep.acc32_out = acc_all.data() + (size_t)r * slots;
      const int *wc = &R[r].ctl.overflow;
      emu::launch((unsigned)((slots + 31) / 32), 32, [&] {
        walk_group_kernel<1, false, false, false>(E.node, world * stride, tv, slots, root.data(), eps2, inv_theta2, 3000,
                                                  ep, nullptr, wc);
      });
    }
    const int64_t base = n / world, rem = n % world;
    int64_t ib = 0;
    for (int r = 0; r < world; r++) {  // owners
      const int64_t ni = base + (r < rem ? 1 : 0);
      Epilogue ep;
      std::memset(&ep, 0, sizeof(ep));
      ep.mode = EP_ACC;
      ep.acc_out = acc_out + 3 * ib;
      std::vector<int> inv(ni, -1);
      emu::launch(nblk((int64_t)world * n, 256), 256, [&] {
        dist_inverse_kernel(sidx_all.data(), &R[r].ctl, world, (int)n, ib, ni, inv.data());
      }, true);
      emu::launch(nblk(ni, 256), 256, [&] {
        dist_epilogue_kernel(inv.data(), &R[r].ctl, acc_all.data(), world, (int)n, blk, T, ni, ep);
      }, true);
      ib += ni;
    }
  }
  for (int j = 0; j <= world; j++) split_next_out[j] = R[0].ctl.split_next[j];
  std::memcpy(root_out, root.data(), sizeof(double) * ROOT_DOUBLES);
  maxlevel_out[0] = maxlevel;
  maxlevel_out[1] = R[0].ctl.first;  // index of the root entry (first non-empty rank's segment)
  return ofs == n ? rc : 3;
}
}
