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

#include <climits>
#include <cstdlib>
#include <cstring>

namespace gh {

static constexpr int LEVELS_HI = 21;
static constexpr int LEVELS_MAX = 42;

// ---- source / target accessors --------------------------------------------------------------
struct Src64 {
  const double *pos;
  const double *mass;
  __device__ __forceinline__ void get(int64_t j, double &x, double &y, double &z) const {
    x = pos[3 * j]; y = pos[3 * j + 1]; z = pos[3 * j + 2];
  }
  __device__ __forceinline__ double m(int64_t j) const { return mass[j]; }
};
struct Src32 {
  const float4 *p;
  __device__ __forceinline__ void get(int64_t j, double &x, double &y, double &z) const {
    float4 t = p[j]; x = t.x; y = t.y; z = t.z;
  }
  __device__ __forceinline__ double m(int64_t j) const { return p[j].w; }
};

// root[0..2] centre, root[3] side, root[4..6] min, root[7..9] max
static constexpr int ROOT_DOUBLES = 10;

// ---- K3 bbox ----------------------------------------------------------------------------------
template <class Src>
__global__ void bbox_stage1(Src src, int64_t n, double *__restrict__ part) {
  __shared__ double sh[6][256];
  double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    double p[3];
    src.get(i, p[0], p[1], p[2]);
#pragma unroll
    for (int k = 0; k < 3; k++) { mn[k] = fmin(mn[k], p[k]); mx[k] = fmax(mx[k], p[k]); }
  }
  for (int k = 0; k < 3; k++) { sh[k][threadIdx.x] = mn[k]; sh[3 + k][threadIdx.x] = mx[k]; }
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
      for (int k = 0; k < 3; k++) {
        sh[k][threadIdx.x] = fmin(sh[k][threadIdx.x], sh[k][threadIdx.x + s]);
        sh[3 + k][threadIdx.x] = fmax(sh[3 + k][threadIdx.x], sh[3 + k][threadIdx.x + s]);
      }
    }
    __syncthreads();
  }
  if (threadIdx.x < 6) part[blockIdx.x * 6 + threadIdx.x] = sh[threadIdx.x][0];
}

__global__ void bbox_stage2(const double *__restrict__ part, int nblocks, double eps,
                            double *__restrict__ root) {
  __shared__ double sh[6][256];
  double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
  for (int b = threadIdx.x; b < nblocks; b += blockDim.x)
    for (int k = 0; k < 3; k++) {
      mn[k] = fmin(mn[k], part[b * 6 + k]);
      mx[k] = fmax(mx[k], part[b * 6 + 3 + k]);
    }
  for (int k = 0; k < 3; k++) { sh[k][threadIdx.x] = mn[k]; sh[3 + k][threadIdx.x] = mx[k]; }
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
      for (int k = 0; k < 3; k++) {
        sh[k][threadIdx.x] = fmin(sh[k][threadIdx.x], sh[k][threadIdx.x + s]);
        sh[3 + k][threadIdx.x] = fmax(sh[3 + k][threadIdx.x], sh[3 + k][threadIdx.x + s]);
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    double mnv[3], mxv[3];
    for (int k = 0; k < 3; k++) { mnv[k] = sh[k][0]; mxv[k] = sh[3 + k][0]; }
    // _jbgrav.c:764-769: the un-padded extent is compared with the padded running value
    double boxsize = __dadd_rn(__dadd_rn(mxv[0], -mnv[0]), eps);
    for (int k = 1; k < 3; k++) {
      double ext = __dadd_rn(mxv[k], -mnv[k]);
      if (ext > boxsize) boxsize = __dadd_rn(ext, eps);
    }
    for (int k = 0; k < 3; k++) {
      root[k] = __dmul_rn(0.5, __dadd_rn(mnv[k], mxv[k]));  // :770-772
      root[4 + k] = mnv[k];
      root[7 + k] = mxv[k];
    }
    root[3] = boxsize;
  }
}

// ---- K4 keys ----------------------------------------------------------------------------------
// One descent step of gravoct_calc_subnode/_branchnum + the child-centre update (:406,:441-462).
__device__ __forceinline__ unsigned descend(const double p[3], double c[3], double quarter) {
  unsigned d = 0;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    if (p[k] > c[k]) { d |= (1u << k); c[k] = __dadd_rn(c[k], quarter); }
    else c[k] = __dadd_rn(c[k], -quarter);
  }
  return d;
}

template <class Src>
__global__ void keys_kernel(Src src, int64_t n, const double *__restrict__ root, int levels,
                            uint64_t *__restrict__ hi, uint64_t *__restrict__ lo,
                            int *__restrict__ idx) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double p[3], c[3] = {root[0], root[1], root[2]};
  src.get(i, p[0], p[1], p[2]);
  double size = root[3];
  uint64_t kh = 0, kl = 0;
  for (int l = 1; l <= LEVELS_HI; l++) {
    double quarter = __dmul_rn(0.5, __dmul_rn(0.5, size));  // 0.5 * halfsize (:406)
    kh = (kh << 3) | descend(p, c, quarter);
    size = __dmul_rn(0.5, size);
  }
  if (levels > LEVELS_HI) {
    for (int l = LEVELS_HI + 1; l <= LEVELS_MAX; l++) {
      double quarter = __dmul_rn(0.5, __dmul_rn(0.5, size));
      kl = (kl << 3) | descend(p, c, quarter);
      size = __dmul_rn(0.5, size);
    }
  }
  hi[i] = kh;
  if (lo) lo[i] = kl;
  idx[i] = (int)i;
}

__global__ void gather_u64(const uint64_t *__restrict__ in, const int *__restrict__ idx, int64_t n,
                           uint64_t *__restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[idx[i]];
}

// ---- K6a common levels --------------------------------------------------------------------------
__device__ __forceinline__ int common_levels(uint64_t h0, uint64_t l0, uint64_t h1, uint64_t l1,
                                             int levels) {
  uint64_t x = h0 ^ h1;
  if (x) return __clzll((long long)(x << 1)) / 3;
  if (levels <= LEVELS_HI) return LEVELS_HI;
  x = l0 ^ l1;
  if (x) return LEVELS_HI + __clzll((long long)(x << 1)) / 3;
  return LEVELS_MAX;
}

// cnt[p] = (cells opened at sorted position p) + 1 leaf;  clev[p] = c[p] (c[n-1] = -1)
__global__ void levels_kernel(const uint64_t *__restrict__ hi, const uint64_t *__restrict__ lo,
                              int64_t n, int levels, signed char *__restrict__ clev,
                              int *__restrict__ cnt) {
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  int cprev = -1, c = -1;
  if (p > 0) cprev = common_levels(hi[p - 1], lo ? lo[p - 1] : 0, hi[p], lo ? lo[p] : 0, levels);
  if (p + 1 < n) c = common_levels(hi[p], lo ? lo[p] : 0, hi[p + 1], lo ? lo[p + 1] : 0, levels);
  clev[p] = (signed char)c;
  int open = c - cprev;
  cnt[p] = (open > 0 ? open : 0) + 1;
}

// ---- K7 double-double moments -------------------------------------------------------------------
// Inclusive scans of m, m x, m y, m z over the Morton-sorted particles, in double-double
// arithmetic (hi + lo, ~106 bits), so that the moments of a cell covering sorted particles
// [p, b] are P[b+1] - P[p] without cancellation (errors ~1e-30 of the total).  The products m x
// are formed exactly: hi = fl(m x), lo = fma(m, x, -hi).
// sources gathered once into Morton order: (x, y, z, m) as double4, so that the moment scans,
// the emit kernel and the walk's target loads are all coalesced
template <class Src>
__global__ void gather_sorted_kernel(Src src, const int *__restrict__ idx, int64_t n,
                                     double4 *__restrict__ out) {
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const int64_t j = idx[p];
  double x, y, z;
  src.get(j, x, y, z);
  out[p] = make_double4(x, y, z, src.m(j));
}
// The scan itself is chunked_scan<DD4> (sortscan.cuh): deterministic, coalesced, warp-contiguous
// (fixed summation order -> bitwise reproducible run to run, unlike a decoupled-look-back scan
// with a non-associative operator).
__device__ __forceinline__ DD4 dd4_of(const double4 q) {
  DD4 r;
  r.c[0].h = q.w;
  r.c[0].l = 0.0;
  const double x[3] = {q.x, q.y, q.z};
#pragma unroll
  for (int k = 0; k < 3; k++) {
    r.c[1 + k].h = __dmul_rn(q.w, x[k]);
    r.c[1 + k].l = fma(q.w, x[k], -r.c[1 + k].h);
  }
  return r;
}
struct InParticles {  // element q = (m, m x, m y, m z) of sorted particle q, products exact
  const double4 *sp;
  __device__ __forceinline__ DD4 operator()(int64_t q) const { return dd4_of(sp[q]); }
};
// fp32 tree: plain double moments of (x - root centre).  A cell's moments are P[b+1] - P[p]; the
// rounding error of a prefix is ~1e-16 of the running total, so the centre of mass of even a
// two-particle cell is off by < 1e-16 N |x| m / m_cell ~ 1e-9 kpc at N = 10M -- two orders of
// magnitude below the fp32 resolution (6e-8 |x|) the entry is stored with.  Half the scan traffic
// of the double-double form and none of its error-free transformations (fp64 keeps DD4: there the
// moments must reproduce the reference's to 1e-12).
struct InParticlesRel {
  const double4 *sp;
  const double *root;
  __device__ __forceinline__ D4 operator()(int64_t q) const {
    const double4 t = sp[q];
    D4 r;
    r.c[0] = t.w;
    r.c[1] = t.w * (t.x - root[0]);
    r.c[2] = t.w * (t.y - root[1]);
    r.c[3] = t.w * (t.z - root[2]);
    return r;
  }
};
// moments of sorted particles [p, b]: mass and first moments (fp64: absolute coordinates,
// double-double difference; fp32: relative to the root centre, plain difference)
__device__ __forceinline__ void moment_diff(const DD4 *__restrict__ P, int64_t p, int64_t b, double mh[4]) {
  const DD4 pe = P[b + 1], ps = P[p];
  double rl;
#pragma unroll
  for (int k = 0; k < 4; k++) dd_add(pe.c[k].h, pe.c[k].l, -ps.c[k].h, -ps.c[k].l, mh[k], rl);
}
__device__ __forceinline__ void moment_diff(const D4 *__restrict__ P, int64_t p, int64_t b, double mh[4]) {
  const D4 pe = P[b + 1], ps = P[p];
#pragma unroll
  for (int k = 0; k < 4; k++) mh[k] = pe.c[k] - ps.c[k];
}
template <class Real> struct MomentOf { using type = DD4; };
template <> struct MomentOf<float> { using type = D4; };

// ---- K6b emit -----------------------------------------------------------------------------------
template <class Real>
struct Entries {
  Node<Real> *node;
  int *skip;  // pre-order index after this entry's subtree
};

__device__ __forceinline__ bool same_prefix(uint64_t h, uint64_t l, uint64_t h0, uint64_t l0,
                                            int level) {
  if (level <= LEVELS_HI) {
    int sh = 3 * (LEVELS_HI - level);
    return sh >= 64 ? true : ((h >> sh) == (h0 >> sh));  // level 0: sh = 63
  }
  if (h != h0) return false;
  int sh = 3 * (LEVELS_MAX - level);
  return (l >> sh) == (l0 >> sh);
}

// every third bit of a 63-bit Morton key, compacted (the 21-bit index along one axis)
__device__ __forceinline__ uint64_t compact3(uint64_t x) {
  x &= 0x1249249249249249ull;
  x = (x ^ (x >> 2)) & 0x10c30c30c30c30c3ull;
  x = (x ^ (x >> 4)) & 0x100f00f00f00f00full;
  x = (x ^ (x >> 8)) & 0x001f0000ff0000ffull;
  x = (x ^ (x >> 16)) & 0x001f00000000ffffull;
  x = (x ^ (x >> 32)) & 0x00000000001fffffull;
  return x;
}

// GH_EMIT_MINBLOCKS: resident 128-thread CTAs per SM the register allocation is capped for
// (scripts/build_variants.py; ncu: 66 registers -> 33 % of the warp slots active, latency bound)
#ifdef GH_EMIT_MINBLOCKS
#define GH_EMIT_BOUNDS __launch_bounds__(128, GH_EMIT_MINBLOCKS)
#else
#define GH_EMIT_BOUNDS
#endif
template <class Src, class Real>
__global__ void GH_EMIT_BOUNDS emit_kernel(const double4 *__restrict__ sp, const uint64_t *__restrict__ hi,
                            const uint64_t *__restrict__ lo, const signed char *__restrict__ clev,
                            const int *__restrict__ base /* n+1, exclusive scan of cnt */,
                            const typename MomentOf<Real>::type *__restrict__ P, int64_t n,
                            const double *__restrict__ root, bool rel_origin, double inv_theta2,
                            Entries<Real> E, int *__restrict__ maxlevel) {
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  using V4 = typename Vec4<Real>::type;
  const double ox = rel_origin ? root[0] : 0.0, oy = rel_origin ? root[1] : 0.0,
               oz = rel_origin ? root[2] : 0.0;
  const int c = clev[p];
  const int cprev = (p > 0) ? clev[p - 1] : -1;
  const double4 self = sp[p];
  const double x[3] = {self.x, self.y, self.z};
  int e = base[p];
  if (c > cprev) {
    const uint64_t h0 = hi[p], l0 = lo ? lo[p] : 0;
    double cc[3] = {root[0], root[1], root[2]};
    double size = root[3];
    int deepest = 0;
    int level0 = 0;
    if (sizeof(Real) == 4 && cprev >= 0) {
      // fp32 mode does not need the reference's bit-exact centre chain: jump straight to the
      // first level this particle opens with the closed form
      //   centre_L = root - side/2 + (i_L + 1/2) side / 2^L,  i_L = top L bits of the axis index
      level0 = cprev + 1;
      const double sL = ldexp(size, -level0);
#pragma unroll
      for (int k = 0; k < 3; k++) {
        const uint64_t ik = compact3(h0 >> k) >> (LEVELS_HI - level0);
        cc[k] = root[k] - 0.5 * size + ((double)ik + 0.5) * sL;
      }
      size = sL;
    }
    for (int level = level0; level <= c; level++) {
      if (level > cprev) {
        // this cell (level, centre cc, side size) starts at p.  Galloping + binary search for the
        // last sorted particle b sharing `level` octant levels with p (p+1 does, since c >= level).
        int64_t lo_i = p + 1, step = 1, hi_i;
        // most cells hold a handful of particles: look at the next few common-level bytes first
        // (sequential, cached) -- the cell ends at the first q > p with clev[q] < level
        bool found = false;
        for (int t = 0; t < 12 && lo_i < n; t++) {
          if (clev[lo_i] < level) { found = true; break; }
          lo_i++;
        }
        if (lo_i >= n) { lo_i = n - 1; found = true; }
        hi_i = lo_i;
        if (!found) for (;;) {
          int64_t q = lo_i + step;
          if (q >= n) { hi_i = n - 1; break; }
          if (same_prefix(hi[q], lo ? lo[q] : 0, h0, l0, level)) { lo_i = q; step <<= 1; }
          else { hi_i = q - 1; break; }
        }
        while (!found && lo_i < hi_i) {
          int64_t mid = (lo_i + hi_i + 1) >> 1;
          if (same_prefix(hi[mid], lo ? lo[mid] : 0, h0, l0, level)) lo_i = mid;
          else hi_i = mid - 1;
        }
        const int64_t b = lo_i;
        double mh[4];
        moment_diff(P, p, b, mh);
        V4 com, cen;
        if (sizeof(Real) == 4) {  // moments already relative to the root centre
          com.x = (Real)(mh[1] / mh[0]);
          com.y = (Real)(mh[2] / mh[0]);
          com.z = (Real)(mh[3] / mh[0]);
        } else {
          com.x = (Real)(mh[1] / mh[0] - ox);  // gravoct_finalize :477-479
          com.y = (Real)(mh[2] / mh[0] - oy);
          com.z = (Real)(mh[3] / mh[0] - oz);
        }
        com.w = (Real)mh[0];
        cen.x = (Real)(cc[0] - ox);
        cen.y = (Real)(cc[1] - oy);
        cen.z = (Real)(cc[2] - oz);
        // (size / dist) < theta  <=>  size^2 / theta^2 < dist^2   (theta = 0: inf, never accepted)
        if (sizeof(Real) == 4) {
          cen.w = (Real)__int_as_float((int)(((unsigned)level << SKIP_BITS) | (unsigned)base[b + 1]));
        } else {
          cen.w = (Real)(__dmul_rn(__dmul_rn(size, size), inv_theta2));
          E.skip[e] = base[b + 1];
        }
        pack_node(E.node[e], cen, com);
        e++;
        deepest = level;
      }
      if (level < c) {  // descend one level along p's key (:406,:441-462)
        const int l = level + 1;
        unsigned d;
        if (l <= LEVELS_HI) d = (unsigned)((h0 >> (3 * (LEVELS_HI - l))) & 7u);
        else d = (unsigned)((l0 >> (3 * (LEVELS_MAX - l))) & 7u);
        double quarter = __dmul_rn(0.5, __dmul_rn(0.5, size));
#pragma unroll
        for (int k = 0; k < 3; k++)
          cc[k] = __dadd_rn(cc[k], ((d >> k) & 1u) ? quarter : -quarter);
        size = __dmul_rn(0.5, size);
      }
    }
    atomicMax(maxlevel, deepest);
  }
  // the particle's own leaf: COM = particle position (:473-475), always accepted (:502)
  V4 com, cen;
  com.x = (Real)(x[0] - ox);
  com.y = (Real)(x[1] - oy);
  com.z = (Real)(x[2] - oz);
  com.w = (Real)self.w;
  cen.x = cen.y = cen.z = (Real)0;
  if (sizeof(Real) == 4) {
    cen.w = (Real)__int_as_float((int)(((unsigned)LEAF_LEVEL << SKIP_BITS) | (unsigned)(e + 1)));
  } else {
    cen.w = (Real)-1;
    E.skip[e] = e + 1;
  }
  pack_node(E.node[e], cen, com);
}


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
