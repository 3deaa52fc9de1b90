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
template <class Real> struct Vec4;
template <> struct Vec4<double> { using type = double4; };
template <> struct Vec4<float> { using type = float4; };

template <class Real>
struct alignas(sizeof(Real) * 8) Node {
  // centre and centre of mass interleaved component by component, so that the fp32 walk forms
  // (centre - x, COM - x) with ONE packed FADD2 per axis and (|.|^2, |.|^2 + eps^2) with packed
  // FFMA2s:  a = (cx, mx, cy, my),  b = (cz, mz, s2, m)
  // c* = cell centre - origin, m* = centre of mass - origin, m = mass.
  // fp64: s2 = side^2 / theta^2 (-1 marks a leaf); the skip link lives in a separate int array.
  // fp32: the s2 slot holds the bits of (level << 27 | skip) instead (level 31 marks a leaf), so
  //       one 32-byte load (LDG.256) brings everything the walk needs about an entry;
  //       s2 = side_root^2 / theta^2 * 4^-level is rebuilt with one multiply, bit-identically.
  typename Vec4<Real>::type a;
  typename Vec4<Real>::type b;
};
static constexpr int SKIP_BITS = 27;
static constexpr int LEAF_LEVEL = 31;
template <class Real, class V4>
__device__ __forceinline__ void pack_node(Node<Real> &nd, const V4 &cen, const V4 &com) {
  nd.a.x = cen.x; nd.a.y = com.x; nd.a.z = cen.y; nd.a.w = com.y;
  nd.b.x = cen.z; nd.b.y = com.z; nd.b.z = cen.w; nd.b.w = com.w;
}
// entry i of the fp32 array: both halves with one 256-bit load
__device__ __forceinline__ void load_node32(const Node<float> *__restrict__ nodes, int i, float4 &a,
                                            float4 &b) {
  asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
      : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
      : "l"(nodes + i));
}
// (level, skip) -> s2 and skip;  pow4[l] would be a table, the exponent arithmetic is cheaper
__device__ __forceinline__ void unpack32(float packed, float s2root, float &s2, int &sk) {
  const unsigned u = (unsigned)__float_as_int(packed);
  const unsigned level = u >> SKIP_BITS;
  sk = (int)(u & ((1u << SKIP_BITS) - 1u));
  const float scale = __int_as_float((int)((127u - 2u * level) << 23));  // 4^-level
  s2 = (level == (unsigned)LEAF_LEVEL) ? -1.f : s2root * scale;
}
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

// ---- K8 walk ------------------------------------------------------------------------------------
__device__ __forceinline__ double rsqrt64_t(double s) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(s));
  double t = s * y;
  double e = fma(-t, y, 1.0);
  double p = fma(0.375, e, 0.5);
  double q = e * p;
  return fma(y, q, y);
}
template <bool GUARD>
__device__ __forceinline__ double inv_cube(double s) {
  double y = rsqrt64_t(s);
  if (GUARD) y = (s > 0.0) ? y : 0.0;  // _jbgrav.c:517-518 (only reachable when eps == 0)
  return y * y * y;
}
template <bool GUARD>
__device__ __forceinline__ float inv_cube(float s) {
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(s));
  if (GUARD) y = (s > 0.f) ? y : 0.f;
  return y * y * y;
}

struct TargetsView {
  const double4 *sorted;  // non-null: target p IS sorted source p (self evaluation of all sources)
  const double *pos64;   // (ni,3) or null
  const float4 *pos32;   // (ni) or null  (already relative to the f32 engine origin)
  const int *order;      // sorted position -> local target index (null = identity)
  int64_t order_offset;  // subtracted from order[] values (self case with a slice)
};

// One warp per 32 Morton-consecutive targets.  `i` (warp-uniform) runs through the pre-order
// entry array.  Per lane: `until` = pre-order index up to which this lane is covered by a cell it
// already accepted.  A lane is active at entry i iff i >= until; an active lane accepts iff
// s2 < |centre - x|^2 (leaves carry s2 = -1), adds the monopole and sets until = skip[i];
// otherwise it must open the cell.  The warp advances to min over lanes of (open ? i+1 : until)
// with one REDUX: the scan only touches entries some lane still needs.  Branch-free body.
// fp32: plain fp32 accumulation (<= ~1e3 accepted terms per target; error ~1e-6, far below the
// monopole error).
// Loads target p of the warp's 32 (relative to the fp32 origin when rel_origin).
template <class Real>
__device__ __forceinline__ void load_target(const TargetsView &tv, int64_t p, bool valid,
                                            const double *__restrict__ root, bool rel_origin,
                                            int64_t &ti, Real &x, Real &y, Real &z) {
  ti = 0;
  x = y = z = 0;
  if (!valid) return;
  ti = tv.order ? (int64_t)tv.order[p] - tv.order_offset : p;
  const double ox = rel_origin ? root[0] : 0.0, oy = rel_origin ? root[1] : 0.0,
               oz = rel_origin ? root[2] : 0.0;
  if (tv.sorted) {
    const double4 q = tv.sorted[p];
    x = (Real)(q.x - ox);
    y = (Real)(q.y - oy);
    z = (Real)(q.z - oz);
  } else if (tv.pos64) {
    x = (Real)(tv.pos64[3 * ti] - ox);
    y = (Real)(tv.pos64[3 * ti + 1] - oy);
    z = (Real)(tv.pos64[3 * ti + 2] - oz);
  } else {
    float4 t = tv.pos32[ti];
    x = (Real)((double)t.x - ox);
    y = (Real)((double)t.y - oy);
    z = (Real)((double)t.z - oz);
  }
}

// The per-target scan of one warp (see above): every lane applies the reference's own opening
// test, so the accepted node set is the reference's.
template <class Real, bool STATS, bool GUARD, bool PREFETCH>
__device__ __forceinline__ void lane_scan(const Node<Real> *__restrict__ nodes,
                                          const int *__restrict__ skips, Real s2root, int nentries,
                                          bool valid, Real x, Real y, Real z, Real eps2, Real &ax, Real &ay,
                                          Real &az, unsigned long long &nacc,
                                          unsigned long long &nvis, unsigned long long &niter) {
  int until = valid ? 0 : INT_MAX;
  int i = 0;
  while (i < nentries) {
    if (STATS) niter++;
    const auto na = nodes[i].a;
    auto nb = nodes[i].b;
    int sk;
    if (sizeof(Real) == 4) {
      float s2;
      unpack32((float)nb.z, (float)s2root, s2, sk);
      nb.z = (Real)s2;
    } else {
      sk = skips[i];
    }
    if (PREFETCH) {
      // optional L2 prefetch hint of the entry after this subtree.  Measured on B200 (N = 4M):
      // +19 % time when all 131k warps run (issue bound), -6 % with 16k warps; a register
      // double-buffer prefetch of entry i+1 was 85 % slower.  Off by default.
      asm volatile("prefetch.global.L2 [%0];" ::"l"(nodes + (sk < nentries ? sk : i)));
    }
    const bool active = i >= until;
    bool pass;
    Real ex, ey, ez, s;
    if (sizeof(Real) == 4) {
      // packed: lane .x of every pair is the opening test (centre), lane .y the force (COM)
      const float2 nx2 = make_float2(-(float)x, -(float)x), ny2 = make_float2(-(float)y, -(float)y),
                   nz2 = make_float2(-(float)z, -(float)z);
      const float2 dx = __fadd2_rn(make_float2((float)na.x, (float)na.y), nx2);
      const float2 dy = __fadd2_rn(make_float2((float)na.z, (float)na.w), ny2);
      const float2 dz = __fadd2_rn(make_float2((float)nb.x, (float)nb.y), nz2);
      float2 q = __ffma2_rn(dx, dx, make_float2(0.f, (float)eps2));
      q = __ffma2_rn(dy, dy, q);
      q = __ffma2_rn(dz, dz, q);
      pass = (float)nb.z < q.x;
      ex = (Real)dx.y; ey = (Real)dy.y; ez = (Real)dz.y; s = (Real)q.y;
    } else {
      const Real dx = na.x - x, dy = na.z - y, dz = nb.x - z;
      const Real d2 = dx * dx + dy * dy + dz * dz;
      pass = nb.z < d2;
      ex = na.y - x; ey = na.w - y; ez = nb.y - z;
      s = ex * ex + ey * ey + ez * ez + eps2;
    }
    const bool acc = active && pass;
    const bool open = active && !pass;
    const Real w = acc ? nb.w * inv_cube<GUARD>(s) : (Real)0;
    ax += w * ex;
    ay += w * ey;
    az += w * ez;
    until = acc ? sk : until;
    if (STATS) { nvis += active; nacc += acc; }
    const int next = open ? i + 1 : until;
    i = __reduce_min_sync(0xffffffffu, next);
  }
}

template <class Real, bool STATS, bool GUARD, bool PREFETCH>
__global__ void __launch_bounds__(128, 8)
walk_kernel(const Node<Real> *__restrict__ nodes, const int *__restrict__ skips, int nentries,
            TargetsView tv, int64_t ni, const double *__restrict__ root, bool rel_origin, Real eps2,
            double inv_theta2, Epilogue ep, unsigned long long *__restrict__ stats) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t p = warp * 32 + lane;
  const bool valid = p < ni;
  int64_t ti;
  Real x, y, z;
  load_target<Real>(tv, p, valid, root, rel_origin, ti, x, y, z);
  Real ax = 0, ay = 0, az = 0;
  unsigned long long nacc = 0, nvis = 0, niter = 0;
  const Real s2root = (Real)(root[3] * root[3] * inv_theta2);
  lane_scan<Real, STATS, GUARD, PREFETCH>(nodes, skips, s2root, nentries, valid, x, y, z, eps2, ax, ay,
                                          az, nacc, nvis, niter);
  if (valid) apply_epilogue(ep, ti, (double)ax, (double)ay, (double)az);
  if (STATS) {
    for (int o = 16; o > 0; o >>= 1) {
      nacc += __shfl_down_sync(0xffffffffu, nacc, o);
      nvis += __shfl_down_sync(0xffffffffu, nvis, o);
    }
    if (lane == 0) {
      atomicAdd(&stats[0], nacc);
      atomicAdd(&stats[1], nvis);
      atomicAdd(&stats[2], niter);  // entries this warp stepped through (union over its lanes)
      atomicMax(&stats[3], niter);
    }
  }
}

// ---- K8g group walk (fp32) --------------------------------------------------------------------
// The warp walks the tree ONCE for its 32 Morton-consecutive targets instead of once per lane.
// Traversal and force evaluation are separated:
//   traversal   a shared-memory stack holds sibling chains (first, end) of the pre-order array.
//               Each iteration pops up to 32 chains; lane l loads chain l's first entry, tests it
//               against the bounding box of the 32 targets and pushes (a) the rest of the chain
//               (skip[first], end) and (b), if the cell must be opened, the chain of its children
//               (first+1, skip[first]).  The 32 entry loads of an iteration are independent, so
//               the dependent-load chain of the per-target scan (one entry at a time per warp)
//               becomes ~80 iterations of 32 parallel loads.
//   criterion   a cell is accepted for the group only if EVERY point of the targets' bounding box
//               passes the reference's test, s^2/theta^2 < min_{x in box} |centre - x|^2.  That is
//               the reference's criterion made conservative: each target's accepted set is a
//               refinement of the set the reference would accept for it (some cells the reference
//               accepts are opened further), so the force error is never larger in the sense of
//               the opening angle, at the price of a longer list (measured/modelled:
//               scripts/walk_sim.c, median 2.0x the per-target count at N = 4M).
//   evaluation  accepted entries (COM, mass) go to a shared-memory ring in the paired layout of
//               the direct kernel; every 32 entries all lanes evaluate them for their own target
//               with packed FADD2/FFMA2/FMUL2: 14 FP32-pipe instructions + 2 MUFU + 2 LDS.128 per
//               PAIR of interactions (the per-target scan spends 28 instructions per entry on
//               test + force + control).
// Groups whose list grows beyond `list_limit` entries (bounding boxes that straddle a jump of the
// Morton curve; ~5 % of the groups) or whose chain stack would overflow drop what they have and
// run the per-target scan instead, which bounds the cost of any group.
static constexpr int GROUP_STACK = 320;  // chains per warp
static constexpr int GROUP_RING = 64;    // list entries per warp (two chunks of 32)

// HYBRID: also accumulate sabs += m / (|d|^2 + eps^2) over the FIRST pair of the chunk (2 of 32
// entries): a 1/16 sample of the summed magnitude of the contributions, see walk_group_kernel.
template <bool GUARD, bool HYBRID = false>
__device__ __forceinline__ void eval_chunk(const float4 *__restrict__ pairs, float2 nx2, float2 ny2,
                                           float2 nz2, float2 e2, float2 &fx, float2 &fy, float2 &fz,
                                           float2 *sabs = nullptr) {
#pragma unroll
  for (int q = 0; q < 16; q++) {
    const float4 A = pairs[q], B = pairs[16 + q];  // (x0,x1,y0,y1), (z0,z1,m0,m1)
    const float2 dx = __fadd2_rn(make_float2(A.x, A.y), nx2);
    const float2 dy = __fadd2_rn(make_float2(A.z, A.w), ny2);
    const float2 dz = __fadd2_rn(make_float2(B.x, B.y), nz2);
    float2 s = __ffma2_rn(dx, dx, e2);
    s = __ffma2_rn(dy, dy, s);
    s = __ffma2_rn(dz, dz, s);
    float2 r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r.x) : "f"(s.x));
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r.y) : "f"(s.y));
    if (GUARD) { r.x = (s.x > 0.f) ? r.x : 0.f; r.y = (s.y > 0.f) ? r.y : 0.f; }
    const float2 r2 = __fmul2_rn(r, r);
    if (HYBRID && q == 0) *sabs = __ffma2_rn(r2, make_float2(B.z, B.w), *sabs);
    float2 w = __fmul2_rn(r2, r);
    w = __fmul2_rn(w, make_float2(B.z, B.w));
    fx = __ffma2_rn(w, dx, fx);
    fy = __ffma2_rn(w, dy, fy);
    fz = __ffma2_rn(w, dz, fz);
  }
}

// float <-> int with the same ordering (involution), so REDUX.MIN/MAX can reduce floats
__device__ __forceinline__ int f2ord(float f) {
  const int i = __float_as_int(f);
  return i ^ ((i >> 31) & 0x7fffffff);
}
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i ^ ((i >> 31) & 0x7fffffff)); }
__device__ __forceinline__ float warp_min(float v) { return ord2f(__reduce_min_sync(0xffffffffu, f2ord(v))); }
__device__ __forceinline__ float warp_max(float v) { return ord2f(__reduce_max_sync(0xffffffffu, f2ord(v))); }

// centre and padded half extent of the bounding box of the lanes with in == true (the padding of a
// few ulps makes rounding err on the conservative side)
struct Box { float cx, cy, cz, hx, hy, hz; };
__device__ __forceinline__ Box warp_box(bool in, float x, float y, float z) {
  const float inf = __int_as_float(0x7f800000);
  const float lx = warp_min(in ? x : inf), ux = warp_max(in ? x : -inf);
  const float ly = warp_min(in ? y : inf), uy = warp_max(in ? y : -inf);
  const float lz = warp_min(in ? z : inf), uz = warp_max(in ? z : -inf);
  Box b;
  b.cx = 0.5f * (lx + ux); b.cy = 0.5f * (ly + uy); b.cz = 0.5f * (lz + uz);
  const float pad = 1.0f + 1e-6f;
  b.hx = (0.5f * (ux - lx)) * pad + 1e-6f * fabsf(b.cx);
  b.hy = (0.5f * (uy - ly)) * pad + 1e-6f * fabsf(b.cy);
  b.hz = (0.5f * (uz - lz)) * pad + 1e-6f * fabsf(b.cz);
  return b;
}
// squared distance from point c to the box (0 inside)
__device__ __forceinline__ float box_dist2(const Box &b, float cx, float cy, float cz) {
  const float dx = fmaxf(fabsf(cx - b.cx) - b.hx, 0.f);
  const float dy = fmaxf(fabsf(cy - b.cy) - b.hy, 0.f);
  const float dz = fmaxf(fabsf(cz - b.cz) - b.hz, 0.f);
  return dx * dx + dy * dy + dz * dz;
}

// Resident warps per SM the register allocation is capped for: 32 -> 64 registers per thread,
// 24 -> 80, 20 -> 96, 16 -> 128 (scripts/gpu_variants.sh measures the alternatives).
#ifndef GH_GW_WARPS_PER_SM
#define GH_GW_WARPS_PER_SM 32
#endif
// HYBRID (off by default; GH_WALK_HYBRID=<kappa> turns it on): the 32 targets of a group share one
// list, so their truncation errors are one coherent vector; where a target's net force nearly
// cancels (|a| << sum of |contributions|: the softened core of a cusp) that vector does not
// average out the way the per-target walk's errors do, and the relative error of ~0.01 % of the
// particles exceeds the reference tree's.  With HYBRID each lane compares |a| with a 1/16 sample
// of sum m/(d^2+eps^2) over its list; lanes with |a| < kappa * sum repeat the evaluation with the
// reference's own per-target criterion (lane_scan).  CPU model of this rule (test infrastructure):
// kappa = 0.1 flags 0.1 % of the particles (0.2-0.3 % of the groups) at N = 200k...4M and brings
// p99.99 and max of the error distribution back to the reference tree's.
__constant__ float c_hybrid_kappa2;
template <int WPC, bool STATS, bool GUARD, bool HYBRID = false>
__global__ void __launch_bounds__(32 * WPC, GH_GW_WARPS_PER_SM / WPC)
walk_group_kernel(const Node<float> *__restrict__ nodes, int nentries, TargetsView tv, int64_t ni,
                  const double *__restrict__ root, float eps2, double inv_theta2, int list_limit,
                  Epilogue ep, unsigned long long *__restrict__ stats) {
  __shared__ int2 s_stack[WPC][GROUP_STACK];
  __shared__ float4 s_ring[WPC][GROUP_RING];
  const int lane = threadIdx.x & 31;
  const int wic = threadIdx.x >> 5;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t p = warp * 32 + lane;
  const bool valid = p < ni;
  int64_t ti;
  float x, y, z;
  load_target<float>(tv, p, valid, root, true, ti, x, y, z);

  // Two bounding boxes: the 32 targets are cut where Morton-consecutive targets are farthest
  // apart, so a group that straddles a jump of the curve is two compact boxes instead of one
  // huge one (scripts/walk_sim.c: list p99 4770 -> 1507 entries at N = 4M, mean 1434 -> 1133).
  const float xn = __shfl_down_sync(0xffffffffu, x, 1), yn = __shfl_down_sync(0xffffffffu, y, 1),
              zn = __shfl_down_sync(0xffffffffu, z, 1);
  const bool next_valid = lane < 31 && (p + 1 < ni);
  const float gap = next_valid ? (xn - x) * (xn - x) + (yn - y) * (yn - y) + (zn - z) * (zn - z) : -1.f;
  const int gmax = __reduce_max_sync(0xffffffffu, f2ord(gap));
  const int cut = __ffs(__ballot_sync(0xffffffffu, f2ord(gap) == gmax)) - 1;  // box A = lanes <= cut
  const Box A = warp_box(valid && lane <= cut, x, y, z);
  Box B = warp_box(valid && lane > cut, x, y, z);
  if (!(B.hx >= 0.f)) B = A;  // no valid lane beyond the cut

#define stack s_stack[wic]
#define ring4 s_ring[wic]
  const unsigned lt = (1u << lane) - 1u;
  const unsigned gt = ~lt & ~(1u << lane);
  const float2 nx2 = make_float2(-x, -x), ny2 = make_float2(-y, -y), nz2 = make_float2(-z, -z);
  const float2 e2 = make_float2(eps2, eps2);
  float2 fx = make_float2(0.f, 0.f), fy = fx, fz = fx;
  float2 sabs = make_float2(0.f, 0.f);
  unsigned long long nacc = 0, nvis = 0, niter = 0, nredo = 0;
  const float s2root = (float)(root[3] * root[3] * inv_theta2);

  if (lane == 0) stack[0] = make_int2(0, nentries);
  int sp = 1;
  int head = 0, tail = 0;  // list entries pushed / evaluated
  bool fallback = false;
  __syncwarp();
  while (sp > 0) {
    if (STATS) niter++;
    const int take = sp < 32 ? sp : 32;
    const bool has = lane < take;
    int first = 0, end = 0;
    if (has) { const int2 it = stack[sp - 1 - lane]; first = it.x; end = it.y; }
    sp -= take;
    // the 32 entry loads of this iteration are issued first ...
    float4 na = make_float4(0.f, 0.f, 0.f, 0.f), nb = na;
    if (has) load_node32(nodes, first, na, nb);
    // ... and the 32 list entries the previous iterations completed are evaluated while they fly
    if (head - tail >= 32) {
      eval_chunk<GUARD, HYBRID>(ring4 + (tail & (GROUP_RING - 1)), nx2, ny2, nz2, e2, fx, fy, fz, &sabs);
      tail += 32;
      if (head > list_limit) { fallback = true; break; }
    }
    __syncwarp();
    float s2;
    int sk;
    unpack32(nb.z, s2root, s2, sk);
    const float d2 = fminf(box_dist2(A, na.x, na.z, nb.x), box_dist2(B, na.x, na.z, nb.x));
    const bool acc = has && (s2 < d2);                       // leaves: s2 = -1
    const bool open = has && !acc && (first + 1 < sk);
    const bool rem = has && (sk < end);
    // rest of each chain first, children on top (depth first); lane 0 held the top of the stack
    const unsigned mr = __ballot_sync(0xffffffffu, rem);
    const unsigned mo = __ballot_sync(0xffffffffu, open);
    if (sp + __popc(mr) + __popc(mo) > GROUP_STACK) { fallback = true; break; }
    if (rem) stack[sp + __popc(mr & gt)] = make_int2(sk, end);
    sp += __popc(mr);
    if (open) stack[sp + __popc(mo & gt)] = make_int2(first + 1, sk);
    sp += __popc(mo);
    const unsigned ma = __ballot_sync(0xffffffffu, acc);
    if (acc) {
      const int slot = (head + __popc(ma & lt)) & (GROUP_RING - 1);
      // chunk of 32 entries = 16 rows (x0,x1,y0,y1) then 16 rows (z0,z1,m0,m1): 2-way bank
      // conflicts on these stores instead of 4-way with the rows interleaved
      float *b = &s_ring[wic][(slot & 32) + ((slot & 31) >> 1)].x + (slot & 1);
      b[0] = na.y; b[2] = na.w; b[64] = nb.y; b[66] = nb.w;
    }
    head += __popc(ma);
    if (STATS) nvis += take;
    __syncwarp();
  }
  float ax, ay, az;
  if (!fallback) {
    if (head - tail >= 32) {
      eval_chunk<GUARD, HYBRID>(ring4 + (tail & (GROUP_RING - 1)), nx2, ny2, nz2, e2, fx, fy, fz, &sabs);
      tail += 32;
      __syncwarp();
    }
    if (head > tail) {  // pad the last chunk with massless entries at the box centre
      for (int k = head + lane; k < tail + 32; k += 32) {
        const int slot = k & (GROUP_RING - 1);
        float *b = &s_ring[wic][(slot & 32) + ((slot & 31) >> 1)].x + (slot & 1);
        b[0] = A.cx; b[2] = A.cy; b[64] = A.cz; b[66] = 0.f;
      }
      __syncwarp();
      eval_chunk<GUARD, HYBRID>(ring4 + (tail & (GROUP_RING - 1)), nx2, ny2, nz2, e2, fx, fy, fz, &sabs);
    }
    ax = fx.x + fx.y; ay = fy.x + fy.y; az = fz.x + fz.y;
    if (STATS) nacc = valid ? (unsigned long long)head : 0ull;
    if (STATS) nvis = valid ? nvis : 0ull;
    if (HYBRID) {
      const float S = 16.f * (sabs.x + sabs.y);
      const bool redo = valid && (ax * ax + ay * ay + az * az < c_hybrid_kappa2 * S * S);
      if (__any_sync(0xffffffffu, redo)) {
        float bx = 0.f, by = 0.f, bz = 0.f;
        unsigned long long c0 = 0, c1 = 0, c2 = 0;
        lane_scan<float, false, GUARD, false>(nodes, nullptr, s2root, nentries, redo, x, y, z, eps2, bx, by, bz,
                                              c0, c1, c2);
        if (redo) { ax = bx; ay = by; az = bz; }
        if (STATS) nredo = redo ? 1ull : 0ull;
      }
    }
  } else {
    ax = ay = az = 0.f;
    nacc = nvis = 0;
    unsigned long long it2 = 0;
    lane_scan<float, STATS, GUARD, false>(nodes, nullptr, s2root, nentries, valid, x, y, z, eps2, ax, ay,
                                          az, nacc, nvis, it2);
  }
  if (valid) apply_epilogue(ep, ti, (double)ax, (double)ay, (double)az);
  if (STATS) {
    for (int o = 16; o > 0; o >>= 1) {
      nacc += __shfl_down_sync(0xffffffffu, nacc, o);
      nvis += __shfl_down_sync(0xffffffffu, nvis, o);
      if (HYBRID) nredo += __shfl_down_sync(0xffffffffu, nredo, o);
    }
    if (lane == 0) {
      atomicAdd(&stats[0], nacc);
      atomicAdd(&stats[1], nvis);
      atomicAdd(&stats[2], niter);                 // traversal iterations (32 entries each)
      // low 32 bits: groups that fell back to the per-target scan; high 32 bits: targets the
      // hybrid rule re-evaluated with the per-target criterion
      atomicAdd(&stats[3], (fallback ? 1ull : 0ull) + (HYBRID ? (nredo << 32) : 0ull));
    }
  }
}
#undef stack
#undef ring4

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
