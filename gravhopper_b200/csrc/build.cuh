// build.cuh -- device code of the tree build (K3-K7, K6b): bounding box, octant keys, common
// levels, moment scans' element functors, entry emit.  Included by tree.cu (which keeps the
// workspace and the launch sequence); a header of its own so that tests/emu can compile the same
// source for the host and run it thread by thread (GH_HOST_EMU), like walk.cuh.
#pragma once
#include "common.cuh"
#include "sortscan.cuh"
#include "walk.cuh"

#include <climits>

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


}  // namespace gh
