// walk.cuh -- device code of the tree walks (K8): entry format, per-target scan (walk_kernel),
// group walk (walk_group_kernel).  Included by tree.cu; kept in its own header so that the same
// source can also be compiled for the host with lockstep-warp shims (tests/emu/walk_emu.cpp,
// GH_HOST_EMU): the CPU test suite then executes the real kernel code, not a restatement of it.
// Under GH_HOST_EMU the three inline-PTX spots (256-bit entry load, rsqrt.approx) use plain C.
#pragma once
#include "common.cuh"

#include <climits>

namespace gh {

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
#ifdef GH_HOST_EMU
  a = nodes[i].a;
  b = nodes[i].b;
#else
  asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
      : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
      : "l"(nodes + i));
#endif
}
// (level, skip) -> s2 and skip;  pow4[l] would be a table, the exponent arithmetic is cheaper
__device__ __forceinline__ void unpack32(float packed, float s2root, float &s2, int &sk) {
  const unsigned u = (unsigned)__float_as_int(packed);
  const unsigned level = u >> SKIP_BITS;
  sk = (int)(u & ((1u << SKIP_BITS) - 1u));
  const float scale = __int_as_float((int)((127u - 2u * level) << 23));  // 4^-level
  s2 = (level == (unsigned)LEAF_LEVEL) ? -1.f : s2root * scale;
}
// ---- K8 walk ------------------------------------------------------------------------------------
__device__ __forceinline__ double rsqrt64_t(double s) {
  double y;
#ifdef GH_HOST_EMU
  y = (double)(float)(1.0 / sqrt(s));  // a ~24-bit seed, like MUFU.RSQ64H's 2^-20
#else
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(s));
#endif
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
#ifdef GH_HOST_EMU
  y = 1.0f / sqrtf(s);
#else
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(s));
#endif
  if (GUARD) y = (s > 0.f) ? y : 0.f;
  return y * y * y;
}

// y = s^-1/2 (the quadrupole term needs y^5 and y^7 besides the monopole's y^3)
template <bool GUARD>
__device__ __forceinline__ double inv_sqrt(double s) {
  double y = rsqrt64_t(s);
  if (GUARD) y = (s > 0.0) ? y : 0.0;
  return y;
}
template <bool GUARD>
__device__ __forceinline__ float inv_sqrt(float s) {
  float y;
#ifdef GH_HOST_EMU
  y = 1.0f / sqrtf(s);
#else
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(s));
#endif
  if (GUARD) y = (s > 0.f) ? y : 0.f;
  return y;
}

struct TargetsView {
  const double4 *sorted;  // non-null: target p IS sorted source p (self evaluation of all sources)
  const double *pos64;   // (ni,3) or null
  const float4 *pos32;   // (ni) or null  (already relative to the f32 engine origin)
  const int *order;      // sorted position -> local target index (null = identity)
  int64_t order_offset;  // subtracted from order[] values (self case with a slice)
  // distributed walk (dist_sidx != null): the targets are this rank's share of the GLOBAL Morton
  // order.  Rank q's sorted particles sit at dist_sidx[q * dist_ncap + j], j < dist_counts[q]
  // (source indices; positions in pos32); blocks of dist_blk consecutive j are dealt round-robin:
  // block kb of rank q's range belongs to rank (kb + q) % world.  Target slot s of this rank
  // (= the index its acceleration is stored at) enumerates its blocks: q = s / (T blk),
  // t = (s / blk) % T, kb = (rank - q) mod world + world t.
  const int *dist_sidx;
  const int *dist_counts;
  int dist_rank, dist_world, dist_ncap, dist_blk, dist_T;
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
__device__ __forceinline__ void load_target(const TargetsView &tv, int64_t p, bool &valid,
                                            const double *__restrict__ root, bool rel_origin,
                                            int64_t &ti, Real &x, Real &y, Real &z) {
  ti = 0;
  x = y = z = 0;
  if (!valid) return;
  if (tv.dist_sidx) {
    const int per = tv.dist_T * tv.dist_blk;
    const int q = (int)(p / per), rem = (int)(p % per);
    const int t = rem / tv.dist_blk, o = rem % tv.dist_blk;
    const int kb = (tv.dist_rank - q + tv.dist_world) % tv.dist_world + tv.dist_world * t;
    const int64_t j = (int64_t)kb * tv.dist_blk + o;
    if (j >= tv.dist_counts[q]) { valid = false; return; }
    const float4 tq = tv.pos32[tv.dist_sidx[(int64_t)q * tv.dist_ncap + j]];
    ti = p;
    const double ox = rel_origin ? root[0] : 0.0, oy = rel_origin ? root[1] : 0.0, oz = rel_origin ? root[2] : 0.0;
    x = (Real)((double)tq.x - ox);
    y = (Real)((double)tq.y - oy);
    z = (Real)((double)tq.z - oz);
    return;
  }
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
// QUAD (opt-in): every accepted CELL also contributes its traceless quadrupole
// (quad[6 i ..], written by emit_kernel): a += -(Q e) y^5 + 5/2 (e.Q.e) y^7 e, e = COM - x.
template <class Real, bool STATS, bool GUARD, bool PREFETCH, bool QUAD = false>
__device__ __forceinline__ void lane_scan(const Node<Real> *__restrict__ nodes,
                                          const int *__restrict__ skips, Real s2root, int nentries,
                                          bool valid, Real x, Real y, Real z, Real eps2, Real &ax, Real &ay,
                                          Real &az, unsigned long long &nacc,
                                          unsigned long long &nvis, unsigned long long &niter,
                                          int first_entry = 0, const Real *__restrict__ quad = nullptr) {
  int until = valid ? 0 : INT_MAX;
  int i = first_entry;
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
#ifndef GH_HOST_EMU
      asm volatile("prefetch.global.L2 [%0];" ::"l"(nodes + (sk < nentries ? sk : i)));
#endif
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
    if (QUAD) {
      const Real yy = inv_sqrt<GUARD>(s);
      const Real y2 = yy * yy, y3 = y2 * yy;
      const Real w = acc ? nb.w * y3 : (Real)0;
      ax += w * ex;
      ay += w * ey;
      az += w * ez;
      if (nb.z >= (Real)0) {  // a cell (leaves carry s2 = -1): warp-uniform, one broadcast load
        const Real *q = quad + 6 * (size_t)i;
        const Real qxx = q[0], qyy = q[1], qzz = q[2], qxy = q[3], qxz = q[4], qyz = q[5];
        const Real qex = qxx * ex + qxy * ey + qxz * ez, qey = qxy * ex + qyy * ey + qyz * ez,
                   qez = qxz * ex + qyz * ey + qzz * ez;
        const Real eqe = ex * qex + ey * qey + ez * qez;
        const Real y5 = acc ? y3 * y2 : (Real)0, y7 = y5 * y2;
        const Real g = (Real)2.5 * eqe * y7;
        ax += g * ex - qex * y5;
        ay += g * ey - qey * y5;
        az += g * ez - qez * y5;
      }
    } else {
    const Real w = acc ? nb.w * inv_cube<GUARD>(s) : (Real)0;
    ax += w * ex;
    ay += w * ey;
    az += w * ez;
    }
    until = acc ? sk : until;
    if (STATS) { nvis += active; nacc += acc; }
    const int next = open ? i + 1 : until;
    i = __reduce_min_sync(0xffffffffu, next);
  }
}

template <class Real, bool STATS, bool GUARD, bool PREFETCH, bool QUAD = false>
__global__ void __launch_bounds__(128, 8)
walk_kernel(const Node<Real> *__restrict__ nodes, const int *__restrict__ skips, int nentries,
            TargetsView tv, int64_t ni, const double *__restrict__ root, bool rel_origin, Real eps2,
            double inv_theta2, Epilogue ep, unsigned long long *__restrict__ stats,
            const int *__restrict__ walkctl = nullptr, const Real *__restrict__ quad = nullptr) {
  // `nentries` is the index every chain ends at (the capacity of the entry array, not its fill:
  // the build never tells the host how many entries it wrote).  walkctl (BuildCtl::overflow,
  // ::first): a build whose entries did not fit sets [0]; the walk then leaves the state alone and
  // the host reports it at its next sync.  [1] is the index of the root entry (the first
  // non-empty rank's segment start in a distributed build; 0 otherwise).
  if (walkctl && walkctl[0]) return;
  const int first_entry = walkctl ? walkctl[1] : 0;
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t p = warp * 32 + lane;
  bool valid = p < ni;
  int64_t ti;
  Real x, y, z;
  load_target<Real>(tv, p, valid, root, rel_origin, ti, x, y, z);
  Real ax = 0, ay = 0, az = 0;
  unsigned long long nacc = 0, nvis = 0, niter = 0;
  const Real s2root = (Real)(root[3] * root[3] * inv_theta2);
  lane_scan<Real, STATS, GUARD, PREFETCH, QUAD>(nodes, skips, s2root, nentries, valid, x, y, z, eps2, ax, ay,
                                                az, nacc, nvis, niter, first_entry, quad);
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
#ifdef GH_HOST_EMU
    r.x = 1.0f / sqrtf(s.x);
    r.y = 1.0f / sqrtf(s.y);
#else
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r.x) : "f"(s.x));
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r.y) : "f"(s.y));
#endif
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
// HYBRID (the default, kappa = 0.10; GH_WALK_HYBRID=0 turns it off): the 32 targets of a group share one
// list, so their truncation errors are one coherent vector; where a target's net force nearly
// cancels (|a| << sum of |contributions|: the softened core of a cusp) that vector does not
// average out the way the per-target walk's errors do, and the relative error of ~0.01 % of the
// particles exceeds the reference tree's.  With HYBRID each lane compares |a| with a 1/16 sample
// of sum m/(d^2+eps^2) over its list; lanes with |a| < kappa * sum repeat the evaluation with the
// reference's own per-target criterion (lane_scan over the flagged lanes).  Measured on B200 over all particles of the N = 4M Hernquist sphere
// (profiles/r02_hybrid_sweep_N4M.json): kappa = 0.1 re-evaluates 0.10 % of the targets (p99.99 of
// the error 1.015x the reference tree's, max equal), 0.15 0.30 % (1.002x), 0.2 1.3 % (1.000x).
__constant__ float c_hybrid_kappa2;
template <int WPC, bool STATS, bool GUARD, bool HYBRID = false>
__global__ void __launch_bounds__(32 * WPC, GH_GW_WARPS_PER_SM / WPC)
walk_group_kernel(const Node<float> *__restrict__ nodes, int nentries, TargetsView tv, int64_t ni,
                  const double *__restrict__ root, float eps2, double inv_theta2, int list_limit,
                  Epilogue ep, unsigned long long *__restrict__ stats,
                  const int *__restrict__ walkctl = nullptr) {
  __shared__ int2 s_stack[WPC][GROUP_STACK];
  __shared__ float4 s_ring[WPC][GROUP_RING];
  if (walkctl && walkctl[0]) return;  // see walk_kernel
  // index of the root entry: re-read where it is needed (start, rare re-evaluations) rather than
  // kept in a register for the whole walk
#define GH_FIRST_ENTRY (walkctl ? walkctl[1] : 0)
  const int lane = threadIdx.x & 31;
  const int wic = threadIdx.x >> 5;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t p = warp * 32 + lane;
  bool valid = p < ni;
  int64_t ti;
  float x, y, z;
  load_target<float>(tv, p, valid, root, true, ti, x, y, z);
  // a warp without any target (tail of the last CTA when WPC > 1) has nothing to do; without this
  // its NaN boxes would open every cell until the list limit sends it to the per-target scan
  if (__ballot_sync(0xffffffffu, valid) == 0u) return;

  // Two bounding boxes: the 32 targets are cut where Morton-consecutive targets are farthest
  // apart, so a group that straddles a jump of the curve is two compact boxes instead of one
  // huge one (scripts/walk_sim.c: list p99 4770 -> 1507 entries at N = 4M, mean 1434 -> 1133).
  const float xn = __shfl_down_sync(0xffffffffu, x, 1), yn = __shfl_down_sync(0xffffffffu, y, 1),
              zn = __shfl_down_sync(0xffffffffu, z, 1);
  const bool next_valid = (__shfl_down_sync(0xffffffffu, valid ? 1 : 0, 1) != 0) && valid && lane < 31;
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

  if (lane == 0) stack[0] = make_int2(GH_FIRST_ENTRY, nentries);
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
        // one serial scan for all flagged lanes of the group.  (Measured alternative, rejected: the
        // whole warp on one flagged target at a time with the chain-stack traversal -- a single
        // target's frontier is narrow, ~6 busy lanes over ~120 dependent iterations, 3x the cost.)
        float bx = 0.f, by = 0.f, bz = 0.f;
        unsigned long long c0 = 0, c1 = 0, c2 = 0;
        lane_scan<float, false, GUARD, false>(nodes, nullptr, s2root, nentries, redo, x, y, z, eps2, bx, by, bz,
                                              c0, c1, c2, GH_FIRST_ENTRY);
        if (redo) { ax = bx; ay = by; az = bz; }
        if (STATS) nredo = redo ? 1ull : 0ull;
      }
    }
  } else {
    ax = ay = az = 0.f;
    nacc = nvis = 0;
    unsigned long long it2 = 0;
    lane_scan<float, STATS, GUARD, false>(nodes, nullptr, s2root, nentries, valid, x, y, z, eps2, ax, ay,
                                          az, nacc, nvis, it2, GH_FIRST_ENTRY);
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
#undef GH_FIRST_ENTRY

}  // namespace gh
