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

// ---- distributed build (fp32 tree) -----------------------------------------------------
// With P ranks the Morton key space is cut into P ranges by splitters; rank r sorts, scans and emits
// only the particles whose keys fall into [split[r], split[r+1]).  The global pre-order entry array
// is the concatenation of the ranks' segments, rank r's at [r * stride, r * stride + count_r):
// indices are "virtual" (gaps between segments are never visited: every skip link that leaves a
// segment points at the start of the next non-empty one), so no rank needs another rank's entry
// count.  A cell is emitted by the rank that holds its FIRST particle; the few cells that continue
// beyond that rank's range (at most one per level and rank boundary) take their end -- skip link
// and moment prefix -- from a small table every rank publishes about the cells that contain ITS
// first particle (RankRec2::tab).  Two small all-gathers carry the boundary keys (RankRec1) and
// the tables, totals and key samples (RankRec2); one large all-gather carries the entries.
static constexpr int DIST_SAMPLES = 64;
static constexpr int DIST_MAX_RANKS = 64;
struct RankRec1 {
  uint64_t kfirst, klast;  // smallest / largest key of the rank's range (valid when n > 0)
  int n;                   // particles in the rank's range
  int pad[3];
};
struct CellEnd {
  int bend;  // local index one past the last particle of the level-l cell containing local particle 0
  int base;  // base_local[bend]  (pre-order offset inside the rank's segment)
  D4 P;      // P_local[bend]     (moment prefix inside the rank)
};
struct RankRec2 {
  D4 Mtot;        // P_local[n]
  int nentries;   // base_local[n]
  int pad;
  CellEnd tab[LEVELS_HI + 1];
  uint64_t samples[DIST_SAMPLES];  // sorted local keys at (j + 1/2) n / 64: next step's splitters
};
// per-workspace control block on the device (also used with one rank: rank 0 of 1)
struct BuildCtl {
  int n_local;   // particles in this rank's range (== n with one rank)
  int rank, world;
  int seg;       // first pre-order index of this rank's segment (rank * stride)
  int seg_next;  // first index of the next non-empty rank's segment, or `end`
  int stride;    // entries a segment can hold
  int end;       // world * stride: the index every chain of the walk ends at
  int cprev;     // octant levels shared with the previous non-empty rank's last key (-1: none)
  int cnext;     // ... with the next non-empty rank's first key (-1: none)
  int overflow;  // set by emit_kernel when a segment would overflow; the walk then does nothing
  int first;     // pre-order index of the root entry: the first non-empty rank's segment start
                 // (overflow, first) are what the walk kernels read: `walkctl`
  int nentries;  // entries of this rank's segment (base_local[n_local])
  int maxent;    // largest segment fill over all ranks (identical on every rank: from the gathered records)
  int pad;
  D4 gP;         // moment prefix of all earlier ranks
  int xskip[LEVELS_HI + 1];   // cells of level l that start here and continue beyond this rank:
  D4 xP[LEVELS_HI + 1];       //   their skip link and the moment prefix at their end
  int counts[DIST_MAX_RANKS];               // particles per rank (from the gathered RankRec1)
  uint64_t split[DIST_MAX_RANKS + 1];       // key ranges of this step
  uint64_t split_next[DIST_MAX_RANKS + 1];  // key ranges of the next step (from the gathered samples)
};

// one rank: the whole array is one segment of `cap` entries (virtual end = cap)
__global__ void ctl_init_single(BuildCtl *ctl, int n, int cap) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  ctl->n_local = n; ctl->rank = 0; ctl->world = 1;
  ctl->seg = 0; ctl->seg_next = cap; ctl->stride = cap; ctl->end = cap;
  ctl->cprev = -1; ctl->cnext = -1; ctl->overflow = 0; ctl->first = 0; ctl->nentries = 0; ctl->maxent = 0;
  for (int k = 0; k < 4; k++) ctl->gP.c[k] = 0.0;
}

// distributed: rank / world / stride; the splitters stay (tree_splitters sets them)
__global__ void ctl_init_dist(BuildCtl *ctl, int rank, int world, int stride) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  ctl->n_local = 0; ctl->rank = rank; ctl->world = world;
  ctl->seg = rank * stride; ctl->seg_next = world * stride; ctl->stride = stride; ctl->end = world * stride;
  ctl->cprev = -1; ctl->cnext = -1; ctl->overflow = 0; ctl->first = 0; ctl->nentries = 0; ctl->maxent = 0;
  for (int k = 0; k < 4; k++) ctl->gP.c[k] = 0.0;
}
__global__ void iota_kernel(int *__restrict__ v, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = (int)i;
}
// bootstrap of the distributed build: equal-count key ranges from fully sorted keys
__global__ void splitters_from_sorted_kernel(const uint64_t *__restrict__ shi, int64_t n, int world,
                                             BuildCtl *ctl) {
  const int k = threadIdx.x;
  if (blockIdx.x != 0 || k > world || k > DIST_MAX_RANKS) return;
  uint64_t v = (k == 0) ? 0ull : (k == world ? ~0ull : shi[(n * k) / world]);
  ctl->split_next[k] = v;
  ctl->split[k] = v;
}

// ---- select: keys of this rank's range, in source order (stable) ------------------------------------
static constexpr int SEL_THREADS = 256;
static constexpr int SEL_ROUNDS = 8;
static constexpr int SEL_TILE = SEL_THREADS * SEL_ROUNDS;
// counts per tile of 2048 keys; the tile offsets are chunked_scan<int> of them
__global__ void __launch_bounds__(SEL_THREADS)
select_count_kernel(const uint64_t *__restrict__ keys, int64_t n, const BuildCtl *__restrict__ ctl,
                    int *__restrict__ tilecnt) {
  __shared__ int wsum[SEL_THREADS / 32];
  const uint64_t lo = ctl->split[ctl->rank], hi = ctl->split[ctl->rank + 1];
  const bool last = ctl->rank == ctl->world - 1;
  const int64_t base = (int64_t)blockIdx.x * SEL_TILE;
  int c = 0;
#pragma unroll
  for (int r = 0; r < SEL_ROUNDS; r++) {
    const int64_t q = base + r * SEL_THREADS + threadIdx.x;
    if (q < n) { const uint64_t k = keys[q]; c += (k >= lo && (last || k < hi)) ? 1 : 0; }
  }
  for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int k = 0; k < SEL_THREADS / 32; k++) t += wsum[k];
    tilecnt[blockIdx.x] = t;
  }
}
// element order inside a tile: warp w owns 256 consecutive keys (8 rounds of 32), so the compaction
// keeps the source order
__global__ void __launch_bounds__(SEL_THREADS)
select_compact_kernel(const uint64_t *__restrict__ keys, int64_t n, BuildCtl *__restrict__ ctl,
                      const int *__restrict__ tileoff /* exclusive scan of tilecnt, ntiles + 1 */, int ntiles,
                      uint64_t *__restrict__ kout, int *__restrict__ vout, int64_t ncap) {
  __shared__ int wcnt[SEL_THREADS / 32];
  const uint64_t lo = ctl->split[ctl->rank], hi = ctl->split[ctl->rank + 1];
  const bool last = ctl->rank == ctl->world - 1;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t seg0 = (int64_t)blockIdx.x * SEL_TILE + (int64_t)w * (32 * SEL_ROUNDS);
  uint64_t k[SEL_ROUNDS];
  unsigned m[SEL_ROUNDS];
  int c = 0;
#pragma unroll
  for (int r = 0; r < SEL_ROUNDS; r++) {
    const int64_t q = seg0 + r * 32 + lane;
    k[r] = (q < n) ? keys[q] : 0;
    const bool in = (q < n) && k[r] >= lo && (last || k[r] < hi);
    m[r] = __ballot_sync(0xffffffffu, in);
    c += __popc(m[r]);
  }
  if (lane == 0) wcnt[w] = c;
  __syncthreads();
  int off = tileoff[blockIdx.x];
  for (int j = 0; j < w; j++) off += wcnt[j];
#pragma unroll
  for (int r = 0; r < SEL_ROUNDS; r++) {
    const int64_t q = seg0 + r * 32 + lane;
    if (m[r] & (1u << lane)) {
      const int dst = off + __popc(m[r] & ((1u << lane) - 1u));
      if (dst < ncap) {
        kout[dst] = k[r];
        vout[dst] = (int)q;
      }
    }
    off += __popc(m[r]);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    // more particles in this rank's key range than its arrays hold: build what fits, raise the flag
    // (the walk then does nothing and the host reports it at its next synchronisation)
    const int tot = tileoff[ntiles];
    if (tot > ncap) ctl->overflow = 1;
    ctl->n_local = tot > ncap ? (int)ncap : tot;
  }
}

// boundary keys of this rank's sorted range -> its slot of the gathered RankRec1 array
__global__ void rec1_kernel(const uint64_t *__restrict__ shi, const BuildCtl *__restrict__ ctl,
                            RankRec1 *__restrict__ all) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  RankRec1 r;
  r.n = ctl->n_local;
  r.kfirst = r.n > 0 ? shi[0] : 0;
  r.klast = r.n > 0 ? shi[r.n - 1] : 0;
  // pad[0]: this rank's key range held more particles than its arrays (select_compact_kernel): every
  // rank must know, or the others would walk a tree with particles missing
  r.pad[0] = ctl->overflow;
  r.pad[1] = r.pad[2] = 0;
  all[ctl->rank] = r;
}

__device__ __forceinline__ int common_levels_hi(uint64_t a, uint64_t b) {
  const uint64_t x = a ^ b;
  return x ? __clzll((long long)(x << 1)) / 3 : LEVELS_HI;
}
// after the RankRec1 all-gather: octant levels shared with the neighbouring non-empty ranks
__global__ void neighbours_kernel(const RankRec1 *__restrict__ all, BuildCtl *__restrict__ ctl) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const int r = ctl->rank, P = ctl->world;
  int cprev = -1, cnext = -1, seg_next = ctl->end;
  if (all[r].n > 0) {
    for (int q = r - 1; q >= 0; q--)
      if (all[q].n > 0) { cprev = common_levels_hi(all[q].klast, all[r].kfirst); break; }
    for (int q = r + 1; q < P; q++)
      if (all[q].n > 0) { cnext = common_levels_hi(all[r].klast, all[q].kfirst); seg_next = q * ctl->stride; break; }
  }
  ctl->cprev = cprev;
  ctl->cnext = cnext;
  ctl->seg_next = seg_next;
  int first = ctl->end;  // the root entry is the first entry of the first non-empty rank
  for (int q = P - 1; q >= 0; q--) {
    ctl->counts[q] = all[q].n;
    if (all[q].n > 0) first = q * ctl->stride;
    if (all[q].pad[0]) ctl->overflow = 1;  // some rank's range overflowed: nobody walks this step
  }
  ctl->first = first;
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

// ---- level-min tables: where a cell ends, without touching the keys ------------------------------------
// The cell of level L that starts at sorted particle p ends at the first q >= p with clev[q] < L.
// lm.t[0][q] = clev[q] + 1 as an unsigned byte (0 beyond the last particle); lm.t[k][j] = the
// minimum of lm.t[k-1][32 j .. 32 j + 31]: 1/31 of a byte per particle in total.  cell_end climbs
// until a group to the right holds a smaller value and descends to it: two 16-byte loads and ~60
// integer instructions per table level, <= 2 * levels of them (emit_kernel's gallop + bisection over
// the 8-byte keys costs up to 40 dependent loads for the big cells, and a warp waits for its
// slowest lane).
static constexpr int LM_MAX_TABLES = 7;  // 32^7 particles
struct LevelMin {
  unsigned char *t[LM_MAX_TABLES];
  int64_t n[LM_MAX_TABLES];  // entries of table k (n[0] = particles, capacity); allocated in multiples of 32
  int ntab;                  // 0: no tables
};
static inline __host__ __device__ int64_t lm_padded(int64_t n) { return ((n + 31) / 32) * 32; }
// sizes of the tables for n particles; returns the bytes needed when every table starts at a multiple of
// 256 (table 0 is padded to the levels_kernel grid: a multiple of 256 entries)
static inline size_t lm_layout(int64_t n, LevelMin &lm, size_t off[LM_MAX_TABLES]) {
  size_t bytes = 0;
  lm.ntab = 0;
  int64_t m = n;
  for (int k = 0; k < LM_MAX_TABLES; k++) {
    lm.n[k] = m;
    off[k] = bytes;
    const int64_t padded = (k == 0) ? ((m + 255) / 256) * 256 : lm_padded(m);
    bytes += (size_t)((padded + 255) / 256) * 256;
    lm.ntab = k + 1;
    if (m <= 32) break;
    m = (m + 31) / 32;
  }
  return bytes;
}
// first entry j >= from of group g (32 entries) of a table with value < t (t4 = t in every byte), or 32
__device__ __forceinline__ int lm_first(const unsigned char *__restrict__ tab, int64_t g, int from, unsigned t4) {
  const uint4 *gp = reinterpret_cast<const uint4 *>(tab + g * 32);
  const uint4 a = gp[0], b = gp[1];
  const unsigned w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  unsigned bits = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    // values are <= 43 < 128: (u | 0x80) - t keeps bit 7 of a byte exactly when u >= t, no borrows
    const unsigned x = (w[i] | 0x80808080u) - t4;
    const unsigned y = (~x & 0x80808080u) >> 7;            // 0x01 in the bytes with u < t
    bits |= ((y * 0x01020408u) >> 24) << (4 * i);          // byte j -> bit j
  }
  bits = from < 32 ? (bits & ~((1u << from) - 1u)) : 0u;
  return bits ? __ffs((int)bits) - 1 : 32;
}
// last sorted particle of the level-`level` cell that starts at p (n: particles of this rank)
__device__ __forceinline__ int64_t cell_end(const LevelMin &lm, int64_t p, int level, int64_t n) {
  const unsigned t4 = (unsigned)(level + 1) * 0x01010101u;
  int q = lm_first(lm.t[0], p >> 5, (int)(p & 31) + 1, t4);
  int64_t idx;
  if (q < 32) {
    idx = (p & ~(int64_t)31) + q;
  } else {
    idx = p >> 5;
    int k = 1;
    for (;;) {
      if (k >= lm.ntab) return n - 1;  // nothing smaller to the right
      const int64_t g = idx >> 5;
      q = lm_first(lm.t[k], g, (int)(idx & 31) + 1, t4);
      if (q < 32) { idx = (g << 5) + q; break; }
      idx = g;
      k++;
    }
    while (k > 0) {
      k--;
      q = lm_first(lm.t[k], idx, 0, t4);
      if (q >= 32) return n - 1;  // (inconsistent tables: cannot happen)
      idx = (idx << 5) + q;
    }
  }
  return idx < n ? idx : n - 1;
}
// tables 2 .. ntab-1 from table 1 (one CTA; the sizes shrink by 32 per level)
__global__ void levelmin_top_kernel(const __grid_constant__ LevelMin lm) {
  for (int k = 2; k < lm.ntab; k++) {
    const unsigned char *in = lm.t[k - 1];
    unsigned char *out = lm.t[k];
    const int64_t nk = lm.n[k], np = lm_padded(nk);
    for (int64_t j = threadIdx.x; j < np; j += blockDim.x) {
      unsigned m = 0;
      if (j < nk) {  // minimum of 32 bytes: two 16-byte loads
        const uint4 *gp = reinterpret_cast<const uint4 *>(in + 32 * j);
        const uint4 a = gp[0], b = gp[1];
        const unsigned w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        m = 255;
#pragma unroll
        for (int i = 0; i < 8; i++) {
#pragma unroll
          for (int s = 0; s < 32; s += 8) { const unsigned v = (w[i] >> s) & 0xffu; m = v < m ? v : m; }
        }
      }
      out[j] = (unsigned char)m;
    }
    __syncthreads();
  }
}

// cnt[p] = (cells opened at sorted position p) + 1 leaf;  clev[p] = c[p] (c[n-1] = -1)
// ctl (nullable): this rank's particle count and the levels shared across its range boundaries
// lm (ntab > 0): also the level-min tables 0 and 1 (every thread of the grid takes part: the grid
// covers a multiple of 256 positions, positions beyond the last particle hold 0)
__global__ void levels_kernel(const uint64_t *__restrict__ hi, const uint64_t *__restrict__ lo,
                              int64_t n, int levels, signed char *__restrict__ clev,
                              int *__restrict__ cnt, const BuildCtl *__restrict__ ctl = nullptr,
                              const __grid_constant__ LevelMin lm = LevelMin{{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr},
                                                                             {0, 0, 0, 0, 0, 0, 0}, 0}) {
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (ctl) n = ctl->n_local;
  int c = -1;
  if (p < n) {
    int cprev = ctl ? ctl->cprev : -1;
    c = ctl ? ctl->cnext : -1;
    if (p > 0) cprev = common_levels(hi[p - 1], lo ? lo[p - 1] : 0, hi[p], lo ? lo[p] : 0, levels);
    if (p + 1 < n) c = common_levels(hi[p], lo ? lo[p] : 0, hi[p + 1], lo ? lo[p + 1] : 0, levels);
    clev[p] = (signed char)c;
    int open = c - cprev;
    cnt[p] = (open > 0 ? open : 0) + 1;
  }
  if (lm.ntab > 0) {
    lm.t[0][p] = (unsigned char)(c + 1);
    if (lm.ntab > 1) {
      const int m = __reduce_min_sync(0xffffffffu, c + 1);
      if ((threadIdx.x & 31) == 0) lm.t[1][p >> 5] = (unsigned char)m;
    }
  }
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
                                     double4 *__restrict__ out, const BuildCtl *__restrict__ ctl = nullptr,
                                     int *__restrict__ sidx_slot = nullptr) {
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (ctl) n = ctl->n_local;
  if (p >= n) return;
  const int64_t j = idx[p];
  if (sidx_slot) sidx_slot[p] = (int)j;  // this rank's slot of the gathered sorted-index array
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
// opt-in quadrupoles: second moments about the root centre, plain double.  A cell's central second
// moment is the difference of two prefixes minus M c c^T: for the smallest cells that difference
// cancels to ~1e-16 N m R^2 / (m_cell size^2) relative accuracy -- percent level for a two-particle
// cell at N = 4M, whose quadrupole term is itself ~(size/d)^2 < theta^2/4 of a monopole that is one
// of ~600: far below the monopole truncation error the extension removes.
struct InSecondRel {
  const double4 *sp;
  const double *root;
  __device__ __forceinline__ D6 operator()(int64_t q) const {
    const double4 t = sp[q];
    const double dx = t.x - root[0], dy = t.y - root[1], dz = t.z - root[2];
    D6 r;
    r.c[0] = t.w * dx * dx; r.c[1] = t.w * dy * dy; r.c[2] = t.w * dz * dz;
    r.c[3] = t.w * dx * dy; r.c[4] = t.w * dx * dz; r.c[5] = t.w * dy * dz;
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
// a cell that starts at local particle p and continues beyond this rank's range (fp32 tree only):
// global prefix at its end (stitch_kernel) minus the global prefix at p
__device__ __forceinline__ void moment_diff_beyond(const D4 *__restrict__ P, int64_t p, const BuildCtl *ctl,
                                                   int level, double mh[4]) {
  const D4 pe = ctl->xP[level], ps = P[p], g = ctl->gP;
#pragma unroll
  for (int k = 0; k < 4; k++) mh[k] = pe.c[k] - __dadd_rn(g.c[k], ps.c[k]);
}
__device__ __forceinline__ void moment_diff_beyond(const DD4 *, int64_t, const BuildCtl *, int, double mh[4]) {
  mh[0] = mh[1] = mh[2] = mh[3] = 0.0;  // the fp64 tree is never distributed (cnext is always -1)
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

// ---- distributed build: what a rank publishes about its range, and the stitch --------------------
// One warp.  Lane l: the level-l cell that contains local particle 0 ends at local index bend (one
// past its last particle; n when the whole range shares the prefix) -- binary search over the
// sorted keys.  Lanes also take the key samples; lane 0 the totals.
__global__ void rec2_kernel(const uint64_t *__restrict__ shi, const int *__restrict__ base,
                            const D4 *__restrict__ P, const BuildCtl *__restrict__ ctl,
                            RankRec2 *__restrict__ all) {
  const int lane = threadIdx.x;
  if (blockIdx.x != 0 || lane >= 32) return;
  const int n = ctl->n_local;
  RankRec2 *rec = &all[ctl->rank];
  if (lane <= LEVELS_HI) {
    int bend = 0;
    if (n > 0) {
      const uint64_t h0 = shi[0];
      int lo_i = 1, hi_i = n;  // first index in [1, n] whose key leaves the prefix (n: none)
      while (lo_i < hi_i) {
        const int mid = (lo_i + hi_i) >> 1;
        if (same_prefix(shi[mid], 0, h0, 0, lane)) lo_i = mid + 1; else hi_i = mid;
      }
      bend = lo_i;
    }
    rec->tab[lane].bend = bend;
    rec->tab[lane].base = base[bend];
    rec->tab[lane].P = P[bend];
  }
  for (int j = lane; j < DIST_SAMPLES; j += 32) {
    int q = (int)(((int64_t)(2 * j + 1) * n) / (2 * DIST_SAMPLES));
    if (q > n - 1) q = n - 1;
    rec->samples[j] = n > 0 ? shi[q] : 0;
  }
  if (lane == 0) {
    rec->Mtot = P[n];
    rec->nentries = base[n];
    rec->pad = 0;
  }
}

__device__ __forceinline__ D4 d4_add(const D4 &a, const D4 &b) {
  D4 r;
#pragma unroll
  for (int k = 0; k < 4; k++) r.c[k] = __dadd_rn(a.c[k], b.c[k]);
  return r;
}
// After the RankRec2 all-gather (one thread; P <= 64 ranks, <= 22 levels): moment prefix of the
// earlier ranks, the ends of the cells that leave this rank's range, and the next step's splitters.
__global__ void stitch_kernel(const RankRec1 *__restrict__ r1, const RankRec2 *__restrict__ r2,
                              BuildCtl *__restrict__ ctl, int64_t n_total) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const int r = ctl->rank, P = ctl->world, stride = ctl->stride;
  D4 g;
  for (int k = 0; k < 4; k++) g.c[k] = 0.0;
  for (int q = 0; q < r; q++) g = d4_add(g, r2[q].Mtot);
  ctl->gP = g;
  ctl->nentries = r2[r].nentries;
  // a segment that does not fit anywhere stops the walk on EVERY rank (all see the same records)
  int maxent = 0;
  for (int q = 0; q < P; q++) maxent = r2[q].nentries > maxent ? r2[q].nentries : maxent;
  if (ctl->overflow) maxent = 0x7fffffff;  // (a particle-count overflow, see rec1_kernel) the host's signal
  ctl->maxent = maxent;
  if (maxent > stride) ctl->overflow = 1;
  // cells of level l <= cnext that contain my last particle continue into the next non-empty rank
  for (int l = 0; l <= ctl->cnext && l <= LEVELS_HI; l++) {
    D4 gq = d4_add(g, r2[r].Mtot);  // moment prefix at the start of rank q
    int q = r + 1;
    while (q < P && r1[q].n == 0) q++;
    int skip = ctl->end;
    D4 pend = gq;
    while (q < P) {
      const CellEnd t = r2[q].tab[l];
      if (t.bend < r1[q].n) {  // the cell ends inside rank q
        skip = q * stride + t.base;
        pend = d4_add(gq, t.P);
        break;
      }
      // the cell covers all of rank q: does it continue into the following non-empty rank?
      int q2 = q + 1;
      while (q2 < P && r1[q2].n == 0) q2++;
      const D4 gq2 = d4_add(gq, r2[q].Mtot);
      if (q2 < P && common_levels_hi(r1[q].klast, r1[q2].kfirst) >= l) { q = q2; gq = gq2; continue; }
      skip = (q2 < P) ? q2 * stride : ctl->end;
      pend = gq2;
      break;
    }
    ctl->xskip[l] = skip;
    ctl->xP[l] = pend;
  }
  // next step's key ranges: equal counts, from the ranks' key samples (sample j of rank q sits at
  // global sorted position G_q + (j + 1/2) n_q / 64)
  ctl->split_next[0] = 0;
  ctl->split_next[P] = ~0ull;
  int64_t G[DIST_MAX_RANKS + 1];
  G[0] = 0;
  for (int q = 0; q < P; q++) G[q + 1] = G[q] + r1[q].n;
  for (int k = 1; k < P; k++) {
    const int64_t t = (n_total * k) / P;
    int q = 0;
    while (q < P - 1 && (G[q + 1] <= t || r1[q].n == 0)) q++;
    uint64_t sk = ctl->split[k];
    if (r1[q].n > 0) {
      int64_t j = ((t - G[q]) * DIST_SAMPLES) / r1[q].n;
      if (j < 0) j = 0;
      if (j > DIST_SAMPLES - 1) j = DIST_SAMPLES - 1;
      sk = r2[q].samples[j];
    }
    if (sk < ctl->split_next[k - 1]) sk = ctl->split_next[k - 1];
    ctl->split_next[k] = sk;
  }
}
// adopt the splitters the previous step computed (start of a step)
__global__ void splitters_advance_kernel(BuildCtl *ctl, int world) {
  const int k = threadIdx.x;
  if (blockIdx.x == 0 && k <= world && k <= DIST_MAX_RANKS) ctl->split[k] = ctl->split_next[k];
}

// GH_EMIT_MINBLOCKS: resident 128-thread CTAs per SM the register allocation is capped for
// (scripts/build_variants.py; ncu: 66 registers -> 33 % of the warp slots active, latency bound)
#ifdef GH_EMIT_MINBLOCKS
#define GH_EMIT_BOUNDS __launch_bounds__(128, GH_EMIT_MINBLOCKS)
#else
#define GH_EMIT_BOUNDS
#endif
// ctl: segment placement (seg, seg_next, stride), the cross-rank cell ends and, with `dist`, the
// particle count of this rank.  Every skip link that would leave the segment is seg_next.
template <class Src, class Real>
__global__ void GH_EMIT_BOUNDS emit_kernel(const double4 *__restrict__ sp, const uint64_t *__restrict__ hi,
                            const uint64_t *__restrict__ lo, const signed char *__restrict__ clev,
                            const int *__restrict__ base /* n+1, exclusive scan of cnt */,
                            const typename MomentOf<Real>::type *__restrict__ P, int64_t n,
                            const double *__restrict__ root, bool rel_origin, double inv_theta2,
                            Entries<Real> E, int *__restrict__ maxlevel, BuildCtl *__restrict__ ctl,
                            bool dist, const D6 *__restrict__ P2 = nullptr, Real *__restrict__ quad = nullptr) {
  // P2 / quad (nullable, single rank only): the opt-in quadrupole extension.  quad[6 e ..] receives
  // the traceless quadrupole (xx, yy, zz, xy, xz, yz) of cell entry e about its centre of mass,
  // zeros for leaves.
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (dist) n = ctl->n_local;
  if (p >= n) return;
  using V4 = typename Vec4<Real>::type;
  const int seg = ctl->seg, seg_next = ctl->seg_next, seg_cap = ctl->seg + ctl->stride;
  const double ox = rel_origin ? root[0] : 0.0, oy = rel_origin ? root[1] : 0.0,
               oz = rel_origin ? root[2] : 0.0;
  const int c = clev[p];
  const int cprev = (p > 0) ? clev[p - 1] : ctl->cprev;
  const double4 self = sp[p];
  const double x[3] = {self.x, self.y, self.z};
  int e = seg + base[p];
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
        // a cell that holds this rank's last particle and whose prefix the next rank's first key
        // shares continues beyond the range: its end comes from the stitched table
        const bool beyond = (b == n - 1) && (ctl->cnext >= level);
        double mh[4];
        int skipidx;
        if (beyond) {
          moment_diff_beyond(P, p, ctl, level, mh);
          skipidx = ctl->xskip[level];
        } else {
          moment_diff(P, p, b, mh);
          skipidx = (b + 1 < n) ? seg + base[b + 1] : seg_next;
        }
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
        if (quad && e < seg_cap) {
          const D6 se = P2[b + 1], ss = P2[p];
          // centre of mass relative to the root centre (fp32 moments already are)
          const double c0 = mh[1] / mh[0] - (sizeof(Real) == 4 ? 0.0 : root[0]);
          const double c1 = mh[2] / mh[0] - (sizeof(Real) == 4 ? 0.0 : root[1]);
          const double c2 = mh[3] / mh[0] - (sizeof(Real) == 4 ? 0.0 : root[2]);
          const double cxx = (se.c[0] - ss.c[0]) - mh[0] * c0 * c0, cyy = (se.c[1] - ss.c[1]) - mh[0] * c1 * c1,
                       czz = (se.c[2] - ss.c[2]) - mh[0] * c2 * c2;
          const double tr = cxx + cyy + czz;
          Real *qd = quad + 6 * (size_t)e;
          qd[0] = (Real)(3.0 * cxx - tr);
          qd[1] = (Real)(3.0 * cyy - tr);
          qd[2] = (Real)(3.0 * czz - tr);
          qd[3] = (Real)(3.0 * ((se.c[3] - ss.c[3]) - mh[0] * c0 * c1));
          qd[4] = (Real)(3.0 * ((se.c[4] - ss.c[4]) - mh[0] * c0 * c2));
          qd[5] = (Real)(3.0 * ((se.c[5] - ss.c[5]) - mh[0] * c1 * c2));
        }
        cen.x = (Real)(cc[0] - ox);
        cen.y = (Real)(cc[1] - oy);
        cen.z = (Real)(cc[2] - oz);
        if (e < seg_cap) {
          // (size / dist) < theta  <=>  size^2 / theta^2 < dist^2   (theta = 0: inf, never accepted)
          if (sizeof(Real) == 4) {
            cen.w = (Real)__int_as_float((int)(((unsigned)level << SKIP_BITS) | (unsigned)skipidx));
          } else {
            cen.w = (Real)(__dmul_rn(__dmul_rn(size, size), inv_theta2));
            E.skip[e] = skipidx;
          }
          pack_node(E.node[e], cen, com);
        } else {
          ctl->overflow = 1;
        }
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
  const int after = (p + 1 < n) ? e + 1 : seg_next;
  if (e >= seg_cap) { ctl->overflow = 1; return; }
  if (quad) {
    Real *qd = quad + 6 * (size_t)e;
#pragma unroll
    for (int k = 0; k < 6; k++) qd[k] = (Real)0;
  }
  if (sizeof(Real) == 4) {
    cen.w = (Real)__int_as_float((int)(((unsigned)LEAF_LEVEL << SKIP_BITS) | (unsigned)after));
  } else {
    cen.w = (Real)-1;
    E.skip[e] = after;
  }
  pack_node(E.node[e], cen, com);
}


// ---- K6b emit, fp32 tree, warp-cooperative ("warp" form) ---------------------------------------------
// emit_kernel gives every particle a thread that loops over the cells the particle opens: ncu at
// N = 4M counts 6.4 active lanes per instruction (only a third of the particles open cells, and
// their loops have different lengths), 1170 instructions per warp.  Here a warp owns 32 sorted
// particles, writes their 32 leaves together, numbers the cells they open (a warp scan of the
// counts) and deals those cells to its lanes -- one cell per lane, the same arithmetic as
// emit_kernel's (centre: closed form at the first level the owner opens, then the owner's descent;
// end of the cell: the same search; moments: the same differences), so the entries are
// bit-identical (tests/emu compares the two kernels' arrays).  fp32 entries only, no quadrupoles.
__global__ void __launch_bounds__(128)
emit32_warp_kernel(const double4 *__restrict__ sp, const uint64_t *__restrict__ hi,
                   const signed char *__restrict__ clev, const int *__restrict__ base /* n+1 */,
                   const D4 *__restrict__ P, int64_t n, const double *__restrict__ root, Entries<float> E,
                   int *__restrict__ maxlevel, BuildCtl *__restrict__ ctl, bool dist,
                   const __grid_constant__ LevelMin lm) {
  const int lane = threadIdx.x & 31;
  const int64_t wbase = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 32;
  if (dist) n = ctl->n_local;
  if (wbase >= n) return;  // the whole warp
  const int64_t p = wbase + lane;
  const bool valid = p < n;
  const int seg = ctl->seg, seg_next = ctl->seg_next, seg_cap = ctl->seg + ctl->stride;
  const int c = valid ? (int)clev[p] : -1;
  int cprev = __shfl_up_sync(0xffffffffu, c, 1);
  if (lane == 0) cprev = (p > 0) ? (int)clev[p - 1] : ctl->cprev;
  const int ncell = (valid && c > cprev) ? c - cprev : 0;
  const int e0 = valid ? seg + base[p] : 0;
  const uint64_t h0 = valid ? hi[p] : 0;
  if (valid) {
    // the particle's own leaf: COM = particle position (:473-475), always accepted (:502)
    const double4 self = sp[p];
    float4 com, cen;
    com.x = (float)(self.x - root[0]);
    com.y = (float)(self.y - root[1]);
    com.z = (float)(self.z - root[2]);
    com.w = (float)self.w;
    cen.x = cen.y = cen.z = 0.f;
    const int e = e0 + ncell;
    const int after = (p + 1 < n) ? e + 1 : seg_next;
    if (e >= seg_cap) {
      ctl->overflow = 1;
    } else {
      cen.w = __int_as_float((int)(((unsigned)LEAF_LEVEL << SKIP_BITS) | (unsigned)after));
      pack_node(E.node[e], cen, com);
    }
  }
  const int wmax = __reduce_max_sync(0xffffffffu, ncell > 0 ? c : 0);
  if (lane == 0 && wmax > 0) atomicMax(maxlevel, wmax);

  // the cells: task t of the warp = cell number t - off of the lane whose inclusive count first exceeds t
  const int incl = warp_scan<int>(ncell, lane);
  const int off = incl - ncell;
  const int T = __shfl_sync(0xffffffffu, incl, 31);
  for (int t0 = 0; t0 < T; t0 += 32) {
    const int t = t0 + lane;
    const bool act = t < T;
    int lo_l = 0, hi_l = 31;
#pragma unroll
    for (int s = 0; s < 5; s++) {
      const int mid = (lo_l + hi_l) >> 1;
      const int v = __shfl_sync(0xffffffffu, incl, mid);
      if (v > t) hi_l = mid; else lo_l = mid + 1;
    }
    const int owner = lo_l < 31 ? lo_l : 31;
    const int o_off = __shfl_sync(0xffffffffu, off, owner);
    const int o_cprev = __shfl_sync(0xffffffffu, cprev, owner);
    const int o_e0 = __shfl_sync(0xffffffffu, e0, owner);
    const uint64_t o_h = __shfl_sync(0xffffffffu, h0, owner);
    if (!act) continue;
    const int level = o_cprev + 1 + (t - o_off);
    const int64_t pt = wbase + owner;
    // centre and side of the cell: as emit_kernel forms them for its particle
    double cc[3] = {root[0], root[1], root[2]};
    double size = root[3];
    int l0 = 0;
    if (o_cprev >= 0) {
      l0 = o_cprev + 1;
      const double sL = ldexp(size, -l0);
#pragma unroll
      for (int k = 0; k < 3; k++) {
        const uint64_t ik = compact3(o_h >> k) >> (LEVELS_HI - l0);
        cc[k] = root[k] - 0.5 * size + ((double)ik + 0.5) * sL;
      }
      size = sL;
    }
    for (int l = l0 + 1; l <= level; l++) {  // the owner's descent (:406,:441-462)
      const unsigned d = (unsigned)((o_h >> (3 * (LEVELS_HI - l))) & 7u);
      const double quarter = __dmul_rn(0.5, __dmul_rn(0.5, size));
#pragma unroll
      for (int k = 0; k < 3; k++) cc[k] = __dadd_rn(cc[k], ((d >> k) & 1u) ? quarter : -quarter);
      size = __dmul_rn(0.5, size);
    }
    // last sorted particle b sharing `level` octant levels with pt: from the level-min tables, or
    // (no tables) emit_kernel's search over the keys
    int64_t lo_i = pt + 1, step = 1, hi_i;
    bool found = false;
    if (lm.ntab > 0) { lo_i = cell_end(lm, pt, level, n); found = true; }
    else
    for (int u = 0; u < 12 && lo_i < n; u++) {
      if (clev[lo_i] < level) { found = true; break; }
      lo_i++;
    }
    if (lo_i >= n) { lo_i = n - 1; found = true; }
    hi_i = lo_i;
    if (!found) for (;;) {
      const int64_t q = lo_i + step;
      if (q >= n) { hi_i = n - 1; break; }
      if (same_prefix(hi[q], 0, o_h, 0, level)) { lo_i = q; step <<= 1; }
      else { hi_i = q - 1; break; }
    }
    while (!found && lo_i < hi_i) {
      const int64_t mid = (lo_i + hi_i + 1) >> 1;
      if (same_prefix(hi[mid], 0, o_h, 0, level)) lo_i = mid;
      else hi_i = mid - 1;
    }
    const int64_t b = lo_i;
    const bool beyond = (b == n - 1) && (ctl->cnext >= level);
    double mh[4];
    int skipidx;
    if (beyond) {
      moment_diff_beyond(P, pt, ctl, level, mh);
      skipidx = ctl->xskip[level];
    } else {
      moment_diff(P, pt, b, mh);
      skipidx = (b + 1 < n) ? seg + base[b + 1] : seg_next;
    }
    float4 com, cen;
    com.x = (float)(mh[1] / mh[0]);
    com.y = (float)(mh[2] / mh[0]);
    com.z = (float)(mh[3] / mh[0]);
    com.w = (float)mh[0];
    cen.x = (float)(cc[0] - root[0]);
    cen.y = (float)(cc[1] - root[1]);
    cen.z = (float)(cc[2] - root[2]);
    const int e = o_e0 + (level - (o_cprev + 1));
    if (e < seg_cap) {
      cen.w = __int_as_float((int)(((unsigned)level << SKIP_BITS) | (unsigned)skipidx));
      pack_node(E.node[e], cen, com);
    } else {
      ctl->overflow = 1;
    }
  }
}


// ---- distributed walk: kick and drift of the OWNED particles -------------------------------------
// After the all-gather of the accelerations.  Sorted particle j of rank q's range was walked by rank
// (kb + q) % world for block kb = j / blk; its acceleration sits in that rank's buffer at slot
// q T blk + (kb / world) blk + j % blk (TargetsView, walk.cuh).  Two kernels so that the 170 bytes
// of state the epilogue touches per particle are accessed in OWNED-index order (coalesced): the
// first inverts the sorted-index array for the owned particles (one scattered 4-byte store each),
// the second runs over the owned particles and fetches each one's acceleration (one scattered
// 16-byte load each).
__global__ void dist_inverse_kernel(const int *__restrict__ sidx_all, const BuildCtl *__restrict__ ctl, int world,
                                    int ncap, int64_t ib, int64_t ni, int *__restrict__ inv /* ni */) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= (int64_t)world * ncap) return;
  const int q = (int)(v / ncap), j = (int)(v % ncap);
  if (j >= ctl->counts[q]) return;
  const int64_t gi = sidx_all[v];
  if (gi >= ib && gi < ib + ni) inv[gi - ib] = (int)v;
}
__global__ void dist_epilogue_kernel(const int *__restrict__ inv, const BuildCtl *__restrict__ ctl,
                                     const float4 *__restrict__ acc_all, int world, int ncap, int blk, int T,
                                     int64_t ni, Epilogue ep) {
  if (ctl->overflow) return;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ni) return;
  const int v = inv[i];
  const int q = v / ncap, j = v % ncap;
  const int kb = j / blk, o = j % blk;
  const int w = (kb + q) % world, t = kb / world;
  const int64_t slots = (int64_t)world * T * blk;  // target slots per rank
  const float4 a = acc_all[(int64_t)w * slots + (int64_t)q * T * blk + (int64_t)t * blk + o];
  apply_epilogue(ep, i, (double)a.x, (double)a.y, (double)a.z);
}

}  // namespace gh
