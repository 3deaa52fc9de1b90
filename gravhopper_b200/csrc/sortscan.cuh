// sortscan.cuh -- hand-written device primitives of the tree build (sm_100a):
//   * chunked_scan<T>: deterministic inclusive scan (P[q+1] = x_0 (+) ... (+) x_q, P[0] = identity)
//     over 256-element warp chunks, recursive; used for the moment sums (T = DD4: double-double,
//     fp64 tree; T = D4: plain double, fp32 tree; non-associative operators, fixed order) and for
//     the pre-order offsets / radix digit offsets (T = int);
//   * radix_sort_pairs: stable LSD radix sort of (uint64 key, int value) pairs, 8 bits per pass
//     (K5 of SURVEY 2.4).  Per pass: per-CTA digit histogram (+ global digit totals) -> per-digit
//     row scan of the (digit, CTA) table -> scatter.  In the scatter each warp ranks its 256 keys
//     in order with MATCH.ANY ballots (stable), the CTA's 2048 keys are placed in shared memory in
//     digit order, and written out as contiguous runs per digit (coalesced).
#pragma once
#include "common.cuh"

namespace gh {

static constexpr int SCAN_THREADS = 256;
static constexpr int SCAN_ROUNDS = 8;
static constexpr int SCAN_WARP_ELEMS = 32 * SCAN_ROUNDS;

// ---- element types ---------------------------------------------------------------------------
struct DD {
  double h, l;
};
struct DD4 {
  DD c[4];
};
__device__ __forceinline__ void dd_add(double ah, double al, double bh, double bl, double &rh,
                                       double &rl) {
  double s = __dadd_rn(ah, bh);
  double bb = __dadd_rn(s, -ah);
  double e = __dadd_rn(__dadd_rn(ah, -__dadd_rn(s, -bb)), __dadd_rn(bh, -bb));
  e = __dadd_rn(e, __dadd_rn(al, bl));
  rh = __dadd_rn(s, e);
  rl = __dadd_rn(e, -__dadd_rn(rh, -s));
}

template <class T> struct ScanOps;
template <> struct ScanOps<int> {
  static __device__ __forceinline__ int zero() { return 0; }
  static __device__ __forceinline__ int add(int a, int b) { return a + b; }
  static __device__ __forceinline__ int shfl_up(int v, int d) { return __shfl_up_sync(0xffffffffu, v, d); }
  static __device__ __forceinline__ int shfl(int v, int src) { return __shfl_sync(0xffffffffu, v, src); }
};
template <> struct ScanOps<DD4> {
  static __device__ __forceinline__ DD4 zero() {
    DD4 r;
#pragma unroll
    for (int k = 0; k < 4; k++) r.c[k].h = r.c[k].l = 0.0;
    return r;
  }
  static __device__ __forceinline__ DD4 add(const DD4 &a, const DD4 &b) {
    DD4 r;
#pragma unroll
    for (int k = 0; k < 4; k++) dd_add(a.c[k].h, a.c[k].l, b.c[k].h, b.c[k].l, r.c[k].h, r.c[k].l);
    return r;
  }
  static __device__ __forceinline__ DD4 shfl_up(const DD4 &v, int d) {
    DD4 r;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      r.c[k].h = __shfl_up_sync(0xffffffffu, v.c[k].h, d);
      r.c[k].l = __shfl_up_sync(0xffffffffu, v.c[k].l, d);
    }
    return r;
  }
  static __device__ __forceinline__ DD4 shfl(const DD4 &v, int src) {
    DD4 r;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      r.c[k].h = __shfl_sync(0xffffffffu, v.c[k].h, src);
      r.c[k].l = __shfl_sync(0xffffffffu, v.c[k].l, src);
    }
    return r;
  }
};

// plain double moments (m, m x, m y, m z): the fp32 tree's scan element (half the bytes of DD4, no
// error-free transformations; see moment_diff in tree.cu for why that is enough there)
struct D4 {
  double c[4];
};
template <> struct ScanOps<D4> {
  static __device__ __forceinline__ D4 zero() {
    D4 r;
#pragma unroll
    for (int k = 0; k < 4; k++) r.c[k] = 0.0;
    return r;
  }
  static __device__ __forceinline__ D4 add(const D4 &a, const D4 &b) {
    D4 r;
#pragma unroll
    for (int k = 0; k < 4; k++) r.c[k] = __dadd_rn(a.c[k], b.c[k]);
    return r;
  }
  static __device__ __forceinline__ D4 shfl_up(const D4 &v, int d) {
    D4 r;
#pragma unroll
    for (int k = 0; k < 4; k++) r.c[k] = __shfl_up_sync(0xffffffffu, v.c[k], d);
    return r;
  }
  static __device__ __forceinline__ D4 shfl(const D4 &v, int src) {
    D4 r;
#pragma unroll
    for (int k = 0; k < 4; k++) r.c[k] = __shfl_sync(0xffffffffu, v.c[k], src);
    return r;
  }
};

// second moments (m dx dx, m dy dy, m dz dz, m dx dy, m dx dz, m dy dz) about the root centre: the scan
// element of the opt-in quadrupole extension (plain double, see emit_kernel)
struct D6 {
  double c[6];
};
template <> struct ScanOps<D6> {
  static __device__ __forceinline__ D6 zero() {
    D6 r;
#pragma unroll
    for (int k = 0; k < 6; k++) r.c[k] = 0.0;
    return r;
  }
  static __device__ __forceinline__ D6 add(const D6 &a, const D6 &b) {
    D6 r;
#pragma unroll
    for (int k = 0; k < 6; k++) r.c[k] = __dadd_rn(a.c[k], b.c[k]);
    return r;
  }
  static __device__ __forceinline__ D6 shfl_up(const D6 &v, int d) {
    D6 r;
#pragma unroll
    for (int k = 0; k < 6; k++) r.c[k] = __shfl_up_sync(0xffffffffu, v.c[k], d);
    return r;
  }
  static __device__ __forceinline__ D6 shfl(const D6 &v, int src) {
    D6 r;
#pragma unroll
    for (int k = 0; k < 6; k++) r.c[k] = __shfl_sync(0xffffffffu, v.c[k], src);
    return r;
  }
};

// inclusive scan across the 32 lanes (fixed order: Hillis-Steele, lower lanes on the left)
template <class T>
__device__ __forceinline__ T warp_scan(T v, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    T o = ScanOps<T>::shfl_up(v, d);
    if (lane >= d) v = ScanOps<T>::add(o, v);
  }
  return v;
}

template <class T>
struct InArray {
  const T *a;
  __device__ __forceinline__ T operator()(int64_t q) const { return a[q]; }
};

// `ndev` (nullable): the element count lives on the device (distributed tree build: a rank does not
// know on the host how many particles fall into its key range); `n` is then the capacity the
// launch was sized for, and chunks beyond the real count contribute zeros.
template <class T, class In>
__global__ void __launch_bounds__(SCAN_THREADS)
scan_phase1(In in, int64_t n, T *__restrict__ warpsum, const int *__restrict__ ndev = nullptr) {
  const int lane = threadIdx.x & 31;
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t base = w * SCAN_WARP_ELEMS;
  if (base >= n) return;
  if (ndev) {
    n = *ndev;
    if (base >= n) { if (lane == 0) warpsum[w] = ScanOps<T>::zero(); return; }
  }
  T acc = ScanOps<T>::zero();
#pragma unroll 2
  for (int r = 0; r < SCAN_ROUNDS; r++) {  // element order inside the warp: round-major
    const int64_t q = base + r * 32 + lane;
    T v = (q < n) ? in(q) : ScanOps<T>::zero();
    v = warp_scan<T>(v, lane);  // same association as phase 3
    acc = ScanOps<T>::add(acc, ScanOps<T>::shfl(v, 31));
  }
  if (lane == 0) warpsum[w] = acc;
}
// P[q + 1] = offset of the warp (+) inclusive scan inside the warp; P[0] = identity.
// warpoff == nullptr: single-warp launch over the whole (short) array.
template <class T, class In>
__global__ void __launch_bounds__(SCAN_THREADS)
scan_phase3(In in, int64_t n, const T *__restrict__ warpoff, T *__restrict__ P /* n + 1 */,
            const int *__restrict__ ndev = nullptr) {
  const int lane = threadIdx.x & 31;
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t base = w * SCAN_WARP_ELEMS;
  if (w == 0 && lane == 0) P[0] = ScanOps<T>::zero();
  if (ndev) n = *ndev;
  if (base >= n) return;
  T carry = warpoff ? warpoff[w] : ScanOps<T>::zero();
  for (int r = 0; r < SCAN_ROUNDS; r++) {
    const int64_t q = base + r * 32 + lane;
    T v = (q < n) ? in(q) : ScanOps<T>::zero();
    v = warp_scan<T>(v, lane);
    const T out = ScanOps<T>::add(carry, v);
    if (q < n) P[q + 1] = out;
    carry = ScanOps<T>::add(carry, ScanOps<T>::shfl(v, 31));
  }
}

#ifndef GH_HOST_EMU  // launch sequence: the host emulation (tests/emu) has its own
// P[0..n]: P[0] = identity, P[q+1] = in(0) (+) ... (+) in(q).  Recursion over 256-element warp
// chunks (depth 3 at n = 16.7M, 4 beyond); `levels` supplies scratch for the per-level totals.
// `ndev` (nullable): real element count on the device, n = capacity (see scan_phase1).
template <class T, class In>
static int chunked_scan(In in, int64_t n, T *P, DeviceBuffer *levels, int depth, cudaStream_t st,
                        const int *ndev = nullptr) {
  if (n <= 0) return GH_OK;
  if (n <= SCAN_WARP_ELEMS) {
    scan_phase3<T, In><<<1, 32, 0, st>>>(in, n, nullptr, P, ndev);
    GH_LAUNCH_CHECK();
    return GH_OK;
  }
  if (depth >= 4) { set_error("chunked_scan: too many levels"); return GH_EINVAL; }
  const int64_t nw = (n + SCAN_WARP_ELEMS - 1) / SCAN_WARP_ELEMS;
  const unsigned nsb = (unsigned)((nw * 32 + SCAN_THREADS - 1) / SCAN_THREADS);
  GH_TRY(levels[depth].reserve(sizeof(T) * (size_t)(2 * nw + 1)));  // totals[nw] + prefix[nw + 1]
  T *totals = levels[depth].as<T>();
  T *prefix = totals + nw;
  scan_phase1<T, In><<<nsb, SCAN_THREADS, 0, st>>>(in, n, totals, ndev);
  GH_LAUNCH_CHECK();
  GH_TRY((chunked_scan<T, InArray<T>>(InArray<T>{totals}, nw, prefix, levels, depth + 1, st)));
  scan_phase3<T, In><<<nsb, SCAN_THREADS, 0, st>>>(in, n, prefix, P, ndev);
  GH_LAUNCH_CHECK();
  return GH_OK;
}
#endif  // GH_HOST_EMU

// ---- radix sort ------------------------------------------------------------------------------
// lanes of the warp that hold the same digit (0 .. 255; values >= 0x100 mark lanes without a pair:
// they match nobody).  GH_RS_MATCH = 0: one MATCH.ANY; 1 (default): nine ballots and eight selects.
// ncu of the ranking kernels shows MATCH.ANY as the limiter -- sm__throughput 85 % at 18 % issue
// utilisation, `short_scoreboard` the top stall, ~90 cycles of an SM per MATCH by the kernels'
// MATCH counts -- while the ballot form costs ~30 issue slots the kernels have to spare.
#ifndef GH_RS_MATCH
#define GH_RS_MATCH 1
#endif
__device__ __forceinline__ unsigned warp_match_digit(unsigned d) {
#if GH_RS_MATCH == 0
  return __match_any_sync(0xffffffffu, d);
#else
  const bool valid = d < 0x100u;
  unsigned peers = __ballot_sync(0xffffffffu, valid);
#pragma unroll
  for (int k = 0; k < 8; k++) {
    const bool bit = (d >> k) & 1u;
    const unsigned bk = __ballot_sync(0xffffffffu, bit);
    peers &= bit ? bk : ~bk;
  }
  return valid ? peers : (1u << (threadIdx.x & 31));
#endif
}

static constexpr int RS_THREADS = 256;
static constexpr int RS_WARPS = RS_THREADS / 32;
// measured at N = 4M (profiles/r02_sort_ab.txt): 12 / 10 / 8 / 6 / 5 / 4 keys per thread -> build
// 1.72 / 1.59 / 1.46 / 1.40 / 1.59 / 1.53 ms with the splitter sort (the classic sort is flat
// between 6 and 8)
#ifndef GH_RS_ROUNDS
#define GH_RS_ROUNDS 6
#endif
// ranking of a tile's keys: 0 = the leader of every digit group reads, adds and stores the running
// count round by round; 1 = all ballots first, then one shared-memory atomic per round (see
// tile_rank in bucketsort.cuh).  Measured (profiles/r02_sort_ab.txt): no gain (-0.04 ... +0.08 ms of build).
#ifndef GH_RS_RANK
#define GH_RS_RANK 0
#endif
static constexpr int RS_ROUNDS = GH_RS_ROUNDS;           // keys per thread (6: 1536-key tiles)
static constexpr int RS_TILE = RS_THREADS * RS_ROUNDS;   // 2048 keys per CTA
static constexpr int RS_RADIX = 256;

__global__ void __launch_bounds__(RS_THREADS)
rs_hist_kernel(const uint64_t *__restrict__ keys, int64_t n, int shift, int *__restrict__ hist,
               int nblocks, int *__restrict__ gtot /* [256], zeroed */, const int *__restrict__ ndev = nullptr) {
  __shared__ int h[RS_RADIX];
  if (ndev) n = *ndev;  // real count on the device; the grid covers the capacity
  h[threadIdx.x] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * RS_TILE;
  const int lane = threadIdx.x & 31;
  // all loads first (independent), then one shared-memory atomic per DISTINCT digit of a warp's 32
  // keys (MATCH.ANY): the high digits of Morton keys are the same for a whole tile, and 2048
  // atomics on one address would serialise
  uint64_t k[RS_ROUNDS];
#pragma unroll
  for (int r = 0; r < RS_ROUNDS; r++) {
    const int64_t q = base + r * RS_THREADS + threadIdx.x;
    k[r] = (q < n) ? keys[q] : 0;
  }
#pragma unroll
  for (int r = 0; r < RS_ROUNDS; r++) {
    const int64_t q = base + r * RS_THREADS + threadIdx.x;
    const bool valid = q < n;
    const unsigned d = valid ? (unsigned)((k[r] >> shift) & 0xff) : (0x100u + (unsigned)lane);
    const unsigned peers = warp_match_digit(d);
    if (valid && lane == __ffs(peers) - 1) atomicAdd(&h[d], __popc(peers));
  }
  __syncthreads();
  const int c = h[threadIdx.x];
  hist[(int64_t)threadIdx.x * nblocks + blockIdx.x] = c;  // digit-major table
  if (c) atomicAdd(&gtot[threadIdx.x], c);                // integer adds: order does not matter
}

// CTA d turns row d of the table (counts of digit d per CTA of the sort) into exclusive offsets
// within the digit: offs[d][b] = sum_{b' < b} hist[d][b'].
__global__ void __launch_bounds__(RS_THREADS)
rs_rowscan_kernel(int *__restrict__ hist, int nblocks) {
  __shared__ int wtot[RS_WARPS];
  int *row = hist + (int64_t)blockIdx.x * nblocks;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int carry = 0;
  for (int b0 = 0; b0 < nblocks; b0 += RS_THREADS) {
    const int b = b0 + threadIdx.x;
    const int v = (b < nblocks) ? row[b] : 0;
    const int incl = warp_scan<int>(v, lane);
    if (lane == 31) wtot[w] = incl;
    __syncthreads();
    int woff = 0, tot = 0;
#pragma unroll
    for (int k = 0; k < RS_WARPS; k++) {
      woff += (k < w) ? wtot[k] : 0;
      tot += wtot[k];
    }
    if (b < nblocks) row[b] = carry + woff + incl - v;
    carry += tot;
    __syncthreads();
  }
}

// GH_RS_VARIANT / GH_RS_MINBLOCKS: scatter-kernel alternatives measured with
// scripts/build_variants.py + scripts/gpu_variants2.sh (see DESIGN 4.3 for the outcome).
#ifndef GH_RS_VARIANT
#define GH_RS_VARIANT 0
#endif
#ifdef GH_RS_MINBLOCKS
#define GH_RS_BOUNDS __launch_bounds__(RS_THREADS, GH_RS_MINBLOCKS)
#else
#define GH_RS_BOUNDS __launch_bounds__(RS_THREADS)
#endif
__global__ void GH_RS_BOUNDS
rs_scatter_kernel(const uint64_t *__restrict__ kin, const int *__restrict__ vin,
                  uint64_t *__restrict__ kout, int *__restrict__ vout, int64_t n, int shift,
                  const int *__restrict__ offs /* per-digit exclusive offsets (rs_rowscan_kernel) */,
                  const int *__restrict__ gtot /* [256] keys per digit */, int nblocks,
                  const int *__restrict__ ndev = nullptr) {
  __shared__ int whist[RS_WARPS][RS_RADIX];  // per-warp digit counts, then per-warp digit bases
  __shared__ int dstart[RS_RADIX];           // start of digit d in the CTA-local sorted tile
  __shared__ int gbase[RS_RADIX];            // global start of this CTA's keys of digit d
  __shared__ int wtot[RS_WARPS];
  __shared__ uint64_t skey[RS_TILE];
  __shared__ int sval[RS_TILE];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int64_t tile0 = (int64_t)blockIdx.x * RS_TILE;
  const int64_t seg0 = tile0 + (int64_t)w * (32 * RS_ROUNDS);
  if (ndev) n = *ndev;
  if (tile0 >= n) return;  // whole CTA (capacity-sized grid)
#pragma unroll
  for (int k = 0; k < RS_WARPS; k++) whist[k][tid] = 0;
  __syncthreads();

  uint64_t key[RS_ROUNDS];
  int val[RS_ROUNDS], lrank[RS_ROUNDS];
#if GH_RS_RANK == 1
  unsigned peers[RS_ROUNDS], dg[RS_ROUNDS];
  int old[RS_ROUNDS];
#pragma unroll
  for (int r = 0; r < RS_ROUNDS; r++) {
    const int64_t q = seg0 + r * 32 + lane;
    const bool valid = q < n;
    key[r] = valid ? kin[q] : 0;
    val[r] = valid ? vin[q] : 0;
  }
#pragma unroll
  for (int r = 0; r < RS_ROUNDS; r++) {
    const bool valid = seg0 + r * 32 + lane < n;
    dg[r] = valid ? (unsigned)((key[r] >> shift) & 0xff) : (0x100u + (unsigned)lane);
    peers[r] = warp_match_digit(dg[r]);
  }
#pragma unroll
  for (int r = 0; r < RS_ROUNDS; r++) {
    const bool valid = seg0 + r * 32 + lane < n;
    old[r] = 0;
    if (valid && lane == __ffs(peers[r]) - 1) old[r] = atomicAdd(&whist[w][dg[r]], __popc(peers[r]));
    __syncwarp();
  }
#pragma unroll
  for (int r = 0; r < RS_ROUNDS; r++) {
    const int base = __shfl_sync(0xffffffffu, old[r], __ffs(peers[r]) - 1);
    lrank[r] = base + __popc(peers[r] & ((1u << lane) - 1u));
  }
#elif GH_RS_VARIANT == 0
  // one loop: load, ballot, update the running per-digit count, round by round
#pragma unroll
  for (int r = 0; r < RS_ROUNDS; r++) {
    const int64_t q = seg0 + r * 32 + lane;
    const bool valid = q < n;
    key[r] = valid ? kin[q] : 0;
    val[r] = valid ? vin[q] : 0;
    // invalid lanes get private pseudo-digits so that they match nobody
    const unsigned d = valid ? (unsigned)((key[r] >> shift) & 0xff) : (0x100u + (unsigned)lane);
    const unsigned peers = warp_match_digit(d);
    const int leader = __ffs(peers) - 1;
    int old = 0;
    if (lane == leader && valid) {
      old = whist[w][d];
      whist[w][d] = old + __popc(peers);
    }
    old = __shfl_sync(0xffffffffu, old, leader);
    lrank[r] = old + __popc(peers & ((1u << lane) - 1u));
    __syncwarp();
  }
#else
  // all key loads and all MATCH.ANY ballots first: they are independent of each other, only the
  // running per-digit counts have to be updated round by round (stability).  ncu showed the warps
  // waiting on one MATCH / one load at a time when the loops were one.  Variant 2 loads the values
  // only after the ranking (they are not needed before the tile is placed in shared memory).
  unsigned peers[RS_ROUNDS];
#pragma unroll
  for (int r = 0; r < RS_ROUNDS; r++) {
    const int64_t q = seg0 + r * 32 + lane;
    key[r] = (q < n) ? kin[q] : 0;
#if GH_RS_VARIANT == 1
    val[r] = (q < n) ? vin[q] : 0;
#endif
  }
#pragma unroll
  for (int r = 0; r < RS_ROUNDS; r++) {
    const bool valid = seg0 + r * 32 + lane < n;
    // invalid lanes get private pseudo-digits so that they match nobody
    const unsigned d = valid ? (unsigned)((key[r] >> shift) & 0xff) : (0x100u + (unsigned)lane);
    peers[r] = warp_match_digit(d);
  }
#if GH_RS_VARIANT == 2
#pragma unroll
  for (int r = 0; r < RS_ROUNDS; r++) {
    const int64_t q = seg0 + r * 32 + lane;
    val[r] = (q < n) ? vin[q] : 0;
  }
#endif
#pragma unroll
  for (int r = 0; r < RS_ROUNDS; r++) {
    const bool valid = seg0 + r * 32 + lane < n;
    const unsigned d = (unsigned)((key[r] >> shift) & 0xff);
    const int leader = __ffs(peers[r]) - 1;
    int old = 0;
    if (lane == leader && valid) {
      old = whist[w][d];
      whist[w][d] = old + __popc(peers[r]);
    }
    old = __shfl_sync(0xffffffffu, old, leader);
    lrank[r] = old + __popc(peers[r] & ((1u << lane) - 1u));
    __syncwarp();
  }
#endif
  __syncthreads();

  // thread d: turn the per-warp counts of digit d into per-warp bases, get the CTA total
  int tot = 0;
#pragma unroll
  for (int k = 0; k < RS_WARPS; k++) {
    const int c = whist[k][tid];
    whist[k][tid] = tot;
    tot += c;
  }
  // exclusive scan of the 256 digit totals across the CTA
  int incl = warp_scan<int>(tot, lane);
  if (lane == 31) wtot[w] = incl;
  __syncthreads();
  int woff = 0;
#pragma unroll
  for (int k = 0; k < RS_WARPS; k++) woff += (k < w) ? wtot[k] : 0;
  dstart[tid] = woff + incl - tot;
  __syncthreads();
  // global start of digit d = number of keys with a smaller digit (scan of gtot across the CTA)
  {
    const int g = gtot[tid];
    const int gincl = warp_scan<int>(g, lane);
    if (lane == 31) wtot[w] = gincl;
    __syncthreads();
    int goff = 0;
#pragma unroll
    for (int k = 0; k < RS_WARPS; k++) goff += (k < w) ? wtot[k] : 0;
    gbase[tid] = goff + gincl - g + offs[(int64_t)tid * nblocks + blockIdx.x];
  }
  __syncthreads();

#pragma unroll
  for (int r = 0; r < RS_ROUNDS; r++) {
    const int64_t q = seg0 + r * 32 + lane;
    if (q < n) {
      const int d = (int)((key[r] >> shift) & 0xff);
      const int pos = dstart[d] + whist[w][d] + lrank[r];
      skey[pos] = key[r];
      sval[pos] = val[r];
    }
  }
  __syncthreads();
  const int64_t left = n - tile0;
  const int ntile = (int)(left < RS_TILE ? left : RS_TILE);
#pragma unroll
  for (int k = 0; k < RS_ROUNDS; k++) {
    const int j = k * RS_THREADS + tid;
    if (j < ntile) {
      const uint64_t kk = skey[j];
      const int d = (int)((kk >> shift) & 0xff);
      const int64_t g = (int64_t)gbase[d] + (j - dstart[d]);
      kout[g] = kk;
      vout[g] = sval[j];
    }
  }
}

#ifndef GH_HOST_EMU  // launch sequences: the host emulation (tests/emu) has its own
struct RadixScratch {
  DeviceBuffer hist, gtot;
  void release() {
    hist.release();
    gtot.release();
  }
};

// Stable LSD sort of (key, value) pairs on key bits [0, nbits).  Ping-pongs between (kA, vA) and
// (kB, vB), starting from A; *result_in_B says where the sorted data ends up (both are clobbered).
// `ndev` (nullable): real pair count on the device, n = capacity the grids are sized for.
static int radix_sort_pairs(uint64_t *kA, int *vA, uint64_t *kB, int *vB, int64_t n, int nbits,
                            RadixScratch &rs, cudaStream_t st, bool *result_in_B, const int *ndev = nullptr) {
  *result_in_B = false;
  if (n <= 1) return GH_OK;
  const int nblocks = (int)((n + RS_TILE - 1) / RS_TILE);
  const int64_t tbl = (int64_t)RS_RADIX * nblocks;
  const int npass = (nbits + 7) / 8;
  GH_TRY(rs.hist.reserve(sizeof(int) * (size_t)tbl));
  GH_TRY(rs.gtot.reserve(sizeof(int) * RS_RADIX * (size_t)npass));
  GH_CUDA(cudaMemsetAsync(rs.gtot.ptr, 0, sizeof(int) * RS_RADIX * (size_t)npass, st));
  uint64_t *kin = kA, *kout = kB;
  int *vin = vA, *vout = vB;
  bool inB = false;
  for (int pass = 0; pass < npass; pass++) {
    const int shift = 8 * pass;
    int *gtot = rs.gtot.as<int>() + RS_RADIX * pass;
    rs_hist_kernel<<<nblocks, RS_THREADS, 0, st>>>(kin, n, shift, rs.hist.as<int>(), nblocks, gtot, ndev);
    GH_LAUNCH_CHECK();
    rs_rowscan_kernel<<<RS_RADIX, RS_THREADS, 0, st>>>(rs.hist.as<int>(), nblocks);
    GH_LAUNCH_CHECK();
    rs_scatter_kernel<<<nblocks, RS_THREADS, 0, st>>>(kin, vin, kout, vout, n, shift, rs.hist.as<int>(), gtot,
                                                     nblocks, ndev);
    GH_LAUNCH_CHECK();
    uint64_t *tk = kin; kin = kout; kout = tk;
    int *tv = vin; vin = vout; vout = tv;
    inB = !inB;
  }
  *result_in_B = inB;
  return GH_OK;
}

#endif  // GH_HOST_EMU
}  // namespace gh
