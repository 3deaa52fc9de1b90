// bucketsort.cuh -- "splitter sort" of the tree build's (Morton key, index) pairs: the sort for the
// steps of a RUNNING simulation, where the key distribution of step n is known from step n-1.
//
// The classic LSD radix sort (sortscan.cuh) moves every pair through global memory eight times
// (63-bit keys, 8 bits per pass, three kernels per pass).  Here the key space is cut into B buckets
// of ~1365 pairs (8/9 of a 1536-pair tile) by splitters taken from the previous step's sorted keys (every (n/B)-th key), and
//   1. two stable partition passes (the classic pass kernels with the digit = low / high byte of
//      the BUCKET id, found by binary search over the splitters) bring every pair into its bucket;
//   2. one kernel sorts every bucket inside shared memory (one CTA per bucket: the same stable
//      tile ranking as the scatter kernel, 8 bits per pass, only over the bits in which the
//      bucket's bounds differ), so the pairs cross global memory 3 times instead of 8.
// Buckets that outgrew the tile (1536 pairs; the splitters are refreshed every step, so this takes
// a violent change of the system within one step) are sorted by their CTA with the classic pass
// structure over global memory, tile after tile: slower, never wrong.  The result is the classic
// sort's, bit for bit (both are stable).  Without valid splitters (first step, stateless calls)
// the build uses the classic sort.
//
// Three forms of step 1 live here (GH_SORT selects; DESIGN 4.3 has the measurements):
//   bucket  (bs_*)   the two partition passes described above;
//   place   (bp_*)   bucket id once per key, counts and slots by GLOBAL atomics -- measured no faster;
//   place2  (bp2_*)  the same through per-CTA shared-memory histograms, no global atomics: the default.
// The place forms leave the order inside a bucket arbitrary; their bucket kernel (bp_bucket_kernel)
// orders by (key, value) instead of relying on stability, ranking only the top bits that can differ.
#pragma once
#include "sortscan.cuh"

namespace gh {

// pairs per bucket the splitters aim at: 8/9 of a tile.  The bucket kernel's cost is per CTA-pass, not
// per pair, so fuller buckets are cheaper (measured: 2/3 of a tile +0.12 ms of build at N = 4M); the
// Poisson fluctuation of a bucket's count from one step to the next is ~3 %
static constexpr int BS_TARGET = (8 * RS_TILE) / 9;
static constexpr int BS_MIN_BUCKETS = 257;  // below this the classic sort is used (launch bound anyway)
static constexpr int BS_MAX_BUCKETS = 65536;

// bucket of a key: the largest b with spl[b] <= key (spl[0] = 0, so every key has one)
struct BucketOf {
  const uint64_t *spl;
  int nb;
  __device__ __forceinline__ int operator()(uint64_t k) const {
    int lo = 0, hi = nb - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (spl[mid] <= k) lo = mid; else hi = mid - 1;
    }
    return lo;
  }
};
struct BucketDigit {  // digit of a partition pass
  BucketOf b;
  int shift;
  __device__ __forceinline__ unsigned operator()(uint64_t k) const { return (unsigned)((b(k) >> shift) & 0xff); }
};
struct BitsDigit {    // digit of a classic pass
  int shift;
  __device__ __forceinline__ unsigned operator()(uint64_t k) const { return (unsigned)((k >> shift) & 0xff); }
};

struct TileSmem {
  int whist[RS_WARPS][RS_RADIX];  // per-warp digit counts, then per-warp digit bases
  int dstart[RS_RADIX];           // start of digit d in the tile's sorted order
  int gbase[RS_RADIX];            // (global passes) running start of digit d in the output
  int wtot[RS_WARPS];
  uint64_t skey[RS_TILE];
  int sval[RS_TILE];
};

// Stable ranking of one tile (<= RS_TILE pairs, thread t / round r holds pair w*256 + r*32 + lane)
// by digit.  On return lrank[r] is the pair's rank among the pairs of its digit inside its warp,
// sm.whist[w][d] the number of pairs of digit d in warps < w, sm.dstart[d] the start of digit d in
// the sorted tile; the return value is the tile's count of digit threadIdx.x.  (The body of
// rs_scatter_kernel, shared here between the partition passes and the in-shared-memory sort.)
// tile_rank_digits: the digits are the caller's (dig[r] of an invalid pair = 0x100 + lane, a private
// pseudo-digit that matches nobody); tile_rank computes them from the keys.
__device__ __forceinline__ int tile_rank_digits(TileSmem &sm, int count, int (&lrank)[RS_ROUNDS],
                                                const unsigned (&dig)[RS_ROUNDS]) {
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
#pragma unroll
  for (int k = 0; k < RS_WARPS; k++) sm.whist[k][tid] = 0;
  __syncthreads();
#if GH_RS_RANK == 1
  // all MATCH.ANY ballots first (independent of each other), then one shared-memory atomic per
  // round by the group's leader: the atomic returns the running count, nothing later in the chain
  // waits for it, so the rounds pipeline (ncu: the leader's load-add-store chain was the
  // short-scoreboard stall of the scatter pass).  __syncwarp orders the rounds' atomics.
  unsigned peers[RS_ROUNDS];
  int old[RS_ROUNDS];
#pragma unroll
  for (int r = 0; r < RS_ROUNDS; r++) {
    peers[r] = warp_match_digit(dig[r]);
  }
#pragma unroll
  for (int r = 0; r < RS_ROUNDS; r++) {
    const bool valid = w * (32 * RS_ROUNDS) + r * 32 + lane < count;
    old[r] = 0;
    if (valid && lane == __ffs(peers[r]) - 1) old[r] = atomicAdd(&sm.whist[w][dig[r]], __popc(peers[r]));
    __syncwarp();
  }
#pragma unroll
  for (int r = 0; r < RS_ROUNDS; r++) {
    const int base = __shfl_sync(0xffffffffu, old[r], __ffs(peers[r]) - 1);
    lrank[r] = base + __popc(peers[r] & ((1u << lane) - 1u));
  }
#else
#pragma unroll
  for (int r = 0; r < RS_ROUNDS; r++) {
    const bool valid = w * (32 * RS_ROUNDS) + r * 32 + lane < count;
    const unsigned d = dig[r];
    const unsigned peers = warp_match_digit(d);
    const int leader = __ffs(peers) - 1;
    int old = 0;
    if (lane == leader && valid) {
      old = sm.whist[w][d];
      sm.whist[w][d] = old + __popc(peers);
    }
    old = __shfl_sync(0xffffffffu, old, leader);
    lrank[r] = old + __popc(peers & ((1u << lane) - 1u));
    __syncwarp();
  }
#endif
  __syncthreads();
  int tot = 0;
#pragma unroll
  for (int k = 0; k < RS_WARPS; k++) {
    const int c = sm.whist[k][tid];
    sm.whist[k][tid] = tot;
    tot += c;
  }
  const int incl = warp_scan<int>(tot, lane);
  if (lane == 31) sm.wtot[w] = incl;
  __syncthreads();
  int woff = 0;
#pragma unroll
  for (int k = 0; k < RS_WARPS; k++) woff += (k < w) ? sm.wtot[k] : 0;
  sm.dstart[tid] = woff + incl - tot;
  __syncthreads();
  return tot;
}
template <class Digit>
__device__ __forceinline__ int tile_rank(TileSmem &sm, const uint64_t (&key)[RS_ROUNDS], int count, Digit dg,
                                         int (&lrank)[RS_ROUNDS], unsigned (&dig)[RS_ROUNDS]) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int r = 0; r < RS_ROUNDS; r++) {
    const bool valid = w * (32 * RS_ROUNDS) + r * 32 + lane < count;
    // invalid lanes get private pseudo-digits so that they match nobody
    dig[r] = valid ? dg(key[r]) : (0x100u + (unsigned)lane);
  }
  return tile_rank_digits(sm, count, lrank, dig);
}

// place the tile's pairs into sm.skey / sm.sval in digit order (after tile_rank)
__device__ __forceinline__ void tile_place(TileSmem &sm, const uint64_t (&key)[RS_ROUNDS], const int (&val)[RS_ROUNDS],
                                           int count, const int (&lrank)[RS_ROUNDS], const unsigned (&dig)[RS_ROUNDS]) {
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
#pragma unroll
  for (int r = 0; r < RS_ROUNDS; r++) {
    if (w * (32 * RS_ROUNDS) + r * 32 + lane < count) {
      const int d = (int)dig[r];
      const int pos = sm.dstart[d] + sm.whist[w][d] + lrank[r];
      sm.skey[pos] = key[r];
      sm.sval[pos] = val[r];
    }
  }
  __syncthreads();
}

// ---- partition passes: the classic three kernels with the bucket id's bytes as digits --------------
__global__ void __launch_bounds__(RS_THREADS)
bs_hist_kernel(const uint64_t *__restrict__ keys, int64_t n, BucketDigit dg, int *__restrict__ hist, int nblocks,
               int *__restrict__ gtot /* [256], zeroed */, const int *__restrict__ ndev) {
  __shared__ int h[RS_RADIX];
  if (ndev) n = *ndev;
  h[threadIdx.x] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * RS_TILE;
  const int lane = threadIdx.x & 31;
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
    const unsigned d = valid ? dg(k[r]) : (0x100u + (unsigned)lane);
    const unsigned peers = warp_match_digit(d);
    if (valid && lane == __ffs(peers) - 1) atomicAdd(&h[d], __popc(peers));
  }
  __syncthreads();
  const int c = h[threadIdx.x];
  hist[(int64_t)threadIdx.x * nblocks + blockIdx.x] = c;
  if (c) atomicAdd(&gtot[threadIdx.x], c);
}

__global__ void __launch_bounds__(RS_THREADS)
bs_scatter_kernel(const uint64_t *__restrict__ kin, const int *__restrict__ vin, uint64_t *__restrict__ kout,
                  int *__restrict__ vout, int64_t n, BucketDigit dg, const int *__restrict__ offs,
                  const int *__restrict__ gtot, int nblocks, const int *__restrict__ ndev) {
  __shared__ TileSmem sm;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int64_t tile0 = (int64_t)blockIdx.x * RS_TILE;
  if (ndev) n = *ndev;
  if (tile0 >= n) return;
  const int64_t left = n - tile0;
  const int count = (int)(left < RS_TILE ? left : RS_TILE);
  uint64_t key[RS_ROUNDS];
  int val[RS_ROUNDS], lrank[RS_ROUNDS];
  unsigned dig[RS_ROUNDS];
#pragma unroll
  for (int r = 0; r < RS_ROUNDS; r++) {
    const int j = w * (32 * RS_ROUNDS) + r * 32 + lane;
    key[r] = (j < count) ? kin[tile0 + j] : 0;
    val[r] = (j < count) ? vin[tile0 + j] : 0;
  }
  tile_rank(sm, key, count, dg, lrank, dig);
  {  // global start of digit d = pairs with a smaller digit + this CTA's offset inside the digit
    const int g = gtot[tid];
    const int gincl = warp_scan<int>(g, lane);
    if (lane == 31) sm.wtot[w] = gincl;
    __syncthreads();
    int goff = 0;
#pragma unroll
    for (int k = 0; k < RS_WARPS; k++) goff += (k < w) ? sm.wtot[k] : 0;
    sm.gbase[tid] = goff + gincl - g + offs[(int64_t)tid * nblocks + blockIdx.x];
  }
  __syncthreads();
  tile_place(sm, key, val, count, lrank, dig);
#pragma unroll
  for (int k = 0; k < RS_ROUNDS; k++) {
    const int j = k * RS_THREADS + tid;
    if (j < count) {
      const uint64_t kk = sm.skey[j];
      const int d = (int)dg(kk);
      const int64_t g = (int64_t)sm.gbase[d] + (j - sm.dstart[d]);
      kout[g] = kk;
      vout[g] = sm.sval[j];
    }
  }
}

// boff[b] = first position of the partitioned array whose bucket is >= b (b = 0 .. nb; boff[nb] = n)
__global__ void bs_offsets_kernel(const uint64_t *__restrict__ keys, int64_t n, BucketOf bo, int *__restrict__ boff,
                                  const int *__restrict__ ndev) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (ndev) n = *ndev;
  if (b > bo.nb) return;
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (bo(keys[mid]) < b) lo = mid + 1; else hi = mid;
  }
  boff[b] = (int)lo;
}

// ---- the buckets: one CTA each ---------------------------------------------------------------------
// keys/vals: the partitioned array (sorted in place); kscr/vscr: scratch of the same size (only
// the oversize path uses it).  Bits that are equal in the bucket's bounds need no pass.
__global__ void __launch_bounds__(RS_THREADS)
bs_bucket_kernel(uint64_t *__restrict__ keys, int *__restrict__ vals, uint64_t *__restrict__ kscr,
                 int *__restrict__ vscr, const int *__restrict__ boff, const uint64_t *__restrict__ spl, int nb,
                 int nbits) {
  __shared__ TileSmem sm;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int b = blockIdx.x;
  const int64_t s0 = boff[b], s1 = boff[b + 1];
  const int64_t size = s1 - s0;
  if (size <= 1) return;
  // bits in which two keys of this bucket can differ
  const uint64_t klo = spl[b], khi = (b + 1 < nb) ? spl[b + 1] - 1 : ~0ull;
  int topbit = 64 - __clzll((long long)(klo ^ khi));  // 0 when all keys are equal
  if (topbit > nbits) topbit = nbits;
  const int npass = (topbit + 7) / 8;
  if (npass == 0) return;
  uint64_t key[RS_ROUNDS];
  int val[RS_ROUNDS], lrank[RS_ROUNDS];
  unsigned dig[RS_ROUNDS];
  if (size <= RS_TILE) {
    const int count = (int)size;
#pragma unroll
    for (int r = 0; r < RS_ROUNDS; r++) {
      const int j = w * (32 * RS_ROUNDS) + r * 32 + lane;
      key[r] = (j < count) ? keys[s0 + j] : 0;
      val[r] = (j < count) ? vals[s0 + j] : 0;
    }
    for (int pass = 0; pass < npass; pass++) {
      tile_rank(sm, key, count, BitsDigit{8 * pass}, lrank, dig);
      tile_place(sm, key, val, count, lrank, dig);
      if (pass + 1 < npass) {
#pragma unroll
        for (int r = 0; r < RS_ROUNDS; r++) {
          const int j = w * (32 * RS_ROUNDS) + r * 32 + lane;
          if (j < count) { key[r] = sm.skey[j]; val[r] = sm.sval[j]; }
        }
        __syncthreads();
      }
    }
#pragma unroll
    for (int k = 0; k < RS_ROUNDS; k++) {
      const int j = k * RS_THREADS + tid;
      if (j < count) { keys[s0 + j] = sm.skey[j]; vals[s0 + j] = sm.sval[j]; }
    }
    return;
  }
  // oversize bucket: the classic pass structure over global memory, by this CTA alone.  An even
  // number of passes, so that the result ends where it started.
  __shared__ int ghist[RS_RADIX];
  const int gp = (npass + 1) & ~1;
  uint64_t *kin = keys + s0, *kout = kscr + s0;
  int *vin = vals + s0, *vout = vscr + s0;
  const int ntiles = (int)((size + RS_TILE - 1) / RS_TILE);
  for (int pass = 0; pass < gp; pass++) {
    const BitsDigit dg{8 * pass};
    ghist[tid] = 0;
    __syncthreads();
    for (int64_t q = tid; q < size; q += RS_THREADS) atomicAdd(&ghist[dg(kin[q])], 1);
    __syncthreads();
    {  // exclusive scan of the digit counts -> running output offsets
      const int g = ghist[tid];
      const int gincl = warp_scan<int>(g, lane);
      if (lane == 31) sm.wtot[w] = gincl;
      __syncthreads();
      int goff = 0;
#pragma unroll
      for (int k = 0; k < RS_WARPS; k++) goff += (k < w) ? sm.wtot[k] : 0;
      __syncthreads();
      ghist[tid] = goff + gincl - g;
    }
    __syncthreads();
    for (int t = 0; t < ntiles; t++) {
      const int64_t t0 = (int64_t)t * RS_TILE;
      const int count = (int)((size - t0) < RS_TILE ? (size - t0) : RS_TILE);
#pragma unroll
      for (int r = 0; r < RS_ROUNDS; r++) {
        const int j = w * (32 * RS_ROUNDS) + r * 32 + lane;
        key[r] = (j < count) ? kin[t0 + j] : 0;
        val[r] = (j < count) ? vin[t0 + j] : 0;
      }
      const int tot = tile_rank(sm, key, count, dg, lrank, dig);
      tile_place(sm, key, val, count, lrank, dig);
#pragma unroll
      for (int k = 0; k < RS_ROUNDS; k++) {
        const int j = k * RS_THREADS + tid;
        if (j < count) {
          const uint64_t kk = sm.skey[j];
          const int d = (int)dg(kk);
          const int64_t g = (int64_t)ghist[d] + (j - sm.dstart[d]);
          kout[g] = kk;
          vout[g] = sm.sval[j];
        }
      }
      __syncthreads();
      ghist[tid] += tot;  // this tile's pairs of digit tid are placed
      __syncthreads();
    }
    uint64_t *tk = kin; kin = kout; kout = tk;
    int *tv = vin; vin = vout; vout = tv;
    __threadfence_block();
    __syncthreads();
  }
}

// ---- "place": the splitter sort with ONE trip into the buckets ---------------------------------------
// The two partition passes above cost 2 x (histogram + row scan + ranked scatter) + the offsets
// search, and every kernel of them repeats the 12-step splitter search per key.  Here:
//   bp_count_kernel   bucket id of every key, ONCE (coarse splitter table in shared memory + one
//                     128-byte line of the full table), stored as 16 bits; counts per bucket with
//                     global reductions (no return value);
//   bp_scan_kernel    exclusive scan of the counts = bucket offsets and running cursors (one CTA);
//   bp_place_kernel   every pair goes straight to cursor[bucket]++ (global atomic with return);
//   bp_bucket_kernel  one CTA per bucket sorts it inside shared memory.
// The atomics make the order INSIDE a bucket arbitrary, so the bucket kernel cannot lean on
// stability; it does not need to.  It ranks by the top BP_SUBBITS bits of (key - bucket start) that can
// differ inside the bucket (<= 3 passes of 8 bits instead of up to 8 over the whole key), then puts the
// rare runs of pairs that tie in those bits into (key, value) order by insertion (one thread per run).
// A bucket with a run longer than BP_MAX_RUN (many particles in a tiny corner of the bucket's key
// range, or coincident ones) is sorted again the long way: LSD passes over the value bits, then over
// every key bit that can differ.  Either way the result is ordered by (key, value): exactly the
// stable sort of pairs whose values ascend in the input, which is what the build feeds it
// (keys_kernel: value = particle index).  Oversize buckets: the global-memory pass structure of
// bs_bucket_kernel with the value passes in front.
static constexpr int BP_THREADS = 256;
static constexpr int BP_ROUNDS = 8;      // pairs per thread in the count / place kernels
static constexpr int BP_COARSE = 16;     // every 16th splitter is staged in shared memory (16 x 8 B = one line)
#ifndef GH_BP_SUBBITS
#define GH_BP_SUBBITS 24
#endif
static constexpr int BP_SUBBITS = GH_BP_SUBBITS;
static constexpr int BP_MAX_RUN = 16;

__device__ __forceinline__ int bp_bucket_of(const uint64_t *coarse, int nc, const uint64_t *__restrict__ spl, int nb,
                                            uint64_t k) {
  int lo = 0, hi = nc - 1;  // largest j with coarse[j] <= k (coarse[0] = spl[0] = 0)
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (coarse[mid] <= k) lo = mid; else hi = mid - 1;
  }
  int b0 = lo * BP_COARSE, b1 = b0 + BP_COARSE - 1;
  if (b1 > nb - 1) b1 = nb - 1;
  while (b0 < b1) {  // largest b in the line with spl[b] <= k
    const int mid = (b0 + b1 + 1) >> 1;
    if (spl[mid] <= k) b0 = mid; else b1 = mid - 1;
  }
  return b0;
}

__global__ void __launch_bounds__(BP_THREADS)
bp_count_kernel(const uint64_t *__restrict__ keys, int64_t n, const uint64_t *__restrict__ spl, int nb,
                unsigned short *__restrict__ bid, int *__restrict__ count /* [nb], zeroed */,
                const int *__restrict__ ndev) {
  __shared__ uint64_t coarse[BS_MAX_BUCKETS / BP_COARSE];
  if (ndev) n = *ndev;
  const int nc = (nb + BP_COARSE - 1) / BP_COARSE;
  for (int j = threadIdx.x; j < nc; j += BP_THREADS) coarse[j] = spl[j * BP_COARSE];
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * (BP_THREADS * BP_ROUNDS);
  uint64_t k[BP_ROUNDS];
#pragma unroll
  for (int r = 0; r < BP_ROUNDS; r++) {
    const int64_t q = base + r * BP_THREADS + threadIdx.x;
    k[r] = (q < n) ? keys[q] : 0;
  }
#pragma unroll
  for (int r = 0; r < BP_ROUNDS; r++) {
    const int64_t q = base + r * BP_THREADS + threadIdx.x;
    if (q < n) {
      const int b = bp_bucket_of(coarse, nc, spl, nb, k[r]);
      bid[q] = (unsigned short)b;
      atomicAdd(&count[b], 1);
    }
  }
}

// boff[b] = pairs in buckets < b (b = 0 .. nb), cursor[b] = boff[b].  One CTA.
__global__ void __launch_bounds__(RS_THREADS)
bp_scan_kernel(const int *__restrict__ count, int nb, int *__restrict__ boff, int *__restrict__ cursor) {
  __shared__ int wtot[RS_WARPS];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int carry = 0;
  for (int b0 = 0; b0 < nb; b0 += RS_THREADS) {
    const int b = b0 + threadIdx.x;
    const int v = (b < nb) ? count[b] : 0;
    const int incl = warp_scan<int>(v, lane);
    if (lane == 31) wtot[w] = incl;
    __syncthreads();
    int woff = 0, tot = 0;
#pragma unroll
    for (int k = 0; k < RS_WARPS; k++) {
      woff += (k < w) ? wtot[k] : 0;
      tot += wtot[k];
    }
    if (b < nb) { const int e = carry + woff + incl - v; boff[b] = e; cursor[b] = e; }
    carry += tot;
    __syncthreads();
  }
  if (threadIdx.x == 0) boff[nb] = carry;
}

__global__ void __launch_bounds__(BP_THREADS)
bp_place_kernel(const uint64_t *__restrict__ kin, const int *__restrict__ vin, const unsigned short *__restrict__ bid,
                int64_t n, int *__restrict__ cursor, uint64_t *__restrict__ kout, int *__restrict__ vout,
                const int *__restrict__ ndev) {
  if (ndev) n = *ndev;
  const int64_t base = (int64_t)blockIdx.x * (BP_THREADS * BP_ROUNDS);
  uint64_t k[BP_ROUNDS];
  int v[BP_ROUNDS], pos[BP_ROUNDS];
#pragma unroll
  for (int r = 0; r < BP_ROUNDS; r++) {
    const int64_t q = base + r * BP_THREADS + threadIdx.x;
    pos[r] = -1;
    if (q < n) {
      k[r] = kin[q];
      v[r] = vin[q];
      pos[r] = atomicAdd(&cursor[bid[q]], 1);
    }
  }
#pragma unroll
  for (int r = 0; r < BP_ROUNDS; r++)
    if (pos[r] >= 0) { kout[pos[r]] = k[r]; vout[pos[r]] = v[r]; }
}

// ---- "place", shared-memory form --------------------------------------------------------------------
// Measured on B200 (N = 4M, 3072 buckets): the global reductions of bp_count_kernel and the global
// atomics of bp_place_kernel run at ~18 G/s -- 1365 updates per address -- and cost 230 + 250 us,
// no better than the two partition passes they replace.  Same idea with the contention moved into
// shared memory: G CTAs each own a contiguous chunk of the input, count their chunk's pairs per
// bucket in a shared-memory histogram (bp2_count_kernel -> one row of a G x nb table), a column
// scan turns the table into every CTA's first slot inside every bucket (bp2_colscan_kernel +
// bp_scan_kernel for the bucket offsets), and bp2_place_kernel hands out slots with shared-memory
// atomics on its own row.  No global atomic at all.  Needs nb <= BP2_MAX_NB (histogram + coarse
// splitters in 48 KB of static shared memory): N <= 14M at 1365 pairs per bucket; beyond that the
// partition passes are used.
static constexpr int BP2_THREADS = 512;
static constexpr int BP2_MAX_NB = 10240;
static constexpr int BP2_UNROLL = 8;

// pairs per CTA: a multiple of the CTA size, G chunks cover n
static inline __host__ __device__ int64_t bp2_chunk(int64_t n, int G) {
  const int64_t c = (n + G - 1) / G;
  return ((c + BP2_THREADS - 1) / BP2_THREADS) * BP2_THREADS;
}

__global__ void __launch_bounds__(BP2_THREADS)
bp2_count_kernel(const uint64_t *__restrict__ keys, int64_t n, const uint64_t *__restrict__ spl, int nb,
                 unsigned short *__restrict__ bid, int *__restrict__ ghist /* [gridDim.x][nb] */,
                 const int *__restrict__ ndev) {
  __shared__ uint64_t coarse[BP2_MAX_NB / BP_COARSE];
  __shared__ int hist[BP2_MAX_NB];
  const int64_t chunk = bp2_chunk(n, (int)gridDim.x);  // from the capacity: the same on every launch of a step
  if (ndev) n = *ndev;
  const int nc = (nb + BP_COARSE - 1) / BP_COARSE;
  for (int j = threadIdx.x; j < nc; j += BP2_THREADS) coarse[j] = spl[j * BP_COARSE];
  for (int j = threadIdx.x; j < nb; j += BP2_THREADS) hist[j] = 0;
  __syncthreads();
  const int64_t q0 = (int64_t)blockIdx.x * chunk;
  int64_t q1 = q0 + chunk;
  if (q1 > n) q1 = n;
  for (int64_t qb = q0; qb < q1; qb += BP2_THREADS * BP2_UNROLL) {
    uint64_t k[BP2_UNROLL];
#pragma unroll
    for (int u = 0; u < BP2_UNROLL; u++) {
      const int64_t q = qb + u * BP2_THREADS + threadIdx.x;
      k[u] = (q < q1) ? keys[q] : 0;
    }
#pragma unroll
    for (int u = 0; u < BP2_UNROLL; u++) {
      const int64_t q = qb + u * BP2_THREADS + threadIdx.x;
      if (q < q1) {
        const int b = bp_bucket_of(coarse, nc, spl, nb, k[u]);
        bid[q] = (unsigned short)b;
        atomicAdd(&hist[b], 1);
      }
    }
  }
  __syncthreads();
  int *row = ghist + (int64_t)blockIdx.x * nb;
  for (int j = threadIdx.x; j < nb; j += BP2_THREADS) row[j] = hist[j];
}

// column b of the table: ghist[c][b] <- pairs of bucket b in CTAs < c; tot[b] = pairs of bucket b.
// A CTA of 256 threads owns 32 columns; warp w sums rows [w R, (w+1) R), R = ceil(G / 8) (lanes =
// columns: every row access is one 128-byte line), the eight partial sums are prefixed through
// shared memory, and the warp walks its rows again writing the running offsets.
__global__ void __launch_bounds__(256) bp2_colscan_kernel(int *__restrict__ ghist, int G, int nb, int *__restrict__ tot) {
  __shared__ int psum[8][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int b = blockIdx.x * 32 + lane;
  const int R = (G + 7) / 8;
  const int c0 = w * R, c1 = (c0 + R < G) ? c0 + R : G;
  int sum = 0;
  if (b < nb) {
    for (int cb = c0; cb < c1; cb += 8) {  // eight rows at a time: independent loads
      int v[8];
#pragma unroll
      for (int u = 0; u < 8; u++) v[u] = (cb + u < c1) ? ghist[(int64_t)(cb + u) * nb + b] : 0;
#pragma unroll
      for (int u = 0; u < 8; u++) sum += v[u];
    }
  }
  psum[w][lane] = sum;
  __syncthreads();
  int run = 0, total = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    const int s = psum[k][lane];
    run += (k < w) ? s : 0;
    total += s;
  }
  if (b < nb) {
    for (int cb = c0; cb < c1; cb += 8) {
      int v[8];
#pragma unroll
      for (int u = 0; u < 8; u++) v[u] = (cb + u < c1) ? ghist[(int64_t)(cb + u) * nb + b] : 0;
#pragma unroll
      for (int u = 0; u < 8; u++)
        if (cb + u < c1) {
          ghist[(int64_t)(cb + u) * nb + b] = run;
          run += v[u];
        }
    }
    if (w == 0) tot[b] = total;
  }
}

__global__ void __launch_bounds__(BP2_THREADS)
bp2_place_kernel(const uint64_t *__restrict__ kin, const int *__restrict__ vin, const unsigned short *__restrict__ bid,
                 int64_t n, const int *__restrict__ ghist, const int *__restrict__ boff, int nb,
                 uint64_t *__restrict__ kout, int *__restrict__ vout, const int *__restrict__ ndev) {
  __shared__ int cur[BP2_MAX_NB];
  const int64_t chunk = bp2_chunk(n, (int)gridDim.x);
  if (ndev) n = *ndev;
  const int *row = ghist + (int64_t)blockIdx.x * nb;
  for (int j = threadIdx.x; j < nb; j += BP2_THREADS) cur[j] = boff[j] + row[j];
  __syncthreads();
  const int64_t q0 = (int64_t)blockIdx.x * chunk;
  int64_t q1 = q0 + chunk;
  if (q1 > n) q1 = n;
  for (int64_t qb = q0; qb < q1; qb += BP2_THREADS * BP2_UNROLL) {
    uint64_t k[BP2_UNROLL];
    int v[BP2_UNROLL], b[BP2_UNROLL];
#pragma unroll
    for (int u = 0; u < BP2_UNROLL; u++) {
      const int64_t q = qb + u * BP2_THREADS + threadIdx.x;
      b[u] = -1;
      if (q < q1) { k[u] = kin[q]; v[u] = vin[q]; b[u] = (int)bid[q]; }
    }
#pragma unroll
    for (int u = 0; u < BP2_UNROLL; u++)
      if (b[u] >= 0) {
        const int pos = atomicAdd(&cur[b[u]], 1);
        kout[pos] = k[u];
        vout[pos] = v[u];
      }
  }
}

#ifdef GH_HOST_EMU
// which way the buckets went (tests/test_tree_emu.py): [0] compact ranking, [1] the long way,
// [2] oversize, [3] runs ordered by insertion
static long long g_bp_stats[4] = {0, 0, 0, 0};
#define GH_BP_STAT(k) __atomic_fetch_add(&g_bp_stats[k], 1ll, __ATOMIC_RELAXED)
#else
#define GH_BP_STAT(k) ((void)0)
#endif
struct SubDigit {  // digit of (key - bucket start)
  uint64_t klo;
  int shift;
  __device__ __forceinline__ unsigned operator()(uint64_t k) const {
    return shift < 64 ? (unsigned)(((k - klo) >> shift) & 0xff) : 0u;
  }
};
// k[0 .. count) is ordered by (k - klo) >> shift.  If a run of pairs that agree in those bits starts
// at j: its length (counting stops at cap + 1), else 0.
__device__ __forceinline__ int bp_run_length(const uint64_t *k, int count, int j, uint64_t klo, int shift, int cap) {
  const uint64_t s = (k[j] - klo) >> shift;
  if (j > 0 && ((k[j - 1] - klo) >> shift) == s) return 0;
  int len = 1;
  while (j + len < count && len <= cap && ((k[j + len] - klo) >> shift) == s) len++;
  return len;
}
// insertion sort of the run [j, j + len) by (key, value)
__device__ __forceinline__ void bp_sort_run(uint64_t *k, int *v, int j, int len) {
  for (int a = 1; a < len; a++) {
    const uint64_t ka = k[j + a];
    const int va = v[j + a];
    int t = a;
    while (t > 0 && (k[j + t - 1] > ka || (k[j + t - 1] == ka && v[j + t - 1] > va))) {
      k[j + t] = k[j + t - 1];
      v[j + t] = v[j + t - 1];
      t--;
    }
    k[j + t] = ka;
    v[j + t] = va;
  }
}

// keys/vals: the placed pairs (sorted in place); kscr/vscr: scratch of the same size (oversize path).
// nbits: key bits in use; vbits: bits of the largest value.
#ifndef GH_BP_MINBLOCKS
#define GH_BP_MINBLOCKS 4  // 64 registers: as many resident CTAs as bs_bucket_kernel (the ranking is latency bound)
#endif
__global__ void __launch_bounds__(RS_THREADS, GH_BP_MINBLOCKS)
bp_bucket_kernel(uint64_t *__restrict__ keys, int *__restrict__ vals, uint64_t *__restrict__ kscr,
                 int *__restrict__ vscr, const int *__restrict__ boff, const uint64_t *__restrict__ spl, int nb,
                 int nbits, int vbits) {
  __shared__ TileSmem sm;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int b = blockIdx.x;
  const int64_t s0 = boff[b], s1 = boff[b + 1];
  const int64_t size = s1 - s0;
  if (size <= 1) return;
  // (key - klo) of this bucket's pairs lies in [0, span]: `sbits` bits can differ
  const uint64_t klo = spl[b];
  const uint64_t kmax = (nbits >= 64) ? ~0ull : ((1ull << nbits) - 1ull);
  const uint64_t khi = (b + 1 < nb) ? spl[b + 1] - 1 : kmax;
  const uint64_t span = khi >= klo ? khi - klo : 0;
  const int sbits = 64 - __clzll((long long)span);  // 0: all keys are equal
  const int nvp = (vbits + 7) / 8;
  uint64_t key[RS_ROUNDS];
  int val[RS_ROUNDS], lrank[RS_ROUNDS];
  unsigned dig[RS_ROUNDS];
  if (size <= RS_TILE) {
    const int count = (int)size;
#pragma unroll
    for (int r = 0; r < RS_ROUNDS; r++) {
      const int j = w * (32 * RS_ROUNDS) + r * 32 + lane;
      key[r] = (j < count) ? keys[s0 + j] : 0;
      val[r] = (j < count) ? vals[s0 + j] : 0;
    }
    bool full = sbits == 0;
    if (!full) {
      const int shift0 = sbits > BP_SUBBITS ? sbits - BP_SUBBITS : 0;
      const int npass = (sbits - shift0 + 7) / 8;
      for (int pass = 0; pass < npass; pass++) {
        tile_rank(sm, key, count, SubDigit{klo, shift0 + 8 * pass}, lrank, dig);
        tile_place(sm, key, val, count, lrank, dig);
        if (pass + 1 < npass) {
#pragma unroll
          for (int r = 0; r < RS_ROUNDS; r++) {
            const int j = w * (32 * RS_ROUNDS) + r * 32 + lane;
            if (j < count) { key[r] = sm.skey[j]; val[r] = sm.sval[j]; }
          }
          __syncthreads();
        }
      }
      // runs that tie in the ranked bits: short ones are ordered here, a long one sends the bucket
      // the long way
      int ok = 1;
      int len[RS_ROUNDS];
#pragma unroll
      for (int k = 0; k < RS_ROUNDS; k++) {
        const int j = k * RS_THREADS + tid;
        len[k] = (j < count) ? bp_run_length(sm.skey, count, j, klo, shift0, BP_MAX_RUN) : 0;
        if (len[k] > BP_MAX_RUN) ok = 0;
      }
      full = !__syncthreads_and(ok);
      if (!full) {
#pragma unroll
        for (int k = 0; k < RS_ROUNDS; k++)
          if (len[k] >= 2) { GH_BP_STAT(3); bp_sort_run(sm.skey, sm.sval, k * RS_THREADS + tid, len[k]); }
        __syncthreads();
      }
      if (full) {
#pragma unroll
        for (int r = 0; r < RS_ROUNDS; r++) {
          const int j = w * (32 * RS_ROUNDS) + r * 32 + lane;
          if (j < count) { key[r] = sm.skey[j]; val[r] = sm.sval[j]; }
        }
        __syncthreads();
      }
    }
    if (tid == 0) GH_BP_STAT(full ? 1 : 0);
    if (full) {  // block-uniform: value passes, then every key bit that can differ
      const int nkp = (sbits + 7) / 8;
      for (int pass = 0; pass < nvp + nkp; pass++) {
#pragma unroll
        for (int r = 0; r < RS_ROUNDS; r++) {
          const bool valid = w * (32 * RS_ROUNDS) + r * 32 + lane < count;
          const unsigned d = pass < nvp ? (unsigned)((val[r] >> (8 * pass)) & 0xff)
                                        : SubDigit{klo, 8 * (pass - nvp)}(key[r]);
          dig[r] = valid ? d : (0x100u + (unsigned)lane);
        }
        tile_rank_digits(sm, count, lrank, dig);
        tile_place(sm, key, val, count, lrank, dig);
        if (pass + 1 < nvp + nkp) {
#pragma unroll
          for (int r = 0; r < RS_ROUNDS; r++) {
            const int j = w * (32 * RS_ROUNDS) + r * 32 + lane;
            if (j < count) { key[r] = sm.skey[j]; val[r] = sm.sval[j]; }
          }
          __syncthreads();
        }
      }
    }
#pragma unroll
    for (int k = 0; k < RS_ROUNDS; k++) {
      const int j = k * RS_THREADS + tid;
      if (j < count) { keys[s0 + j] = sm.skey[j]; vals[s0 + j] = sm.sval[j]; }
    }
    return;
  }
  // oversize bucket: stable passes over global memory, tile after tile, by this CTA alone -- first
  // over the value bits, then over the key bits that can differ.  An even number of passes, so
  // that the result ends where it started (the extra pass sees equal digits).
  __shared__ int ghist[RS_RADIX];
  if (tid == 0) GH_BP_STAT(2);
  const int nkp = (sbits + 7) / 8;
  const int gp = (nvp + nkp + 1) & ~1;
  uint64_t *kin = keys + s0, *kout = kscr + s0;
  int *vin = vals + s0, *vout = vscr + s0;
  const int ntiles = (int)((size + RS_TILE - 1) / RS_TILE);
  for (int pass = 0; pass < gp; pass++) {
    const SubDigit kd{klo, 8 * (pass - nvp)};
    const int vshift = 8 * pass;
    ghist[tid] = 0;
    __syncthreads();
    for (int64_t q = tid; q < size; q += RS_THREADS)
      atomicAdd(&ghist[pass < nvp ? (unsigned)((vin[q] >> vshift) & 0xff) : kd(kin[q])], 1);
    __syncthreads();
    {  // exclusive scan of the digit counts -> running output offsets
      const int g = ghist[tid];
      const int gincl = warp_scan<int>(g, lane);
      if (lane == 31) sm.wtot[w] = gincl;
      __syncthreads();
      int goff = 0;
#pragma unroll
      for (int k = 0; k < RS_WARPS; k++) goff += (k < w) ? sm.wtot[k] : 0;
      __syncthreads();
      ghist[tid] = goff + gincl - g;
    }
    __syncthreads();
    for (int t = 0; t < ntiles; t++) {
      const int64_t t0 = (int64_t)t * RS_TILE;
      const int count = (int)((size - t0) < RS_TILE ? (size - t0) : RS_TILE);
#pragma unroll
      for (int r = 0; r < RS_ROUNDS; r++) {
        const int j = w * (32 * RS_ROUNDS) + r * 32 + lane;
        key[r] = (j < count) ? kin[t0 + j] : 0;
        val[r] = (j < count) ? vin[t0 + j] : 0;
        const unsigned d = pass < nvp ? (unsigned)((val[r] >> vshift) & 0xff) : kd(key[r]);
        dig[r] = (j < count) ? d : (0x100u + (unsigned)lane);
      }
      const int tot = tile_rank_digits(sm, count, lrank, dig);
      tile_place(sm, key, val, count, lrank, dig);
#pragma unroll
      for (int k = 0; k < RS_ROUNDS; k++) {
        const int j = k * RS_THREADS + tid;
        if (j < count) {
          const uint64_t kk = sm.skey[j];
          const int vv = sm.sval[j];
          const int d = (int)(pass < nvp ? (unsigned)((vv >> vshift) & 0xff) : kd(kk));
          const int64_t g = (int64_t)ghist[d] + (j - sm.dstart[d]);
          kout[g] = kk;
          vout[g] = vv;
        }
      }
      __syncthreads();
      ghist[tid] += tot;  // this tile's pairs of digit tid are placed
      __syncthreads();
    }
    uint64_t *tk = kin; kin = kout; kout = tk;
    int *tv = vin; vin = vout; vout = tv;
    __threadfence_block();
    __syncthreads();
  }
}

// spl[b] = the (b n / nb)-th sorted key (b = 1 .. nb-1), spl[0] = 0: the next step's buckets
__global__ void bs_splitters_kernel(const uint64_t *__restrict__ sorted, int64_t n, int nb, uint64_t *__restrict__ spl,
                                    const int *__restrict__ ndev) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (ndev) n = *ndev;
  if (b >= nb) return;
  spl[b] = (b == 0 || n <= 0) ? 0ull : sorted[(n * b) / nb];
}

#ifndef GH_HOST_EMU
struct SplitterState {
  DeviceBuffer spl, boff, bid, cnt;
  int nb = 0;          // buckets the stored splitters describe (0 = none)
  int64_t cap = 0;     // capacity (n) they were taken for
  void release() { spl.release(); boff.release(); bid.release(); cnt.release(); nb = 0; }
};
static inline int bs_buckets_for(int64_t n) {
  int64_t nb = (n + BS_TARGET - 1) / BS_TARGET;
  if (nb > BS_MAX_BUCKETS) nb = BS_MAX_BUCKETS;
  return (int)nb;
}

// Stable sort of (key, value) pairs on key bits [0, nbits) with the buckets of `ss` (valid splitters
// required: ss.nb >= BS_MIN_BUCKETS).  Result in (kA, vA); (kB, vB) is scratch.
static int splitter_sort_pairs(uint64_t *kA, int *vA, uint64_t *kB, int *vB, int64_t n, int nbits, RadixScratch &rs,
                               SplitterState &ss, cudaStream_t st, const int *ndev) {
  if (n <= 1) return GH_OK;
  const int nblocks = (int)((n + RS_TILE - 1) / RS_TILE);
  GH_TRY(rs.hist.reserve(sizeof(int) * (size_t)RS_RADIX * (size_t)nblocks));
  GH_TRY(rs.gtot.reserve(sizeof(int) * RS_RADIX * 8));
  GH_TRY(ss.boff.reserve(sizeof(int) * (size_t)(ss.nb + 2)));
  GH_CUDA(cudaMemsetAsync(rs.gtot.ptr, 0, sizeof(int) * RS_RADIX * 2, st));
  const BucketOf bo{ss.spl.as<uint64_t>(), ss.nb};
  uint64_t *kin = kA, *kout = kB;
  int *vin = vA, *vout = vB;
  for (int pass = 0; pass < 2; pass++) {
    const BucketDigit dg{bo, 8 * pass};
    int *gtot = rs.gtot.as<int>() + RS_RADIX * pass;
    bs_hist_kernel<<<nblocks, RS_THREADS, 0, st>>>(kin, n, dg, rs.hist.as<int>(), nblocks, gtot, ndev);
    GH_LAUNCH_CHECK();
    rs_rowscan_kernel<<<RS_RADIX, RS_THREADS, 0, st>>>(rs.hist.as<int>(), nblocks);
    GH_LAUNCH_CHECK();
    bs_scatter_kernel<<<nblocks, RS_THREADS, 0, st>>>(kin, vin, kout, vout, n, dg, rs.hist.as<int>(), gtot, nblocks, ndev);
    GH_LAUNCH_CHECK();
    uint64_t *tk = kin; kin = kout; kout = tk;
    int *tv = vin; vin = vout; vout = tv;
  }
  // two passes: the partitioned pairs are back in (kA, vA)
  bs_offsets_kernel<<<(ss.nb + 1 + 255) / 256, 256, 0, st>>>(kA, n, bo, ss.boff.as<int>(), ndev);
  GH_LAUNCH_CHECK();
  bs_bucket_kernel<<<ss.nb, RS_THREADS, 0, st>>>(kA, vA, kB, vB, ss.boff.as<int>(), ss.spl.as<uint64_t>(), ss.nb, nbits);
  GH_LAUNCH_CHECK();
  return GH_OK;
}

// The "place" form (see bp_count_kernel).  Result in (kB, vB); (kA, vA) is clobbered (scratch of the
// oversize path).  The values of the input must ascend (vA[q] < vA[q+1]): the result is then the
// stable sort's.
static int splitter_place_sort_pairs(uint64_t *kA, int *vA, uint64_t *kB, int *vB, int64_t n, int nbits,
                                     SplitterState &ss, cudaStream_t st, const int *ndev) {
  if (n <= 1) return GH_OK;
  GH_TRY(ss.bid.reserve(sizeof(unsigned short) * (size_t)n));
  GH_TRY(ss.cnt.reserve(sizeof(int) * (size_t)(2 * ss.nb + 2)));
  GH_TRY(ss.boff.reserve(sizeof(int) * (size_t)(ss.nb + 2)));
  int *count = ss.cnt.as<int>(), *cursor = count + ss.nb + 1;
  GH_CUDA(cudaMemsetAsync(count, 0, sizeof(int) * (size_t)ss.nb, st));
  const unsigned nblocks = (unsigned)((n + BP_THREADS * BP_ROUNDS - 1) / (BP_THREADS * BP_ROUNDS));
  const uint64_t *spl = ss.spl.as<uint64_t>();
  bp_count_kernel<<<nblocks, BP_THREADS, 0, st>>>(kA, n, spl, ss.nb, ss.bid.as<unsigned short>(), count, ndev);
  GH_LAUNCH_CHECK();
  bp_scan_kernel<<<1, RS_THREADS, 0, st>>>(count, ss.nb, ss.boff.as<int>(), cursor);
  GH_LAUNCH_CHECK();
  bp_place_kernel<<<nblocks, BP_THREADS, 0, st>>>(kA, vA, ss.bid.as<unsigned short>(), n, cursor, kB, vB, ndev);
  GH_LAUNCH_CHECK();
  int vbits = 1;
  while (vbits < 31 && (int64_t(1) << vbits) < n) vbits++;
  bp_bucket_kernel<<<ss.nb, RS_THREADS, 0, st>>>(kB, vB, kA, vA, ss.boff.as<int>(), spl, ss.nb, nbits, vbits);
  GH_LAUNCH_CHECK();
  return GH_OK;
}

// The shared-memory form of "place" (bp2_*): same contract as splitter_place_sort_pairs.  G: CTAs of
// the count / place kernels (a couple per SM).
static int splitter_place2_sort_pairs(uint64_t *kA, int *vA, uint64_t *kB, int *vB, int64_t n, int nbits,
                                      SplitterState &ss, int G, cudaStream_t st, const int *ndev) {
  if (n <= 1) return GH_OK;
  if (ss.nb > BP2_MAX_NB) { set_error("place2: %d buckets exceed the shared-memory histogram", ss.nb); return GH_EINVAL; }
  GH_TRY(ss.bid.reserve(sizeof(unsigned short) * (size_t)n));
  GH_TRY(ss.cnt.reserve(sizeof(int) * ((size_t)G * (size_t)ss.nb + (size_t)(2 * ss.nb + 2))));
  GH_TRY(ss.boff.reserve(sizeof(int) * (size_t)(ss.nb + 2)));
  int *ghist = ss.cnt.as<int>(), *tot = ghist + (size_t)G * (size_t)ss.nb, *cursor = tot + ss.nb + 1;
  const uint64_t *spl = ss.spl.as<uint64_t>();
  bp2_count_kernel<<<G, BP2_THREADS, 0, st>>>(kA, n, spl, ss.nb, ss.bid.as<unsigned short>(), ghist, ndev);
  GH_LAUNCH_CHECK();
  bp2_colscan_kernel<<<(ss.nb + 31) / 32, 256, 0, st>>>(ghist, G, ss.nb, tot);
  GH_LAUNCH_CHECK();
  bp_scan_kernel<<<1, RS_THREADS, 0, st>>>(tot, ss.nb, ss.boff.as<int>(), cursor);
  GH_LAUNCH_CHECK();
  bp2_place_kernel<<<G, BP2_THREADS, 0, st>>>(kA, vA, ss.bid.as<unsigned short>(), n, ghist, ss.boff.as<int>(), ss.nb, kB, vB, ndev);
  GH_LAUNCH_CHECK();
  int vbits = 1;
  while (vbits < 31 && (int64_t(1) << vbits) < n) vbits++;
  bp_bucket_kernel<<<ss.nb, RS_THREADS, 0, st>>>(kB, vB, kA, vA, ss.boff.as<int>(), spl, ss.nb, nbits, vbits);
  GH_LAUNCH_CHECK();
  return GH_OK;
}

// after a sort: the next step's splitters from the sorted keys (capacity n, real count possibly on
// the device)
static int splitter_refresh(SplitterState &ss, const uint64_t *sorted, int64_t n, cudaStream_t st, const int *ndev) {
  const int nb = bs_buckets_for(n);
  if (nb < BS_MIN_BUCKETS) { ss.nb = 0; return GH_OK; }
  GH_TRY(ss.spl.reserve(sizeof(uint64_t) * (size_t)(nb + 1)));
  bs_splitters_kernel<<<(nb + 255) / 256, 256, 0, st>>>(sorted, n, nb, ss.spl.as<uint64_t>(), ndev);
  GH_LAUNCH_CHECK();
  ss.nb = nb;
  ss.cap = n;
  return GH_OK;
}
#endif  // GH_HOST_EMU

}  // namespace gh
