// direct.cuh -- device code of the direct-summation force kernels (fp32 / fp64 / finalize) and the
// host-side split heuristic.  Included by direct.cu (which keeps the launches); a header of its
// own so that tests/emu can compile the same source for the host (GH_HOST_EMU), like walk.cuh.
#pragma once
#include "common.cuh"

#include <cstdio>
#include <cstdlib>

namespace gh {

// -------------------------------------------------------------------------------------------
// fp32
// -------------------------------------------------------------------------------------------
// bare MUFU.RSQ (rsqrtf() without -ftz wraps it in three denormal-handling instructions)
__device__ __forceinline__ float rsq_approx(float x) {
  float y;
#ifdef GH_HOST_EMU
  y = 1.0f / sqrtf(x);
#else
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
#endif
  return y;
}

template <int BLOCK>
struct Tile32 {
  float4 a[BLOCK / 2];  // (x'0, x'1, y'0, y'1) per source pair
  float4 b[BLOCK / 2];  // (z'0, z'1, s0, s1)
  float2 c[BLOCK / 2];  // (e0, e1)
};

template <int BLOCK, int MODE>
__device__ __forceinline__ void store_tile32(Tile32<BLOCK> &t, int tid, float4 g, float eps2) {
  // g = (x, y, z, m); m == 0 (padding or a massless tracer) contributes exactly zero
  float s = g.w > 0.f ? rsqrtf(g.w) : 0.f;
  float e = g.w > 0.f ? eps2 * s * s : 1.f;
  if (MODE == 0) { s = 1.f; e = 0.f; }
  float *pa = reinterpret_cast<float *>(&t.a[tid >> 1]);
  float *pb = reinterpret_cast<float *>(&t.b[tid >> 1]);
  float *pc = reinterpret_cast<float *>(&t.c[tid >> 1]);
  int h = tid & 1;
  pa[h] = g.x * s;
  pa[2 + h] = g.y * s;
  pb[h] = g.z * s;
  pb[2 + h] = (MODE == 0) ? g.w : s;
  if (MODE != 0) pc[h] = e;
}

// MODE 0: tile holds (x, y, z, m); d = x_j - x_i (FADD2), w = m r^3 (3 FMUL2): 12 FP32 ops.
// MODE 1: per-source scaling described above: 11 FP32 ops, but d' of a source that coincides
//         with the target is a rounding residue instead of an exact zero -> only usable when
//         the caller knows targets never coincide with sources (kept for measurement).
template <int BLOCK, int KI, bool GUARD, int MODE, int MINB, int UNR>
__global__ void __launch_bounds__(BLOCK, MINB)
direct_f32_kernel(const float4 *__restrict__ src, int64_t nj, const float4 *__restrict__ tgt,
                  int64_t ni, float eps2, int64_t jchunk, double *__restrict__ partial,
                  Epilogue ep) {
  __shared__ Tile32<BLOCK> tile[2];
  const int tid = threadIdx.x;
  const int64_t i0 = (int64_t)blockIdx.x * (BLOCK * KI);
  const int64_t jb = (int64_t)blockIdx.y * jchunk;
  const int64_t je = (jb + jchunk < nj) ? jb + jchunk : nj;
  const int ntiles = (int)((je - jb + BLOCK - 1) / BLOCK);

  float2 nx[KI], ny[KI], nz[KI];
  double ax[KI], ay[KI], az[KI];
#pragma unroll
  for (int k = 0; k < KI; k++) {
    int64_t i = i0 + tid + (int64_t)k * BLOCK;
    float4 t = (i < ni) ? tgt[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    nx[k] = make_float2(-t.x, -t.x);
    ny[k] = make_float2(-t.y, -t.y);
    nz[k] = make_float2(-t.z, -t.z);
    ax[k] = ay[k] = az[k] = 0.0;
  }

  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  // A tile whose BLOCK sources all carry the same mass m0 (the usual case: equal-mass
  // components stored contiguously) skips the per-interaction mass multiply; its fp32 partial sum
  // is scaled by m0 once, when it is promoted to fp64.  12 -> 11 FP32 operations per interaction.
  float m0 = 0.f;
  bool uni = false;
  {
    int64_t j = jb + tid;
    float4 g = (j < je) ? src[j] : zero4;
    m0 = (jb < je) ? src[jb].w : 0.f;
    store_tile32<BLOCK, MODE>(tile[0], tid, g, eps2);
    uni = __syncthreads_and(g.w == m0) != 0;
  }

  for (int t = 0; t < ntiles; t++) {
    float4 g = zero4;
    float m0n = 0.f;
    const bool more = (t + 1 < ntiles);
    if (more) {
      const int64_t jt = jb + (int64_t)(t + 1) * BLOCK;
      int64_t j = jt + tid;
      if (j < je) g = src[j];
      m0n = src[jt].w;
    }
    const Tile32<BLOCK> &T = tile[t & 1];
    float2 fx[KI], fy[KI], fz[KI];
#pragma unroll
    for (int k = 0; k < KI; k++) fx[k] = fy[k] = fz[k] = make_float2(0.f, 0.f);

    if (MODE == 0 && uni) {
#pragma unroll UNR
      for (int p = 0; p < BLOCK / 2; p++) {
        const float4 A = T.a[p];
        const float2 zj = *reinterpret_cast<const float2 *>(&T.b[p]);
        const float2 xj = make_float2(A.x, A.y), yj = make_float2(A.z, A.w);
        const float2 e = make_float2(eps2, eps2);
#pragma unroll
        for (int k = 0; k < KI; k++) {
          float2 dx = __fadd2_rn(xj, nx[k]);
          float2 dy = __fadd2_rn(yj, ny[k]);
          float2 dz = __fadd2_rn(zj, nz[k]);
          float2 q = __ffma2_rn(dx, dx, e);
          q = __ffma2_rn(dy, dy, q);
          q = __ffma2_rn(dz, dz, q);
          float2 r;
          if (GUARD) {
            r.x = q.x > 0.f ? rsq_approx(q.x) : 0.f;
            r.y = q.y > 0.f ? rsq_approx(q.y) : 0.f;
          } else {
            r.x = rsq_approx(q.x);
            r.y = rsq_approx(q.y);
          }
          float2 r2 = __fmul2_rn(r, r);
          float2 r3 = __fmul2_rn(r2, r);
          fx[k] = __ffma2_rn(r3, dx, fx[k]);
          fy[k] = __ffma2_rn(r3, dy, fy[k]);
          fz[k] = __ffma2_rn(r3, dz, fz[k]);
        }
      }
      const double dm0 = (double)m0;
#pragma unroll
      for (int k = 0; k < KI; k++) {
        ax[k] = fma((double)(fx[k].x + fx[k].y), dm0, ax[k]);
        ay[k] = fma((double)(fy[k].x + fy[k].y), dm0, ay[k]);
        az[k] = fma((double)(fz[k].x + fz[k].y), dm0, az[k]);
      }
    } else {
#pragma unroll 4
      for (int p = 0; p < BLOCK / 2; p++) {
        const float4 A = T.a[p];
        const float4 B = T.b[p];
        const float2 xj = make_float2(A.x, A.y), yj = make_float2(A.z, A.w);
        const float2 zj = make_float2(B.x, B.y), s = make_float2(B.z, B.w);
        float2 e;
        if (MODE == 0) e = make_float2(eps2, eps2);
        else e = T.c[p];
#pragma unroll
        for (int k = 0; k < KI; k++) {
          float2 dx, dy, dz;
          if (MODE == 0) {
            dx = __fadd2_rn(xj, nx[k]);
            dy = __fadd2_rn(yj, ny[k]);
            dz = __fadd2_rn(zj, nz[k]);
          } else {
            dx = __ffma2_rn(nx[k], s, xj);
            dy = __ffma2_rn(ny[k], s, yj);
            dz = __ffma2_rn(nz[k], s, zj);
          }
          float2 q = __ffma2_rn(dx, dx, e);
          q = __ffma2_rn(dy, dy, q);
          q = __ffma2_rn(dz, dz, q);
          float2 r;
          if (GUARD) {  // eps == 0: a source exactly at the target contributes zero
            r.x = q.x > 0.f ? rsq_approx(q.x) : 0.f;
            r.y = q.y > 0.f ? rsq_approx(q.y) : 0.f;
          } else {
            r.x = rsq_approx(q.x);
            r.y = rsq_approx(q.y);
          }
          float2 r2 = __fmul2_rn(r, r);
          float2 r3 = __fmul2_rn(r2, r);
          if (MODE == 0) r3 = __fmul2_rn(r3, s);  // s holds the masses in MODE 0
          fx[k] = __ffma2_rn(r3, dx, fx[k]);
          fy[k] = __ffma2_rn(r3, dy, fy[k]);
          fz[k] = __ffma2_rn(r3, dz, fz[k]);
        }
      }
#pragma unroll
      for (int k = 0; k < KI; k++) {  // second accumulation level: fp64 across tiles
        ax[k] += (double)(fx[k].x + fx[k].y);
        ay[k] += (double)(fy[k].x + fy[k].y);
        az[k] += (double)(fz[k].x + fz[k].y);
      }
    }
    if (more) store_tile32<BLOCK, MODE>(tile[(t + 1) & 1], tid, g, eps2);
    uni = __syncthreads_and(more && g.w == m0n) != 0;
    m0 = m0n;
  }

#pragma unroll
  for (int k = 0; k < KI; k++) {
    int64_t i = i0 + tid + (int64_t)k * BLOCK;
    if (i >= ni) continue;
    if (partial) {
      double *o = partial + ((int64_t)blockIdx.y * ni + i) * 3;
      o[0] = ax[k];
      o[1] = ay[k];
      o[2] = az[k];
    } else {
      apply_epilogue(ep, i, ax[k], ay[k], az[k]);
    }
  }
}

// -------------------------------------------------------------------------------------------
// fp64
// -------------------------------------------------------------------------------------------
// s^-1/2 to ~1e-19: MUFU.RSQ64H seed (relative error ~2^-20) + one third-order correction
// y (1 + e/2 + 3e^2/8), e = 1 - s y^2; truncation 5e^3/16 ~ 1e-18.
__device__ __forceinline__ double rsqrt64(double s) {
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

template <int BLOCK>
struct Tile64 {
  double4 s[BLOCK];  // (x, y, z, m)
};

template <int BLOCK, int KI, bool GUARD>
__global__ void __launch_bounds__(BLOCK)
direct_f64_kernel(const double *__restrict__ spos, const double *__restrict__ smass, int64_t nj,
                  const double *__restrict__ tpos, int64_t ni, double eps2, int64_t jchunk,
                  double *__restrict__ partial, Epilogue ep) {
  __shared__ Tile64<BLOCK> tile[2];
  const int tid = threadIdx.x;
  const int64_t i0 = (int64_t)blockIdx.x * (BLOCK * KI);
  const int64_t jb = (int64_t)blockIdx.y * jchunk;
  const int64_t je = (jb + jchunk < nj) ? jb + jchunk : nj;
  const int ntiles = (int)((je - jb + BLOCK - 1) / BLOCK);

  double xi[KI], yi[KI], zi[KI], ax[KI], ay[KI], az[KI];
#pragma unroll
  for (int k = 0; k < KI; k++) {
    int64_t i = i0 + tid + (int64_t)k * BLOCK;
    if (i < ni) {
      xi[k] = tpos[3 * i];
      yi[k] = tpos[3 * i + 1];
      zi[k] = tpos[3 * i + 2];
    } else {
      xi[k] = yi[k] = zi[k] = 0.0;
    }
    ax[k] = ay[k] = az[k] = 0.0;
  }

  auto fetch = [&](int64_t j) -> double4 {
    if (j < je) return make_double4(spos[3 * j], spos[3 * j + 1], spos[3 * j + 2], smass[j]);
    return make_double4(0.0, 0.0, 0.0, 0.0);  // zero mass: contributes exactly zero
  };
  // tiles whose sources all carry the same mass m0 factor it out of the inner loop (17 -> 16 DP
  // operations per interaction); the tile's sum is accumulated separately and scaled once
  double m0 = (jb < je) ? smass[jb] : 0.0;
  bool uni;
  {
    double4 g = fetch(jb + tid);
    tile[0].s[tid] = g;
    uni = __syncthreads_and(g.w == m0) != 0;
  }

  for (int t = 0; t < ntiles; t++) {
    const bool more = (t + 1 < ntiles);
    double4 g = make_double4(0.0, 0.0, 0.0, 0.0);
    double m0n = 0.0;
    if (more) {
      const int64_t jt = jb + (int64_t)(t + 1) * BLOCK;
      g = fetch(jt + tid);
      m0n = smass[jt];
    }
    const Tile64<BLOCK> &T = tile[t & 1];
    if (uni) {
      double tx[KI], ty[KI], tz[KI];
#pragma unroll
      for (int k = 0; k < KI; k++) tx[k] = ty[k] = tz[k] = 0.0;
#pragma unroll 4
      for (int p = 0; p < BLOCK; p++) {
        const double4 sj = T.s[p];
#pragma unroll
        for (int k = 0; k < KI; k++) {
          double dx = sj.x - xi[k];
          double dy = sj.y - yi[k];
          double dz = sj.z - zi[k];
          double s = fma(dx, dx, eps2);
          s = fma(dy, dy, s);
          s = fma(dz, dz, s);
          double y = rsqrt64(s);
          if (GUARD) y = (s > 0.0) ? y : 0.0;
          double w = y * y * y;
          tx[k] = fma(w, dx, tx[k]);
          ty[k] = fma(w, dy, ty[k]);
          tz[k] = fma(w, dz, tz[k]);
        }
      }
#pragma unroll
      for (int k = 0; k < KI; k++) {
        ax[k] = fma(m0, tx[k], ax[k]);
        ay[k] = fma(m0, ty[k], ay[k]);
        az[k] = fma(m0, tz[k], az[k]);
      }
    } else {
#pragma unroll 4
      for (int p = 0; p < BLOCK; p++) {
        const double4 sj = T.s[p];
#pragma unroll
        for (int k = 0; k < KI; k++) {
          double dx = sj.x - xi[k];
          double dy = sj.y - yi[k];
          double dz = sj.z - zi[k];
          double s = fma(dx, dx, eps2);
          s = fma(dy, dy, s);
          s = fma(dz, dz, s);
          double y = rsqrt64(s);
          if (GUARD) y = (s > 0.0) ? y : 0.0;  // _jbgrav.c:327-328
          double w = sj.w * (y * y * y);
          ax[k] = fma(w, dx, ax[k]);
          ay[k] = fma(w, dy, ay[k]);
          az[k] = fma(w, dz, az[k]);
        }
      }
    }
    if (more) tile[(t + 1) & 1].s[tid] = g;
    uni = __syncthreads_and(more && g.w == m0n) != 0;
    m0 = m0n;
  }

#pragma unroll
  for (int k = 0; k < KI; k++) {
    int64_t i = i0 + tid + (int64_t)k * BLOCK;
    if (i >= ni) continue;
    if (partial) {
      double *o = partial + ((int64_t)blockIdx.y * ni + i) * 3;
      o[0] = ax[k];
      o[1] = ay[k];
      o[2] = az[k];
    } else {
      apply_epilogue(ep, i, ax[k], ay[k], az[k]);
    }
  }
}

__global__ void finalize_kernel(const double *__restrict__ partial, int S, int64_t ni,
                                Epilogue ep) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ni) return;
  double ax = 0.0, ay = 0.0, az = 0.0;
  for (int c = 0; c < S; c++) {  // fixed chunk order: results do not depend on scheduling
    const double *o = partial + ((int64_t)c * ni + i) * 3;
    ax += o[0];
    ay += o[1];
    az += o[2];
  }
  apply_epilogue(ep, i, ax, ay, az);
}

// -------------------------------------------------------------------------------------------
// launch heuristics
// -------------------------------------------------------------------------------------------
static int env_int(const char *name, int dflt) {
  const char *s = getenv(name);
  return (s && *s) ? atoi(s) : dflt;
}

struct Split {
  int S;
  int64_t jchunk;
  unsigned itiles;
};

static Split choose_split(int64_t ni, int64_t nj, int itile, int tj) {
  // aim for >= GH_DIRECT_CTAS CTAs (default 148 SMs x 2 resident x 16) so the last wave is a small
  // fraction of the run; never make a chunk shorter than 2 source tiles.
  const int64_t want = env_int("GH_DIRECT_CTAS", 148 * 2 * 16);
  Split sp;
  sp.itiles = (unsigned)((ni + itile - 1) / itile);
  int64_t ntiles = (nj + tj - 1) / tj;
  int64_t S = (want + sp.itiles - 1) / sp.itiles;
  int64_t maxS = (ntiles + 1) / 2;
  if (S > maxS) S = maxS;
  if (S > 65535) S = 65535;
  if (S < 1) S = 1;
  int64_t tiles_per = (ntiles + S - 1) / S;
  sp.jchunk = tiles_per * tj;
  sp.S = (int)((nj + sp.jchunk - 1) / sp.jchunk);
  if (sp.S < 1) sp.S = 1;
  return sp;
}

}  // namespace gh
