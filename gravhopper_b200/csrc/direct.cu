// direct.cu -- direct-summation force kernels for sm_100a (no tensor cores: the pair kernel is
// not a contraction).  Replaces directsummation_workhorse / directsummation_position_workhorse
// (/root/reference/gravhopper/_jbgrav.c:140-193, :299-353).
//
// Decomposition: grid.x tiles the targets (BLOCK*KI per CTA, KI targets per thread held in
// registers), grid.y splits the sources into S contiguous chunks so that small problems still
// fill 148 SMs and large ones have many more CTAs than SM slots (no wave tail).  Sources stream
// through shared memory in tiles of BLOCK, double buffered; every thread reads the same source
// (a shared-memory broadcast).  S == 1: the epilogue (plain store, or the fused kick+drift of
// the leapfrog) runs inside the force kernel.  S > 1: per-chunk partial sums go to scratch and a
// finalize kernel adds them in chunk order (deterministic) and runs the same epilogue.
//
// fp32 kernel: per source the tile holds x' = x s, y' = y s, z' = z s, s = m^-1/2, e = eps^2 s^2,
// so that with d' = x' - x_i s (one FMA) and q = |d'|^2 + e (three FMAs), rsqrt(q)^3 d' =
// m d / (r^2+eps^2)^{3/2}: 11 FP32 operations and one MUFU.RSQ per interaction.  Two sources
// are processed per instruction with Blackwell's packed FFMA2/FMUL2 (one issue slot, two
// lanes-worth of FMA), which is what makes room in the issue stream for the MUFU and LDS.
// Accumulation is two-level: fp32 within a tile of BLOCK sources, fp64 across tiles (SURVEY
// F10: plain fp32 accumulation fails the 1e-5 bar at N = 1M).
//
// fp64 kernel: d = x_j - x_i exactly as the reference forms it, s = |d|^2 + eps^2, and
// s^-3/2 from MUFU.RSQ64H refined by one third-order step (error ~ 1e-19), then m s^-3/2 d
// accumulated with DFMA in source order within a chunk.
#include "common.cuh"
#include "direct.cuh"

#include <cstdio>
#include <cstdlib>

namespace gh {

template <int BLOCK, int KI, int MINB = 1, int UNR = 4>
static int run_f32(const DirectArgs &a, DeviceBuffer &ws, cudaStream_t st, cudaEvent_t *ev) {
  Split sp = choose_split(a.ni, a.nj, BLOCK * KI, BLOCK);
  double *partial = nullptr;
  if (sp.S > 1) {
    GH_TRY(ws.reserve(sizeof(double) * 3 * (size_t)sp.S * (size_t)a.ni));
    partial = ws.as<double>();
  }
  dim3 grid(sp.itiles, sp.S);
  float eps2 = (float)(a.eps * a.eps);
  // fp32: a softening whose square is below ~1e-16 kpc^2 cannot regularise anything at fp32
  // resolution, and m * rsqrt(eps^2)^3 of the self term would overflow to inf (inf * 0 = NaN):
  // treat it as eps = 0 with the zero-distance guard, like the reference's _position guard
  const bool tiny_eps = !(eps2 >= GH_F32_MIN_EPS2);
  if (tiny_eps) eps2 = 0.f;
  if (ev) GH_CUDA(cudaEventRecord(ev[0], st));
  const int mode = env_int("GH_F32_MODE", 0);
  if (tiny_eps)
    direct_f32_kernel<BLOCK, KI, true, 0, MINB, UNR><<<grid, BLOCK, 0, st>>>(a.src32, a.nj, a.tgt32, a.ni, eps2,
                                                                 sp.jchunk, partial, a.ep);
  else if (mode == 1)
    direct_f32_kernel<BLOCK, KI, false, 1, MINB, UNR><<<grid, BLOCK, 0, st>>>(a.src32, a.nj, a.tgt32, a.ni, eps2,
                                                                  sp.jchunk, partial, a.ep);
  else
    direct_f32_kernel<BLOCK, KI, false, 0, MINB, UNR><<<grid, BLOCK, 0, st>>>(a.src32, a.nj, a.tgt32, a.ni, eps2,
                                                                  sp.jchunk, partial, a.ep);
  GH_LAUNCH_CHECK();
  if (ev) GH_CUDA(cudaEventRecord(ev[1], st));
  if (partial) {
    finalize_kernel<<<(unsigned)((a.ni + 255) / 256), 256, 0, st>>>(partial, sp.S, a.ni, a.ep);
    GH_LAUNCH_CHECK();
  }
  return GH_OK;
}

template <int BLOCK, int KI>
static int run_f64(const DirectArgs &a, DeviceBuffer &ws, cudaStream_t st, cudaEvent_t *ev) {
  Split sp = choose_split(a.ni, a.nj, BLOCK * KI, BLOCK);
  double *partial = nullptr;
  if (sp.S > 1) {
    GH_TRY(ws.reserve(sizeof(double) * 3 * (size_t)sp.S * (size_t)a.ni));
    partial = ws.as<double>();
  }
  dim3 grid(sp.itiles, sp.S);
  double eps2 = a.eps * a.eps;
  if (ev) GH_CUDA(cudaEventRecord(ev[0], st));
  if (a.eps == 0.0)
    direct_f64_kernel<BLOCK, KI, true><<<grid, BLOCK, 0, st>>>(a.src_pos, a.src_mass, a.nj, a.tgt_pos,
                                                              a.ni, eps2, sp.jchunk, partial, a.ep);
  else
    direct_f64_kernel<BLOCK, KI, false><<<grid, BLOCK, 0, st>>>(a.src_pos, a.src_mass, a.nj, a.tgt_pos,
                                                               a.ni, eps2, sp.jchunk, partial, a.ep);
  GH_LAUNCH_CHECK();
  if (ev) GH_CUDA(cudaEventRecord(ev[1], st));
  if (partial) {
    finalize_kernel<<<(unsigned)((a.ni + 255) / 256), 256, 0, st>>>(partial, sp.S, a.ni, a.ep);
    GH_LAUNCH_CHECK();
  }
  return GH_OK;
}

int launch_direct(const DirectArgs &a, DeviceBuffer &ws, cudaStream_t st, cudaEvent_t *ev) {
  if (a.ni <= 0) return GH_OK;
  if (a.prec == GH_PREC_F32) {
    int ki = env_int("GH_F32_KI", 0);
    int blk = env_int("GH_F32_BLOCK", 0);
    // measured on B200 (profiles/r01_sweep_direct*.txt): 4 targets/thread x 256 threads and
    // 8 x 128 are within 2 % of each other (76 / 74-76 % of peak); smaller shapes only pay off
    // when there are too few targets to fill the chip
    if (ki == 0) ki = (a.ni >= 131072) ? (a.mixed_mass ? 8 : 4) : (a.ni >= 16384 ? 2 : 1);
    if (blk == 0) blk = (ki == 4) ? 256 : 128;
    const int minb = env_int("GH_F32_MINB", 1);
    const int unr = env_int("GH_F32_UNROLL", 4);
    if (ki == 8 && blk == 128 && unr == 2) return run_f32<128, 8, 1, 2>(a, ws, st, ev);
    if (ki == 8 && blk == 128 && unr == 1) return run_f32<128, 8, 1, 1>(a, ws, st, ev);
    if (ki == 8 && blk == 128 && unr == 8) return run_f32<128, 8, 1, 8>(a, ws, st, ev);
    if (ki == 4 && blk == 256 && unr == 2) return run_f32<256, 4, 1, 2>(a, ws, st, ev);
    if (ki == 4 && blk == 256 && unr == 8) return run_f32<256, 4, 1, 8>(a, ws, st, ev);
    if (ki == 6 && blk == 128) return run_f32<128, 6, 1, 4>(a, ws, st, ev);
    if (ki == 6 && blk == 192) return run_f32<192, 6, 1, 4>(a, ws, st, ev);
    if (ki == 12 && blk == 64) return run_f32<64, 12, 1, 2>(a, ws, st, ev);
    if (ki == 16 && blk == 64) return run_f32<64, 16, 1, 2>(a, ws, st, ev);
    if (ki == 8 && blk == 128 && minb == 3) return run_f32<128, 8, 3>(a, ws, st, ev);
    if (ki == 8 && blk == 128 && minb == 4) return run_f32<128, 8, 4>(a, ws, st, ev);
    if (ki == 4 && blk == 256 && minb == 3) return run_f32<256, 4, 3>(a, ws, st, ev);
    if (ki == 4 && blk == 128 && minb == 6) return run_f32<128, 4, 6>(a, ws, st, ev);
    switch (ki * 1000 + blk) {
      case 8128: return run_f32<128, 8>(a, ws, st, ev);
      case 8064: return run_f32<64, 8>(a, ws, st, ev);
      case 4256: return run_f32<256, 4>(a, ws, st, ev);
      case 4128: return run_f32<128, 4>(a, ws, st, ev);
      case 4512: return run_f32<512, 4>(a, ws, st, ev);
      case 2256: return run_f32<256, 2>(a, ws, st, ev);
      case 2128: return run_f32<128, 2>(a, ws, st, ev);
      case 1128: return run_f32<128, 1>(a, ws, st, ev);
      default:
        set_error("unsupported GH_F32_KI/GH_F32_BLOCK combination %d/%d", ki, blk);
        return GH_EINVAL;
    }
  } else if (a.prec == GH_PREC_F64) {
    int ki = env_int("GH_F64_KI", 0);
    if (ki == 0) ki = (a.ni >= 65536) ? 2 : 1;
    switch (ki) {
      case 4: return run_f64<128, 4>(a, ws, st, ev);
      case 2: return run_f64<256, 2>(a, ws, st, ev);
      default: return run_f64<128, 1>(a, ws, st, ev);
    }
  }
  set_error("launch_direct: bad precision %d", a.prec);
  return GH_EINVAL;
}

// -------------------------------------------------------------------------------------------
// small O(N) kernels
// -------------------------------------------------------------------------------------------
__global__ void pack32_kernel(const double *__restrict__ pos, const double *__restrict__ mass,
                              int64_t n, double ox, double oy, double oz, float4 *__restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = make_float4((float)(pos[3 * i] - ox), (float)(pos[3 * i + 1] - oy),
                       (float)(pos[3 * i + 2] - oz), mass ? (float)mass[i] : 0.f);
}

int launch_pack32(const double *pos, const double *mass, int64_t n, const double origin[3],
                  float4 *out, cudaStream_t st) {
  if (n <= 0) return GH_OK;
  pack32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(pos, mass, n, origin[0], origin[1],
                                                           origin[2], out);
  GH_LAUNCH_CHECK();
  return GH_OK;
}

__global__ void half_drift_kernel(const double *__restrict__ x, const double *__restrict__ v,
                                  const double *__restrict__ mass, int64_t n, double dt,
                                  double *__restrict__ xhalf, float4 *__restrict__ src32, double ox,
                                  double oy, double oz) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double h[3];
#pragma unroll
  for (int k = 0; k < 3; k++) {  // gravhopper.py:409
    double hd = __dmul_rn(__dmul_rn(__dmul_rn(0.5, v[3 * i + k]), dt), GH_KPC_PER_KMS_MYR);
    h[k] = __dadd_rn(x[3 * i + k], hd);
    xhalf[3 * i + k] = h[k];
  }
  if (src32)
    src32[i] = make_float4((float)(h[0] - ox), (float)(h[1] - oy), (float)(h[2] - oz), (float)mass[i]);
}

int launch_half_drift(const double *x, const double *v, const double *mass, int64_t n, double dt,
                      double *xhalf, float4 *src32, const double origin[3], cudaStream_t st) {
  if (n <= 0) return GH_OK;
  half_drift_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x, v, mass, n, dt, xhalf, src32,
                                                               origin[0], origin[1], origin[2]);
  GH_LAUNCH_CHECK();
  return GH_OK;
}

__global__ void potentials_kernel(PotentialSet ps, const double *__restrict__ xhalf, int64_t n,
                                  const double *__restrict__ ext_in, double *__restrict__ ext_out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double x[3] = {xhalf[3 * i], xhalf[3 * i + 1], xhalf[3 * i + 2]};
  double a[3] = {0.0, 0.0, 0.0};
  if (ext_in) { a[0] = ext_in[3 * i]; a[1] = ext_in[3 * i + 1]; a[2] = ext_in[3 * i + 2]; }
  eval_potentials(ps, x, a);
  ext_out[3 * i] = a[0];
  ext_out[3 * i + 1] = a[1];
  ext_out[3 * i + 2] = a[2];
}

int launch_potentials(const PotentialSet &ps, const double *xhalf, int64_t n, const double *ext_in,
                      double *ext_out, cudaStream_t st) {
  if (n <= 0) return GH_OK;
  potentials_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(ps, xhalf, n, ext_in, ext_out);
  GH_LAUNCH_CHECK();
  return GH_OK;
}

// Energy diagnostic (no reference code; consistent with _jbgrav.c:165-166).  Each target sums
// m_j / sqrt(r^2+eps^2) over sources with global index > its own (i < j pairs).
__global__ void energy_kernel(const double *__restrict__ x, const double *__restrict__ v,
                              const double *__restrict__ mt, int64_t ni,
                              const double *__restrict__ spos, const double *__restrict__ smass,
                              int64_t nj, int64_t self_offset, double eps2, double *out2) {
  __shared__ double4 sh[128];
  __shared__ double red[2][128];
  const int tid = threadIdx.x;
  int64_t i = (int64_t)blockIdx.x * 128 + tid;
  double xi = 0, yi = 0, zi = 0;
  if (i < ni) { xi = x[3 * i]; yi = x[3 * i + 1]; zi = x[3 * i + 2]; }
  const int64_t gi = i + self_offset;
  double pot = 0.0;
  // sources with index > the smallest target index of this block are the only ones needed
  int64_t jstart = ((int64_t)blockIdx.x * 128 + self_offset) / 128 * 128;
  for (int64_t j0 = jstart; j0 < nj; j0 += 128) {
    int64_t j = j0 + tid;
    sh[tid] = (j < nj) ? make_double4(spos[3 * j], spos[3 * j + 1], spos[3 * j + 2], smass[j])
                       : make_double4(0, 0, 0, 0);
    __syncthreads();
    for (int p = 0; p < 128; p++) {
      if (j0 + p > gi && j0 + p < nj) {
        double dx = sh[p].x - xi, dy = sh[p].y - yi, dz = sh[p].z - zi;
        pot += sh[p].w / sqrt(dx * dx + dy * dy + dz * dz + eps2);
      }
    }
    __syncthreads();
  }
  double ke = 0.0, pe = 0.0;
  if (i < ni) {
    ke = 0.5 * mt[i] * (v[3 * i] * v[3 * i] + v[3 * i + 1] * v[3 * i + 1] + v[3 * i + 2] * v[3 * i + 2]);
    pe = -GH_G * mt[i] * pot;
  }
  red[0][tid] = ke;
  red[1][tid] = pe;
  __syncthreads();
  for (int s = 64; s > 0; s >>= 1) {
    if (tid < s) { red[0][tid] += red[0][tid + s]; red[1][tid] += red[1][tid + s]; }
    __syncthreads();
  }
  if (tid == 0) { atomicAdd(&out2[0], red[0][0]); atomicAdd(&out2[1], red[1][0]); }
}

int launch_energy(const double *x, const double *v, const double *m_tgt, int64_t ni,
                  const double *src_pos, const double *src_mass, int64_t nj, int64_t self_offset,
                  double eps, double *out2, cudaStream_t st) {
  if (ni <= 0) return GH_OK;
  energy_kernel<<<(unsigned)((ni + 127) / 128), 128, 0, st>>>(x, v, m_tgt, ni, src_pos, src_mass, nj,
                                                            self_offset, eps * eps, out2);
  GH_LAUNCH_CHECK();
  return GH_OK;
}

}  // namespace gh
