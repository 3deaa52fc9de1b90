// ic.cuh -- device code of the initial-condition sampler (Philox, inverse-CDF tables, ic_kernel,
// centring kernels).  Included by ic.cu (which keeps the launches); a header of its own so that
// tests/emu can compile the same source for the host (GH_HOST_EMU), like walk.cuh.
#pragma once
#include "common.cuh"

namespace gh {

// ---- Philox4x32-10 (Salmon et al. 2011), counter = (particle index, draw index), key = seed ----
struct Philox {
  uint32_t c[4], k[2];
  __device__ __forceinline__ void round() {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
    uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
    uint32_t n0 = hi1 ^ c[1] ^ k[0], n1 = lo1, n2 = hi0 ^ c[3] ^ k[1], n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
  }
  __device__ __forceinline__ void run() {
#pragma unroll
    for (int r = 0; r < 10; r++) {
      round();
      k[0] += 0x9E3779B9u;
      k[1] += 0xBB67AE85u;
    }
  }
};

// four 32-bit words for (particle i, block d of draws)
__device__ __forceinline__ void philox4(uint64_t seed, uint64_t i, uint32_t d, uint32_t out[4]) {
  Philox p;
  p.c[0] = (uint32_t)i; p.c[1] = (uint32_t)(i >> 32); p.c[2] = d; p.c[3] = 0x47524156u;  // "GRAV"
  p.k[0] = (uint32_t)seed; p.k[1] = (uint32_t)(seed >> 32);
  p.run();
  for (int k = 0; k < 4; k++) out[k] = p.c[k];
}
// uniform double in [0,1) with 53 random bits from two words
__device__ __forceinline__ double u53(uint32_t a, uint32_t b) {
  return (double)((((uint64_t)a << 21) ^ (uint64_t)(b >> 11)) & ((1ull << 53) - 1)) * (1.0 / 9007199254740992.0);
}
struct Draws {  // eight uniforms per particle
  double u[8];
  __device__ Draws(uint64_t seed, uint64_t i) {
    uint32_t w[4];
    for (uint32_t d = 0; d < 4; d++) {
      philox4(seed, i, d, w);
      u[2 * d] = u53(w[0], w[1]);
      u[2 * d + 1] = u53(w[2], w[3]);
    }
  }
};

// linear interpolation y(x) through a table with non-decreasing xs (np.interp semantics)
__device__ __forceinline__ double interp(const double *xs, const double *ys, int n, double x) {
  if (x <= xs[0]) return ys[0];
  if (x >= xs[n - 1]) return ys[n - 1];
  int lo = 0, hi = n - 1;
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (xs[mid] <= x) lo = mid; else hi = mid;
  }
  double dx = xs[hi] - xs[lo];
  return dx > 0.0 ? ys[lo] + (ys[hi] - ys[lo]) * (x - xs[lo]) / dx : ys[lo];
}

__device__ __forceinline__ void sphere(double r, double uc, double up, double out[3]) {
  const double ct = 2.0 * uc - 1.0, ph = 6.283185307179586 * up;
  const double st = sqrt(fmax(0.0, 1.0 - ct * ct));
  double s, c;
  sincos(ph, &s, &c);
  out[0] = r * st * c;
  out[1] = r * st * s;
  out[2] = r * ct;
}

// kind 1: Plummer (b, M); table = (q_cumprob -> q)          gravhopper.py:1452-1491
// kind 2: Hernquist (a, M, cutoff); table = (E -> cumulative f(E)), used in both directions
//                                                            gravhopper.py:1544-1605
// kind 3: TSIS (maxrad, M)                                   gravhopper.py:1378-1398
__global__ void ic_kernel(int kind, int64_t n, double p0, double p1, double p2, const double *tx,
                          const double *ty, int nt, uint64_t seed, double *__restrict__ pos,
                          double *__restrict__ vel, double *__restrict__ mass) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Draws d(seed, (uint64_t)i);
  double x[3], v[3];
  if (kind == 1) {
    const double b = p0, M = p1;
    double xi = fmax(d.u[0], 1e-300);
    double r = b / sqrt(pow(xi, -2.0 / 3.0) - 1.0);
    sphere(r, d.u[1], d.u[2], x);
    double q = interp(tx, ty, nt, d.u[3]);
    double vmag = q * sqrt(2.0 * GH_G * M / b) * pow(1.0 + (r / b) * (r / b), -0.25);
    sphere(vmag, d.u[4], d.u[5], v);
  } else if (kind == 2) {
    const double a = p0, M = p1, cutoff = p2;
    const double xi_cut = cutoff * cutoff / ((1.0 + cutoff) * (1.0 + cutoff));
    double xi = fmax(d.u[0] * xi_cut, 1e-300);
    double roa = 1.0 / (1.0 / sqrt(xi) - 1.0);
    sphere(roa * a, d.u[1], d.u[2], x);
    const double potential = -1.0 / (1.0 + roa);
    double max_xi = interp(tx, ty, nt, -potential);  // E -> cumulative
    double E = interp(ty, tx, nt, d.u[3] * max_xi);  // cumulative -> E
    double vmag = sqrt(2.0 * fmax(-(E + potential), 0.0) * GH_G * M / a);
    sphere(vmag, d.u[4], d.u[5], v);
  } else {
    const double maxrad = p0, M = p1;
    const double sigma = sqrt(M * GH_G / (2.0 * maxrad));
    sphere(d.u[0] * maxrad, d.u[1], d.u[2], x);
    // Box-Muller: three normals from four uniforms
    double r1 = sqrt(-2.0 * log(fmax(d.u[3], 1e-300))), r2 = sqrt(-2.0 * log(fmax(d.u[5], 1e-300)));
    double s1, c1, s2, c2;
    sincos(6.283185307179586 * d.u[4], &s1, &c1);
    sincos(6.283185307179586 * d.u[6], &s2, &c2);
    v[0] = sigma * r1 * c1; v[1] = sigma * r1 * s1; v[2] = sigma * r2 * c2;
    (void)s2;
  }
  const double M = p1;
  for (int k = 0; k < 3; k++) { pos[3 * i + k] = x[k]; vel[3 * i + k] = v[k]; }
  mass[i] = M / (double)n;
}

// Exponential disk (gravhopper.py:1669-1733): R from the tabulated cumulative mass profile, azimuth
// uniform, z = 2 z0 atanh(u) with a random sign; velocities: mean rotation R Omega(R), dispersions
// sigma_R = sigmaR(Rd) exp((1 - R/Rd) / 4), sigma_phi^2 = sigma_R^2 4 Omega^2 / kappa^2, sigma_z^2 =
// pi G z0 Sigma(R) / 2: exactly the expressions of the reference and of the host generator
// (ic_raw.expdisk), with the two functions that need Bessel functions and the caller's rotation
// curve tabulated by the host on the radial grid tR:  tvphi = R sqrt(Omega^2),
// tratio = 4 Omega^2 / kappa^2 (both smooth and bounded, unlike Omega^2 itself at R -> 0).
// prm = {sigma0 [Msun/kpc^2], Rd [kpc], z0 [kpc], sigmaR(Rd) [km/s]}.
__global__ void ic_expdisk_kernel(int64_t n, double sigma0, double Rd, double z0, double sigR_Rd,
                                  const double *tR, const double *tcum, const double *tvphi,
                                  const double *tratio, int nt, uint64_t seed, double *__restrict__ pos,
                                  double *__restrict__ vel, double *__restrict__ mass) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Draws d(seed, (uint64_t)i);
  const double R = fmax(interp(tcum, tR, nt, d.u[0]), 1e-6 * Rd);
  double sp, cp;
  sincos(6.283185307179586 * d.u[1], &sp, &cp);
  const double zu = fmin(d.u[2], 1.0 - 1e-16);
  const double z = 2.0 * z0 * atanh(zu) * (d.u[3] < 0.5 ? 1.0 : -1.0);
  const double vphi_mean = interp(tR, tvphi, nt, R);
  const double ratio = interp(tR, tratio, nt, R);
  const double sigma_R = sigR_Rd * exp(0.25 * (1.0 - R / Rd));
  const double sigma2_phi = sigma_R * sigma_R * ratio;
  const double sigma2_z = 3.141592653589793 * GH_G * z0 * sigma0 * 0.5 * exp(-R / Rd);
  // Box-Muller: three normals from four uniforms
  const double r1 = sqrt(-2.0 * log(fmax(d.u[4], 1e-300))), r2 = sqrt(-2.0 * log(fmax(d.u[6], 1e-300)));
  double s1, c1, s2, c2;
  sincos(6.283185307179586 * d.u[5], &s1, &c1);
  sincos(6.283185307179586 * d.u[7], &s2, &c2);
  (void)s2;
  const double vphi = sqrt(fmax(sigma2_phi, 0.0)) * (r1 * c1) + vphi_mean;
  const double vR = sigma_R * (r1 * s1);
  const double vz = sqrt(sigma2_z) * (r2 * c2);
  pos[3 * i + 0] = R * cp;
  pos[3 * i + 1] = R * sp;
  pos[3 * i + 2] = z;
  vel[3 * i + 0] = -vphi * sp + vR * cp;
  vel[3 * i + 1] = vphi * cp + vR * sp;
  vel[3 * i + 2] = vz;
  mass[i] = 3.141592653589793 * Rd * Rd * sigma0 / (double)n;
}

// force_centers (gravhopper.py:1768-1785): subtract the unweighted mean (deterministic 2-stage sum)
__global__ void mean_stage1(const double *__restrict__ a, int64_t n, double *__restrict__ part) {
  __shared__ double sh[3][256];
  double s[3] = {0, 0, 0};
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    for (int k = 0; k < 3; k++) s[k] += a[3 * i + k];
  for (int k = 0; k < 3; k++) sh[k][threadIdx.x] = s[k];
  __syncthreads();
  for (int st = 128; st > 0; st >>= 1) {
    if (threadIdx.x < st) for (int k = 0; k < 3; k++) sh[k][threadIdx.x] += sh[k][threadIdx.x + st];
    __syncthreads();
  }
  if (threadIdx.x < 3) part[blockIdx.x * 3 + threadIdx.x] = sh[threadIdx.x][0];
}
__global__ void mean_stage2_shift(double *__restrict__ a, int64_t n, const double *__restrict__ part, int nb) {
  __shared__ double mean[3];
  if (threadIdx.x < 3) {
    double s = 0.0;
    for (int b = 0; b < nb; b++) s += part[b * 3 + threadIdx.x];
    mean[threadIdx.x] = s / (double)n;
  }
  __syncthreads();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    for (int k = 0; k < 3; k++) a[3 * i + k] -= mean[k];
}

}  // namespace gh
