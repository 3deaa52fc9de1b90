// probe.cu -- FP32 FMA peak probe (measurement aid of bench.py's roofline, not on the hot path).
// MEASURED_PEAKS.json holds HBM GB/s and bf16 TF/s but no FP32 figure, so the direct-summation
// roofline's denominator is measured here instead of quoted: an FFMA-only kernel, 8 independent
// dependency chains per thread, 1024 threads x 2 CTAs resident per SM (16 warps per scheduler),
// no memory traffic inside the loop.  Reported as 2 flop x FFMA lane-operations / CUDA-event time.
#include "common.cuh"

namespace gh {

__global__ void __launch_bounds__(1024, 2) ffma_probe_kernel(float a, float b, int iters, float *out) {
  float c0 = threadIdx.x, c1 = c0 + 1.f, c2 = c0 + 2.f, c3 = c0 + 3.f, c4 = c0 + 4.f, c5 = c0 + 5.f, c6 = c0 + 6.f,
        c7 = c0 + 7.f;
#pragma unroll 1
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 16; u++) {
      c0 = fmaf(c0, a, b); c1 = fmaf(c1, a, b); c2 = fmaf(c2, a, b); c3 = fmaf(c3, a, b);
      c4 = fmaf(c4, a, b); c5 = fmaf(c5, a, b); c6 = fmaf(c6, a, b); c7 = fmaf(c7, a, b);
    }
  }
  const float s = ((c0 + c1) + (c2 + c3)) + ((c4 + c5) + (c6 + c7));
  if (s == 123.456f) out[0] = s;  // keeps the chains alive; practically never taken
}

}  // namespace gh

extern "C" int gh_fp32_fma_probe(int repeats, double *tflops) {
  using namespace gh;
  if (!tflops || repeats <= 0) { set_error("gh_fp32_fma_probe: bad arguments"); return GH_EINVAL; }
  *tflops = 0.0;
  int dev = 0, sms = 0;
  GH_CUDA(cudaGetDevice(&dev));
  GH_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  float *out = nullptr;
  GH_CUDA(cudaMalloc(&out, sizeof(float)));
  cudaEvent_t e0, e1;
  GH_CUDA(cudaEventCreate(&e0));
  GH_CUDA(cudaEventCreate(&e1));
  const int iters = 4096, grid = sms * 2 * 4;
  const double flop = 2.0 * 8.0 * 16.0 * (double)iters * 1024.0 * (double)grid;
  double best = 0.0;
  int rc = GH_OK;
  for (int r = 0; r < repeats + 1 && rc == GH_OK; r++) {  // first launch = warm-up
    cudaEventRecord(e0, 0);
    ffma_probe_kernel<<<grid, 1024>>>(0.999f, 1e-3f, iters, out);
    cudaEventRecord(e1, 0);
    cudaError_t ce = cudaEventSynchronize(e1);
    if (ce != cudaSuccess) { set_error("gh_fp32_fma_probe: %s", cudaGetErrorString(ce)); rc = GH_ECUDA; break; }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (r > 0 && ms > 0.f) { const double tf = flop / (ms * 1e-3) / 1e12; if (tf > best) best = tf; }
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  *tflops = best;
  return rc;
}
